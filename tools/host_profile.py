#!/usr/bin/env python
"""cProfile of the HOST side of eager train steps (the path an unmodified Lightning loop takes): where the ~12 ms of enqueue
time per step go.  Diagnostic only."""
import cProfile
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    from radmmm_b200 import decoders, loss as L, synthetic as syn
    from radmmm_b200.common import SequenceLength
    dev = torch.device("cuda", 0)
    dec = decoders.RADMMMFlow(n_speaker_dim=16, use_accent=True, n_accent_dim=8, n_text_dim=520, n_group_size=2,
                              n_mel_channels=80, n_flows=8)
    dec.load_state_dict(syn.synthetic_state_dict())
    dec = dec.to(dev).set_precision("bf16").train()
    bt = {k: v.to(dev) for k, v in syn.synthetic_batch(8, 800, tag="bench.rank0").items()}
    crit = L.RADMMMFlowLoss(1.0, 2)

    def step():
        for p in dec.parameters():
            p.grad = None
        out = dec(bt["mel"], bt["spk_vecs"], bt["context"], SequenceLength(bt["out_lens"], 800), f0=bt["f0"],
                  energy_avg=bt["energy_avg"], accent_vecs=bt["accent_vecs"])
        loss = crit(out, bt["out_lens"])["loss_mel"][0]
        loss.backward()

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(10):
        step()
    pr.disable()
    torch.cuda.synchronize()
    st = pstats.Stats(pr)
    st.sort_stats("cumulative").print_stats(45)
    st.sort_stats("tottime").print_stats(35)


if __name__ == "__main__":
    main()
