"""Stage the UNMODIFIED reference files of the hot path into ``oracle/_ref/`` (git-ignored, rides ``gpurun``).

    python -m oracle.stage_ref            # needs /root/reference (the build container); a no-op elsewhere

The reference is pure Python, so "building" it is copying the files its decoder path imports (decoders.py
and what it pulls in, the flow loss, the RAdam optimizer, the STFT front end) byte for byte, with a manifest of their
SHA-256 digests.  Nothing here is product source and nothing under ``rad-mmm_b200/`` may import it: it is the CPU
baseline of ``bench.py --impl reference`` (``cpu_baseline.kind = "reference"``) and a second checker for the tests.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("RADMMM_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = [
    "decoders.py", "common.py", "splines.py", "partialconv1d.py", "maskedbatchnorm1d.py", "alignment.py", "utils.py",
    "loss.py", "stft_loss.py", "radam.py", "audio_processing.py", "attribute_predictors.py", "models/radmmm.py",
    "vocoders/hifigan_models.py", "vocoders/hifigan_env.py", "vocoders/hifigan_utils.py", "LICENSE",
]


def stage(verbose: bool = True) -> bool:
    if not os.path.isdir(SRC):
        if verbose:
            print(f"stage_ref: {SRC} not present; keeping {DST} as is")
        return os.path.isdir(DST)
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        if not os.path.exists(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "files": manifest}, f, indent=1)
    if verbose:
        print(f"stage_ref: {len(manifest)} reference files -> {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
