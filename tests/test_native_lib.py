"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/radmmm_b200.h declares, the ctypes
struct mirrors match the compiled layouts, and the product path refuses to run without CUDA (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from radmmm_b200 import build
    build.build()
    from radmmm_b200 import _native
    return _native.lib()


def header_functions():
    src = open(os.path.join(ROOT, "include", "radmmm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(radmmm_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    from radmmm_b200 import _native
    names = header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/radmmm_b200.h but not exported"
        assert n in _native.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(_native.SIGNATURES) == names


def test_struct_layouts(lib):
    from radmmm_b200 import _native
    assert lib.radmmm_abi_version() == _native.ABI_VERSION
    assert lib.radmmm_sizeof_flow_desc() == ctypes.sizeof(_native.FlowDesc)
    assert lib.radmmm_sizeof_flow_grads() == ctypes.sizeof(_native.FlowGrads)


def test_size_queries(lib):
    assert lib.radmmm_pitch(400) == 416
    assert lib.radmmm_rows(8, 400) == 3328
    assert lib.radmmm_rows(1, 1) == 256
    for mode in (0, 1, 2):
        p = lib.radmmm_flow_prepared_bytes(mode, 160, 1056, 1024, 4)
        w_train = lib.radmmm_flow_workspace_bytes(mode, 1, 8, 400, 160, 1056, 1024, 4)
        w_inf = lib.radmmm_flow_workspace_bytes(mode, 0, 8, 400, 160, 1056, 1024, 4)
        assert p > 2 * 26.5e6 * (4 if mode == 0 else 2)      # weights + transposes
        assert w_train > w_inf > 0
    assert lib.radmmm_context_rows_bytes(0, 8, 400, 1056) == 3328 * 1152 * 4
    assert lib.radmmm_context_rows_bytes(1, 8, 400, 1056) == 3328 * 1152 * 2
    assert lib.radmmm_context_rows_bytes(2, 8, 400, 1056) == 3328 * 1152 * 2 * 2      # hi + lo planes


def test_argument_errors_are_reported(lib):
    # no GPU work: argument validation happens before any launch
    rc = lib.radmmm_stft_mel(None, None, None, None, 1, 4096, 512, 256, 80, 1e-5, None)
    assert rc < 0 and b"n_fft" in lib.radmmm_last_error()
    rc = lib.radmmm_spline_forward(None, None, None, None, None, 1, 4, 8, 16, -3.0, 3.0, 0, None)
    assert rc < 0 and b"bins" in lib.radmmm_last_error()


def test_no_cpu_fallback():
    """The product path must fail loudly, not fall back, when there is no CUDA device / tensor."""
    from radmmm_b200 import common
    layer = common.AffineTransformationLayer(12, 10, 2, affine_model="wavenet", scaling_fn="tanh", n_channels=128,
                                             use_partial_padding=True)
    z = torch.zeros(1, 12, 8)
    ctx = torch.zeros(1, 10, 8)
    with pytest.raises(RuntimeError, match="CUDA"):
        layer(z, ctx)
    from radmmm_b200 import audio_processing
    stft = audio_processing.TacotronSTFT()
    with pytest.raises(RuntimeError, match="CUDA"):
        stft.mel_spectrogram(torch.zeros(1, 4096))


def test_integration_doc_flowdesc_matches_binding():
    """INTEGRATION.md shows the ctypes struct a maintainer would write: its field list must be the real one (round 1's
    copy had lost `side_stream`) and the ABI version it asserts must be the current one."""
    import re
    from radmmm_b200 import _native
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = text[text.index("class FlowDesc(C.Structure)"):text.index("assert C.sizeof(FlowDesc)")]
    doc_fields = re.findall(r'\("(\w+)",', block)
    assert doc_fields == [f[0] for f in _native.FlowDesc._fields_]
    assert f"radmmm_abi_version() == {_native.ABI_VERSION}" in text
