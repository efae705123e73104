"""The C-ABI library loads and exports every symbol include/ehgr_b200.h declares (no compute)."""
import ctypes
import re
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (REPO / "include" / "ehgr_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ehgr_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_symbols():
    syms = _declared_symbols()
    assert "ehgr_temporal_shift_fwd" in syms and "ehgr_abi_version" in syms


def test_library_exports_every_declared_symbol():
    import ehgr_b200
    lib = ehgr_b200._lib.lib()
    for s in _declared_symbols():
        assert hasattr(lib, s), f"libehgr_b200.so does not export {s}"
    assert lib.ehgr_abi_version() == ehgr_b200._lib.ABI_VERSION == 2
    assert ehgr_b200._lib.lib().ehgr_status_string(-4).decode().startswith("invalid")


def test_binding_table_covers_header():
    import ehgr_b200
    declared = set(_declared_symbols()) - {"ehgr_abi_version", "ehgr_status_string", "ehgr_launch_count", "ehgr_stream_t"}
    bound = set(ehgr_b200._lib.SIGNATURES)
    assert declared == bound, (declared - bound, bound - declared)


def test_argument_errors_do_not_need_a_gpu():
    import ehgr_b200
    lib = ehgr_b200._lib.lib()
    assert lib.ehgr_temporal_shift_fwd(None, None, 1, 8, 8, 4, 1, 0, 0, None) == -1          # null
    buf = ctypes.create_string_buffer(64)
    p = ctypes.addressof(buf)
    assert lib.ehgr_temporal_shift_fwd(p, p + 32, 1, 8, 8, 4, 1, 7, 0, None) == -3             # dtype
    assert lib.ehgr_temporal_shift_fwd(p, p + 32, 1, 0, 8, 4, 1, 0, 0, None) == -4             # shape
    assert lib.ehgr_temporal_shift_fwd(p, p + 32, 1, 8, 8, 4, 5, 0, 0, None) == -4             # 2*fold > c
    assert lib.ehgr_temporal_shift_fwd(p + 1, p + 32, 1, 8, 8, 4, 1, 0, 0, None) == -2         # alignment
    assert lib.ehgr_temporal_shift_fwd(p, p + 32, 0, 8, 8, 4, 1, 0, 0, None) == 0              # empty batch


def test_built_objects_contain_the_blackwell_instructions():
    """SASS evidence (no GPU needed): the GEMMs issue tcgen05 MMAs (UTCHMMA) fed by TMA boxes (UTMALDG: plain / affine /
    gated operands, weight slices, 3x3 im2col) and cp.async gathers (LDGSTS: temporal shift), the depthwise and stem
    kernels load their tiles with TMA, and the FP32 inner loops use packed FFMA2."""
    import shutil
    import subprocess
    from pathlib import Path
    if shutil.which("cuobjdump") is None:
        import pytest
        pytest.skip("cuobjdump not available")
    import ehgr_b200
    build = Path(ehgr_b200._lib.LIB_PATH).parent / "build"
    if not (build / "pw_tc.o").exists():
        ehgr_b200.build.build()
    want = {"pw_tc.o": ("UTCHMMA", "LDGSTS", "UTCBAR", "UTMALDG"), "pw_tc_wgrad.o": ("UTCHMMA", "LDGSTS", "UTMALDG"),
            "dw_sw.o": ("UTMALDG", "FFMA2"), "stem.o": ("UTMALDG", "FFMA2"), "bn.o": ("FFMA2",)}
    for obj, mnemonics in want.items():
        sass = subprocess.run(["cuobjdump", "-sass", str(build / obj)], capture_output=True, text=True, timeout=300).stdout
        for m in mnemonics:
            assert m in sass, (obj, m)
