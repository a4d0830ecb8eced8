"""CPU suite: the C-ABI library loads and exports every symbol include/mars5_b200.h declares (no compute calls)."""
import os
import re
import subprocess

from mars5_tts_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "mars5_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(m5_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    declared = _header_symbols()
    assert len(declared) >= 20
    nm = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (m5_[a-z0-9_]+)", nm))
    missing = [s for s in declared if s not in exported]
    assert not missing, missing
    for s in declared:
        assert hasattr(lib, s)


def test_ctypes_table_matches_header():
    assert sorted(capi.exported_symbols()) == _header_symbols()


def test_struct_sizes_match_header_layout():
    import ctypes as C
    # m5_model_cfg: 8 int32 + 2 float + 9 int32 + 5 float + 9 int32 ; m5_ar_cfg: 12 x 4 bytes ; m5_nar_cfg: 6 x 4 + pointer
    assert C.sizeof(capi.ModelCfg) == 4 * (8 + 2 + 9 + 5 + 9)
    assert C.sizeof(capi.ArCfg) == 48
    assert C.sizeof(capi.NarCfg) == 32
    assert C.sizeof(capi.Tensor) == 32


def test_built_for_sm100a_with_tcgen05_and_tma():
    """The shipped SASS contains the Blackwell instructions the design claims (UTCHMMA = tcgen05.mma, UTMALDG = TMA, LDTM)."""
    out = subprocess.run(["cuobjdump", "-sass", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out or "SM100a" in out.upper() or "sm_100" in out
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in out, mnemonic
