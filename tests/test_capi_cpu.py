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


def test_struct_sizes_match_header_layout(tmp_path):
    """sizeof / offsetof of every struct of the header as gcc lays them out == the ctypes mirrors in capi.py."""
    import ctypes as C
    src = tmp_path / "sz.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "mars5_b200.h"\n'
        'int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(m5_model_cfg), sizeof(m5_ar_cfg), sizeof(m5_nar_cfg), '
        'sizeof(m5_tensor), offsetof(m5_ar_cfg, typical_p), offsetof(m5_nar_cfg, schedule), offsetof(m5_nar_cfg, jump_len)); return 0; }\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert got == [C.sizeof(capi.ModelCfg), C.sizeof(capi.ArCfg), C.sizeof(capi.NarCfg), C.sizeof(capi.Tensor),
                   capi.ArCfg.typical_p.offset, capi.NarCfg.schedule.offset, capi.NarCfg.jump_len.offset], got
    assert C.sizeof(capi.ModelCfg) == 4 * (8 + 2 + 9 + 5 + 9)


def test_built_for_sm100a_with_tcgen05_and_tma():
    """The shipped SASS contains the Blackwell instructions the design claims (UTCHMMA = tcgen05.mma kind::f16, UTCQMMA = kind::f8f6f4 -- the
    fp8 lo pass of the mixed8 numerics --, .2CTA = cta_group::2, UTMALDG = TMA, LDTM = tcgen05.ld, LDGSTS = cp.async of the decode kernel's weight ring)."""
    out = subprocess.run(["cuobjdump", "-sass", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out or "SM100a" in out.upper() or "sm_100" in out
    for mnemonic in ("UTCHMMA", "UTCHMMA.2CTA", "UTCQMMA.2CTA", "UTMALDG.2D", "LDTM", "LDGSTS"):
        assert mnemonic in out, mnemonic
