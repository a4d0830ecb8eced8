"""CPU suite: csrc/philox.cuh (the production randomness of the AR sampler and the NAR Gumbel draws) against the published
known-answer vectors of Philox4x32-10 (Random123 kat_vectors: zero, all-ones and the pi-digit counter / key).  The header is
compiled as host C++ exactly as the CUDA sources include it."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KAT = [((0, 0), (0, 0, 0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
       ((0xffffffff, 0xffffffff), (0xffffffff,) * 4, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
       ((0xa4093822, 0x299f31d0), (0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs a host C++ compiler")
def test_philox4x32_10_known_answers(tmp_path):
    src = tmp_path / "kat.cpp"
    src.write_text('#define __host__\n#define __device__\n#include <cstdio>\n#include "philox.cuh"\n'
                   "int main() { unsigned o[4];\n" +
                   "".join(f"  m5::philox4x32({k[0]}u, {k[1]}u, {c[0]}u, {c[1]}u, {c[2]}u, {c[3]}u, o); "
                           'std::printf("%08x %08x %08x %08x\\n", o[0], o[1], o[2], o[3]);\n' for k, c, _ in KAT) + "  return 0; }\n")
    exe = tmp_path / "kat"
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "mars5-tts_b200", "csrc"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    for line, (_, _, want) in zip(out, KAT):
        assert [int(x, 16) for x in line.split()] == list(want), line
