#!/bin/bash
# Same-box A/B of the tcgen05 attention kernels: working tree vs mars5-tts_b200/lib/variants/libmars5_b200_base.so
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python tools/attn_check.py > gpurun_out/attn_check.log 2>&1; echo "attn_check rc=$?"; tail -8 gpurun_out/attn_check.log
timeout 200 python -m pytest tests/test_kernels_gpu.py tests/test_zzz_fullsize_gpu.py -m gpu -q -s -k "attn or attention" > gpurun_out/attn_tests.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/attn_tests.log
for i in 1 2; do
  timeout 100 python tools/attn_bench.py split 2>&1 | sed 's/^/new  /'
  M5_LIB_PATH=mars5-tts_b200/lib/variants/libmars5_b200_base.so timeout 100 python tools/attn_bench.py split 2>&1 | sed 's/^/base /'
done | tee gpurun_out/attn_ab.log
