#!/bin/bash
# Same-box A/B of the fused AR decode kernel: working tree vs mars5-tts_b200/lib/variants/libmars5_b200_base.so
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_pipeline_gpu.py tests/test_zz_tts_gpu.py -m gpu -q -x > gpurun_out/ar_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/ar_tests.log
export M5_AR_PROFILE=1 M5_AR_BENCH_REPS=2
timeout 200 python tools/ar_decode_bench.py --P 1200 --N 130 2>&1 | grep -E "profile|rep " | sed 's/^/new  /' | tee gpurun_out/ar_ab.log
M5_LIB_PATH=mars5-tts_b200/lib/variants/libmars5_b200_base.so timeout 200 python tools/ar_decode_bench.py --P 1200 --N 130 2>&1 | grep -E "profile|rep " | sed 's/^/base /' | tee -a gpurun_out/ar_ab.log
