#!/bin/bash
# One gpurun call: GPU test suite (parity figures printed), attention micro-benchmark of the pair / single-key kernels, one ncu
# capture of the single-key kernel, one bench step per numerics mode on the same box.  Outputs under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s)
timeout 600 python -m pytest tests -m gpu -q -s --durations=8 > gpurun_out/tests.log 2>&1; echo "tests rc=$? $(( $(date +%s) - t0 )) s"
tail -5 gpurun_out/tests.log
timeout 180 python tools/attn_bench.py split > gpurun_out/attn_split.log 2>&1; echo "attn rc=$?"; cat gpurun_out/attn_split.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:flash_tc5 -s 16 -c 1 -f -o gpurun_out/flash_ksingle \
  python tools/attn_bench.py split > gpurun_out/ncu_flash.log 2>&1; echo "ncu rc=$?"
timeout 420 python bench.py --steps 1 --warmup 2 --mode mixed8k --also mixed8 --no-cpu-baseline --no-e2e \
  > gpurun_out/bench_mixed8k.json 2> gpurun_out/bench_mixed8k.err; echo "bench rc=$? $(( $(date +%s) - t0 )) s"
cat gpurun_out/bench_mixed8k.json | cut -c1-1500
