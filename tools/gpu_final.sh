#!/bin/bash
# End-of-round verification on one box: GPU suite, smoke(), the default bench line, a launch list of one NAR step.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$(date +%s)
timeout 400 python -m pytest tests -m gpu -q -s > gpurun_out/final_tests.log 2>&1; echo "tests rc=$? $(( $(date +%s) - t0 )) s"; tail -3 gpurun_out/final_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final_smoke.log
timeout 520 python bench.py --steps 2 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$? $(( $(date +%s) - t0 )) s"
cut -c1-400 gpurun_out/final_bench.json
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'flash|gemm|norm|posterior|renoise|embed|istft|voc|rope|sample' --csv \
  --log-file gpurun_out/final_launches.csv python bench.py --steps 1 --warmup 1 --T 1 --no-e2e --no-cpu-baseline --no-also-fast > gpurun_out/final_ncu_bench.log 2>&1
echo "ncu rc=$? $(( $(date +%s) - t0 )) s"; wc -l gpurun_out/final_launches.csv
