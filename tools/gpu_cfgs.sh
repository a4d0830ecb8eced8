#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 90 python bench.py --config c2 --steps 2 --warmup 2 --no-cpu-baseline --no-also-fast > gpurun_out/final_c2.json 2> gpurun_out/final_c2.err; echo "c2 rc=$?"; cut -c1-200 gpurun_out/final_c2.json
timeout 140 python bench.py --config c5 --steps 1 --warmup 2 --no-cpu-baseline --no-also-fast > gpurun_out/final_c5.json 2> gpurun_out/final_c5.err; echo "c5 rc=$?"; cut -c1-200 gpurun_out/final_c5.json
