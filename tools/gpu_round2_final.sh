#!/bin/bash
# Final-state evidence in one gpurun call: all GPU tests, smoke, bench (+ reference arm), step timing.
#   gpurun --timeout 1800 -- 'bash tools/gpu_round2_final.sh r02Y'
tag=${1:-r02Y}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $out/${tag}_gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest_gpu.log
tail -3 $out/${tag}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/${tag}_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
cut -c1-300 $out/${tag}_bench.json
timeout 500 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err; echo "bench ref rc=$?"
cut -c1-300 $out/${tag}_bench_ref.json
timeout 300 python tools/time_step.py > $out/${tag}_time_step.log 2>&1; tail -3 $out/${tag}_time_step.log
