#!/bin/bash
# Round-2 evidence in one gpurun call: GPU tests, smoke, bench (+ reference arm), per-kernel profile of the sweep,
# ncu launch lists and --set full captures of the top kernels. Everything lands in gpurun_out/<tag>_*.
#   gpurun --timeout 2400 -- 'bash tools/gpu_round2.sh r02f'
tag=${1:-r02f}
out=gpurun_out
mkdir -p $out
WL=synthetic-1000taxa-1Mpat-5000trees
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $out/${tag}_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest_gpu.log
tail -3 $out/${tag}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/${tag}_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
cut -c1-300 $out/${tag}_bench.json
timeout 500 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err; echo "bench ref rc=$?"
cut -c1-300 $out/${tag}_bench_ref.json
timeout 300 python tools/profile_sweep.py > $out/${tag}_profile_sweep.log 2>&1; cat $out/${tag}_profile_sweep.log
timeout 300 python tools/time_step.py > $out/${tag}_time_step.log 2>&1; tail -3 $out/${tag}_time_step.log
# launch list of one pass + one batched sweep at full size (graphs off): durations only (one ncu pass, no replay)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $out/${tag}_launches_step.csv \
  python profiles/prof_pass.py $WL 125000 1 sweep > $out/${tag}_launches_step.log 2>&1; echo "ncu list rc=$?"
# the same with DRAM bytes at 40 000 patterns (ncu's save/restore of 127 GB makes multi-pass metrics at full size take minutes per launch)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv \
  --log-file $out/${tag}_launches_step_40k.csv python profiles/prof_pass.py $WL 40000 1 sweep > $out/${tag}_launches_step_40k.log 2>&1; echo "ncu list 40k rc=$?"
# full captures at 40 000 patterns
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_node -s 1 -c 2 -f -o $out/${tag}_k_node \
  python profiles/prof_pass.py $WL 40000 1 > $out/${tag}_ncu_node.log 2>&1; echo "ncu node rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_likelihood -c 1 -f -o $out/${tag}_k_likelihood \
  python profiles/prof_pass.py $WL 40000 1 > $out/${tag}_ncu_lik.log 2>&1; echo "ncu lik rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_opt_eval_model|k_opt_prepare_ratio" -c 3 -f -o $out/${tag}_k_opt \
  python profiles/prof_pass.py $WL 40000 1 sweep > $out/${tag}_ncu_opt.log 2>&1; echo "ncu opt rc=$?"
ls -la $out | tail -24
