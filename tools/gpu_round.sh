#!/bin/bash
# One gpurun call's worth of evidence for the current build: GPU parity tests, bench line,
# ncu launch list (with DRAM bytes) of one pass + sweep, and --set full captures of the top kernels.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r01d'
tag=${1:-r01x}
out=gpurun_out
mkdir -p $out
WL=${2:-synthetic-1000taxa-1Mpat-5000trees}
PAT=${3:-125000}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $out/${tag}_gpu.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest.log
tail -3 $out/${tag}_pytest.log
timeout 300 python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/${tag}_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
cat $out/${tag}_bench.json | cut -c1-600
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err; echo "bench ref rc=$?"
cat $out/${tag}_bench_ref.json | cut -c1-400
timeout 300 python tools/time_small.py > $out/${tag}_time_small.log 2>&1; cat $out/${tag}_time_small.log
timeout 300 python tools/profile_sweep.py $WL $PAT > $out/${tag}_profile_sweep.log 2>&1; tail -12 $out/${tag}_profile_sweep.log
# launch list: one pass + one batched sweep, graphs off
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6000 --csv \
  --log-file $out/${tag}_launches_pass.csv python profiles/prof_pass.py $WL $PAT 1 > $out/${tag}_prof_pass.log 2>&1; echo "ncu list rc=$?"
# launch list of the reference's Gauss-Seidel sweep at 100k patterns (3000 launches from inside the sweep)
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 3000 --csv --log-file $out/${tag}_launches_gs.csv \
  python profiles/prof_pass.py synthetic-200taxa-100kpat-1000trees 100000 1 gs > $out/${tag}_prof_gs.log 2>&1; echo "ncu gs list rc=$?"
[ -n "$QUICK" ] && { ls -la $out; exit 0; }
# full captures (at 40k / 20k patterns so that ncu's save/restore of device memory stays small): the two
# largest k_node launches (first rootward levels), the likelihood kernel, the sweep objective
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_node -s 1 -c 2 -f -o $out/${tag}_k_node \
  python profiles/prof_pass.py $WL 40000 1 > $out/${tag}_ncu_node.log 2>&1; echo "ncu node rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_likelihood -c 1 -f -o $out/${tag}_k_likelihood \
  python profiles/prof_pass.py $WL 20000 1 > $out/${tag}_ncu_lik.log 2>&1; echo "ncu lik rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_opt -s 4 -c 4 -f -o $out/${tag}_k_opt \
  python profiles/prof_pass.py $WL 40000 1 sweep > $out/${tag}_ncu_opt.log 2>&1; echo "ncu opt rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_opt_cluster -s 200 -c 2 -f -o $out/${tag}_k_opt_cluster \
  python profiles/prof_pass.py synthetic-200taxa-100kpat-1000trees 100000 1 gs > $out/${tag}_ncu_cluster.log 2>&1; echo "ncu cluster rc=$?"
ls -la $out
