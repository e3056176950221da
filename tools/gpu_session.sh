#!/bin/bash
# One gpurun call of round 2: tests, bench, sweep variants, ncu captures. Everything lands in gpurun_out/<tag>_*.
# usage: tools/gpu_session.sh <tag> [steps...]   steps: tests bench variants ncu_sweep ncu_pass
tag=$1; shift
steps="${@:-tests bench variants}"
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $out/${tag}_gpu.txt 2>&1
for s in $steps; do
  case $s in
    tests)
      timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log ;;
    tests_all)
      timeout 1800 python -m pytest tests -m gpu -q > $out/${tag}_pytest_all.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_all.log ;;
    tests_new)
      timeout 1200 python -m pytest tests/test_round2_gpu.py tests/test_pybito_gpu.py -m gpu -q > $out/${tag}_pytest_new.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_new.log ;;
    bench)
      timeout 900 python bench.py --steps 10 --warmup 3 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?" >> $out/${tag}_bench.err ;;
    bench_ref)
      timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err ;;
    variants)
      SWEEP_VARIANTS="${SWEEP_VARIANTS:-0,16,p,p16,p16x512,p12,p9x512,p8x1024,p6x1024}" timeout 900 python tools/sweep_variants.py synthetic-1000taxa-1Mpat-5000trees > $out/${tag}_sweep_variants_1000.log 2>&1 ;;
    variants200)
      SWEEP_VARIANTS="${SWEEP_VARIANTS200:-0,p,p8,p16}" timeout 600 python tools/sweep_variants.py synthetic-200taxa-100kpat-1000trees > $out/${tag}_sweep_variants_200.log 2>&1 ;;
    ncu_sweep)
      BITO_GP_OPT_SCHEME=3 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_opt_cluster|k_opt_prepare_cluster" -s 2 -c 4 \
        -o $out/${tag}_k_opt_pipelined -f python profiles/prof_pass.py synthetic-1000taxa-1Mpat-5000trees 40000 1 sweep > $out/${tag}_ncu_sweep.log 2>&1 ;;
    node_variants)
      for v in ${NODE_VARIANTS:-X=0 BITO_GP_NODE_PREFETCH=1 BITO_GP_NODE_PREFETCH=2 BITO_GP_NODE_PREFETCH=4 BITO_GP_PREP_PREFETCH=1 BITO_GP_PREP_PREFETCH=2}; do
        env $v timeout 300 python tools/time_pass.py synthetic-1000taxa-1Mpat-5000trees - 5 sweep >> $out/${tag}_time_pass.log 2>&1
      done ;;
    ncu_small)
      BITO_GP_OPT_SCHEME=3 BITO_GP_OPT_CLUSTER=16 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_opt_cluster|k_opt_prepare_cluster" -c 2 \
        -o $out/${tag}_k_opt_pipelined -f python profiles/prof_pass.py synthetic-small 125000 1 sweep > $out/${tag}_ncu_small.log 2>&1
      BITO_GP_OPT_SCHEME=3 BITO_GP_OPT_CLUSTER=16 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none --csv \
        --log-file $out/${tag}_launches_small.csv python profiles/prof_pass.py synthetic-small 125000 1 sweep > $out/${tag}_launches_small.log 2>&1 ;;
    ncu_pass)
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_node" -s 40 -c 2 \
        -o $out/${tag}_k_node -f python profiles/prof_pass.py synthetic-1000taxa-1Mpat-5000trees 40000 1 > $out/${tag}_ncu_pass.log 2>&1 ;;
    launches)
      timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
        --log-file $out/${tag}_launches_step.csv python profiles/prof_pass.py synthetic-1000taxa-1Mpat-5000trees 125000 1 sweep > $out/${tag}_launches_step.log 2>&1 ;;
  esac
done
ls -la $out | tail -20
