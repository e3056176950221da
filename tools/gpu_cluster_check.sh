#!/bin/bash
# GPU check of the cluster-resident optimiser: its parity tests first (bounded: a hung cluster barrier must not
# eat the box), then the whole GPU suite, then sweep timings of every optimiser scheme on the bench workloads.
#   gpurun --timeout 1500 -- 'bash tools/gpu_cluster_check.sh r01f'
tag=${1:-r01x}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $out/${tag}_gpu.txt 2>&1
timeout 420 python -m pytest tests/test_gp_engine_gpu.py -m gpu -x -q -k "cluster or scheme" > $out/${tag}_pytest_cluster.log 2>&1
rc=$?; echo "cluster tests rc=$rc"; tail -15 $out/${tag}_pytest_cluster.log
if [ $rc -ne 0 ]; then export SWEEP_VARIANTS=0; export BITO_GP_OPT_CLUSTER=0; fi
timeout 600 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $out/${tag}_pytest.log
WL=synthetic-200taxa-100kpat-1000trees
SWEEP_VARIANTS=${SWEEP_VARIANTS:-0,auto} timeout 300 python tools/sweep_variants.py $WL > $out/${tag}_sweep_variants_200.log 2>&1
echo "variants200 rc=$?"; tail -4 $out/${tag}_sweep_variants_200.log
SWEEP_VARIANTS=${SWEEP_VARIANTS:-0,1,16,4x1024,16x1024,auto} timeout 400 python tools/sweep_variants.py $WL 20000 gauss_seidel > $out/${tag}_sweep_variants_gs20k.log 2>&1
echo "variantsGS20k rc=$?"; tail -8 $out/${tag}_sweep_variants_gs20k.log
SWEEP_VARIANTS=${SWEEP_VARIANTS:-0,16,8x1024,16x1024,auto} timeout 600 python tools/sweep_variants.py $WL - gauss_seidel > $out/${tag}_sweep_variants_gs100k.log 2>&1
echo "variantsGS100k rc=$?"; tail -7 $out/${tag}_sweep_variants_gs100k.log
timeout 300 python tools/time_small.py > $out/${tag}_time_small.log 2>&1; cat $out/${tag}_time_small.log
