#!/bin/bash
# A/B of programmatic dependent launch (BITO_GP_PDL=0|1): parity tests with it on, then the Gauss-Seidel sweep,
# the pass + batched sweep and the small real-data shapes with it off and on.
#   gpurun --timeout 1500 -- 'bash tools/gpu_pdl_ab.sh r02P'
tag=${1:-r02P}
out=gpurun_out
mkdir -p $out
WL=synthetic-1000taxa-1Mpat-5000trees
timeout 900 python -m pytest tests/test_gp_engine_gpu.py tests/test_round2_gpu.py tests/test_models_gpu.py -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest.log
tail -3 $out/${tag}_pytest.log
SWEEP_VARIANTS="auto@PDL=0,auto" timeout 400 python tools/sweep_variants.py $WL - gauss_seidel > $out/${tag}_sweep_variants_gs.log 2>&1; cat $out/${tag}_sweep_variants_gs.log
for v in 0 1; do
  BITO_GP_PDL=$v timeout 300 python tools/time_pass.py $WL - 5 sweep >> $out/${tag}_time_pass.log 2>&1
  BITO_GP_PDL=$v timeout 300 python tools/time_small.py ds1_config1 fluA >> $out/${tag}_time_small.log 2>&1
done
cat $out/${tag}_time_pass.log $out/${tag}_time_small.log
