mkdir -p gpurun_out
SWEEP_VARIANTS="auto" timeout 900 python tools/sweep_variants.py synthetic-1000taxa-1Mpat-5000trees - gauss_seidel > gpurun_out/r02H_sweep_variants_gs.log 2>&1
cat gpurun_out/r02H_sweep_variants_gs.log
SWEEP_VARIANTS="auto" timeout 900 python tools/sweep_variants.py synthetic-200taxa-100kpat-1000trees - gauss_seidel >> gpurun_out/r02H_sweep_variants_gs.log 2>&1
tail -2 gpurun_out/r02H_sweep_variants_gs.log
