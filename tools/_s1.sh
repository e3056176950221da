mkdir -p gpurun_out
SWEEP_VARIANTS="auto@OPT_MODEL=0,auto,16x512" timeout 900 python tools/sweep_variants.py synthetic-1000taxa-1Mpat-5000trees - gauss_seidel > gpurun_out/r02x_sweep_variants_gs.log 2>&1
cat gpurun_out/r02x_sweep_variants_gs.log
