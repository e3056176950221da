mkdir -p gpurun_out
SWEEP_VARIANTS="0,0@OPT_SEGMENTS=4,0@OPT_SEGMENTS=2,0@OPT_FIRST_CHECK=6,0@OPT_FIRST_CHECK=8,0@OPT_EVAL_OCC=3" timeout 600 python tools/sweep_variants.py synthetic-1000taxa-1Mpat-5000trees > gpurun_out/r02C_sweep_variants_1000.log 2>&1
cat gpurun_out/r02C_sweep_variants_1000.log
