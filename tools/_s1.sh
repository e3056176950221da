mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gp_engine_gpu.py -m gpu -x -q -k "opt or sweep or brent or scheme or branch" > gpurun_out/r02r_pytest_opt.log 2>&1; echo "rc=$?" >> gpurun_out/r02r_pytest_opt.log
tail -4 gpurun_out/r02r_pytest_opt.log
SWEEP_VARIANTS="0@PREP_LEAN=0,0" timeout 600 python tools/sweep_variants.py synthetic-1000taxa-1Mpat-5000trees > gpurun_out/r02r_sweep_variants_1000.log 2>&1
cat gpurun_out/r02r_sweep_variants_1000.log
timeout 600 python tools/profile_sweep.py > gpurun_out/r02r_profile_sweep.log 2>&1; cat gpurun_out/r02r_profile_sweep.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_opt_prepare_ratio" -s 0 -c 1 -o gpurun_out/r02r_k_opt_prepare -f python profiles/prof_pass.py synthetic-1000taxa-1Mpat-5000trees 40000 1 sweep > gpurun_out/r02r_ncu.log 2>&1
