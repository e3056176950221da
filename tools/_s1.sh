mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gp_engine_gpu.py tests/test_round2_gpu.py tests/test_host_shim_gpu.py tests/test_nni_parity_gpu.py -m gpu -q -x > gpurun_out/r02G_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02G_pytest.log
tail -4 gpurun_out/r02G_pytest.log
for v in "BITO_GP_FUSE_SMALL_LEVELS=0" "X=1"; do echo "== $v"; env $v timeout 300 python tools/time_small.py ds1_config1 fluA 2>&1 | grep cuda; done > gpurun_out/r02G_time_small.log 2>&1; cat gpurun_out/r02G_time_small.log
SWEEP_VARIANTS="16x512@FUSE_SMALL_LEVELS=0,16x512,auto" timeout 900 python tools/sweep_variants.py synthetic-1000taxa-1Mpat-5000trees - gauss_seidel > gpurun_out/r02G_sweep_variants_gs.log 2>&1
cat gpurun_out/r02G_sweep_variants_gs.log
