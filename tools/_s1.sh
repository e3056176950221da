mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_tp_parity.py -m gpu -q > gpurun_out/r02A_pytest_tp.log 2>&1; echo "rc=$?" >> gpurun_out/r02A_pytest_tp.log
tail -6 gpurun_out/r02A_pytest_tp.log | cut -c1-250
