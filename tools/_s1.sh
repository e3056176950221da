mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_round2_gpu.py -m gpu -q -k "taylor" > gpurun_out/r02y_pytest_taylor.log 2>&1; echo "rc=$?" >> gpurun_out/r02y_pytest_taylor.log
tail -12 gpurun_out/r02y_pytest_taylor.log | cut -c1-300
