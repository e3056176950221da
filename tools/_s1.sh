mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_tp_search_parity.py -m gpu -q -s > gpurun_out/r02D_pytest_tp_search.log 2>&1; echo "rc=$?" >> gpurun_out/r02D_pytest_tp_search.log
grep -E "^ok|^FAIL|passed|failed|rc=" gpurun_out/r02D_pytest_tp_search.log | cut -c1-230
