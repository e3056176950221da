mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_round2_gpu.py -m gpu -q -x -k "taylor_model_sweep_matches_oracle and 14-4099" > gpurun_out/r02J_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02J_racecheck.log
grep -E "passed|failed|RACECHECK SUMMARY|rc=" gpurun_out/r02J_racecheck.log | cut -c1-200
grep -E "^=========     at " gpurun_out/r02J_racecheck.log | sort | uniq -c | sort -rn | head -5 | cut -c1-220
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gp_engine_gpu.py -m gpu -q -x -k "cluster_optimizer_sweeps_match_reference and five_taxon" > gpurun_out/r02J_racecheck_old.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02J_racecheck_old.log
grep -E "passed|failed|RACECHECK SUMMARY|rc=" gpurun_out/r02J_racecheck_old.log | cut -c1-200
grep -E "^=========     at " gpurun_out/r02J_racecheck_old.log | sort | uniq -c | sort -rn | head -5 | cut -c1-220
