mkdir -p gpurun_out
timeout 600 python tools/time_step.py > gpurun_out/r02u_time_step.log 2>&1; cat gpurun_out/r02u_time_step.log
