mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_parity.py > gpurun_out/r02v_parity2.log 2>&1; echo "parity rc=$?" >> gpurun_out/r02v_parity2.log
tail -12 gpurun_out/r02v_parity2.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02v_bench2.json 2> gpurun_out/r02v_bench2.err; echo "bench rc=$?" >> gpurun_out/r02v_bench2.err
tail -3 gpurun_out/r02v_bench2.err; cut -c1-400 gpurun_out/r02v_bench2.json
