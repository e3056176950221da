mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02z_bench2.json 2> gpurun_out/r02z_bench2.err; echo "bench rc=$?" >> gpurun_out/r02z_bench2.err
tail -3 gpurun_out/r02z_bench2.err
