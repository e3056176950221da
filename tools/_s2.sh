mkdir -p gpurun_out
N=${NG:-4}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02B_bench$N.json 2> gpurun_out/r02B_bench$N.err; echo "bench rc=$?" >> gpurun_out/r02B_bench$N.err
tail -2 gpurun_out/r02B_bench$N.err
