#!/bin/bash
# Multi-GPU session: tools/gpu_session_multi.sh <tag> <n_gpus>
tag=$1; n=$2; out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > $out/${tag}_gpu.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29521 tools/multi_gpu_parity.py > $out/${tag}_parity${n}.log 2>&1; echo "parity rc=$?" >> $out/${tag}_parity${n}.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $n --steps 5 --warmup 3 > $out/${tag}_bench${n}.json 2> $out/${tag}_bench${n}.err; echo "bench rc=$?" >> $out/${tag}_bench${n}.err
tail -3 $out/${tag}_parity${n}.log; tail -3 $out/${tag}_bench${n}.err
