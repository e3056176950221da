out=gpurun_out; mkdir -p $out; tag=r01n
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/multi_gpu_parity.py > $out/${tag}_parity_peer.log 2>&1; echo "parity peer rc=$?"; tail -4 $out/${tag}_parity_peer.log
BITO_GP_PEER_ALLREDUCE=0 timeout 600 $TR tools/multi_gpu_parity.py > $out/${tag}_parity_nccl.log 2>&1; echo "parity nccl rc=$?"; tail -3 $out/${tag}_parity_nccl.log
timeout 600 $TR tools/multi_gpu_time.py > $out/${tag}_time_peer.log 2>&1; echo "time peer rc=$?"; tail -2 $out/${tag}_time_peer.log
BITO_GP_PEER_ALLREDUCE=0 timeout 600 $TR tools/multi_gpu_time.py > $out/${tag}_time_nccl.log 2>&1; echo "time nccl rc=$?"; tail -2 $out/${tag}_time_nccl.log
