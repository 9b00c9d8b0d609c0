# multi-GPU checks (gpurun --gpus N): gathered-list parity vs the oracle, then the bench line at N GPUs
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_verify_gather.py 2>&1 | grep -v "^W\|warn" | tail -12 | tee gpurun_out/multi_verify_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 50 --warmup 5 2> gpurun_out/multi_bench_n$N.err | tail -2 | tee gpurun_out/multi_bench_n$N.json
tail -3 gpurun_out/multi_bench_n$N.err
