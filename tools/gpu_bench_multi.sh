# Multi-GPU evidence on N GPUs of one box: sharded pytest, bench (ours + reference arm) under torchrun.   gpurun --gpus N -- bash tools/gpu_bench_multi.sh N
set -x
N=${1:-2}
O=gpurun_out/multi
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q -k "shard" > $O/pytest_gpu_${N}gpu_sharded.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_${N}gpu_sharded.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_n${N}.json 2> $O/bench_n${N}.err; echo "bench rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --impl reference --steps 3 --warmup 1 > $O/bench_n${N}_reference_arm.json 2> $O/bench_n${N}_ref.err; echo "ref rc=$?"
tail -5 $O/bench_n${N}.err
