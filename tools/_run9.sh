mkdir -p gpurun_out
nvidia-smi -L | wc -l
for cfg in "8 peer" "8 nccl" "4 peer"; do set -- $cfg
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $1 --steps 10 --warmup 3 --exchange $2 > gpurun_out/bench_n$1_$2.json 2> gpurun_out/bench_n$1_$2.err; tail -3 gpurun_out/bench_n$1_$2.err; cat gpurun_out/bench_n$1_$2.json
done
