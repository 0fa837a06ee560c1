mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}_peer_shm.json 2> gpurun_out/bench_n${N}_peer_shm.err; grep -v "^W\|OMP\|\*\*\*" gpurun_out/bench_n${N}_peer_shm.err | tail -3; python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read()); print('value %.4e ms %.3f kern %.3f e2e %.4e (%.3f ms) equal=%s %s'%(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'], d['e2e']['ms_per_step'], d['config']['e2e_equals_device_run'], d['config']['e2e_host_output']))" gpurun_out/bench_n${N}_peer_shm.json
