# Round-2 evidence run on one B200 (everything lands in gpurun_out/evidence; summaries are copied to profiles/ by hand).
set -x
O=gpurun_out/evidence
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
for n in 64 1024 4096; do python tools/latency.py --tracks $n --mode auto >> $O/realtime_latency.log 2>&1; done
python tools/latency.py --tracks 1024 --mode exact >> $O/realtime_latency.log 2>&1
WBX_FIR=fft python tools/kbench.py --tracks 256 --blocks 1 --reverb 65536 --fpl 4 --iters 50 >> $O/kbench_cfg5_realtime.log 2>&1
WBX_FIR=fft python tools/kbench.py --tracks 256 --blocks 4 --reverb 65536 --fpl 4 --iters 50 >> $O/kbench_cfg5_realtime.log 2>&1
cat $O/realtime_latency.log $O/kbench_cfg5_realtime.log
timeout 900 python tools/bounce_bench.py --tracks 1024 --minutes 10 > $O/bounce_bench.log 2>&1; tail -8 $O/bounce_bench.log
NCU="ncu --set full --clock-control none --import-source on -c 1"
timeout 400 $NCU -k regex:mix_kernel --launch-skip 2 -o $O/mix_cfg2 python tools/kbench.py --tracks 1024 --blocks 4096 --fpl 16 --iters 1 > $O/ncu_mix_cfg2.out 2>&1
timeout 400 $NCU -k regex:mix_kernel --launch-skip 2 -o $O/mix_cfg3 python tools/kbench.py --tracks 1024 --blocks 4096 --rate 44100 --fpl 16 --iters 1 > $O/ncu_mix_cfg3.out 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^(?!.*interleave_sample).*" -c 1500 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 --cold 0 --min-seconds 0 --sub-seconds 0 --sub-steps 1 > $O/launches_bench.out 2>&1
tail -3 $O/launches_bench.out; wc -l $O/launches_bench.csv
