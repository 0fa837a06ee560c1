mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"mix_kernel|expand|level|clamp|ingest|shard" -c 60 --csv --log-file gpurun_out/r01b_launches_bench.csv python bench.py --steps 2 --warmup 1 --cold 0 > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mix_kernel -s 2 -c 1 -o gpurun_out/mix_full python tools/kbench.py --tracks 1024 --blocks 4096 --fpl 16 --iters 1 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mix_kernel -s 2 -c 1 -o gpurun_out/lin_full python tools/kbench.py --tracks 1024 --blocks 1024 --rate 44100 --fpl 16 --iters 1 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:effects_kernel -s 1 -c 1 -o gpurun_out/fx_full python tools/kbench.py --tracks 512 --blocks 256 --fx 1 --fpl 16 --iters 1 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
