set -x
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
timeout 600 python bench.py > $O/bench_n1_default.json 2> $O/bench_n1_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > $O/bench_n1_reference_arm.json 2> $O/bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 --cold 0 --min-seconds 0 --sub-seconds 0 --sub-steps 1 > $O/launches_bench.out 2>&1
NCU="ncu --set full --clock-control none --import-source on -c 1"
timeout 400 $NCU -k regex:mix_kernel --launch-skip 2 -o $O/mix_cfg2 python tools/kbench.py --tracks 1024 --blocks 4096 --fpl 16 --iters 1 > $O/ncu_mix_cfg2.out 2>&1
timeout 400 $NCU -k regex:mix_kernel --launch-skip 2 -o $O/mix_cfg3 python tools/kbench.py --tracks 1024 --blocks 4096 --rate 44100 --fpl 16 --iters 1 > $O/ncu_mix_cfg3.out 2>&1
timeout 400 $NCU -k regex:fx_chain_kernel --launch-skip 1 -o $O/fx_cfg4 python tools/kbench.py --tracks 512 --blocks 1024 --fx 1 --fpl 16 --iters 1 > $O/ncu_fx_cfg4.out 2>&1
timeout 400 $NCU -k regex:fir_tc_kernel --launch-skip 1 -o $O/fir_cfg5 python tools/kbench.py --tracks 256 --blocks 64 --reverb 65536 --fpl 16 --iters 1 > $O/ncu_fir_cfg5.out 2>&1
python tools/kbench.py --tracks 1024 --blocks 4096 --rate 44100 --fpl 16 --iters 5 > $O/kbench_cfg3.log 2>&1
python tools/kbench.py --tracks 512 --blocks 1024 --fx 1 --fpl 16 --iters 5 > $O/kbench_cfg4.log 2>&1
python tools/kbench.py --tracks 4096 --blocks 256 --fx 1 --fpl 16 --iters 5 >> $O/kbench_cfg4.log 2>&1
python tools/kbench.py --tracks 256 --blocks 64 --reverb 65536 --fpl 16 --iters 5 > $O/kbench_cfg5.log 2>&1
ls -la $O
