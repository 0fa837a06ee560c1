set -x
O=gpurun_out/r02g
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -8 $O/pytest_gpu.log
WBX_FIR=fft python tools/kbench.py --tracks 256 --blocks 64 --reverb 65536 --fpl 4 --iters 5 >> $O/kbench_cfg5.log 2>&1
WBX_FIR=fft WBX_FFT_P=512 python tools/kbench.py --tracks 256 --blocks 64 --reverb 65536 --fpl 4 --iters 5 >> $O/kbench_cfg5.log 2>&1
WBX_FIR=fft python tools/kbench.py --tracks 256 --blocks 1 --reverb 65536 --fpl 4 --iters 20 >> $O/kbench_cfg5.log 2>&1
cat $O/kbench_cfg5.log
WBX_FIR=fft timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fft_|fir_|mix_kernel|render_tracks|patch_fx" -c 40 --csv --log-file $O/launches_cfg5_fft.csv python tools/kbench.py --tracks 256 --blocks 64 --reverb 65536 --fpl 4 --iters 2 > $O/launches_cfg5_fft.out 2>&1
timeout 600 python bench.py > $O/bench_n1_default.json 2> $O/bench_n1_default.err; tail -6 $O/bench_n1_default.err
