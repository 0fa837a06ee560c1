set -x
O=gpurun_out/r02l
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q -k "reverb" > $O/pytest_gpu_reverb.log 2>&1; echo "rc=$?" >> $O/pytest_gpu_reverb.log; tail -3 $O/pytest_gpu_reverb.log
WBX_FIR=fft python tools/kbench.py --tracks 256 --blocks 64 --reverb 65536 --fpl 4 --iters 5 > $O/kbench_cfg5.log 2>&1
WBX_FIR=fft WBX_FFT_MAC_REG=1 python tools/kbench.py --tracks 256 --blocks 64 --reverb 65536 --fpl 4 --iters 5 >> $O/kbench_cfg5.log 2>&1
WBX_FIR=fft python tools/kbench.py --tracks 32 --blocks 64 --reverb 65536 --fpl 4 --iters 5 >> $O/kbench_cfg5.log 2>&1
cat $O/kbench_cfg5.log
WBX_FIR=fft timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fft_|fir_|mix_kernel|render_tracks|patch_fx" -c 40 --csv --log-file $O/launches_cfg5_fft.csv python tools/kbench.py --tracks 256 --blocks 64 --reverb 65536 --fpl 4 --iters 2 > $O/launches_cfg5_fft.out 2>&1
NCU="ncu --set full --clock-control none --import-source on -c 1"
WBX_FIR=fft timeout 300 $NCU -k regex:fft_mac --launch-skip 2 -o $O/fft_mac python tools/kbench.py --tracks 256 --blocks 64 --reverb 65536 --fpl 4 --iters 2 > $O/ncu2.out 2>&1
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "reverb_cfg5_tap_count or reverb_extension" > $O/racecheck_fft.log 2>&1; echo "rc=$?" >> $O/racecheck_fft.log; tail -3 $O/racecheck_fft.log
