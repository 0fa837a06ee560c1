# compute-sanitizer (racecheck, memcheck) over the parity tests of the kernels new in round 2 + the golden / fuzz / sharded sessions; then an ncu launch list of the cfg 5 chain stage.   gpurun -- bash tools/gpu_sanitizer.sh
set -x
O=gpurun_out/sanitizer
mkdir -p $O
SEL="effects_every_kernel_shape or effects_time_parallel or reverb_extension or reverb_odd or warm_equals_cold or sharded_engines_on_one or golden_sharded or fused or tree or golden_scenario or fuzz"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "$SEL" > $O/compute_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/compute_sanitizer_racecheck.log
tail -4 $O/compute_sanitizer_racecheck.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "$SEL" > $O/compute_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/compute_sanitizer_memcheck.log
tail -3 $O/compute_sanitizer_memcheck.log
WBX_FIR=fft timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fft_|fir_|mix_kernel|render_tracks|patch_fx" -c 40 --csv --log-file $O/launches_cfg5_fft.csv python tools/kbench.py --tracks 256 --blocks 64 --reverb 65536 --fpl 4 --iters 2 > $O/launches_cfg5_fft.out 2>&1
WBX_FIR=fft python tools/kbench.py --tracks 256 --blocks 64 --reverb 65536 --fpl 4 --iters 5 > $O/kbench_cfg5.log 2>&1; cat $O/kbench_cfg5.log
