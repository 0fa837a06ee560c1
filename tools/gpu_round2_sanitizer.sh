# compute-sanitizer over the kernels that are new in round 2 (time-parallel effect chain in every shape, FFT reverb, tensor-core
# reverb with tap split, fused shard exchange on one GPU), plus the new reverb tests without the sanitizer.
set -x
O=gpurun_out/r02j
mkdir -p $O
SEL="effects_every_kernel_shape or effects_time_parallel or reverb_extension or reverb_odd or warm_equals_cold or sharded_engines_on_one or golden_sharded or fused"
timeout 600 python -m pytest tests -m gpu -x -q -k "reverb" > $O/pytest_gpu_reverb.log 2>&1; echo "rc=$?" >> $O/pytest_gpu_reverb.log; tail -3 $O/pytest_gpu_reverb.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "$SEL" > $O/compute_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/compute_sanitizer_memcheck.log
tail -4 $O/compute_sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "$SEL" > $O/compute_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/compute_sanitizer_racecheck.log
tail -4 $O/compute_sanitizer_racecheck.log
