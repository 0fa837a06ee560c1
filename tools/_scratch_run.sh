set -x
O=gpurun_out/final3
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
timeout 900 python bench.py > $O/bench_n1_default.json 2> $O/bench_n1_default.err; echo "bench rc=$?"; tail -4 $O/bench_n1_default.err
python tools/kbench.py --tracks 1024 --blocks 4096 --fpl 16 --iters 5 > $O/kbench.log 2>&1
python tools/kbench.py --tracks 1024 --blocks 4096 --rate 44100 --fpl 16 --iters 5 >> $O/kbench.log 2>&1
python tools/kbench.py --tracks 1024 --blocks 1024 --rate 44100 --poly 1 --fpl 16 --iters 5 >> $O/kbench.log 2>&1
cat $O/kbench.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "golden_scenario or fuzz or tree or sharded_engines_on_one or polyphase or fades" > $O/racecheck_mix.log 2>&1; echo "racecheck rc=$?" >> $O/racecheck_mix.log; tail -3 $O/racecheck_mix.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "golden_scenario or fuzz or tree or sharded_engines_on_one or polyphase or fades or ragged" > $O/memcheck_mix.log 2>&1; echo "memcheck rc=$?" >> $O/memcheck_mix.log; tail -3 $O/memcheck_mix.log
