set -x
O=gpurun_out/r02o
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -5 $O/pytest_gpu.log
python tools/kbench.py --tracks 1024 --blocks 4096 --fpl 16 --iters 5 > $O/kbench.log 2>&1
python tools/kbench.py --tracks 1024 --blocks 4096 --rate 44100 --fpl 16 --iters 5 >> $O/kbench.log 2>&1
python tools/kbench.py --tracks 1024 --blocks 1024 --rate 44100 --poly 1 --fpl 16 --iters 5 >> $O/kbench.log 2>&1
python tools/kbench.py --tracks 1024 --blocks 4096 --offset 1 --fpl 16 --iters 5 >> $O/kbench.log 2>&1
cat $O/kbench.log
for n in 64 1024; do python tools/latency.py --tracks $n --mode auto >> $O/latency.log 2>&1; done; cat $O/latency.log
