set -x
O=gpurun_out/r02b
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
python tools/kbench.py --tracks 1024 --blocks 4096 --rate 44100 --fpl 16 --iters 5 > $O/kbench_cfg3.log 2>&1
python tools/kbench.py --tracks 1024 --blocks 4096 --fpl 16 --iters 5 > $O/kbench_cfg2.log 2>&1
for sh in 4,2,1,1 4,2,2,1 2,2,1,2 2,2,2,2 1,2,1,4 1,2,2,4 1,4,2,2 2,4,1,1; do
  echo "== shape $sh" >> $O/kbench_cfg4_shapes.log
  WBX_FX_SHAPE=$sh python tools/kbench.py --tracks 512 --blocks 1024 --fx 1 --fpl 16 --iters 5 2>&1 | grep "fx submit" >> $O/kbench_cfg4_shapes.log
  WBX_FX_SHAPE=$sh python tools/kbench.py --tracks 4096 --blocks 256 --fx 1 --fpl 8 --iters 5 2>&1 | grep "fx submit" >> $O/kbench_cfg4_shapes.log
done
cat $O/kbench_cfg3.log $O/kbench_cfg2.log $O/kbench_cfg4_shapes.log
