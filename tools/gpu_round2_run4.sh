set -x
O=gpurun_out/r02d
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "effects or polyphase" > $O/pytest_gpu_effects.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_effects.log
tail -5 $O/pytest_gpu_effects.log
for sh in 4,2,1,1 1,2,1,4; do
  echo "== shape $sh" >> $O/kbench_cfg4.log
  WBX_FX_SHAPE=$sh python tools/kbench.py --tracks 512 --blocks 1024 --fx 1 --fpl 16 --iters 5 2>&1 | grep "fx submit" >> $O/kbench_cfg4.log
  WBX_FX_SHAPE=$sh python tools/kbench.py --tracks 4096 --blocks 256 --fx 1 --fpl 8 --iters 5 2>&1 | grep "fx submit" >> $O/kbench_cfg4.log
done
cat $O/kbench_cfg4.log
