set -x
O=gpurun_out/r02n
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "effects or polyphase or fades" > $O/pytest_gpu_effects.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_effects.log
tail -3 $O/pytest_gpu_effects.log
for sh in 4,2,1,1 2,2,1,2; do
  echo "== shape $sh" >> $O/kbench_cfg4.log
  WBX_FX_SHAPE=$sh python tools/kbench.py --tracks 512 --blocks 1024 --fx 1 --fpl 16 --iters 5 2>&1 | grep "fx submit" >> $O/kbench_cfg4.log
  WBX_FX_SHAPE=$sh python tools/kbench.py --tracks 4096 --blocks 256 --fx 1 --fpl 8 --iters 5 2>&1 | grep "fx submit" >> $O/kbench_cfg4.log
done
cat $O/kbench_cfg4.log
NCU="ncu --set full --clock-control none --import-source on -c 1"
timeout 400 $NCU -k regex:fx_chain_kernel --launch-skip 1 -o $O/fx_cfg4 python tools/kbench.py --tracks 512 --blocks 1024 --fx 1 --fpl 16 --iters 1 > $O/ncu_fx_cfg4.out 2>&1
