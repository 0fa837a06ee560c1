set -x
O=gpurun_out/r02p
mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on -c 1"
timeout 400 $NCU -k regex:mix_kernel --launch-skip 2 -o $O/mix_cfg2 python tools/kbench.py --tracks 1024 --blocks 4096 --fpl 16 --iters 1 > $O/ncu_mix_cfg2.out 2>&1
timeout 400 $NCU -k regex:mix_kernel --launch-skip 2 -o $O/mix_cfg3 python tools/kbench.py --tracks 1024 --blocks 4096 --rate 44100 --fpl 16 --iters 1 > $O/ncu_mix_cfg3.out 2>&1
timeout 400 $NCU -k regex:mix_kernel --launch-skip 2 -o $O/mix_poly python tools/kbench.py --tracks 1024 --blocks 1024 --rate 44100 --poly 1 --fpl 16 --iters 1 > $O/ncu_poly.out 2>&1
