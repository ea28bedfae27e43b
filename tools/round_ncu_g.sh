#!/bin/bash
# ncu --set full captures of the kernels added in r1_g (own Adam, multi-tensor staging), one launch each; raw csv pages only.
set -u
O=gpurun_out/ncu_g; mkdir -p $O
CMD="python bench.py --steps 1 --warmup 3 --no-graph --cpu-iters 0"
for k in adam_mt_kernel stage_multi_kernel mt_reg_grad_kernel; do
  timeout 150 ncu --set full --clock-control none --cache-control none --import-source on -k regex:$k -s 4 -c 1 -f -o $O/$k $CMD > $O/$k.log 2>&1
  ncu -i $O/$k.ncu-rep --page raw --csv > $O/${k}_raw.csv 2>/dev/null
  rm -f $O/$k.ncu-rep
done
ls -la $O
