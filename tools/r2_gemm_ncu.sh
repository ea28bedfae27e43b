#!/bin/bash
O=gpurun_out/gemm_ncu; mkdir -p $O
for cfg in "dec.logits 1256" "rec.dW_hh 2256" "rec.dW_hh 1256" "rec.out 1256" "rec.out 2256"; do
  set -- $cfg
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -f -o /tmp/${1}_$2 python tools/gemm_one.py $1 $2 > $O/${1}_$2.log 2>&1
  python tools/ncu_hot.py /tmp/${1}_$2.ncu-rep 30 > $O/${1}_$2.txt 2>&1
  ncu -i /tmp/${1}_$2.ncu-rep --page raw --csv > $O/${1}_$2_raw.csv 2>/dev/null
done
ls -la $O
