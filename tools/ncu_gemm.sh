#!/bin/bash
O=gpurun_out
python tools/gemm_bench.py > $O/r2p_gemm.log 2>&1 || exit 1
for c in mixer step; do
  ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 5 -c 1 -o /tmp/r2_gemm_$c -f python tools/gemm_bench.py $c > $O/ncu_gemm_$c.log 2>&1
  ncu -i /tmp/r2_gemm_$c.ncu-rep --page raw --csv > $O/r2_gemm_$c.raw.csv 2>/dev/null
  ncu -i /tmp/r2_gemm_$c.ncu-rep --page source --csv > $O/r2_gemm_$c.source.csv 2>/dev/null
done
cat $O/r2p_gemm.log
