#!/bin/bash
# Final round-2 captures with the finished kernels (one GPU): launch list, ncu --set full of the hot kernels.
export CTCB_NO_GRAPH=1
NCU="ncu --clock-control none"
B="python bench.py --headline-only --no-cpu-baseline"
timeout 600 $NCU --metrics gpu__time_duration.sum -s 150 -c 400 --csv --log-file gpurun_out/launches_r2.csv $B --steps 3 --warmup 3 > gpurun_out/launches_r2.out 2>&1
timeout 900 $NCU --set full --import-source on -k regex:"sweep_cluster_kernel|ctc_par_kernel|gemm_tc_kernel" --launch-skip 36 --launch-count 12 -o gpurun_out/prof_c2_r2 -f $B --steps 2 --warmup 3 > gpurun_out/prof_c2_r2.out 2>&1
timeout 900 $NCU --set full --import-source on -k regex:"gemm_tc_kernel" --launch-skip 45 --launch-count 8 -o gpurun_out/prof_c3_r2 -f $B --config c3 --steps 1 --warmup 3 > gpurun_out/prof_c3_r2.out 2>&1
CTCB_SWEEP_TC_COOP=0 timeout 600 $NCU --set full --import-source on -k regex:"sweep_tc_kernel" --launch-skip 4 --launch-count 2 -o gpurun_out/prof_sweeptc_nocoop_r2 -f $B --config c3 --steps 1 --warmup 3 > gpurun_out/prof_sweeptc_nocoop_r2.out 2>&1
ls -la gpurun_out/*.ncu-rep
