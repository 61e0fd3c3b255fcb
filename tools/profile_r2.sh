#!/bin/bash
# Round-2 evidence run (one GPU): launch list of the benchmarked step, ncu --set full captures of the hot kernels,
# compute-sanitizer logs.  Outputs under gpurun_out/ (summarised into profiles/ by tools/make_profiles.py).
export CTCB_NO_GRAPH=1
NCU="ncu --clock-control none"
B="python bench.py --headline-only --no-cpu-baseline"
# 1. every launch of three C2 steps with its device time (cold-cache, serialised: compare shares)
timeout 600 $NCU --metrics gpu__time_duration.sum -s 150 -c 400 --csv --log-file gpurun_out/launches_r2.csv $B --steps 3 --warmup 3 > gpurun_out/launches_r2.out 2>&1
# 2. full captures: one C2 step's hot kernels, the isolated CTC kernel, the C3 tensor-core sweep and GEMMs
timeout 900 $NCU --set full --import-source on -k regex:"sweep_cluster_kernel|ctc_pair_kernel|gemm_tc_kernel" --launch-skip 36 --launch-count 12 -o gpurun_out/prof_c2_r2 -f $B --steps 2 --warmup 3 > gpurun_out/prof_c2_r2.out 2>&1
timeout 600 $NCU --set full --import-source on -k regex:ctc_warp_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/prof_ctc_r2 -f $B --steps 1 --warmup 3 > gpurun_out/prof_ctc_r2.out 2>&1
timeout 900 $NCU --set full --import-source on -k regex:"sweep_tc_kernel|gemm_tc_kernel" --launch-skip 45 --launch-count 15 -o gpurun_out/prof_c3_r2 -f $B --config c3 --steps 1 --warmup 3 > gpurun_out/prof_c3_r2.out 2>&1
# 3. sanitizers on the tiny configuration (cluster sweep, tcgen05 GEMM, two-warp CTC) and on the tensor-core sweep
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/sanitizer_memcheck_smoke_r2.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck_smoke_r2.out 2>&1
timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/sanitizer_racecheck_smoke_r2.log python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_racecheck_smoke_r2.out 2>&1
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/sanitizer_memcheck_sweep_tc_r2.log python tools/sweep_time.py 6 20 1024 > gpurun_out/sanitizer_memcheck_sweep_tc_r2.out 2>&1
timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/sanitizer_racecheck_sweep_tc_r2.log python tools/sweep_time.py 6 20 1024 > gpurun_out/sanitizer_racecheck_sweep_tc_r2.out 2>&1
ls -la gpurun_out/*r2*
tail -3 gpurun_out/*sanitizer*r2.log
