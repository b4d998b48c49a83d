set -x
cd $GRAFT_REPO_ROOT
B2="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B2 > gpurun_out/plain_cfg2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02c_launches_bench_cfg2.csv $B2 > gpurun_out/ncu_l2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sinkhorn_.wd_kernel -s 6 -c 2 -f -o gpurun_out/r02c_cfg2_sinkhorn $B2 > gpurun_out/ncu_f2.log 2>&1; echo rc=$?
T1="python tools/time_ot_kernels.py 4x1024"
$T1 > gpurun_out/plain_lean1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:lean_kernel -s 10 -c 2 -f -o gpurun_out/r02c_lean_4x1024 $T1 > gpurun_out/ncu_f3.log 2>&1; echo rc=$?
T2="python tools/time_ot_kernels.py 32x256"
$T2 > gpurun_out/plain_lean2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:lean_kernel -s 10 -c 2 -f -o gpurun_out/r02c_lean_32x256 $T2 > gpurun_out/ncu_f4.log 2>&1; echo rc=$?
B3="python bench.py --config cfg3 --steps 2 --warmup 3 --no-cpu-baseline"
$B3 > gpurun_out/plain_cfg3.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02c_launches_bench_cfg3.csv $B3 > gpurun_out/ncu_l3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"segmented_sort_trim|circular_wp_kernel" -s 6 -c 3 -f -o gpurun_out/r02c_cfg3_sliced $B3 > gpurun_out/ncu_f5.log 2>&1; echo rc=$?
ls -la gpurun_out/*.ncu-rep
