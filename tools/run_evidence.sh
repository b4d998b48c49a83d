# Evidence at HEAD on one B200 (gpurun -- "TAG=r02k bash tools/run_evidence.sh"): GPU tests, smoke, bench lines, reference arm, launch lists.
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests -x -q -m gpu -s 2>&1 | tail -260 > $O/${TAG:-r02i}_pytest_gpu.log; tail -3 $O/${TAG:-r02i}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/${TAG:-r02i}_smoke.log
python bench.py > $O/${TAG:-r02i}_bench_cfg2_n1.json 2> $O/${TAG:-r02i}_bench_cfg2_n1.err; tail -c 400 $O/${TAG:-r02i}_bench_cfg2_n1.json; echo
for c in cfg1 cfg3 cfg4 cfg5; do python bench.py --config $c > $O/${TAG:-r02i}_bench_${c}_n1.json 2> $O/${TAG:-r02i}_bench_${c}_n1.err; head -c 250 $O/${TAG:-r02i}_bench_${c}_n1.json; echo; done
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG:-r02i}_bench_reference.json 2>&1; tail -c 300 $O/${TAG:-r02i}_bench_reference.json; echo
B2="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG:-r02i}_launches_bench_cfg2.csv $B2 > $O/ncu_l2.log 2>&1; echo launches rc=$?
B3="python bench.py --config cfg3 --steps 2 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/${TAG:-r02i}_launches_bench_cfg3.csv $B3 > $O/ncu_l3.log 2>&1; echo launches3 rc=$?
python tools/time_ot_kernels.py > $O/${TAG:-r02i}_ot_kernels.md 2>&1; tail -12 $O/${TAG:-r02i}_ot_kernels.md
python tools/time_planar.py > $O/${TAG:-r02i}_planar.md 2>&1; tail -4 $O/${TAG:-r02i}_planar.md
python tools/time_max_ssw.py > $O/${TAG:-r02i}_max_ssw.md 2>&1; cat $O/${TAG:-r02i}_max_ssw.md
python tools/time_pseudo_max.py > $O/${TAG:-r02i}_pseudo_max.md 2>&1; cat $O/${TAG:-r02i}_pseudo_max.md
