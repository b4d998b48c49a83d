# Round-2 evidence at HEAD on one B200 (gpurun): GPU tests, smoke, every bench config, the reference arm, launch lists, ncu captures.
cd $GRAFT_REPO_ROOT
O=gpurun_out
python -m pytest tests -x -q -m gpu -s 2>&1 | tail -220 > $O/r02g_pytest_gpu.log; tail -3 $O/r02g_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/r02g_smoke.log
python bench.py > $O/r02g_bench_cfg2_n1.json 2> $O/r02g_bench_cfg2_n1.err; tail -c 400 $O/r02g_bench_cfg2_n1.json; echo
for c in cfg1 cfg3 cfg4 cfg5; do python bench.py --config $c > $O/r02g_bench_${c}_n1.json 2> $O/r02g_bench_${c}_n1.err; head -c 250 $O/r02g_bench_${c}_n1.json; echo; done
python bench.py --impl reference --steps 3 --warmup 1 > $O/r02g_bench_reference.json 2>&1; tail -c 300 $O/r02g_bench_reference.json; echo
B2="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02g_launches_bench_cfg2.csv $B2 > $O/ncu_l2.log 2>&1; echo launches rc=$?
python tools/one_auction.py && ncu --set full --clock-control none --import-source on -k regex:auction_kernel -s 1 -c 1 -f -o $O/r02g_auction python tools/one_auction.py > $O/ncu_au.log 2>&1; echo ncu auction rc=$?
python tools/time_shwd_step.py > $O/r02g_training_step.txt 2>&1; cat $O/r02g_training_step.txt
python tools/profile_auction.py 2>&1 | cut -d"|" -f1 > $O/r02g_auction_times.txt; SHWD_B200_LIB=$PWD/tools/variants/au_prof.so python tools/profile_auction.py >> $O/r02g_auction_times.txt 2>&1; python tools/time_dense_emd2.py >> $O/r02g_auction_times.txt 2>&1; tail -5 $O/r02g_auction_times.txt
python tools/kernel_times_sliced.py 2>&1 | grep -v Warn | grep -v _warn > $O/r02g_kernel_times_sliced_cfg3.txt
python tools/time_chamfer.py > $O/r02g_chamfer.txt 2>&1; tail -6 $O/r02g_chamfer.txt
python tools/time_graphs.py > $O/r02g_graphed_loss.md 2>&1; tail -8 $O/r02g_graphed_loss.md
