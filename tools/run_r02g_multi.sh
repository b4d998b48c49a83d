# Round-2 multi-GPU evidence (gpurun --gpus N -- 'bash tools/run_r02g_multi.sh N'): cfg2 weak and strong scaling, cfg3 batch-sharded.
cd $GRAFT_REPO_ROOT
N=$1; O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
$TR bench.py --gpus $N --no-cpu-baseline > $O/r02g_bench_cfg2_weak_n$N.json 2> $O/err_weak_n$N.log; head -c 200 $O/r02g_bench_cfg2_weak_n$N.json; echo
$TR bench.py --gpus $N --scaling strong --no-cpu-baseline > $O/r02g_bench_cfg2_strong_n$N.json 2> $O/err_strong_n$N.log; head -c 200 $O/r02g_bench_cfg2_strong_n$N.json; echo
$TR bench.py --gpus $N --config cfg3 --no-cpu-baseline > $O/r02g_bench_cfg3_n$N.json 2> $O/err_cfg3_n$N.log; head -c 200 $O/r02g_bench_cfg3_n$N.json; echo
