# Round-2 multi-GPU evidence at HEAD for cfg3 (gpurun --gpus N -- 'bash tools/run_r02j_multi.sh N'): batch-sharded sliced loss.
cd $GRAFT_REPO_ROOT
N=$1; O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
$TR bench.py --gpus $N --config cfg3 --no-cpu-baseline > $O/r02j_bench_cfg3_n$N.json 2> $O/err_cfg3_n$N.log; head -c 300 $O/r02j_bench_cfg3_n$N.json; echo
