for i in 1 2 3; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$i bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('clocks', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['e2e']['ms_per_step'])"
SHWD_BENCH_NO_CLOCKS=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$i bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('noclk ', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['e2e']['ms_per_step'])"
done
