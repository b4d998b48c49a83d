set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_r01h.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke_r01h.log
python bench.py > gpurun_out/bench_r01h_n1.json 2> gpurun_out/bench_r01h_n1.err; tail -c 600 gpurun_out/bench_r01h_n1.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r01h_reference.json 2>&1; tail -c 400 gpurun_out/bench_r01h_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_r01h.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_r01h.log 2>&1; echo ncu rc=$?
python tools/sweep.py > gpurun_out/sweep_r01h.md 2> gpurun_out/sweep_r01h.err; echo sweep rc=$?; tail -8 gpurun_out/sweep_r01h.md
python tools/bench_hbm_stages.py > gpurun_out/hbm_stages_r01h.md 2>&1; tail -9 gpurun_out/hbm_stages_r01h.md
