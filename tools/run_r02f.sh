cd $GRAFT_REPO_ROOT
python -m pytest tests -x -q -m gpu -s 2>&1 | tail -150 > gpurun_out/r02f_pytest_gpu.log; tail -3 gpurun_out/r02f_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02f_smoke.log
python bench.py > gpurun_out/r02f_bench_cfg2_n1.json 2> gpurun_out/r02f_bench_cfg2_n1.err; tail -c 1500 gpurun_out/r02f_bench_cfg2_n1.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02f_bench_reference.json 2>&1; tail -c 600 gpurun_out/r02f_bench_reference.json
python tools/entropic_gap.py > gpurun_out/r02e_entropic_gap.md 2>&1; cat gpurun_out/r02e_entropic_gap.md
