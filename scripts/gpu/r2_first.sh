# round 2, first contact of the nested-dissection plan (one GPU): quick checks, full GPU suite, bench
mkdir -p gpurun_out
timeout 600 python tests/sparse_gpu_check.py 150 12000 12 2 1 > gpurun_out/plan_small.log 2>&1; echo "plan small rc=$?"; tail -3 gpurun_out/plan_small.log
FEBA_VERBOSE=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_nocpu.json 2> gpurun_out/bench_nocpu.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_nocpu.err; python scripts/bench_summary.py gpurun_out/bench_nocpu.json
timeout 3000 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
