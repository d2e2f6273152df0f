# round 2, second GPU run: plan fix check, timing variants of the plan's task graph, launch list, GPU suite
mkdir -p gpurun_out
export FEBA_BENCH_CACHE=/tmp/feba_cache
timeout 600 python tests/sparse_gpu_check.py 150 12000 12 2 1 > gpurun_out/plan_small.log 2>&1; echo "plan small rc=$?"; tail -2 gpurun_out/plan_small.log
B="bench.py --steps 5 --warmup 3 --no-cpu"
run() { name=$1; shift; env "$@" timeout 600 python $B > gpurun_out/v_$name.json 2> gpurun_out/v_$name.err; echo "$name rc=$?"; }
run default FEBA_VERBOSE=1
run eager FEBA_SOLVE_GRAPH=0
run s32 FEBA_DAG_STREAMS=32
run s8 FEBA_DAG_STREAMS=8
run t4 FEBA_TILE_MAX=4
run t12 FEBA_TILE_MAX=12
run leaf48 FEBA_ND_LEAF=48
run dense FEBA_PLAN=-1
python scripts/bench_summary.py gpurun_out/v_*.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 9000 -c 2600 --csv --log-file gpurun_out/launches_r2b.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
timeout 3000 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
