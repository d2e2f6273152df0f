# exact-dependency graphs: correctness, then timing variants (one GPU)
mkdir -p gpurun_out
export FEBA_BENCH_CACHE=/tmp/feba_cache
timeout 600 python tests/sparse_gpu_check.py 150 12000 12 2 1 > gpurun_out/plan_small.log 2>&1; echo "plan small rc=$?"; tail -2 gpurun_out/plan_small.log
timeout 900 python tests/sparse_gpu_check.py 1100 110000 96 8 0 > gpurun_out/plan_1100.log 2>&1; echo "plan 1100 rc=$?"; tail -2 gpurun_out/plan_1100.log
B="bench.py --steps 5 --warmup 3 --no-cpu"
run() { name=$1; shift; env "$@" timeout 600 python $B > gpurun_out/w_$name.json 2> gpurun_out/w_$name.err; echo "$name rc=$?"; tail -2 gpurun_out/w_$name.err; }
run default FEBA_VERBOSE=1
run t12 FEBA_TILE_MAX=12
run t16 FEBA_TILE_MAX=16
run t6 FEBA_TILE_MAX=6
run noexact FEBA_EXACT_DEPS=0 FEBA_DAG_STREAMS=32
run nomorton FEBA_POINT_ORDER=0
python scripts/bench_summary.py gpurun_out/w_*.json
