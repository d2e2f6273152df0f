# left-looking column kernel + wider fused TRSM + batch graph: correctness, timing, full GPU suite
mkdir -p gpurun_out
export FEBA_BENCH_CACHE=/tmp/feba_cache
timeout 600 python tests/sparse_gpu_check.py 150 12000 12 2 1 > gpurun_out/plan_small.log 2>&1; echo "plan small rc=$?"; tail -2 gpurun_out/plan_small.log
timeout 900 python tests/sparse_gpu_check.py 1100 110000 96 6 0 > gpurun_out/plan_1100.log 2>&1; echo "plan 1100 rc=$?"; tail -2 gpurun_out/plan_1100.log
B="bench.py --steps 5 --warmup 3 --no-cpu"
run() { name=$1; shift; env "$@" timeout 600 python $B > gpurun_out/d_$name.json 2> gpurun_out/d_$name.err; echo "$name rc=$?"; tail -1 gpurun_out/d_$name.err; }
run columns FEBA_VERBOSE=1
run recursive FEBA_CHOL_COLUMNS=0
run columns_t8 FEBA_TILE_MAX=8
run columns_t4 FEBA_TILE_MAX=4
python scripts/bench_summary.py gpurun_out/d_*.json
timeout 600 python bench.py --workload config5 --steps 5 --warmup 3 --no-cpu > gpurun_out/d_config5.json 2> gpurun_out/d_config5.err; echo "config5 rc=$?"; tail -2 gpurun_out/d_config5.err; python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/d_config5.json') if l.startswith('{')][0]
print('config5: batch graph', d['ms_per_step'], 'ms; per-handle launches', d['one_launch_pair_per_block_ms_per_step'], 'ms; one at a time', d['one_block_at_a_time_ms_per_step'], 'ms; e2e', d['e2e']['ms_per_step'])"
timeout 3000 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
