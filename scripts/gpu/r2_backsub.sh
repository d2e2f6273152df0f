# record-based back-substitution: parity subset, plan check, timing against the recomputing kernel
mkdir -p gpurun_out
export FEBA_BENCH_CACHE=/tmp/feba_cache
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "not full_size" > gpurun_out/bs_parity.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/bs_parity.log
timeout 600 python tests/sparse_gpu_check.py 150 12000 12 2 1 > gpurun_out/bs_plan_small.log 2>&1; echo "plan small rc=$?"; tail -2 gpurun_out/bs_plan_small.log
B="bench.py --steps 10 --warmup 3 --no-cpu"
run() { name=$1; shift; env "$@" timeout 600 python $B > gpurun_out/q_$name.json 2> gpurun_out/q_$name.err; echo "$name rc=$?"; tail -1 gpurun_out/q_$name.err; }
run rec FEBA_VERBOSE=0
run recompute FEBA_BACKSUB_REC=0
python scripts/bench_summary.py gpurun_out/q_rec.json gpurun_out/q_recompute.json
python - <<PY
import json
for n in ("rec","recompute"):
    d=[json.loads(l) for l in open(f"gpurun_out/q_{n}.json") if l.startswith("{")][0]
    print(n, d["adjustment"]["iterations"], d["adjustment"]["sigma02"], d["adjustment"]["xhat_l2"], d["adjustment"]["deltasum"][:6])
PY
