# group form on N GPUs: check + bench only
N=${1:-8}
mkdir -p gpurun_out
export FEBA_BENCH_CACHE=/tmp/feba_cache
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561"
timeout 600 $TR scripts/group_check.py --images 900 --points 90000 > gpurun_out/group_check_n$N.log 2>&1; echo "group_check rc=$?"; grep "^\[free\]\|^\[mixed\]\|group form" gpurun_out/group_check_n$N.log | tail -6
python -c "
import sys; sys.path.insert(0,'.')
import bench
bench.make_workload('config4', 1.0)" 2>/dev/null
FEBA_VERBOSE=1 timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_group_n$N.json 2> gpurun_out/bench_group_n$N.err; echo "bench group rc=$?"; grep "feba\]" gpurun_out/bench_group_n$N.err | head -2; tail -3 gpurun_out/bench_group_n$N.err
python scripts/bench_summary.py gpurun_out/bench_group_n$N.json
python - <<PY
import json
for f in ("gpurun_out/bench_group_n$N.json",):
    try:
        d=[json.loads(l) for l in open(f) if l.startswith("{")][0]
        print(d["config"]["parallelism"]); print("group_check", d["group_check"]); print("adjustment", {k:v for k,v in d["adjustment"].items() if k!="deltasum"}, d["adjustment"]["deltasum"][:7]); print("exchange_ms", d["kernels"]["cholesky"]["exchange_ms"], "create_s", d["create_s"])
    except Exception as e: print(f, e)
PY
