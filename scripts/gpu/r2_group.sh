# group form (feba_create_shard) on N GPUs of one box:  gpurun --gpus N -- 'bash scripts/gpu/r2_group.sh N'
N=${1:-2}
mkdir -p gpurun_out
export FEBA_BENCH_CACHE=/tmp/feba_cache
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561"
timeout 900 $TR scripts/group_check.py > gpurun_out/group_check_n$N.log 2>&1; echo "group_check rc=$?"; grep -v "^W\|^\[W\|Warning" gpurun_out/group_check_n$N.log | tail -8
python -c "
import sys; sys.path.insert(0,'.')
import bench, argparse
a=argparse.Namespace(workload='config4', scale=1.0)
bench.make_workload('config4', 1.0)" 2>/dev/null
FEBA_VERBOSE=1 timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_group_n$N.json 2> gpurun_out/bench_group_n$N.err; echo "bench group rc=$?"; grep "feba\]" gpurun_out/bench_group_n$N.err | head -4; tail -3 gpurun_out/bench_group_n$N.err
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 --replicated > gpurun_out/bench_repl_n$N.json 2> gpurun_out/bench_repl_n$N.err; echo "bench replicated rc=$?"; tail -2 gpurun_out/bench_repl_n$N.err
python scripts/bench_summary.py gpurun_out/bench_group_n$N.json gpurun_out/bench_repl_n$N.json
python - <<PY
import json
for f in ("gpurun_out/bench_group_n$N.json",):
    try:
        d=[json.loads(l) for l in open(f) if l.startswith("{")][0]
        print(d["config"]["parallelism"]); print("group_check", d["group_check"]); print("adjustment", {k:v for k,v in d["adjustment"].items() if k!="deltasum"}, d["adjustment"]["deltasum"][:7]); print("exchange_ms", d["kernels"]["cholesky"]["exchange_ms"], "create_s", d["create_s"])
    except Exception as e: print(f, e)
PY
