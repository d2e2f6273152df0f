# group form on N GPUs: bench (20 steps) + group check
N=${1:-8}
mkdir -p gpurun_out
export FEBA_BENCH_CACHE=/tmp/feba_cache
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561"
python -c "
import sys; sys.path.insert(0,'.')
import bench
bench.make_workload('config4', 1.0)" 2>/dev/null
FEBA_VERBOSE=1 timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err; echo "bench group rc=$?"; tail -2 gpurun_out/scale_n$N.err
python scripts/bench_summary.py gpurun_out/scale_n$N.json
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/scale_n$N.json") if l.startswith("{")][0]
print("group_check", d["group_check"]); print("adjustment", {k:v for k,v in d["adjustment"].items() if k!="deltasum"}); print("exchange_ms", d["kernels"]["cholesky"]["exchange_ms"], "create_s", d["create_s"], "clocks", d["clocks"])
PY
if [ "$2" = "config5" ]; then
timeout 900 $TR bench.py --gpus $N --workload config5 --steps 10 --warmup 3 --no-cpu > gpurun_out/scale_config5_n$N.json 2> gpurun_out/scale_config5_n$N.err; echo "config5 rc=$?"; python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/scale_config5_n$N.json') if l.startswith('{')][0]
print('config5 n=$N: batch', d['ms_per_step'], 'ms per step; e2e', d['e2e']['ms_per_step'])"
fi
