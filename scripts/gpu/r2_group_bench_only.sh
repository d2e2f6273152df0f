# group form on N GPUs: bench only (its JSON line carries group_check and the adjustment checksums)
N=${1:-4}
mkdir -p gpurun_out
export FEBA_BENCH_CACHE=/tmp/feba_cache
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561"
python -c "
import sys; sys.path.insert(0,'.')
import bench
bench.make_workload('config4', 1.0)" 2>/dev/null
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_group_n$N.json 2> gpurun_out/bench_group_n$N.err; echo "bench group rc=$?"; tail -2 gpurun_out/bench_group_n$N.err
python scripts/bench_summary.py gpurun_out/bench_group_n$N.json
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/bench_group_n$N.json") if l.startswith("{")][0]
print("group_check", d["group_check"]); print("sigma02", d["adjustment"]["sigma02"], "exchange_ms", d["kernels"]["cholesky"]["exchange_ms"])
PY
