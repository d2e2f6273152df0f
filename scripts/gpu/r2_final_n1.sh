# one GPU: default bench with the CPU-port parity object, reference arm (short), sanitizer logs
mkdir -p gpurun_out
export FEBA_BENCH_CACHE=/tmp/feba_cache
FEBA_VERBOSE=1 timeout 1200 python bench.py > gpurun_out/bench_n1_default.json 2> gpurun_out/bench_n1_default.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_n1_default.err
python scripts/bench_summary.py gpurun_out/bench_n1_default.json
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/bench_n1_default.json") if l.startswith("{")][0]
print("parity", d["parity"]); print("cpu", d["cpu_baseline"]); print("roofline", {k:v for k,v in d["roofline"].items() if k not in ("kernel","peak_source")}); print("residual_stage", d["residual_stage"]); print("clocks", d["clocks"])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err; echo "reference rc=$?"; cut -c1-600 gpurun_out/bench_reference_arm.json
bash scripts/gpu/r2_sanitize.sh
