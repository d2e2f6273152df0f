# end state, one GPU: whole GPU suite, smoke, default bench (with the CPU-port parity object), reference arm (short)
mkdir -p gpurun_out
export FEBA_BENCH_CACHE=/tmp/feba_cache
timeout 3000 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_final.log
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_final.log
FEBA_VERBOSE=1 timeout 1200 python bench.py > gpurun_out/bench_n1_final.json 2> gpurun_out/bench_n1_final.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_n1_final.err
python scripts/bench_summary.py gpurun_out/bench_n1_final.json
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/bench_n1_final.json") if l.startswith("{")][0]
print("parity", d["parity"]); print("cpu", d["cpu_baseline"]); print("roofline", {k:v for k,v in d["roofline"].items() if k not in ("kernel","peak_source")}); print("residual_stage", d["residual_stage"]); print("clocks", d["clocks"]); print("e2e", d["e2e"])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_final.json 2> gpurun_out/bench_reference_final.err; echo "reference rc=$?"; cut -c1-500 gpurun_out/bench_reference_final.json
