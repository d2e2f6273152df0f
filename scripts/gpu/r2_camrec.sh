# camera block from the records (k_cam_rec): timing + whole-adjustment checksums, then the parity subset
mkdir -p gpurun_out
export FEBA_BENCH_CACHE=/tmp/feba_cache
timeout 50 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/q_camrec.json 2> gpurun_out/q_camrec.err; echo "bench rc=$?"
python scripts/bench_summary.py gpurun_out/q_camrec.json
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/q_camrec.json") if l.startswith("{")][0]
print("camrec", d["adjustment"]["iterations"], d["adjustment"]["sigma02"], d["adjustment"]["xhat_l2"], d["adjustment"]["xhat_cam_l2"], d["adjustment"]["deltasum"][:6])
PY
timeout 65 python -m pytest tests/test_gpu_parity.py -x -q -k "not full_size and not baseline_configs and not covariance and not multi_camera" > gpurun_out/camrec_parity.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/camrec_parity.log
