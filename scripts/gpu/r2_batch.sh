mkdir -p gpurun_out
timeout 900 python bench.py --workload config5 --steps 5 --warmup 3 --no-cpu > gpurun_out/e_config5.json 2> gpurun_out/e_config5.err; echo "config5 rc=$?"; tail -2 gpurun_out/e_config5.err; python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/e_config5.json') if l.startswith('{')][0]
print('config5: batch', d['ms_per_step'], 'ms; per-handle launches', d['one_launch_pair_per_block_ms_per_step'], 'ms; one at a time', d['one_block_at_a_time_ms_per_step'], 'ms; e2e', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'])"
FEBA_BATCH_FUSED=0 timeout 900 python bench.py --workload config5 --steps 5 --warmup 3 --no-cpu > gpurun_out/e_config5_unfused.json 2> gpurun_out/e_config5_unfused.err; python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/e_config5_unfused.json') if l.startswith('{')][0]
print('config5 unfused: batch', d['ms_per_step'], 'ms; launches', d['gpu_launches'])"
timeout 3000 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
