# round-2 first contact of the block-sparse reduced system (FEBA_SPARSE=1), one GPU:
#   gpurun --timeout 900 -- 'bash scripts/gpu/sparse_first_run.sh'
# 1. parity of the sparse form against the dense form and the oracle on small networks (forced task graph),
# 2. the same at the size where the task graph is the default (u_c = 6,610),
# 3. timing at configs[3] for a few supertile sizes, next to the dense form.
mkdir -p gpurun_out
FEBA_VERBOSE=1 timeout 300 python tests/sparse_gpu_check.py 150 12000 2 > gpurun_out/sparse_check_small.log 2>&1; echo "small rc=$?"; tail -4 gpurun_out/sparse_check_small.log
FEBA_VERBOSE=1 timeout 600 python tests/sparse_gpu_check.py 1100 110000 8 > gpurun_out/sparse_check_1100.log 2>&1; echo "1100 rc=$?"; tail -4 gpurun_out/sparse_check_1100.log
B="bench.py --steps 5 --warmup 3 --no-cpu"
timeout 300 python $B > gpurun_out/sp_dense.json 2> gpurun_out/sp_dense.err
for T in 6 8 10 14; do
  FEBA_SPARSE=1 FEBA_DAG_TILE=$T timeout 300 python $B > gpurun_out/sp_T$T.json 2> gpurun_out/sp_T$T.err
done
python scripts/bench_summary.py gpurun_out/sp_*.json
