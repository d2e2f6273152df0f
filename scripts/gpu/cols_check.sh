# column form: single-GPU parity test, single-GPU timing (eager / graph), two-GPU parity + timing
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "task_graph" 2>&1 | tail -3
B1="bench.py --steps 5 --warmup 3 --no-cpu"
FEBA_DAG_FORM=cols timeout 200 python $B1 > gpurun_out/c1_cols_eager.json 2>gpurun_out/c1_cols_eager.err
FEBA_DAG_FORM=cols FEBA_SOLVE_GRAPH=1 timeout 200 python $B1 > gpurun_out/c1_cols_graph.json 2>/dev/null
FEBA_DAG_FORM=cols FEBA_GREEN_SMS=16 timeout 200 python $B1 > gpurun_out/c1_cols_eager_g16.json 2>/dev/null
python scripts/bench_summary.py gpurun_out/c1_*.json
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29801 scripts/dist_check.py --workload config3 --tile 6 > gpurun_out/dist_check_cols.log 2>&1; echo "dist_check rc=$?"; tail -1 gpurun_out/dist_check_cols.log
B="bench.py --gpus 2 --steps 5 --warmup 3"
p=29810
run() { name=$1; shift; p=$((p+1)); env "$@" timeout 300 $TR --master-port $p $B > gpurun_out/c2_$name.json 2> gpurun_out/c2_$name.err; }
run eager X=1
run eager_g16 FEBA_GREEN_SMS=16
run eager_g32 FEBA_GREEN_SMS=32
run graph FEBA_SOLVE_GRAPH=1
run eager_t10_g32 FEBA_GREEN_SMS=32 FEBA_DAG_TILE=10
python scripts/bench_summary.py gpurun_out/c2_*.json
tail -3 gpurun_out/c2_eager.err
fi
