timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "task_graph" 2>&1 | tail -2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29801 scripts/dist_check.py --workload config3 --tile 6 > gpurun_out/dist_check_cols2.log 2>&1; echo "dist_check rc=$?"; tail -1 gpurun_out/dist_check_cols2.log | cut -c1-200
B="bench.py --gpus 2 --steps 5 --warmup 3"
p=29810
run() { name=$1; shift; p=$((p+1)); env "$@" timeout 300 $TR --master-port $p $B > gpurun_out/c3_$name.json 2> gpurun_out/c3_$name.err; }
run eager X=1
run eager_g16 FEBA_GREEN_SMS=16
run eager_g32 FEBA_GREEN_SMS=32
run eager_t10_g32 FEBA_GREEN_SMS=32 FEBA_DAG_TILE=10
python scripts/bench_summary.py gpurun_out/c3_*.json
