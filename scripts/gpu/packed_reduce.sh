# round-2 first check of the packed exchange (FEBA_PACKED_REDUCE=1): parity at 2 GPUs, then timing
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29801 scripts/dist_check.py --workload config3 --tile 6 --packed > gpurun_out/dist_check_packed.log 2>&1; echo "dist_check rc=$?"; tail -1 gpurun_out/dist_check_packed.log | cut -c1-400
B="bench.py --gpus 2 --steps 5 --warmup 3"
timeout 300 $TR --master-port 29811 $B > gpurun_out/p2_full.json 2> gpurun_out/p2_full.err
FEBA_PACKED_REDUCE=1 timeout 300 $TR --master-port 29812 $B > gpurun_out/p2_packed.json 2> gpurun_out/p2_packed.err
python scripts/bench_summary.py gpurun_out/p2_*.json
