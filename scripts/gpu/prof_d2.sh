TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
B="bench.py --gpus 2 --steps 5 --warmup 3"
FEBA_DIST_PROF=1 timeout 300 $TR --master-port 29701 $B > gpurun_out/d2_gprof.json 2> gpurun_out/d2_gprof.err
FEBA_DIST_PROF=1 FEBA_GREEN_SMS=32 timeout 300 $TR --master-port 29702 $B > gpurun_out/d2_gprof_green32.json 2> gpurun_out/d2_gprof_green32.err
python scripts/bench_summary.py gpurun_out/d2_gprof*.json
grep "feba dist prof" gpurun_out/d2_gprof.err | sort -s -k4,4n -k5,5n
