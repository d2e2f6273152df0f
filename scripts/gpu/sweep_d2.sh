TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
B="bench.py --gpus 2 --steps 5 --warmup 3"
p=29600
run() { name=$1; shift; p=$((p+1)); env "$@" timeout 300 $TR --master-port $p $B > gpurun_out/d2_$name.json 2> gpurun_out/d2_$name.err; }
run green16 FEBA_GREEN_SMS=16
run green32 FEBA_GREEN_SMS=32
run green48 FEBA_GREEN_SMS=48
run green32_s12 FEBA_GREEN_SMS=32 FEBA_DAG_STREAMS=12
run green32_u FEBA_GREEN_SMS=32 FEBA_UPD1_BULK=1
run green32_prof FEBA_GREEN_SMS=32 FEBA_NO_GRAPH=1 FEBA_DIST_PROF=1
python scripts/bench_summary.py gpurun_out/d2_green*.json
grep "feba dist prof" gpurun_out/d2_green32_prof.err | sort -k4,4n -k5,5n | head -40
