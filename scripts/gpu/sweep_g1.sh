B="bench.py --steps 5 --warmup 3 --no-cpu"
export FEBA_VERBOSE=1
for g in 0 8 16 24 32; do FEBA_GREEN_SMS=$g timeout 200 python $B > gpurun_out/g1_green$g.json 2> gpurun_out/g1_green$g.err; grep "feba\]" gpurun_out/g1_green$g.err | head -2; done
FEBA_GREEN_SMS=16 FEBA_UPD1_BULK=1 timeout 200 python $B > gpurun_out/g1_green16_u.json 2>/dev/null
FEBA_GREEN_SMS=16 FEBA_DAG_TILE=10 timeout 200 python $B > gpurun_out/g1_green16_t10.json 2>/dev/null
FEBA_GREEN_SMS=16 FEBA_DAG_TILE=12 timeout 200 python $B > gpurun_out/g1_green16_t12.json 2>/dev/null
FEBA_GREEN_SMS=16 FEBA_DAG_STREAMS=12 timeout 200 python $B > gpurun_out/g1_green16_s12.json 2>/dev/null
python scripts/bench_summary.py gpurun_out/g1_*.json
tail -3 gpurun_out/g1_green16.err
