mkdir -p gpurun_out
export FEBA_BENCH_CACHE=/tmp/feba_cache
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_chunk_reduce" --launch-skip 3 -c 1 -o gpurun_out/ncu_chunk_reduce -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_chunk.log 2>&1; echo "ncu rc=$?"
