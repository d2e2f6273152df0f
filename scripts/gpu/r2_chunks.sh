# chunk form of the assembly: correctness (block tests, small plan check), then timing against the image-major form
mkdir -p gpurun_out
export FEBA_BENCH_CACHE=/tmp/feba_cache
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "reduced_system_blocks or single_iteration or small_synthetic or flag_compaction or edge_cases or multi_camera or setting_variants" > gpurun_out/pytest_chunks.log 2>&1; echo "pytest subset rc=$?"; tail -5 gpurun_out/pytest_chunks.log
timeout 600 python tests/sparse_gpu_check.py 150 12000 12 2 1 > gpurun_out/plan_small.log 2>&1; echo "plan small rc=$?"; tail -2 gpurun_out/plan_small.log
B="bench.py --steps 5 --warmup 3 --no-cpu"
run() { name=$1; shift; env "$@" timeout 600 python $B > gpurun_out/c_$name.json 2> gpurun_out/c_$name.err; echo "$name rc=$?"; tail -2 gpurun_out/c_$name.err; }
run chunks FEBA_VERBOSE=1
run imagemajor FEBA_CHUNKS=0
python scripts/bench_summary.py gpurun_out/c_*.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_point_pass|k_chunk_reduce|k_sum_img|k_sum_blk|k_cam_reduce|k_backsub|k_residuals" --launch-skip 30 -c 12 --csv --log-file gpurun_out/ncu_asm_r2d.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_asm.log 2>&1; echo "ncu rc=$?"
