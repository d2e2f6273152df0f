# final-state evidence on one GPU: launch list of a steady-state iteration, DRAM traffic of the hot kernels,
# one ncu --set full capture of the assembly kernels and the factorisation kernels
mkdir -p gpurun_out
export FEBA_BENCH_CACHE=/tmp/feba_cache
B="bench.py --steps 2 --warmup 3 --no-cpu"
timeout 600 python $B > gpurun_out/g_plain.json 2> gpurun_out/g_plain.err; echo "plain rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 5400 -c 1800 --csv --log-file gpurun_out/launches_r2g.csv python $B > gpurun_out/ncu_g1.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_point_pass|k_image_pass|k_pair_pass|k_cam_reduce|k_backsub|k_residuals|k_border_scale|k_clear_blocks" --launch-skip 24 -c 16 --csv --log-file gpurun_out/traffic_r2g.csv python $B > gpurun_out/ncu_g2.log 2>&1; echo "traffic rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_gemm_nt|k_chol_column|k_trsm_fused|k_backstep" --launch-skip 3000 -c 600 --csv --log-file gpurun_out/traffic_chol_r2g.csv python $B > gpurun_out/ncu_g3.log 2>&1; echo "traffic chol rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_pair_pass|k_image_pass|k_point_pass" --launch-skip 9 -c 3 -o gpurun_out/ncu_assembly_r2g -f python $B > gpurun_out/ncu_g4.log 2>&1; echo "full asm rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_chol_column" --launch-skip 400 -c 2 -o gpurun_out/ncu_column_r2g -f python $B > gpurun_out/ncu_g5.log 2>&1; echo "full column rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv
