# end-state evidence on one GPU: DRAM traffic of the assembly / update kernels, launch list of one steady-state iteration
mkdir -p gpurun_out
export FEBA_BENCH_CACHE=/tmp/feba_cache
B="bench.py --steps 2 --warmup 3 --no-cpu"
timeout 150 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_point_pass|k_image_pass|k_pair_pass|k_cam_reduce|k_cam_direct|k_backsub|k_border_scale|k_clear_blocks" --launch-skip 28 -c 21 --csv --log-file gpurun_out/traffic_r2h.csv python $B > gpurun_out/ncu_h2.log 2>&1; echo "traffic rc=$?"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 5400 -c 1800 --csv --log-file gpurun_out/launches_r2h.csv python $B > gpurun_out/ncu_h1.log 2>&1; echo "launch list rc=$?"
ls -la gpurun_out/*r2h.csv
