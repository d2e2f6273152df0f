# compute-sanitizer memcheck / racecheck on small problems: dense task graph, nested-dissection plan with
# exact-dependency capture, chunk form of the assembly, multi-camera atomics path; logs -> gpurun_out/sanitize_*.log
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
run() { tool=$1; name=$2; shift 2; env "$@" timeout 900 $S --tool $tool --error-exitcode 9 --launch-timeout 0 python tests/sanity_small.py free mixed > gpurun_out/sanitize_${tool}_$name.log 2>&1; echo "$tool $name rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|delta err" gpurun_out/sanitize_${tool}_$name.log | tail -4; }
run memcheck default
run memcheck plan FEBA_PLAN=1 FEBA_PLAN_MIN_BLOCKS=0 FEBA_ND_LEAF=4 FEBA_TILE_MAX=1
run memcheck densedag FEBA_PLAN=-1 FEBA_DAG_TILE=1
run racecheck default
run racecheck plan FEBA_PLAN=1 FEBA_PLAN_MIN_BLOCKS=0 FEBA_ND_LEAF=4 FEBA_TILE_MAX=1
run racecheck densedag FEBA_PLAN=-1 FEBA_DAG_TILE=1
