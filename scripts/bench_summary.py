"""Print 'label ms_per_step phases' for bench output files (skips non-JSON lines)."""
import json
import sys

for path in sys.argv[1:]:
    try:
        for line in open(path):
            if line.startswith("{"):
                d = json.loads(line)
                k = {n: round(v["ms"], 2) for n, v in d.get("kernels", {}).items()}
                print(f"{path}: n={d['n_gpus']} {d['ms_per_step']:.2f} ms  e2e {d['e2e']['ms_per_step']:.2f} ms  "
                      f"launches {d['gpu_launches']}  {k}")
    except OSError as exc:
        print(path, exc)
