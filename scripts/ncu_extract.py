"""profiles/r2_ncu_traffic.json from an ncu metrics pass (csv):

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        -k regex:"k_point_pass|k_image_pass|k_pair_pass|k_gemm_nt|k_chol_column|k_trsm_fused|k_backsub|k_residuals" \
        --launch-skip N -c M --csv --log-file launches.csv python bench.py --steps 2 --warmup 3 --no-cpu
    python scripts/ncu_extract.py launches.csv profiles/r2_ncu_traffic.json

Per kernel: launches seen, mean duration, mean DRAM read + write bytes per launch.  The entry
"assembly (k_point_pass + k_image_pass + k_pair_pass)" adds the assembly kernels of ONE iteration (these three plus
the small camera kernels k_cam_direct / k_cam_reduce when they were captured): that is what
bench.py prints as roofline.traffic when the assembly is the dominant phase."""
import collections
import csv
import json
import sys


def main(src, dst):
    lines = [ln for ln in open(src) if not ln.startswith("==")]
    per = collections.OrderedDict()
    for r in csv.DictReader(lines):
        name = r["Kernel Name"].split("(")[0].replace("void ", "").replace("feba::", "")
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        d = per.setdefault(name, collections.defaultdict(list))
        d[r["Metric Name"]].append(v * scale)
    out = {"source": f"ncu metrics pass, {src}", "kernels": {}}
    for name, d in per.items():
        n = len(d.get("gpu__time_duration.sum", []))
        if not n:
            continue
        rd = sum(d.get("dram__bytes_read.sum", [0])) / n
        wr = sum(d.get("dram__bytes_write.sum", [0])) / n
        out["kernels"][name] = {"launches": n, "ms_per_launch": sum(d["gpu__time_duration.sum"]) / n,
                                "dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr}
    asm = [k for k in out["kernels"] if k.startswith(("k_point_pass", "k_image_pass", "k_pair_pass", "k_cam_direct", "k_cam_reduce"))]
    if asm:
        out["kernels"]["assembly (k_point_pass + k_image_pass + k_pair_pass)"] = {
            "launches": 1, "ms_per_launch": sum(out["kernels"][k]["ms_per_launch"] for k in asm),
            "dram_bytes_per_launch": sum(out["kernels"][k]["dram_bytes_per_launch"] for k in asm),
            "parts": asm}
    json.dump(out, open(dst, "w"), indent=1)
    for k, v in out["kernels"].items():
        print(f"{k}: {v['launches']} launches, {v['ms_per_launch']:.3f} ms, {v['dram_bytes_per_launch'] / 1e9:.3f} GB per launch")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
