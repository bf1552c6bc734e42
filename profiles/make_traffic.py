"""Writes profiles/r02_raster_traffic.json from ncu per-kernel DRAM captures of THIS tree:

    python profiles/make_traffic.py bunny:128=gpurun_out/x_dram_per_kernel.csv [crates:8=...csv]

Each csv comes from
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off \
        -k regex:k_ -c <one step's launches> --csv --log-file <csv> python bench.py --workload W --frames F --steps 1 --warmup 3 --kernel-only
bench.py prints `roofline.traffic` from this file only while the hash of retrofire_b200/csrc matches (a stale figure is never shown).
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

out = {"csrc_sha": bench.csrc_sha(), "captures": {}}
for a in sys.argv[1:]:
    key, path = a.split("=")
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = rows[0]
    ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
    per = {}
    for r in rows[1:]:
        per.setdefault((r[ii], r[ki].split("(")[0]), {})[r[mi]] = float(r[vi].replace(",", ""))
    tot = sum(m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"] for m in per.values())
    ras = [m for (i, k), m in per.items() if "k_raster" in k]
    assert len(ras) == 1, "capture exactly one step (one k_raster launch)"
    out["captures"][key] = {"dram_bytes_read": ras[0]["dram__bytes_read.sum"], "dram_bytes_write": ras[0]["dram__bytes_write.sum"],
                            "raster_us_under_ncu": ras[0]["gpu__time_duration.sum"] / 1e3, "pass_dram_bytes": tot,
                            "kernels": {k: {"us": m["gpu__time_duration.sum"] / 1e3, "dram_read": m["dram__bytes_read.sum"], "dram_write": m["dram__bytes_write.sum"]}
                                        for (i, k), m in per.items()}}
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_raster_traffic.json"), "w"), indent=1)
print(json.dumps({k: (v["dram_bytes_read"] + v["dram_bytes_write"], v["pass_dram_bytes"]) for k, v in out["captures"].items()}))
