"""Per-kernel tables of a record session (scratch/r02_final.sh) from its ncu csv files:

    python profiles/make_tables.py <launches.csv> <dram_per_kernel.csv> [title]

prints (1) the launch list folded per kernel (count, average duration, share of the summed kernel time) and (2) DRAM / L2 bytes
of every kernel of ONE step. Durations are ncu's (cold caches, kernels serialised): the SHARES are what the bench line's
`kernel_time_share` must agree with, not the absolute times.
"""
import collections
import csv
import sys


def rows(path):
    rs = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    h = rs[0]
    ix = {n: h.index(n) for n in ("ID", "Kernel Name", "Metric Name", "Metric Value")}
    per = collections.OrderedDict()
    for r in rs[1:]:
        try:
            per.setdefault((r[ix["ID"]], r[ix["Kernel Name"]].split("(")[0]), {})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
        except ValueError:
            pass
    return per


launches, dram = rows(sys.argv[1]), rows(sys.argv[2])
print("# " + (sys.argv[3] if len(sys.argv) > 3 else "record session"))
print("# launch list: ncu --metrics gpu__time_duration.sum --clock-control none, bench.py --steps 4 --warmup 3 --kernel-only")
fold = collections.OrderedDict()
for (i, k), m in launches.items():
    fold.setdefault(k, []).append(m["gpu__time_duration.sum"] / 1e3)
tot = sum(sum(v) for v in fold.values())
for k, v in sorted(fold.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:<50s} n={len(v):3d} avg={sum(v) / len(v):9.1f} us share={sum(v) / tot:.3f}")
print()
print("# DRAM / L2 bytes of every kernel of ONE step: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_{read,write}.sum")
print(f"{'kernel':<34s}{'us':>9s}{'dram rd MB':>12s}{'dram wr MB':>12s}{'L2 rd MB':>10s}{'L2 wr MB':>10s}")
tr = tw = 0.0
for (i, k), m in dram.items():
    rd, wr = m["dram__bytes_read.sum"], m["dram__bytes_write.sum"]
    tr += rd; tw += wr
    print(f"{k:<34s}{m['gpu__time_duration.sum'] / 1e3:9.1f}{rd / 1e6:12.1f}{wr / 1e6:12.1f}"
          f"{m['lts__t_sectors_op_read.sum'] * 32 / 1e6:10.1f}{m['lts__t_sectors_op_write.sum'] * 32 / 1e6:10.1f}")
print(f"total DRAM per step: read {tr / 1e9:.2f} GB + write {tw / 1e9:.2f} GB = {(tr + tw) / 1e9:.2f} GB")
