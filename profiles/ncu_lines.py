"""Per-source-line table of an ncu report, per kernel: executed warp-instructions, stall samples and the dominant stall
reasons of every hot line.

    python profiles/ncu_lines.py <file.ncu-rep> [top_n]

Uses `ncu --page source --print-source cuda,sass --csv` (the plain `cuda` view of this ncu version carries no metric
columns); the rows whose "Line No" is set are ncu's own per-line aggregates of the SASS rows that follow them.
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
kernels = {}  # function -> {(file, line): [inst, samples, {stall: n}, source, shared excessive wavefronts]}
path = func = hdr = None
for r in csv.reader(out.splitlines()):
    if len(r) >= 2 and r[0] == "File Path":
        path = r[1].split("/")[-1]
        continue
    if len(r) >= 2 and r[0] == "Function Name":
        func = r[1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if not hdr or len(r) != len(hdr) or not r[0].strip().isdigit():
        continue
    ix = {n: i for i, n in enumerate(hdr)}  # duplicate names ("Source"): the later (SASS) column wins, unused here

    def num(name):
        try:
            return int(r[ix[name]] or 0) if name in ix else 0
        except ValueError:
            return 0

    stalls = {n[6:]: num(n) for n in hdr if n.startswith("stall_") and "(" not in n}
    e = kernels.setdefault(func, {}).setdefault((path, int(r[0])), [0, 0, {}, r[1].strip(), 0])
    e[0] += num("Instructions Executed")
    e[1] += num("# Samples")
    e[4] += num("L1 Wavefronts Shared Excessive")
    for k, v in stalls.items():
        e[2][k] = e[2].get(k, 0) + v

for func, lines in kernels.items():
    tot = sum(e[0] for e in lines.values()) or 1
    stot = sum(e[1] for e in lines.values()) or 1
    print(f"\nkernel: {func}\n  {tot} executed warp-instructions, {stot} stall samples; lines sorted by samples")
    agg = {}
    for e in lines.values():
        for k, v in e[2].items():
            agg[k] = agg.get(k, 0) + v
    print("  stall samples by reason: " + ", ".join(f"{k} {100 * v / stot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    for (f, ln), e in sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]:
        why = " ".join(f"{k}:{100 * v / max(e[1], 1):.0f}" for k, v in sorted(e[2].items(), key=lambda kv: -kv[1])[:3] if v)
        bank = f" bankx{e[4] / 1e6:.1f}M" if e[4] > 1e5 else ""
        print(f"  {100 * e[1] / stot:5.1f}% smp {100 * e[0] / tot:5.1f}% inst  {f}:{ln:<4d} [{why}]{bank}  {e[3][:100]}")
