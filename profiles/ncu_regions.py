"""Instruction share and lane efficiency of code regions of a kernel, from an ncu report with source info:
    python profiles/ncu_regions.py <file.ncu-rep> [file:lo-hi=name ...]     (default regions: per source file)
"""
import collections, csv, subprocess, sys
rep = sys.argv[1]
regions = []
for a in sys.argv[2:]:
    loc, name = a.split('=')
    f, r = loc.split(':'); lo, hi = r.split('-')
    regions.append((f, int(lo), int(hi), name))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
path = hdr = None
rows = []
for r in csv.reader(out.splitlines()):
    if len(r) >= 2 and r[0] == "File Path": path = r[1].split('/')[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if not hdr or len(r) != len(hdr) or not r[0].strip().isdigit(): continue
    ix = {n: i for i, n in enumerate(hdr)}
    def num(n):
        try: return int(r[ix[n]] or 0)
        except ValueError: return 0
    rows.append((path, int(r[0]), num("Instructions Executed"), num("Thread Instructions Executed"), num("# Samples") if "# Samples" in ix else 0, r[1].strip()))
tot = sum(x[2] for x in rows)
b = collections.Counter(); bt = collections.Counter()
def region(p, l):
    for f, lo, hi, name in regions:
        if p == f and lo <= l <= hi: return name
    return p
for p, l, i, t, smp, src in rows:
    b[region(p, l)] += i; bt[region(p, l)] += t
print(f"total {tot} warp instructions")
for k, v in b.most_common(): print(f"{v / tot * 100:5.1f}% inst  lane-eff {bt[k] / max(v, 1) / 32 * 100:4.0f}%  {k}")
print('--- top lines by instructions')
for p, l, i, t, smp, src in sorted(rows, key=lambda x: -x[2])[:int(__import__('os').environ.get('TOP', 30))]:
    print(f"{i / tot * 100:5.1f}% eff {t / max(i, 1) / 32 * 100:3.0f}% {p}:{l}  {src[:120]}")
