"""Per-source-line instruction counts from an ncu report: python profiles/ncu_source.py <rep> [top_n]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
files, cur, hdr = {}, None, None
for r in rows:
    if len(r) == 2 and r[0] == 'File Name':
        cur = r[1]; hdr = None; continue
    if r and r[0] == 'Line No':
        hdr = r; continue
    if hdr and cur and len(r) == len(hdr):
        files.setdefault(cur, []).append(dict(zip(hdr, r)))
tot = 0; items = []
for f, ls in files.items():
    for l in ls:
        try: n = int(l.get('Instructions Executed', '0') or 0)
        except ValueError: n = 0
        try: smp = int(l.get('# Samples', '0') or 0)
        except ValueError: smp = 0
        tot += n; items.append((n, smp, f.split('/')[-1], l['Line No'], l['Source'].strip()[:110]))
stot = sum(i[1] for i in items) or 1
print(f"total warp-instructions {tot}, samples {stot}")
for n, smp, f, ln, src in sorted(items, reverse=True)[:top]:
    print(f"{100*n/max(tot,1):5.1f}% inst {100*smp/stot:5.1f}% smp  {f}:{ln:>4s}  {src}")
