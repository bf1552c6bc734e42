"""Hot SASS regions of an ncu report: python profiles/ncu_sass.py <rep> [min_pct]
Prints instructions in address order whose executed-count share >= min_pct (default 0.4 %), with running block totals."""
import csv, subprocess, sys
rep = sys.argv[1]; minp = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(r for r in rows if r and r[0] == 'Address')
i0 = rows.index(hdr)
ix = {n: i for i, n in enumerate(hdr)}
ins = []
for r in rows[i0 + 1:]:
    if len(r) != len(hdr): continue
    ins.append((r[ix['Source']].strip(), int(r[ix['Instructions Executed']] or 0), int(r[ix['Thread Instructions Executed']] or 0), int(r[ix['# Samples']] or 0)))
tot = sum(i[1] for i in ins); ttot = sum(i[2] for i in ins); stot = sum(i[3] for i in ins) or 1
print(f"SASS instructions {len(ins)}, executed warp-instr {tot}, avg active threads {ttot/max(tot,1):.1f}, samples {stot}")
for k, (src, n, tn, smp) in enumerate(ins):
    if 100 * n / tot >= minp or 100 * smp / stot >= 2 * minp:
        print(f"{k:5d} {100*n/tot:5.2f}% {100*smp/stot:5.2f}%smp thr={tn/max(n,1):4.1f}  {src[:90]}")
