"""Executed-instruction share by SASS block: python profiles/ncu_blocks.py <rep> [block=40]"""
import csv, subprocess, sys
rep = sys.argv[1]; B = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(r for r in rows if r and r[0] == 'Address'); i0 = rows.index(hdr); ix = {n: i for i, n in enumerate(hdr)}
ins = [(r[ix['Source']].strip(), int(r[ix['Instructions Executed']] or 0), int(r[ix['# Samples']] or 0)) for r in rows[i0 + 1:] if len(r) == len(hdr)]
tot = sum(i[1] for i in ins); st = sum(i[2] for i in ins) or 1
print(f"{len(ins)} SASS instructions, {tot} executed warp-instructions, {st} samples")
for b in range(0, len(ins), B):
    blk = ins[b:b + B]; n = sum(i[1] for i in blk); s = sum(i[2] for i in blk)
    ops = {}
    for src, c, _ in blk:
        t = src.split()
        op = t[1] if t[0].startswith('@') else t[0]
        ops[op.split('.')[0]] = ops.get(op.split('.')[0], 0) + c
    top = sorted(ops.items(), key=lambda kv: -kv[1])[:6]
    if n / tot > 0.008 or s / st > 0.02:
        print(f"{b:5d}-{b+B:5d} {100*n/tot:5.1f}% inst {100*s/st:5.1f}% smp  max-exec {max(i[1] for i in blk)/1e3:8.0f}k  " + ' '.join(f"{k}:{v/1e6:.1f}M" for k, v in top))
