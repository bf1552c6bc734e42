"""Top SASS instructions of one kernel by stall samples, with their source line: python profiles/ncu_top_sass.py <rep> <kernel substring> <n>"""
import csv, subprocess, sys
rep=sys.argv[1]; kern=sys.argv[2]; top=int(sys.argv[3])
out = subprocess.run(["ncu","-i",rep,"--page","source","--print-source","cuda,sass","--csv"],capture_output=True,text=True).stdout
func=None; hdr=None; cur_line=None; path=None
seen={}
for r in csv.reader(out.splitlines()):
    if len(r)>=2 and r[0]=="File Path": path=r[1].split('/')[-1]; continue
    if len(r)>=2 and r[0]=="Function Name": func=r[1]; continue
    if r and r[0]=="Line No": hdr=r; continue
    if not hdr or len(r)!=len(hdr): continue
    if kern not in func: continue
    if r[0].strip().isdigit(): cur_line=(path,int(r[0])); continue
    if r[2].startswith("0x"):
        addr=int(r[2],16)
        if addr in seen: continue
        ix={n:i for i,n in enumerate(hdr)}
        smp=int(r[6] or 0); inst=int(r[7] or 0)
        stalls={n[6:]:int(r[i] or 0) for i,n in enumerate(hdr) if n.startswith("stall_") and "(" not in n}
        seen[addr]=(smp,inst,r[3].strip(),cur_line,stalls)
base=min(seen)
tot=sum(v[0] for v in seen.values())
print("total samples",tot,"n sass",len(seen))
for addr,(smp,inst,src,line,st) in sorted(seen.items(), key=lambda kv:-kv[1][0])[:top]:
    why=" ".join(f"{k}:{v}" for k,v in sorted(st.items(), key=lambda kv:-kv[1])[:2] if v)
    print(f"{100*smp/tot:5.2f}% {inst/1e6:7.2f}M  +{addr-base:05x}  {line[0]}:{line[1]:<4d} {src[:60]:60s} [{why}]")
