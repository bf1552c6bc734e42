"""L2 sectors per source line of one kernel: python profiles/ncu_l2_lines.py <rep> <kernel substring>"""
import csv, subprocess, sys
rep=sys.argv[1]; kern=sys.argv[2]
out = subprocess.run(["ncu","-i",rep,"--page","source","--print-source","cuda,sass","--csv"],capture_output=True,text=True).stdout
func=None; hdr=None; path=None; rows={}
for r in csv.reader(out.splitlines()):
    if len(r)>=2 and r[0]=="File Path": path=r[1].split('/')[-1]; continue
    if len(r)>=2 and r[0]=="Function Name": func=r[1]; continue
    if r and r[0]=="Line No": hdr=r; continue
    if not hdr or len(r)!=len(hdr) or kern not in func: continue
    if r[0].strip().isdigit():
        ix={n:i for i,n in enumerate(hdr)}
        def num(n):
            try: return int(r[ix[n]] or 0)
            except: return 0
        e=rows.setdefault((path,int(r[0])),[0,0,0,0,r[1].strip()[:90]])
        e[0]+=num("L2 Theoretical Sectors Global"); e[1]+=num("L2 Theoretical Sectors Global Excessive"); e[2]+=num("L2 Theoretical Sectors Local"); e[3]+=num("L1 Tag Requests Global")
tot=sum(e[0] for e in rows.values()); print("total L2 theoretical sectors global", tot, "local", sum(e[2] for e in rows.values()))
for k,e in sorted(rows.items(), key=lambda kv:-kv[1][0])[:25]:
    print(f"{e[0]/1e6:8.2f}M sectors  excess {e[1]/1e6:8.2f}M  local {e[2]/1e6:6.2f}M  tagreq {e[3]/1e6:7.2f}M  {k[0]}:{k[1]} {e[4]}")
