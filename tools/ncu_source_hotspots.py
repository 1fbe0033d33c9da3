"""Share of a kernel's stall samples per SASS instruction from an ncu capture taken with --import-source on (source page):
    python tools/ncu_source_hotspots.py <file.ncu-rep> <kernel regex> <launch index among the matches> [top n]"""
import csv,collections,subprocess,sys,io
rep,kern,skip=sys.argv[1],sys.argv[2],sys.argv[3]
raw=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--kernel-name","regex:"+kern,"--launch-skip",skip,"--launch-count","1"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw)))
print(rows[0][1][:120])
hdr=rows[1]; data=[r for r in rows[2:] if len(r)==len(hdr)]
ia=hdr.index("Source"); isamp=hdr.index("Warp Stall Sampling (All Samples)"); iex=hdr.index("Instructions Executed")
stall_cols=[i for i,h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
seen=set(); uniq=[]
for r in data:
    if r[0] in seen: continue
    seen.add(r[0]); uniq.append(r)
tot=sum(int(r[isamp]) for r in uniq if r[isamp].isdigit())
print("unique instr",len(uniq),"samples",tot)
agg={hdr[i]:0 for i in stall_cols}
for r in uniq:
    for i in stall_cols:
        if r[i].isdigit(): agg[hdr[i]]+=int(r[i])
s=sum(agg.values()); print(sorted(((round(100*v/s,1),k) for k,v in agg.items() if v),reverse=True)[:8])
top=sorted(uniq,key=lambda r:-int(r[isamp]) if r[isamp].isdigit() else 0)[:int(sys.argv[4]) if len(sys.argv)>4 else 16]
for r in top:
    st=sorted(((int(r[i]) if r[i].isdigit() else 0, hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
    print(f"{100*int(r[isamp])/tot:5.2f}% ex {r[iex]:>8} {r[ia].strip()[:60]:60s} {st}")
