import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--kernel-name","regex:"+pat],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[1]; ix={h:i for i,h in enumerate(hdr)}
data=[]
for r in rows[2:]:
    if r and r[0]=='Kernel Name': break
    data.append(r)
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot={s:0 for s in stalls}; samples=0; lines=[]
for r in data:
    try: n=int(r[ix['# Samples']])
    except: continue
    samples+=n
    best=('',0)
    for s in stalls:
        try:
            v=int(r[ix[s]]); tot[s]+=v
            if v>best[1]: best=(s,v)
        except: pass
    lines.append((n,r[ix['Source']][:90], int(r[ix['Instructions Executed']] or 0), best[0]))
print("samples",samples,"warp-inst",sum(l[2] for l in lines))
for s,v in sorted(tot.items(), key=lambda x:-x[1])[:7]: print(f"  {s:26s} {100*v/samples:5.1f}%")
lines.sort(key=lambda x:-x[0])
for n,src,ie,st in lines[:top]: print(f"{100*n/samples:5.1f}% inst={ie:9d} {st[6:18]:12s} {src}")
