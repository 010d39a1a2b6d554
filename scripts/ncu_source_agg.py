"""Stall samples and executed instructions per CUDA source line of a capture.
usage: ncu -i <rep> --page source --csv --print-source cuda,sass > src.csv ; ncu_source_agg.py src.csv [top] [samp|instr]"""
import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
top=int(sys.argv[2]) if len(sys.argv)>2 else 45
key=sys.argv[3] if len(sys.argv)>3 else 'samp'
cur=None; hdr=None
agg=[]
for r in rows:
    if len(r)==2 and r[0]=="File Path": cur=r[1].split('/')[-1]; continue
    if len(r)>5 and r[0]=="Line No": hdr=r; continue
    if hdr and len(r)==len(hdr) and r[0]!="":
        try:
            s=int(r[hdr.index('# Samples')] or 0); e=int(r[hdr.index('Instructions Executed')] or 0)
        except: continue
        agg.append((s,e,cur,r[0],r[1].strip()[:100], r))
ts=sum(a[0] for a in agg); te=sum(a[1] for a in agg)
print(ts,te)
ks=['stall_long_sb','stall_wait','stall_short_sb','stall_barrier','stall_branch_resolving','stall_no_inst','stall_math','stall_mio','stall_lg','stall_not_selected','stall_selected','stall_dispatch']
tot={k:sum(int(a[5][hdr.index(k)] or 0) for a in agg) for k in ks}
print({k:round(100*v/ts,1) for k,v in tot.items()})
for s,e,f,l,src,r in sorted(agg,key=lambda a:-(a[0] if key=='samp' else a[1]))[:top]:
    st=sorted(((int(r[hdr.index(k)] or 0),k) for k in ks),reverse=True)[:2]
    print("%5.2f%% samp %5.2f%% instr %s:%s  %s  [%s]"%(100*s/ts,100*e/te,f,l,src," ".join("%s=%d%%"%(k[6:],100*v/max(s,1)) for v,k in st)))
