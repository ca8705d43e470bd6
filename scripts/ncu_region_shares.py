"""Share of samples and instructions per code region (file, first line, last
line) of one kernel in an .ncu-rep captured with --import-source on
(development aid).

    python scripts/ncu_region_shares.py REP KERNEL_REGEX "{'name': ('file.cuh', lo, hi), ...}"
"""
import csv, subprocess, sys
rep, kernel = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass,cuda','-k','regex:'+kernel],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
cur=None; agg={}
for r in rows:
    if r and r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if len(r)>8 and r[2]=='-' and r[0].isdigit():
        try: agg[(cur,int(r[0]))]=(int(r[6]),int(r[7]))
        except ValueError: pass
ts=sum(v[0] for v in agg.values()); ti=sum(v[1] for v in agg.values())
regions=eval(sys.argv[3])
out={}
for (f,l),(s,n) in agg.items():
    name='other:'+f
    for rn,(rf,a,b) in regions.items():
        if f==rf and a<=l<=b: name=rn; break
    o=out.setdefault(name,[0,0]); o[0]+=s; o[1]+=n
for k,(s,n) in sorted(out.items(), key=lambda kv:-kv[1][0]):
    print(f'{k:34s} samples {100*s/ts:5.1f}%  inst {100*n/ti:5.1f}%')
