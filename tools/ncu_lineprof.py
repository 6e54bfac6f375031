"""Executed warp instructions and stall samples per CUDA source line of an .ncu-rep captured with
--import-source on.  usage: ncu_lineprof.py rep n_samples [min instructions per sample to list]"""
import csv, sys, subprocess
rep=sys.argv[1]; nq=float(sys.argv[2]); thr=float(sys.argv[3]) if len(sys.argv)>3 else 4
out=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
cur=None; agg=[]; hdr=None
for r in rows:
    if not r: continue
    if r[0]=="File Path": cur=r[1].split('/')[-1]; continue
    if r[0]=="Line No": hdr=r; continue
    if r[0] in ("Function Name","Kernel Name"): continue
    if hdr and len(r)>20 and r[2]=="-" and r[0].isdigit():
        d=dict(zip(hdr,r)); d['file']=cur; agg.append(d)
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot=sum(int(d['# Samples']) for d in agg); ti=sum(int(d['Instructions Executed']) for d in agg)
print("instr/sample",ti/nq,"samples",tot)
tots={s:sum(int(d[s] or 0) for d in agg) for s in stalls}
print({k[6:]:round(100*v/tot,1) for k,v in sorted(tots.items(), key=lambda x:-x[1])[:9]})
import os
_root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "3d-reconstruction-from-point-cloud_b200", "csrc")
_text = {}
def _line(fn, no):                      # the line of the repo's own source (the capture carries none in CSV)
    if fn not in _text:
        try: _text[fn] = open(os.path.join(_root, fn)).read().split("\n")
        except OSError: _text[fn] = []
    t = _text[fn]
    return t[no - 1].strip() if 0 < no <= len(t) else ""
for d in agg:
    if d['Source'].strip() in ("", "-"): d['Source'] = _line(d['file'], int(d['Line No']))
for d in sorted(agg,key=lambda d:(d['file'],int(d['Line No']))):
    ie=int(d['Instructions Executed'])/nq; sm=int(d['# Samples'])
    if ie>=thr or sm>=0.008*tot:
        top=sorted(((int(d[s] or 0),s[6:]) for s in stalls),reverse=True)[:2]
        print(f"{d['file'][:14]:14s} {d['Line No']:>4s} ins {ie:6.1f} smp {100*sm/tot:4.1f}% {top[0][1]:>12s} {d['Source'].strip()[:70]}")
