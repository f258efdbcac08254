import csv, subprocess, collections, sys
rep=sys.argv[1]
out=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
# find line ranges from the source
import re
core=open("/root/repo/minigrid-rl_b200/csrc/mgrl_core.cuh").read().splitlines()
def find(pat):
    for i,l in enumerate(core,1):
        if pat in l: return i
    return 0
L_philox=find("MGRL_HD void philox4x32_10"); L_step=find("MGRL_HD StepOut env_step"); L_obs=find("// -------------------------------------------------------------------------- observation")
L_full=find("MGRL_HD void encode_full"); L_gen=find("MGRL_HD void generate("); L_adopt=find("MGRL_HD void adopt_layout")
agg=collections.defaultdict(lambda:[0,0,0]); cur=None; kernel=first=None; col={}
for r in rows:
    if not r: continue
    if r[0]=="File Path": cur=r[1].split("/")[-1]; continue
    if r[0]=="Function Name": kernel=r[1]; first=first or kernel; continue
    if r[0]=="Line No": col={n:i for i,n in enumerate(r)}; continue
    if kernel!=first: continue
    try: line=int(r[0])
    except: continue
    def grp(f,l):
        if f=="mgrl_core.cuh":
            if L_philox<=l<L_step: return "philox"
            if L_step<=l<L_obs: return "env_step"
            if L_obs<=l<L_full: return "encode"
            if L_gen<=l<L_adopt: return "generate"
            if l>=L_adopt: return "adopt"
            return "core_helpers"
        if f=="mgrl_kernels.cu": return "kernel.cu"
        return f
    try: vals=[float(r[col[k]] or 0) for k in ("Instructions Executed","Thread Instructions Executed","# Samples")]
    except (ValueError,IndexError): continue
    a=agg[grp(cur,line)]
    a[0]+=vals[0]; a[1]+=vals[1]; a[2]+=vals[2]
ti=sum(a[0] for a in agg.values()); ts=sum(a[2] for a in agg.values())
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][0]):
    print(f"{k:28s} warp-inst {a[0]:12.0f} ({100*a[0]/ti:5.1f}%) lanes {a[1]/max(a[0],1):5.1f} samples {100*a[2]/ts:5.1f}%")
print("total warp-inst", ti)
