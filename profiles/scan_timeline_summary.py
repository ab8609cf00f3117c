"""Summary of the TL lines a -DSCAN_TIMELINE build prints (see scan_timeline.py).  usage: python profiles/scan_timeline_summary.py <log>"""
import numpy as np, collections, sys
runs=[];cur=None
for l in open(sys.argv[1]):
    if l.startswith('RUN'): cur=[]; runs.append(cur)
    elif l.startswith('TL') and cur is not None:
        p=l.split(); cur.append([int(x) for x in p[1:]])
a=np.array(runs[-1])
t0=a[:,2].min()
first=(a[:,3]-t0)/1e3; end=(a[:,4]-t0)/1e3
print("warps",len(a),"first data us median %.1f"%np.median(first[first>0]))
print("end us: min %.1f p10 %.1f median %.1f p90 %.1f max %.1f"%(end.min(),np.percentile(end,10),np.median(end),np.percentile(end,90),end.max()))
print("strips per warp: min %d median %d max %d"%(a[:,5].min(),np.median(a[:,5]),a[:,5].max()))
