import sys; sys.path.insert(0,'tests')
import numpy as np, torch
import fastdem_b200 as fd
from fastdem_b200 import synthetic as syn
names=["prologue","entry","tma","sorted","reduced","chunks_done","compacted","estimated","listed","jobs_done","exit"]
for wname in ["c2_lidar64_local","c3_rgbd_p2","c1_vlp16_local"]:
    wl=syn.WORKLOADS[wname]
    m=fd.ElevationMap(wl.map_width,wl.map_height,wl.resolution); d=fd.FastDEM(m,wl.config())
    flush=torch.empty(256<<20,dtype=torch.uint8,device='cuda')
    rows=[]
    for k in range(12):
        s=syn.make_scan(wl,k%4)
        c=fd.PointCloud(torch.from_numpy(s['xyzw']).cuda(), None if s['intensity'] is None else torch.from_numpy(s['intensity']).cuda(), None if s['rgb'] is None else torch.from_numpy(s['rgb']).cuda())
        flush.fill_(k); torch.cuda.synchronize()
        d.integrate_stats(c,*syn.pose(wl,k))
        if k>=4:
            rows.append(d.debug_phase_clocks())
            ct=d.debug_cta_times().astype(np.int64); ct=ct[ct[:,0]>0]
            t0=ct[:,0].min(); starts=(ct[:,0]-t0)/1e3; ends=(ct[:,1]-t0)/1e3; dur=ends-starts
            if k==11: print(wname,'CTAs',len(ct),'span %.1f us'%ends.max(),'start p50/p90/max %.1f/%.1f/%.1f'%(np.percentile(starts,50),np.percentile(starts,90),starts.max()),'dur p50/p90/max %.1f/%.1f/%.1f'%(np.percentile(dur,50),np.percentile(dur,90),dur.max()))
    a=np.array(rows,dtype=np.float64)
    med=np.median(a,axis=0)
    print(wname,'records in CTA0 bucket (median)',med[15])
    prev=0
    for i,nm in enumerate(names):
        print(f"   {nm:12s} t={med[i]/1.965e3:7.2f} us  (+{(med[i]-prev)/1.965e3:6.2f})"); prev=med[i]
