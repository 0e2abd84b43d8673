#!/usr/bin/env python
"""CPU simulation of the shared-memory bank conflicts of the staged NCHW forward's blend on cfg1's RoI draw: wavefronts per warp-level LDS for
the compact row-segment layout the kernel uses, for a dense x + S*y placement with the narrowest odd stride, and with the best of 33 strides per tile
(compact 2.40, odd stride 2.17, best stride 1.33).  Uses the oracle for the sample centres (development tool)."""
import sys, numpy as np
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, ROOT)
import workloads as WL
from oracle import rroi_oracle as O
H,W,PH,PW=180,320,8,64
def wavefronts(addr, valid):
    # addr: [32] int float-index; valid mask; wavefronts = max over banks of distinct addresses in that bank
    a=addr[valid]
    if a.size==0: return 0
    banks=a%32
    w=0
    for b in np.unique(banks):
        w=max(w,len(np.unique(a[banks==b])))
    return w
tot_c=tot_d={}; 
res={'compact':[0,0]}
Ss=[None]
cands=list(range(33,97,2))
res.update({'dense_best':[0,0],'dense_w+pad1':[0,0]})
for img in range(6):
    r=WL.random_rois(img,64,0)
    f=np.zeros((1,1,H,W),np.float32)
    out,ix,iy=O.forward(f,r,PH,PW,0.25)
    ix=ix[:,0]; iy=iy[:,0]
    valid=WL.valid_counts(r,PH,PW)//PH
    for n in range(64):
        for tile in range(2):
            pw0=tile*32
            m=(np.arange(32)+pw0)[None,:]<valid[n]
            cx=ix[n][:,pw0:pw0+32]; cy=iy[n][:,pw0:pw0+32]
            taps=[]
            for fx,fy in ((np.floor,np.floor),(np.ceil,np.floor),(np.ceil,np.ceil),(np.floor,np.ceil)):
                x=fx(cx).astype(int); y=fy(cy).astype(int)
                ok=m&(x>0)&(x<W-1)&(y>0)&(y<H-1)
                taps.append((x,y,ok))
            allok=np.zeros_like(m)
            ys=[];xs=[]
            for x,y,ok in taps:
                ys.append(y[ok]); xs.append(x[ok])
            if sum(len(a) for a in ys)==0: continue
            yy=np.concatenate(ys); xx=np.concatenate(xs)
            y0=yy.min(); y1=yy.max(); nrows=y1-y0+1
            rlo=np.full(nrows,10**9); rhi=np.full(nrows,-1)
            for a,b in zip(yy,xx):
                rlo[a-y0]=min(rlo[a-y0],b); rhi[a-y0]=max(rhi[a-y0],b)
            gcnt=np.where(rhi>=rlo,((rhi|3)-(rlo&~3)+1)>>2,0)
            goff=np.concatenate([[0],np.cumsum(gcnt)])
            xmin=(xx.min()&~3); width=((xx.max()|3)-xmin+1)
            # compact
            for name in res:
                pass
            def run(addr_fn):
                w=0;c=0
                for x,y,ok in taps:
                    for ph in range(8):
                        v=ok[ph]
                        if not v.any(): continue
                        addr=addr_fn(x[ph],y[ph])
                        w+=wavefronts(addr,v); c+=1
                return w,c
            w,c=run(lambda x,y: goff[np.clip(y-y0,0,nrows-1)]*4+(x-(rlo[np.clip(y-y0,0,nrows-1)]&~3)))
            res['compact'][0]+=w; res['compact'][1]+=c
            S1=width+ (1 if width%2==0 else 0)   # odd stride
            w,c=run(lambda x,y: (y-y0)*S1+(x-xmin))
            res['dense_w+pad1'][0]+=w; res['dense_w+pad1'][1]+=c
            best=None
            for S in range(width, width+33):
                w,c=run(lambda x,y: (y-y0)*S+(x-xmin))
                if best is None or w<best[0]: best=(w,c,S)
            res['dense_best'][0]+=best[0]; res['dense_best'][1]+=best[1]
for k,(w,c) in res.items(): print(k, "avg wavefronts per warp LDS: %.2f"%(w/max(c,1)))
