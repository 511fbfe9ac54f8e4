"""CPU analysis of the headline frame (no GPU): how many (Gaussian, tile) intersections and (warp rectangle, Gaussian)
candidates can contribute at all, i.e. what tile-level / bounding-box culling could remove from the blend kernels.
Uses the test-infrastructure projection of oracle/ (development aid, not part of the product)."""
import sys, math, time, numpy as np, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from oracle import oracle as O
torch.set_num_threads(8)
raw, viewmat, K, target, n_total = bench.build_workload(0,1,3)
means=raw["means"]; quats=raw["quats"]; scales=torch.exp(raw["scales"]); opac=torch.sigmoid(raw["opacities"])
W,H=1920,1080
with torch.no_grad():
    radii, m2, depths, conics, _ = O.projection(means,None,quats,scales,viewmat,K,W,H,0.3,0.01,100.0,3.0)
radii=radii[0].numpy(); m2=m2[0].numpy(); conics=conics[0].numpy(); op=opac.numpy()
vis=np.nonzero(radii>0)[0]
print("visible",len(vis))
ts=16; tw,th=120,68
r=radii[vis].astype(np.float32)/ts; cx=m2[vis,0]/ts; cy=m2[vis,1]/ts
x0=np.clip(np.floor(cx-r),0,tw).astype(int); x1=np.clip(np.ceil(cx+r),0,tw).astype(int)
y0=np.clip(np.floor(cy-r),0,th).astype(int); y1=np.clip(np.ceil(cy+r),0,th).astype(int)
cnt=(x1-x0)*(y1-y0)
print("isects",cnt.sum())
# expand
gi=np.repeat(np.arange(len(vis)),cnt)
off=np.concatenate([[0],np.cumsum(cnt)[:-1]])
k=np.arange(cnt.sum())-np.repeat(off,cnt)
wx=np.repeat(x1-x0,cnt)
tx=np.repeat(x0,cnt)+k%wx; ty=np.repeat(y0,cnt)+k//wx
g=vis[gi]
a=conics[g,0].astype(np.float64); b=conics[g,1].astype(np.float64); c=conics[g,2].astype(np.float64)
mx=m2[g,0].astype(np.float64); my=m2[g,1].astype(np.float64)
o=op[g].astype(np.float64)
qmax=2*np.log(np.maximum(255*o,1e-30))
def minq(x0,x1,y0,y1):
    lx0=x0-mx; lx1=x1-mx; ly0=y0-my; ly1=y1-my
    inside=(lx0<=0)&(lx1>=0)&(ly0<=0)&(ly1>=0)
    q=np.full(len(mx),np.inf)
    for lx in (lx0,lx1):
        dy=np.clip(-b*lx/c,ly0,ly1); q=np.minimum(q,a*lx*lx+2*b*lx*dy+c*dy*dy)
    for ly in (ly0,ly1):
        dx=np.clip(-b*ly/a,lx0,lx1); q=np.minimum(q,a*dx*dx+2*b*dx*ly+c*ly*ly)
    q[inside]=0
    return q
# tile-level (pixel centres 0.5..15.5)
qt=minq(tx*16+0.5,tx*16+15.5,ty*16+0.5,ty*16+15.5)
hit_tile=(qt<=qmax)&(o*255>=1)
print("tile-level hits: %.3f of isects"%hit_tile.mean())
# bbox-level reject for the tile (axis-aligned extents of the 1/255 ellipse)
det=a*c-b*b
ex=np.sqrt(np.maximum(qmax,0)*c/det); ey=np.sqrt(np.maximum(qmax,0)*a/det)
bb=(mx+ex>=tx*16+0.5)&(mx-ex<=tx*16+15.5)&(my+ey>=ty*16+0.5)&(my-ey<=ty*16+15.5)&(o*255>=1)
print("tile-level bbox pass: %.3f ; exact among bbox-pass: %.3f"%(bb.mean(),(hit_tile&bb).sum()/bb.sum()))
# warp-rect level: 8 rects of 8x4 per tile
tot=0; tot_bb=0
for wy in range(4):
    for wxx in range(2):
        X0=tx*16+wxx*8+0.5; X1=X0+7; Y0=ty*16+wy*4+0.5; Y1=Y0+3
        q=minq(X0,X1,Y0,Y1); h=(q<=qmax)&(o*255>=1); tot+=h.sum()
        bbw=(mx+ex>=X0)&(mx-ex<=X1)&(my+ey>=Y0)&(my-ey<=Y1)&(o*255>=1); tot_bb+=bbw.sum()
n=len(mx)
print("warp-rect survivors: %.2f M of %.2f M candidates (%.3f); bbox-pass %.2f M (%.3f)"%(tot/1e6,8*n/1e6,tot/(8*n),tot_bb/1e6,tot_bb/(8*n)))
print("warp-rect survivors per tile-hit isect: %.2f"%(tot/hit_tile.sum()))
# ---- round 2: survivors per rectangle size (sizes the pixels-per-thread trade-off of the stream blend) ----------------
for (rw, rh) in ((4, 4), (8, 4), (8, 8), (16, 8), (16, 16)):
    tot = 0
    for wy in range(16 // rh):
        for wxx in range(16 // rw):
            X0 = tx * 16 + wxx * rw + 0.5; X1 = X0 + rw - 1; Y0 = ty * 16 + wy * rh + 0.5; Y1 = Y0 + rh - 1
            ok = (X0 < W) & (Y0 < H)
            q = minq(X0, np.minimum(X1, W - 0.5), Y0, np.minimum(Y1, H - 0.5)); tot += ((q <= qmax) & (o * 255 >= 1) & ok).sum()
    print("rect %2dx%-2d: %.2f M (rect, Gaussian) survivors = %.1f M pixel slots" % (rw, rh, tot / 1e6, tot * rw * rh / 1e6))
# pixels that can pass alpha >= 1/255: area of the ellipse d^T A d <= qmax (unclipped)
gv = vis; av = conics[gv, 0].astype(np.float64); bv = conics[gv, 1].astype(np.float64); cv = conics[gv, 2].astype(np.float64)
qv = 2 * np.log(np.maximum(255 * op[gv].astype(np.float64), 1e-30)); dv = av * cv - bv * bv
area = np.where((qv > 0) & (dv > 0), math.pi * np.maximum(qv, 0) / np.sqrt(np.maximum(dv, 1e-30)), 0)
print("sum of 1/255-ellipse areas: %.1f M pixel pairs (upper bound of contributing pairs before saturation)" % (area.sum() / 1e6))
