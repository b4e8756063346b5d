"""Offline study of cost predictors for the work-list order: per-configuration cycles of a
traced run (RC_TRACE_DUMP of profiles/tools/trace.py -> gpurun_out/trace_dump.npz) against features that
classify_kernel could compute per thread, and a list-scheduling model (4,736 warps, no hand-over) of the
resulting order.  CPU only.  Results of round 1: profiles/predictor_study_r01.txt."""
import numpy as np, sys, heapq
sys.path.insert(0,'/root/repo')
import robocheck_b200 as rc
d=np.load('/root/repo/gpurun_out/trace_dump.npz')
t=d['trace']; cyc=t[:,0].astype(np.int64); confs=d['confs'].astype(np.float64)
listed=np.nonzero(cyc>0)[0]; c=cyc[listed].astype(np.float64); q=confs[listed]
rob,obs=rc.alpha_meshes()
rx,ry,rz,r1,r2,r3=rob.arrays(); ox,oy,oz,o1,o2,o3=obs.arrays()
RV=np.stack([rx,ry,rz],1).astype(np.float64); OV=np.stack([ox,oy,oz],1).astype(np.float64)
def kd_groups(P, k):
    groups=[np.arange(len(P))]
    while len(groups)<k:
        new=[]
        for g in groups:
            ext=P[g].max(0)-P[g].min(0); a=int(np.argmax(ext)); o=g[np.argsort(P[g,a])]; h=len(o)//2
            new+= [o[:h],o[h:]]
        groups=new
    return groups
def boxes(P, groups):
    lo=np.array([P[g].min(0) for g in groups]); hi=np.array([P[g].max(0) for g in groups])
    return (lo+hi)/2,(hi-lo)/2
# obstacle: groups of triangle centroids, boxes over triangle vertices
OT=np.stack([o1,o2,o3],1); ocent=OV[OT].mean(1)
def obs_boxes(K):
    gs=kd_groups(ocent,K); lo=[];hi=[]
    for g in gs:
        v=OV[OT[g].reshape(-1)]; lo.append(v.min(0)); hi.append(v.max(0))
    lo=np.array(lo);hi=np.array(hi); return (lo+hi)/2,(hi-lo)/2
def rot(q):
    x,y,z,pitch,yaw,roll=q.T
    sA,cA,sB,cB,sC,cC=np.sin(yaw),np.cos(yaw),np.sin(pitch),np.cos(pitch),np.sin(roll),np.cos(roll)
    R=np.empty((len(q),3,3))
    R[:,0,0]=cA*cB; R[:,0,1]=cA*sB*sC-sA*cC; R[:,0,2]=cA*sB*cC+sA*sC
    R[:,1,0]=sA*cB; R[:,1,1]=sA*sB*sC+cA*cC; R[:,1,2]=sA*sB*cC-cA*sC
    R[:,2,0]=-sB; R[:,2,1]=cB*sC; R[:,2,2]=cB*cC
    return R,q[:,:3]
R,T=rot(q)
obox_lo,obox_hi=OV.min(0),OV.max(0)
def feature(G,K,sat6=False):
    gc,gh=boxes(RV,kd_groups(RV,G))     # robot groups (vertices)
    wc=np.einsum('nij,gj->ngi',R,gc)+T[:,None,:]
    wh=np.einsum('nij,gj->ngi',np.abs(R),gh)
    if K==0:
        oc=((obox_lo+obox_hi)/2)[None]; oh=((obox_hi-obox_lo)/2)[None]
    else:
        oc,oh=obs_boxes(K)
    cnt=np.zeros(len(q))
    for k in range(len(oc)):
        ov=np.all(np.abs(wc-oc[k])<=wh+oh[k],axis=2)   # world-axes test only
        cnt+=ov.sum(1)
    return cnt
lc=np.log(c)
W=148*32
def makespan(order):
    h=[0.0]*W; heapq.heapify(h); end=0
    for j in order:
        s=heapq.heappop(h); e=s+c[j]; heapq.heappush(h,e)
        end=max(end,e)
    return end
bound=c.sum()/W
rng=np.random.default_rng(0)
def report(name,f,ncls=8):
    r=np.corrcoef(f,lc)[0,1]; rs=np.corrcoef(np.argsort(np.argsort(f)),np.argsort(np.argsort(c)))[0,1]
    # classes by quantiles of f (ties fine)
    qs=np.quantile(f,np.linspace(0,1,ncls+1)[1:-1]); cls=np.searchsorted(qs,f,side='right')
    order=np.concatenate([rng.permutation(np.nonzero(cls==k)[0]) for k in range(ncls-1,-1,-1)])
    print(f'{name:28s} r(log c) {r:.3f} spearman {rs:.3f}  makespan/bound: {ncls} classes {makespan(order)/bound:.3f}  sorted {makespan(np.argsort(-f, kind="stable"))/bound:.3f}')
report('G32 x obstacle AABB (current)',feature(32,0))
for G,K in ((32,8),(32,16),(32,32),(16,8),(16,16),(64,16),(32,64)):
    report(f'G{G} x K{K} world-axes',feature(G,K))
print('--- more features')
def feature2(G,K,mode):
    gc,gh=boxes(RV,kd_groups(RV,G))
    wc=np.einsum('nij,gj->ngi',R,gc)+T[:,None,:]
    wh=np.einsum('nij,gj->ngi',np.abs(R),gh)
    oc,oh=obs_boxes(K)
    out=np.zeros(len(q))
    for k in range(len(oc)):
        gap=np.abs(wc-oc[k])-(wh+oh[k])          # n,g,3  (negative = overlap depth)
        ov=np.all(gap<=0,axis=2)
        if mode=='vol':
            depth=np.clip(-gap,0,None); depth=np.minimum(depth,2*np.minimum(wh,oh[k]))
            out+=(depth.prod(2)*ov).sum(1)
        elif mode=='sat6':
            # robot axes too: project obstacle box on robot group axes
            dloc=np.einsum('nji,ngj->ngi',R,oc[k]-wc)      # obstacle centre in group frame (n,g,3)
            ohl=np.einsum('nji,j->ni',np.abs(R),oh[k])[:,None,:]
            ov2=np.all(np.abs(dloc)<=gh[None]+ohl,axis=2)
            out+=(ov&ov2).sum(1)
        elif mode=='mindepth':
            depth=np.clip(-gap,0,None).min(2)
            out+=(depth*ov).sum(1)
    return out
for G,K,mode in ((16,8,'sat6'),(32,16,'sat6'),(16,8,'vol'),(16,8,'mindepth'),(32,32,'sat6')):
    report(f'G{G} x K{K} {mode}',feature2(G,K,mode))
f0=feature(32,0)
print('current feature histogram', np.bincount(f0.astype(int)))
for lo in (28,30,32):
    m=f0>=lo; print(lo,'share of configs',m.mean(),'share of cycles',c[m].sum()/c.sum(),'share of top1% configs',(m&(c>np.percentile(c,99))).sum()/(c>np.percentile(c,99)).sum())
