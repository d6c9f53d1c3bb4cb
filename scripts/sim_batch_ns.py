"""CPU (numpy) model of the batched-replacement nested-sampling loop on the C2 problem, used to locate the
logZ bias of rwalk chains.  Diagnostic, not part of the product or of the test suite.

usage: python scripts/sim_batch_ns.py VARIANT [NLIVE SEED BATCH WALKS]
  VARIANT  exact    uniform draws from {L > L*} (iso-likelihood ellipsoid, cube rejection): checks the scheme
           base     rwalk chains, proposal ellipsoid estimated from the live points (what dynesty does)
           truecov  rwalk chains, proposal ellipsoid = the true iso-likelihood shape
"""
import sys, math, time, json
import numpy as np
n=50; N=int(sys.argv[2]) if len(sys.argv)>2 else 2000
variant=sys.argv[1]; seed=int(sys.argv[3]) if len(sys.argv)>3 else 1
K=int(sys.argv[4]) if len(sys.argv)>4 else 200
walks=int(sys.argv[5]) if len(sys.argv)>5 else 70
rng=np.random.default_rng(seed)
C=np.full((n,n),0.4); np.fill_diagonal(C,1.0); Cinv=np.linalg.inv(C); Lc=np.linalg.cholesky(C)
lnorm=-0.5*(n*math.log(2*math.pi)+np.linalg.slogdet(C)[1])
def logl(u):
    v=10*u-5; return -0.5*np.einsum('ij,jk,ik->i',v,Cinv,v)+lnorm
def bound_root(u):
    if variant=='truecov':
        # proposal shaped by the TRUE iso-likelihood ellipsoid through the worst live point
        v=10*u-5; r2=np.einsum('ij,jk,ik->i',v,Cinv,v).max()
        return (Lc*math.sqrt(r2)/10.0)*1.25**(1.0/n)
    ctr=u.mean(0); cov=np.cov(u,rowvar=False); d=u-ctr
    am=np.linalg.inv(cov); f=np.einsum('ij,jk,ik->i',d,am,d).max()
    cov=cov*f/(1-1e-3)
    A=np.linalg.cholesky(cov)*1.25**(1.0/n)
    return A
# initial live points: uniform in cube
u=rng.random((N,n)); l=logl(u)
logvol=0.0; logz=-1e300; lprev=-1e300
def lae(a,b):
    hi,lo=max(a,b),min(a,b); return hi+math.log1p(math.exp(lo-hi)) if lo>-1e299 else hi
ncall=N; it=0; scale=1.0; A=None; ncall_last=0; rounds=0
dlogz=1e-3*(N-1)+0.01
t0=time.time()
# phase 1: unit cube sampling, serial semantics (vectorised in chunks), until eff<10% & ncall>=2N
while True:
    worst=np.argmin(l); lw=l[worst]
    lmax=l.max()
    if lae(0.0,lmax+logvol-logz)<dlogz: break
    # replace worst by uniform draw from cube with l>lw
    while True:
        x=rng.random((64,n)); lx=logl(x); ncall+=1
        ok=np.nonzero(lx>lw)[0]
        # count calls until first success
        if len(ok): ncall+=ok[0]; x=x[ok[0]]; lx=lx[ok[0]]; break
        ncall+=63
    logvol-=math.log((N+1.)/N)
    logz=lae(logz, lae(lw,lprev)+logvol+math.log(0.5*(math.exp(math.log((N+1.)/N))-1)))
    lprev=lw; u[worst]=x; l[worst]=lx; it+=1
    if ncall>=2*N and 100.*it/ncall<10.: break
it1=it
A=bound_root(u); ncall_last=ncall
upd=walks*N
if variant.startswith('serial'): K=1
while True:
    order=np.argsort(l,kind='stable'); sl=l[order]
    if lae(0.0,sl[-1]+logvol-logz)<dlogz: break
    thr=sl[K-1]
    surv=order[K:]
    if variant=='exact':
        # exact uniform draws from {L>thr}: ellipsoid v'Cinv v < r2
        r2=-2*(thr-lnorm)
        cu=np.empty((K,n)); todo=np.arange(K); nc=0
        while len(todo):
            z=rng.standard_normal((len(todo),n)); z*= (rng.random(len(todo))**(1./n)/np.linalg.norm(z,axis=1))[:,None]
            v=math.sqrt(r2)*z@Lc.T; x=(v+5)/10; nc+=len(todo)
            ok=np.all((x>0)&(x<1),axis=1)
            cu[todo[ok]]=x[ok]; todo=todo[~ok]
        cl=logl(cu); na=K; nr=0
    else:
        st=surv[rng.integers(len(surv),size=K)]
        cu=u[st].copy(); cl=l[st].copy(); na=0; nr=0
        for s in range(walks):
            z=rng.standard_normal((K,n)); z*=(rng.random(K)**(1./n)/np.linalg.norm(z,axis=1))[:,None]
            p=cu+scale*z@A.T
            inc=np.all((p>0)&(p<1),axis=1)
            lp=np.where(inc,logl(np.clip(p,0,1)),-np.inf)
            acc=lp>thr
            cu[acc]=p[acc]; cl[acc]=lp[acc]; na+=acc.sum(); nr+=K-acc.sum()
        nc=K*walks
        scale*=math.exp((na/(na+nr)-0.5)/n/0.5)
    ws=np.empty(K)
    for j in range(K):
        L=sl[j]; Lp=sl[j-1] if j else lprev
        lv=logvol+math.log((N-j)/(N+1.))
        ws[j]=lae(L,Lp)+lv+math.log(0.5/(N-j))
    m=ws.max(); logz=lae(logz,m+math.log(np.exp(ws-m).sum()))
    logvol+=math.log((N-K+1)/(N+1.)); lprev=thr
    u[order[:K]]=cu; l[order[:K]]=cl; it+=K; ncall+=nc; rounds+=1
    if ncall>=ncall_last+upd:
        A=bound_root(u); ncall_last=ncall
# add live
order=np.argsort(l); sl=l[order]
for i in range(N):
    lv=logvol+math.log(1.-(i+1.)/(N+1.))
    dlv_prev=(logvol if i==0 else logvol+math.log(1.-i/(N+1.)))
    dv=math.log(math.exp(dlv_prev)-math.exp(lv))
    logz=lae(logz, lae(sl[i], lprev)+dv+math.log(0.5)); lprev=sl[i]
print(json.dumps(dict(variant=variant,N=N,K=K,walks=walks,seed=seed,logz=round(logz,3),it=it,it1=it1,ncall=int(ncall),rounds=rounds,scale=round(scale,4),wall=round(time.time()-t0,1))),flush=True)
