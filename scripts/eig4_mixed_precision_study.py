"""Numpy study of the mixed-precision inner eigen-solve (eig4_run in svd.cu): Jacobi recurrence in FP32 on a scaled copy\nof the Gram matrix, diagonal tracked in FP64, rotations made exactly unitary in FP64; inside a block one-sided Jacobi\nSVD it needs the same number of outer sweeps as the FP64 inner solver (random / graded / rank-deficient inputs).\nStudy script (CPU), not part of the package."""
import numpy as np, sys
BSZ=16; N=32; NP=16
def circle_pair(r,k,n):
    n1=n-1
    if k==0: a,b=r,n1
    else: a,b=(r+k)%n1,(r-k+n1)%n1
    return (a,b) if a<b else (b,a)
SCHED=[[circle_pair(r,k,N) for k in range(NP)] for r in range(N-1)]
def pairs(cross,r):
    return [(k,BSZ+((k+r)&15)) for k in range(NP)] if cross else SCHED[r]

def inner_fp64(G,cross,max_inner,tol2):
    g=G.copy(); q=np.eye(N,dtype=complex); nr=BSZ if cross else N-1; anyrot=False
    for sw in range(max_inner):
        sany=False
        for r in range(nr):
            R=np.eye(N,dtype=complex)
            for (p,qq) in pairs(cross,r):
                a=g[p,p].real;b=g[qq,qq].real;gpq=g[p,qq];mag2=abs(gpq)**2
                if a>0 and b>0 and mag2>tol2*a*b:
                    dd=0.5*(b-a);hh=dd*dd+mag2;den=abs(dd)+np.sqrt(hh);Rr=1/np.sqrt(den*den+mag2);s=np.copysign(Rr,dd)
                    c=den*Rr;o=-s*gpq
                    R[p,p]=c;R[p,qq]=o;R[qq,p]=-np.conj(o);R[qq,qq]=c;sany=True
            if sany or True:
                g=R@g@R.conj().T; q=R@q
        anyrot|=sany
        if not sany: break
    return q,anyrot

def inner_mixed(G,cross,max_inner,tol2):
    f32=np.float32
    dmax=np.max(np.diag(G).real); sc=1.0/dmax
    gf=(G*sc).astype(np.complex64); hd=0.5*sc*np.diag(G).real.copy()
    q=np.eye(N,dtype=complex); nr=BSZ if cross else N-1; anyrot=False
    for sw in range(max_inner):
        sany=False
        for r in range(nr):
            R32=np.eye(N,dtype=np.complex64); R64=np.eye(N,dtype=complex)
            for (p,qq) in pairs(cross,r):
                ha,hb=hd[p],hd[qq]; gpq=gf[p,qq]
                if ha>0 and hb>0:
                    ra=f32(1)/np.sqrt(f32(2)*f32(ha)); rb=f32(1)/np.sqrt(f32(2)*f32(hb))
                    cx=f32(gpq.real)*ra*rb; cy=f32(gpq.imag)*ra*rb
                    if cx*cx+cy*cy>f32(tol2):
                        ddf=f32(hb-ha)
                        m=max(abs(ddf),abs(gpq.real),abs(gpq.imag))
                        ex=int(np.floor(np.log2(m))) if m>0 else -126
                        s2=f32(2.0**(-ex))
                        ddf=ddf*s2; gx=f32(gpq.real)*s2; gy=f32(gpq.imag)*s2
                        mag2=gx*gx+gy*gy; hh=ddf*ddf+mag2; den=abs(ddf)+np.sqrt(hh); inv=np.copysign(f32(1)/den,ddf)
                        tx=gx*inv; ty=gy*inv
                        c32=f32(1)/np.sqrt(f32(1)+tx*tx+ty*ty); o32=np.complex64(complex(-tx*c32,-ty*c32))
                        T2=0.5*float(tx*f32(gpq.real)+ty*f32(gpq.imag))
                        hd[p]=ha-T2; hd[qq]=hb+T2
                        R32[p,p]=c32;R32[p,qq]=o32;R32[qq,p]=-np.conj(o32);R32[qq,qq]=c32
                        txd,tyd=float(tx),float(ty); x=1+txd*txd+tyd*tyd; c=1/np.sqrt(x); o=complex(-txd*c,-tyd*c)
                        R64[p,p]=c;R64[p,qq]=o;R64[qq,p]=-np.conj(o);R64[qq,qq]=c; sany=True
            gf=(R32@gf@R32.conj().T).astype(np.complex64)
            for (p,qq) in pairs(cross,r):
                pass
            q=R64@q
        anyrot|=sany
        if not sany: break
    return q,anyrot

def block_jacobi(A,inner,tol=1e-14,max_sweeps=30):
    W=A.copy(); nv=W.shape[0]; nb=nv//BSZ; tol2=tol*tol
    for sweep in range(max_sweeps):
        notconv=0; mx=0.0
        for r in range(nb-1):
            for k in range(nb//2):
                bi,bj=circle_pair(r,k,nb)
                rows=list(range(bi*BSZ,bi*BSZ+BSZ))+list(range(bj*BSZ,bj*BSZ+BSZ))
                Wp=W[rows]; G=Wp@Wp.conj().T
                d=np.diag(G).real; rel=np.abs(G)**2/np.outer(d,d); np.fill_diagonal(rel,0)
                if rel.max()>tol2:
                    notconv+=1; mx=max(mx,rel.max())
                    intra=max(rel[:16,:16].max(),rel[16:,16:].max()); crossm=rel[:16,16:].max()
                    cross = sweep>0 and intra<=crossm
                    q,anyrot=inner(G,cross,2 if sweep==0 else 1,tol2)
                    if anyrot: W[rows]=q@Wp
        if notconv==0 or mx<=1e-18: return sweep+1,W
    return max_sweeps,W

rng=np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 0)
for kind in ['random','graded','rankdef']:
    nv,ln=128,160
    A=rng.standard_normal((nv,ln))+1j*rng.standard_normal((nv,ln))
    if kind=='graded':
        u,_,vh=np.linalg.svd(A,full_matrices=False); A=(u*np.logspace(0,-7,nv))@vh
    if kind=='rankdef':
        A=(rng.standard_normal((nv,40))+1j*rng.standard_normal((nv,40)))@(rng.standard_normal((40,ln))+1j*rng.standard_normal((40,ln)))
    sref=np.linalg.svd(A,compute_uv=False)
    for name,inner in [('fp64',inner_fp64),('mixed',inner_mixed)]:
        sw,W=block_jacobi(A,inner)
        s=np.sort(np.linalg.norm(W,axis=1))[::-1]
        big=sref>1e-9*sref[0]
        print(kind,name,'sweeps',sw,'max rel err (kept)',np.max(np.abs(s[big]-sref[big])/sref[big]),'abs',np.max(np.abs(s-sref))/sref[0])
