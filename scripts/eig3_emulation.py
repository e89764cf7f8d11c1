"""Numpy emulation of the k_eig3 round structure (svd.cu): double-buffered G, Hermitian block updates, pivot blocks\nrotated twice with identical values, early stop -- checks that every entry is written each round and Q G Q^H stays\nconsistent.  Study script (CPU), not part of the package."""
import numpy as np
BSZ=16; N=32; NP=16
def circle_pair(r,k,n):
    n1=n-1
    if k==0: a,b=r,n1
    else: a,b=(r+k)%n1,(r-k+n1)%n1
    return (a,b) if a<b else (b,a)
def run(G, cross, max_inner, tol2=1e-28):
    sched={}; slot={}
    for r in range(N-1):
        for k in range(NP):
            a,b=circle_pair(r,k,N); sched[(r,k)]=(a,b); slot[(r,a)]=k; slot[(r,b)]=k
    def pair(r,k):
        return (k, BSZ+((k+r)&15)) if cross else sched[(r,k)]
    def slot_of(r,i):
        if cross: return i if i<BSZ else ((i-BSZ-r)&15)
        return slot[(r,i)]
    nrounds = BSZ if cross else N-1
    total=max_inner*nrounds
    g=[G.copy(), np.full((N,N),np.nan+0j)]; q=np.eye(N,dtype=complex)
    rot=[[None]*NP,[None]*NP]
    def make_rot(gb,rr,j):
        p,qq=pair(rr%nrounds,j)
        a=gb[p,p].real; b=gb[qq,qq].real; gpq=gb[p,qq]; mag2=abs(gpq)**2
        c=1.0;o=0j;act=0
        if a>0 and b>0 and mag2>tol2*a*b:
            dd=0.5*(b-a); hh=dd*dd+mag2; den=abs(dd)+np.sqrt(hh); R=1/np.sqrt(den*den+mag2); s=np.copysign(R,dd)
            c=den*R; o=-s*gpq; act=1
        rot[rr&1][j]=(c,o,act); return act
    def block(gi,go,pk,qk,pl,ql,ck,ok,cl,ol,diag,act):
        g00,g01,g10,g11=gi[pk,pl],gi[pk,ql],gi[qk,pl],gi[qk,ql]
        a00=g00*ck+ok*g10; a01=g01*ck+ok*g11; a10=g10*ck-np.conj(ok)*g00; a11=g11*ck-np.conj(ok)*g01
        b00=a00*cl+a01*np.conj(ol); b01=a01*cl-ol*a00; b10=a10*cl+a11*np.conj(ol); b11=a11*cl-ol*a10
        if diag:
            b00=b00.real+0j; b11=b11.real+0j
            if act: b01=0j; b10=0j
            go[pk,pl]=b00; go[pk,ql]=b01; go[qk,pl]=b10; go[qk,ql]=b11
        else:
            go[pk,pl]=b00; go[pk,ql]=b01; go[qk,pl]=b10; go[qk,ql]=b11
            go[pl,pk]=np.conj(b00); go[ql,pk]=np.conj(b01); go[pl,qk]=np.conj(b10); go[ql,qk]=np.conj(b11)
    sweep_any=0
    for j in range(NP): sweep_any|=make_rot(g[0],0,j)
    stop=False; rr=0; nrot=0
    while rr<total:
        r=rr%nrounds; cur=rr&1; gi=g[cur]; go=g[cur^1]; go[:]=np.nan
        have_next = rr+1<total
        # warp 0 (diag + pivot)
        for lane in range(32):
            diag=lane<NP
            if diag: k=l=lane
            else:
                p2,q2=pair((rr+1)%nrounds, lane-NP); k1=slot_of(r,p2); k2=slot_of(r,q2); k=min(k1,k2); l=max(k1,k2)
            if diag or (have_next and k!=l):
                pk,qk=pair(r,k); pl,ql=pair(r,l)
                block(gi,go,pk,qk,pl,ql,rot[cur][k][0],rot[cur][k][1],rot[cur][l][0],rot[cur][l][1],diag,rot[cur][k][2]!=0)
        if have_next:
            # check the entries the next rotations need are already written
            if (rr+1)%nrounds==0:
                if not sweep_any: stop=True
                sweep_any=0
            for j in range(NP):
                p2,q2=pair((rr+1)%nrounds,j)
                assert not np.isnan(go[p2,p2]) and not np.isnan(go[q2,q2]) and not np.isnan(go[p2,q2]), (rr,j)
                sweep_any|=make_rot(go,rr+1,j)
        # B2
        for k in range(NP):
            for l in range(k+1,NP):
                pk,qk=pair(r,k); pl,ql=pair(r,l)
                block(gi,go,pk,qk,pl,ql,rot[cur][k][0],rot[cur][k][1],rot[cur][l][0],rot[cur][l][1],False,False)
        # Q
        for k in range(NP):
            c,o,act=rot[cur][k]
            if act:
                nrot+=1
                pk,qk=pair(r,k); x=q[pk].copy(); y=q[qk].copy()
                q[pk]=x*c+o*y; q[qk]=y*c-np.conj(o)*x
        assert not np.isnan(go).any(), rr
        rr+=1
        if stop: break
    return g[rr&1], q, rr, nrot

rng=np.random.default_rng(0)
W=rng.standard_normal((32,200))+1j*rng.standard_normal((32,200))
G=W@W.conj().T
for cross,inner in [(False,6),(False,2)]:
    gf,q,rr,nrot=run(G,cross,inner)
    D=q@G@q.conj().T
    off=np.abs(D-np.diag(np.diag(D))).max()/np.abs(np.diag(D)).max()
    print('cross',cross,'rounds',rr,'rot',nrot,'offdiag',off,'unitary',np.abs(q@q.conj().T-np.eye(32)).max(),'g consistency',np.abs(gf-D).max()/np.abs(D).max(), 'herm', np.abs(gf-gf.conj().T).max())
# cross-only on a matrix whose 16x16 diagonal blocks are already diagonal
Wa=np.linalg.qr((rng.standard_normal((200,16))+1j*rng.standard_normal((200,16))))[0].T*rng.random(16)[:,None]
Wb=np.linalg.qr((rng.standard_normal((200,16))+1j*rng.standard_normal((200,16))))[0].T*rng.random(16)[:,None]
W=np.vstack([Wa,Wb]); G=W@W.conj().T
gf,q,rr,nrot=run(G,True,4)
D=q@G@q.conj().T
print('cross rounds',rr,'rot',nrot,'offdiag',np.abs(D-np.diag(np.diag(D))).max()/np.abs(np.diag(D)).max(),'g consistency',np.abs(gf-D).max())
