"""Numpy study (CPU, no GPU code): sweeps and rounds of the one-sided block Jacobi SVD against the block size.

Question for the next round: k_round's round is Gram (DMMA) -> 32 x 32 eigen-solve -> update (DMMA) for pairs of
16-row blocks; the DMMA phases run at the ZGEMM kernel's per-SM rate, the eigen-solve (~19 us of a ~47 us round) and
the synchronisation (~8 us) do not shrink.  Would 32-row blocks (64 x 64 eigen-problems) pay?

Every pair is diagonalised exactly here (numpy eigh), round-robin ordering, uniform random complex matrices with
positive parts (the bench distribution), convergence = largest |cos| between rows of a pair below 1e-9 at the start
of a sweep.  Result (this container):

  rows  block  sweeps  rounds/sweep  total rounds
  256     8      11        31           341
  256    16       9        15           135
  256    32       8         7            56
  256    64       5         3            15
  512     8      14        63           882
  512    16      11        31           341      <- the device today (11 sweeps measured at 512 rows)
  512    32       9        15           135
  512    64       8         7            56

DMMA work per sweep does not depend on the block size (n^2/(2 b^2) pairs x (2b)^2 x len), so 32-row blocks save 2 of
11 sweeps of DMMA time (-18 %) and 60 % of the rounds, i.e. of the per-round synchronisation; the eigen-solve chain
per SVD stays about the same if a 64 x 64 solve costs ~2.5x a 32 x 32 one (135 x 50 us against 341 x 19 us).  Estimate
for a 512-row SVD: 16.4 -> ~13.7 ms (-16 %): worth it only together with an eigen-solve spread over the pair's CTAs.

usage: python scripts/svd_blocksize_study.py [rows]
"""
import numpy as np, sys, time
def rr_rounds(n):
    idx=list(range(n)); out=[]
    for _ in range(n-1):
        out.append([(min(idx[i],idx[n-1-i]),max(idx[i],idx[n-1-i])) for i in range(n//2)])
        idx=[idx[0]]+[idx[-1]]+idx[1:-1]
    return out
def block_jacobi(W,bs,tol=1e-9,maxs=30):
    W=W.copy(); n=W.shape[0]; nb=n//bs; rounds=rr_rounds(nb); hist=[]
    for s in range(maxs):
        mx=0
        for rnd in rounds:
            for p,q in rnd:
                idx=np.r_[p*bs:(p+1)*bs,q*bs:(q+1)*bs]
                X=W[idx]; G=X@X.conj().T
                d=np.sqrt(np.abs(np.diag(G)))+1e-300
                C=np.abs(G)/np.outer(d,d); np.fill_diagonal(C,0); mx=max(mx,C.max())
                w,Q=np.linalg.eigh(G)
                W[idx]=Q[:,::-1].conj().T@X
        hist.append(mx)
        if mx<tol: break
    return s+1,hist
rng=np.random.default_rng(0)
n=int(sys.argv[1]) if len(sys.argv)>1 else 256
v=rng.random(n*n)+1j*rng.random(n*n); A=v.reshape(n,n)/np.linalg.norm(v)
for bs in (8,16,32,64):
    t=time.time(); s,h=block_jacobi(A,bs); print(f"n={n} block {bs}: sweeps {s} rounds/sweep {n//bs-1} total rounds {s*(n//bs-1)}", ["%.0e"%x for x in h], f"{time.time()-t:.1f}s")
