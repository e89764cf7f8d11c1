"""BASELINE config 4: 24-qubit statevector (256 MB) -> exact TT-SVD -> truncation to chi=1024.
Size-independent checks: untruncated round trip, truncation error == discarded Schmidt weight."""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qmprs_b200 import host
from qmprs_b200.kernels import get_kernels

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=24); ap.add_argument("--chi", type=int, default=1024)
a = ap.parse_args()
K = get_kernels("cuda:0")
rng = np.random.default_rng(0)
v = rng.random(2 ** a.n) + 1j * rng.random(2 ** a.n); v /= np.linalg.norm(v)
psi = K.from_host(v)
def T():
    torch.cuda.synchronize(); return time.perf_counter()
host.build_mps(K, K.from_host(v[:256] / np.linalg.norm(v[:256])), 8, 16)
K.prof_begin()
t0 = T(); A = host.from_dense(K, psi, a.n); t1 = T()
print(f"from_dense {t1-t0:.3f}s bonds {host.bond_dims(A)} svd_sweeps {K.svd_sweeps}")
d = host.to_dense(K, A); t2 = T()
err = float(torch.linalg.vector_norm(d - psi).item())
print(f"to_dense {t2-t1:.3f}s  round-trip |to_dense(A)-psi| = {err:.3e}")
spec = []
At = host.canonicalize_truncate(K, A, a.chi, spec); t3 = T()
print(f"canonicalize_truncate {t3-t2:.3f}s bonds {host.bond_dims(At)}")
dt = host.to_dense(K, At)
err_t = float(torch.linalg.vector_norm(dt - psi).item()) ** 2
nrm = float(torch.linalg.vector_norm(dt).item()) ** 2
print(f"truncation error |psi_chi - psi|^2 = {err_t:.6e}   1 - |psi_chi|^2 = {1-nrm:.6e}")
p = K.prof_end()
for kname, (ms, cnt, work) in sorted(p.items(), key=lambda kv: -kv[1][0])[:6]:
    print(f"  {kname:12s} {ms:10.2f} ms {cnt:8d} launches  rate {work/max(ms,1e-9)/1e9:.2f} T/s")
print(f"TOTAL tt_svd+truncate {t1-t0 + t3-t2:.3f} s")
