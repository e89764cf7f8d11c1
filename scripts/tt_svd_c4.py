"""BASELINE config 4: 24-qubit statevector (256 MB) -> TT-SVD with truncation to chi=1024
(one Schmidt-form pass, host.from_dense_truncated).  Size-independent check: the truncated state's
error equals the discarded weight 1 - |psi_chi|^2.  Prints per-split timings."""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qmprs_b200 import host
from qmprs_b200.kernels import get_kernels, CUTOFF, MODE_REL

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
# per-split timing (same code as host.from_dense_truncated)
N, chi = a.n, a.chi
A = [None] * N; Tm = psi.reshape(-1, 1); r = 1
t_all = T()
for i in range(N - 1, 0, -1):
    t0 = T(); s0 = K.svd_sweeps
    U, S, Vh = K.svd(Tm.reshape(2 ** i, 2 * r)); k = S.shape[0]
    rank, _ = K.trim(S, k, CUTOFF, MODE_REL, chi); n = K.read_int(rank)
    A[i] = K.scale_copy(Vh[:n]).reshape(n, 2, r)
    Tm = K.scale_copy(U[:, :n], S, None, mode=2, half_power=False); r_old = r; r = n
    t1 = T()
    print(f"split i={i:2d}  matrix {2**i:8d} x {2*r_old:5d} -> rank {n:5d}  {1e3*(t1-t0):9.2f} ms  sweeps {K.svd_sweeps-s0}")
A[0] = Tm.reshape(1, 2, r)
t_end = T()
print(f"TOTAL tt_svd(truncated) {t_end - t_all:.3f} s  bonds {host.bond_dims(A)}")
d = host.to_dense(K, A)
err = float(torch.linalg.vector_norm(d - psi).item()) ** 2
nrm = float(torch.linalg.vector_norm(d).item()) ** 2
print(f"truncation error |psi_chi - psi|^2 = {err:.6e}   1 - |psi_chi|^2 = {1-nrm:.6e}")
