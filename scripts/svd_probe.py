"""Complex128 SVDs of the gate_split shapes (GPU box; also used under ncu).
usage: svd_probe.py m n [m n ...]   (QM_EIG_OLD=1 selects the previous eigen-solve kernel)"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qmprs_b200.kernels import get_kernels
args = [int(x) for x in sys.argv[1:]] or [1024, 1024]
K = get_kernels("cuda:0")
rng = np.random.default_rng(0)
K.svd(K.from_host(rng.random((64, 64)) + 0j))
for m, n in zip(args[0::2], args[1::2]):
    a = rng.random((m, n)) + 1j * rng.random((m, n))
    A = K.from_host(a)
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        U, S, Vh = K.svd(A, _plain=True, backmult=bool(int(os.environ.get("QM_PROBE_BACKMULT", "0"))))
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    s = K.to_host(S)
    sref = np.linalg.svd(a, compute_uv=False)
    u, vh = K.to_host(U), K.to_host(Vh)
    k = s.shape[0]
    rec = np.abs((u[:, :k] * s[None, :]) @ vh[:k] - a).max() / sref[0]
    orth = max(np.abs(u[:, :k].conj().T @ u[:, :k] - np.eye(k)).max(), np.abs(vh[:k] @ vh[:k].conj().T - np.eye(k)).max())
    print(f"svd {m}x{n}: {1e3*best:.2f} ms, sweeps {K.svd_sweeps}, max|s-sref|/s0 {np.abs(s-sref).max()/sref[0]:.2e}, "
          f"|USV-A|/s0 {rec:.1e}, orth {orth:.1e}", flush=True)
