"""One complex128 SVD of the gate_split shape (GPU box; used under ncu)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qmprs_b200.kernels import get_kernels
m = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
K = get_kernels("cuda:0")
rng = np.random.default_rng(0)
a = rng.random((m, n)) + 1j * rng.random((m, n))
A = K.from_host(a)
K.svd(K.from_host(a[:64, :64]))
torch.cuda.synchronize(); t0 = time.perf_counter()
U, S, Vh = K.svd(A)
torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"svd {m}x{n}: {1e3*(t1-t0):.2f} ms, sweeps {K.svd_sweeps}")
