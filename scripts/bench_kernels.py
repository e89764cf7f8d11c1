"""Kernel micro-benchmarks on the GPU box, CUDA events, L2 flushed between repetitions:
  * ZGEMM (this library vs cuBLAS through torch.matmul) at the contraction shapes of SURVEY section 8 A6;
  * SVD (qm_svd vs cuSOLVER through torch.linalg.svd, drivers gesvdj and gesvd) at the gate-split shapes of the
    16- and 20-qubit configurations -- the library bar for the Jacobi SVD.
usage: python scripts/bench_kernels.py [out.json] [--no-gemm] [--no-svd]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qmprs_b200.kernels import get_kernels

K = get_kernels("cuda:0")
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


out = {"zgemm": [], "svd": []}
argv = [a for a in sys.argv[1:] if not a.startswith("--")]
for (m, n, k, ta) in [] if "--no-gemm" in sys.argv else [(1024, 1024, 1024, 0), (512, 2048, 1024, 0), (2048, 512, 1024, 0), (1024, 1024, 2048, 0),
                      (2048, 2048, 2048, 0), (4096, 4096, 4096, 0), (8192, 2048, 2048, 0), (2048, 2048, 8192, 1),
                      (256, 256, 256, 0), (512, 512, 512, 0), (512, 512, 256, 0), (256, 512, 256, 0), (128, 128, 256, 0),
                      (512, 1024, 256, 0), (256, 256, 512, 0)]:
    A = torch.randn((k, m) if ta else (m, k), dtype=torch.complex128, device=dev)
    B = torch.randn(k, n, dtype=torch.complex128, device=dev)
    C = torch.empty(m, n, dtype=torch.complex128, device=dev)
    t_own = timeit(lambda: K.gemm(A, B, out=C, transA=bool(ta)))
    Aop = A.conj().T if ta else A
    t_lib = timeit(lambda: torch.matmul(Aop, B, out=C))
    fl = 8.0 * m * n * k
    row = {"m": m, "n": n, "k": k, "transA": ta, "own_ms": round(t_own, 4), "own_tflops": round(fl / t_own / 1e9, 2),
           "cublas_ms": round(t_lib, 4), "cublas_tflops": round(fl / t_lib / 1e9, 2)}
    print(row, flush=True)
    out["zgemm"].append(row)
import numpy as np
rng = np.random.default_rng(0)
for (m, n) in [] if "--no-svd" in sys.argv else [(64, 64), (128, 512), (256, 256), (256, 1024), (512, 512), (512, 2048), (2048, 512), (1024, 1024)]:
    a = rng.random((m, n)) + 1j * rng.random((m, n))
    A = K.from_host(a)
    s0 = K.svd_sweeps
    t_own = timeit(lambda: K.svd(A, _plain=True), reps=3)
    sweeps = (K.svd_sweeps - s0) / 4
    row = {"m": m, "n": n, "own_ms": round(t_own, 3), "own_sweeps": sweeps}
    for drv in ("gesvdj", "gesvd"):
        if drv == "gesvd" and m * n > 512 * 2048:
            continue                                     # QR-iteration driver: seconds at 1024^2, skipped
        try:
            row[f"cusolver_{drv}_ms"] = round(timeit(lambda: torch.linalg.svd(A, full_matrices=False, driver=drv), reps=2), 3)
        except Exception as ex:
            row[f"cusolver_{drv}_ms"] = repr(ex)[:80]
    sref = np.linalg.svd(a, compute_uv=False)
    row["own_max_rel_err"] = float(np.max(np.abs(K.to_host(K.svd(A, _plain=True)[1]) - sref) / sref))
    sj = torch.linalg.svd(A, full_matrices=False, driver="gesvdj")[1].cpu().numpy()
    row["gesvdj_max_rel_err"] = float(np.max(np.abs(sj - sref) / sref))
    print(row, flush=True)
    out["svd"].append(row)
if argv:
    json.dump(out, open(argv[0], "w"), indent=1)
