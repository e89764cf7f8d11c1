"""Kernel micro-benchmarks on the GPU box: ZGEMM (this library vs cuBLAS through torch.matmul) at the
contraction shapes of SURVEY section 8 A6, CUDA events, L2 flushed between repetitions.
usage: python scripts/bench_kernels.py [out.json]"""
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


out = {"zgemm": []}
for (m, n, k, ta) in [(1024, 1024, 1024, 0), (512, 2048, 1024, 0), (2048, 512, 1024, 0), (1024, 1024, 2048, 0),
                      (2048, 2048, 2048, 0), (4096, 4096, 4096, 0), (8192, 2048, 2048, 0), (2048, 2048, 8192, 1),
                      (256, 256, 256, 0), (512, 512, 512, 0)]:
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
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
