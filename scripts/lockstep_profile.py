"""Per-class kernel time of ONE lock-step group of the batch path (W states through host.prepare_layers_lockstep,
uncaptured, every launch bracketed by CUDA events: qm_prof_begin/end).
usage: python scripts/lockstep_profile.py [n_qubits=12] [chi=64] [layers=10] [width=16]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from qmprs_b200 import host
from qmprs_b200.kernels import CudaKernels

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 12
chi = int(sys.argv[2]) if len(sys.argv) > 2 else 64
L = int(sys.argv[3]) if len(sys.argv) > 3 else 10
W = int(sys.argv[4]) if len(sys.argv) > 4 else 16
K = CudaKernels("cuda:0")
rng = np.random.default_rng(3)
psi = rng.standard_normal((W, 2 ** nq)) + 1j * rng.standard_normal((W, 2 ** nq))
psi /= np.linalg.norm(psi, axis=1, keepdims=True)
flags = torch.zeros(W + 1, dtype=torch.int32, device="cuda:0")
for rep in range(2):
    work = torch.from_numpy(psi).to("cuda:0")
    K.begin_static()
    if rep == 1:
        K.prof_begin()
    n0 = K.launch_count()
    gates, kinds, A = host.prepare_layers_lockstep(K, work, nq, chi, L, 1 - 1e-6, flags)
    tgt = host.to_dense_batch(K, A)
    n1 = K.launch_count()
    if rep == 1:
        prof = K.prof_end()
    K.end_static()
torch.cuda.synchronize()
print(f"{nq} q chi={chi} {L} layers, {W} states in lock step: {n1 - n0} launches, flags {flags.tolist()}")
tot = sum(v[0] for v in prof.values())
for k, (ms, cnt, work) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
    if cnt:
        print(f"  {k:12s} {ms:9.3f} ms  {cnt:5d} launches  avg {1e3 * ms / cnt:8.1f} us  ({100 * ms / tot:4.1f} %)")
print(f"  total {tot:.2f} ms = {tot / W:.2f} ms per state when a group runs alone")
