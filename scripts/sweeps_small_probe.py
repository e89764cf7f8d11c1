"""Time qm_sweeps_small alone (one CTA per state): us per gate-step per CTA wave.
usage: python scripts/sweeps_small_probe.py [n_qubits=12] [layers=10] [sweeps=5] [batch=592]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from qmprs_b200.kernels import get_kernels

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 12
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 10
sweeps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
batch = int(sys.argv[4]) if len(sys.argv) > 4 else 592
K = get_kernels()
dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(1)
t = torch.randn(batch, 1 << nq, dtype=torch.complex128, generator=g)
t = (t / t.norm(dim=1, keepdim=True)).to(dev)
sites, kinds = [], []
for _ in range(layers):                       # a staircase layer: two-qubit gates on (i, i+1), last site first
    for i in range(nq - 2, -1, -1):
        sites.append(i); kinds.append(2)
M = len(sites)
rng = np.random.default_rng(2)                # generic start: Haar-like two-qubit gates (identities would make every
z = rng.standard_normal((batch * M, 4, 4)) + 1j * rng.standard_normal((batch * M, 4, 4))    # environment rank one)
eye = torch.from_numpy(np.linalg.qr(z)[0].reshape(batch * M, 16).copy()).to(dev).contiguous()
ov = torch.zeros(batch, 2, dtype=torch.float64, device=dev)
for rep in range(3):
    gates = eye.clone()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    K.sweeps_small(t, nq, gates, sites, kinds, sweeps, batch=batch, overlaps=ov)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
waves = -(-batch // 148)
print(f"nq={nq} gates={M} sweeps={sweeps} batch={batch} mma={os.environ.get('QM_SMALL_MMA','1')}: {ms:.2f} ms, "
      f"{1e3 * ms / (waves * sweeps * M):.2f} us per gate-step per wave, mean |overlap| {ov.norm(dim=1).mean().item():.6f}")
