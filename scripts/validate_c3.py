"""One-off validation of the headline configuration (20 qubits, chi=512, 15 layers, 50 sweeps):
GPU path vs the canonical oracle on the same state.  The oracle takes ~10-15 minutes of host time."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import qmprs_oracle as O
from qmprs_b200 import host
from qmprs_b200.kernels import get_kernels

n, chi, L, S = 20, 512, 15, int(sys.argv[1]) if len(sys.argv) > 1 else 50
psi = O.random_state(n, 0)
K = get_kernels("cuda:0")
t0 = time.perf_counter(); rd = host.prepare(K, psi, n, chi, L, S); K.synchronize(); t1 = time.perf_counter()
print(f"gpu: {t1-t0:.2f}s layers {rd['n_layers']} fidelity {rd['fidelity']:.12f} bonds {host.bond_dims(rd['mps'])}", flush=True)
t0 = time.perf_counter(); ro = O.prepare(psi, n, chi, L, S, gauge="canonical"); t1 = time.perf_counter()
fo = O.circuit_fidelity(psi, ro["layers"], n)
print(f"oracle: {t1-t0:.1f}s layers {ro['n_layers']} fidelity {fo:.12f} bonds {O.bond_dims(ro['mps'])}")
flat = O.flatten_layers(ro["layers"])
g = rd["gates"].reshape(-1, 16)
kinds = [k for kl in rd["kinds"] for k in kl]
worst = {}
for idx, (li, _, _, site, G) in enumerate(flat):
    worst[li] = max(worst.get(li, 0.0), float(np.abs(g[idx][:G.size] - G.reshape(-1)).max()))
same_struct = len(flat) == len(kinds) and all(kinds[i] == (2 if flat[i][4].shape[0] == 4 else 1) for i in range(len(flat)))
out = {"config": {"n": n, "chi": chi, "layers": L, "sweeps": S, "seed": 0}, "gpu_fidelity": rd["fidelity"],
       "oracle_fidelity": fo, "abs_fidelity_diff": abs(rd["fidelity"] - fo), "same_layer_count": rd["n_layers"] == ro["n_layers"],
       "same_gate_structure": bool(same_struct), "gate_count": len(flat),
       "max_gate_diff_per_layer_application_order": [worst[k] for k in sorted(worst)],
       "bonds_equal": host.bond_dims(rd["mps"]) == O.bond_dims(ro["mps"])}
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/validate_c3.json", "w"), indent=1)
