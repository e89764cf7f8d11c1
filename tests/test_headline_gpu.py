"""Parity at the BASELINE.json sizes themselves (-m gpu): the SVD shapes of the 20-qubit headline
configuration, a 20-qubit / chi=512 run against the oracle, the 12-qubit batch configuration through the
CUDA-graph lanes, and `prepare_mps` fed MPSs in arbitrary gauges.  The oracle legs are sized to finish in
tens of seconds on the box's host cores."""
import numpy as np
import pytest

from oracle import qmprs_oracle as O
from qmprs_b200 import host
from tests.test_kernels_gpu import check_svd, crand

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("m,n", [(1024, 1024), (512, 2048), (2048, 512), (256, 1024)])
def test_svd_headline_gate_split_shapes(K, m, n):
    """theta matrices of the three centre gate-splits of a 20-qubit / chi=512 layer (mps.py:968-971)."""
    rng = np.random.default_rng(m + 3 * n)
    check_svd(K, crand(rng, m, n), tol=5e-10)


def test_svd_first_tt_split_2pow19(K):
    """First split of the 20-qubit TT-SVD (mps.py:242): 2^19 x 2, skinny single-block path."""
    psi = O.random_state(20, 11)
    a = psi.reshape(2 ** 19, 2)
    U, S, Vh = K.svd(K.from_host(a))
    s = K.to_host(S)
    sref = np.linalg.svd(a, compute_uv=False)
    assert np.abs(s - sref).max() <= 1e-12 * sref[0] and np.all(np.abs(s - sref) <= 1e-10 * sref)
    u, vh = K.to_host(U), K.to_host(Vh)
    assert np.abs((u * s[None, :]) @ vh - a).max() <= 1e-12
    assert np.abs(np.conj(u).T @ u - np.eye(2)).max() <= 1e-12


def test_headline_20q_chi512_two_layers_one_sweep(K):
    """BASELINE config 3 shapes (20 qubits, chi=512: TT splits up to 1024 x 1024, truncation 1024 -> 512,
    gate-split SVDs 1024 x 1024 / 512 x 2048 / 2048 x 512) with 2 layers + 1 sweep against the canonical
    oracle.  north_star bars: same bonds / layer count / gate count, spectra 1e-10 relative, extracted
    unitaries 1e-8, fidelity 1e-6."""
    n, chi, L, S = 20, 512, 2, 1
    psi = O.random_state(n, 0)
    rec_o, rec_d = {}, {}
    ro = O.prepare(psi, n, chi, L, S, gauge="canonical", record=rec_o)
    rd = host.prepare(K, psi, n, chi, L, S, record=rec_d, fused=False)
    assert rd["n_layers"] == ro["n_layers"] == L
    assert host.bond_dims(rd["mps"]) == O.bond_dims(ro["mps"])
    assert max(host.bond_dims(rd["mps"])) == chi
    for key in ("tt_svd", "truncate"):                       # A1 and A2 spectra (also config 4's kind of split)
        assert len(rec_o[key]) == len(rec_d[key])
        for so, sd in zip(rec_o[key], rec_d[key]):
            sd = K.to_host(sd)
            assert sd.shape == so.shape and np.abs(sd - so).max() <= 1e-10 * so[0]
    for so, sd in zip(rec_o["gate_split"][0], rec_d["gate_split"][0]):      # A6, first layer
        sd = K.to_host(sd)
        assert sd.shape == so.shape and np.abs(sd - so).max() <= 1e-10 * so[0]
    assert np.abs(np.array(rd["overlaps"]) - np.array(ro["overlaps"])).max() <= 1e-10
    # gates BEFORE the sweep are compared through a second run without sweeps (the sweep rewrites them)
    rd0 = host.prepare(K, psi, n, chi, L, 0, mps=rd["mps"], mps_preconditioned=True)
    ro0 = O.prepare_mps(ro["mps"], L, 0, gauge="canonical")
    flat = O.flatten_layers(ro0["layers"])
    g = rd0["gates"].reshape(-1, 16)
    assert len(flat) == g.shape[0] == L * n
    for idx, (_, _, _, _, G) in enumerate(flat):
        assert np.abs(g[idx][: G.size] - G.reshape(-1)).max() <= 1e-8
    fo = O.circuit_fidelity(psi, ro["layers"], n)
    assert abs(rd["fidelity"] - fo) <= 1e-6
    # default (fused one-pass) build gives the same MPS state and the same fidelity
    rf = host.prepare(K, psi, n, chi, L, S)
    assert host.bond_dims(rf["mps"]) == O.bond_dims(ro["mps"])
    assert abs(rf["fidelity"] - fo) <= 1e-6


def test_config2_16q_chi256_15_layers(K):
    """BASELINE config 2 at its own size (16 qubits, chi=256, 15 layers) against the canonical oracle: layer count,
    gate structure, overlaps after every layer, the first four extracted layers at the 1e-8 bar, and -- continuing
    both sides with 5 optimisation sweeps (the persistent multi-CTA sweep kernel; 50 sweeps would be ~70 s of oracle
    time) -- the fidelity at the 1e-6 bar."""
    n, chi, L, S = 16, 256, 15, 5
    psi = O.random_state(n, 0)
    ro = O.prepare(psi, n, chi, L, 0, gauge="canonical")
    rd = host.prepare(K, psi, n, chi, L, 0, fused=False)
    assert rd["n_layers"] == ro["n_layers"] == L
    assert host.bond_dims(rd["mps"]) == O.bond_dims(ro["mps"])
    dov = np.abs(np.array(rd["overlaps"]) - np.array(ro["overlaps"]))
    assert dov[:8].max() <= 1e-10 and dov.max() <= 1e-6     # measured: 1e-15 .. 1e-12 up to layer 8, 9e-9 at layer 15
    flat = O.flatten_layers(ro["layers"])
    g = rd["gates"].reshape(-1, 16)
    kinds = [k for kl in rd["kinds"] for k in kl]
    assert len(flat) == g.shape[0] == L * n
    for idx, (li, _, _, _, G) in enumerate(flat):
        assert kinds[idx] == (2 if G.shape[0] == 4 else 1)
        if L - 1 - li < 4:                                   # application order is the reverse of extraction order
            assert np.abs(g[idx][: G.size] - G.reshape(-1)).max() <= 1e-8
    assert abs(rd["fidelity"] - O.circuit_fidelity(psi, ro["layers"], n)) <= 1e-6
    for _ in range(S):
        O.sweep(ro["target"], ro["layers"], n, "canonical")
    fo = O.circuit_fidelity(psi, ro["layers"], n)
    rs = host.prepare(K, psi, n, chi, L, S, fused=False)     # the reference's two-pass build, as the oracle's target
    assert abs(rs["fidelity"] - fo) <= 1e-6, (rs["fidelity"], fo)
    # the default one-pass build skips from_dense's intermediate 'rsum2' cut (weight <= 1e-10, i.e. amplitudes ~1e-5:
    # at chi = 256 = full rank it removes the smallest Schmidt value of the centre bond), so its target differs from
    # the oracle's by that truncation noise: measured 2e-6 in fidelity after 5 sweeps
    rf = host.prepare(K, psi, n, chi, L, S)
    assert abs(rf["fidelity"] - fo) <= 1e-5, (rf["fidelity"], fo)


def test_config4_tt_svd_spectra_21q(K):
    """BASELINE config 4 (24-qubit TT-SVD) one size down, spectra of EVERY split against numpy's LAPACK SVD
    (the 24-qubit case itself is covered through properties in test_pipeline_gpu and scripts/tt_svd_c4.py)."""
    n = 21
    psi = O.random_state(n, 9)
    sp_d = []
    A = host.from_dense(K, K.from_host(psi), n, sp_d)
    sp_o = []
    Ao = O.from_dense(psi, n, sp_o)
    assert host.bond_dims(A) == O.bond_dims(Ao)
    for so, sd in zip(sp_o, sp_d):
        sd = K.to_host(sd)
        assert sd.shape == so.shape and np.abs(sd - so).max() <= 1e-10 * so[0]


def test_config5_state_through_graph_lanes(K):
    """BASELINE config 5 per-state configuration (12 qubits, chi=64, 10 layers, 20 sweeps) through
    prepare_state_batch on the captured-graph lanes (what bench.py --gpus N times) vs the oracle."""
    from qmprs_b200 import batch as qb
    n, chi, L, S = 12, 64, 10, 20
    states = np.stack([O.random_state(n, s) for s in range(4)])
    recs = qb.prepare_state_batch(states, chi, L, S, kernels=K, graph_lanes=2)
    assert len(recs) == 4
    for s, r in enumerate(recs):
        ro = O.prepare(states[s], n, chi, L, S, gauge="canonical")
        assert r["n_layers"] == ro["n_layers"] == L
        assert [host.blocks_from_kinds(k) for k in r["kinds"]] == [[(a, b) for a, b, _ in lay] for lay in ro["layers"]]
        fo = O.circuit_fidelity(states[s], ro["layers"], n)
        assert abs(r["fidelity"] - fo) <= 1e-6
        layers_d = []
        for li in range(L):
            gl = [r["gates"][li, i, : (16 if r["kinds"][li][i] == 2 else 4)].reshape((4, 4) if r["kinds"][li][i] == 2 else (2, 2))
                  for i in range(n)]
            layers_d.append([(a, b, gl[a:b + 1]) for a, b in host.blocks_from_kinds(r["kinds"][li])])
        assert np.abs(O.circuit_state(layers_d, n) - O.circuit_state(ro["layers"], n)).max() <= 1e-6


def _scrambled(A, seed, scale=1.0):
    rng = np.random.default_rng(seed)
    B = [a.copy() for a in A]
    for i in range(len(B) - 1):
        r = B[i].shape[2]
        X = np.eye(r) + 0.3 * (rng.standard_normal((r, r)) + 1j * rng.standard_normal((r, r)))
        B[i] = np.einsum("lpr,rs->lps", B[i], X)
        B[i + 1] = np.einsum("sr,rpk->spk", np.linalg.inv(X), B[i + 1])
    B[len(B) // 2] = B[len(B) // 2] * scale
    return B


@pytest.mark.parametrize("kind", ["left", "mixed", "arrays", "scaled"])
def test_prepare_mps_any_gauge(K, kind):
    """sequential.py:360-376 for an MPS that is NOT the right-canonical one this package builds:
    left-canonical, mixed (after apply_unitary_layer), raw arrays in a scrambled gauge, and not normalised.
    Overlaps (early break, hence layer count and depth), gate structure and gates must equal the oracle's."""
    from qmprs.primitives import MPS
    from qmprs.synthesis.mps_encoding import Sequential
    from qmprs_b200 import GateListCircuit
    from qmprs_b200.primitives.mps import DeviceMPS
    n, chi, L, S = 8, 8, 4, 2
    psi = O.random_state(n, 55)
    mps = MPS(statevector=psi, bond_dimension=chi)
    if kind == "left":
        mps.canonicalize("left")
        assert mps.canonical_form == "left"
    elif kind == "mixed":
        layer = mps.generate_bond_D_unitary_layer()
        mps.apply_unitary_layer(layer, inverse=True)
        assert mps.canonical_form == "unknown"
    else:
        arrays = _scrambled(list(mps.mps.arrays), 6, scale=1.0 if kind == "arrays" else 3.0)
        mps = MPS(mps=DeviceMPS.from_arrays(arrays, K))
        assert mps.canonical_form == "unknown"
    A = [np.asarray(a) for a in mps.mps.arrays]
    ro = O.prepare_mps(A, L, S, gauge="canonical")
    enc = Sequential(GateListCircuit)
    circ = enc.prepare_mps(mps, num_layers=L, num_sweeps=S)
    res = enc.last_result
    assert res["n_layers"] == ro["n_layers"]
    assert np.abs(np.array(res["overlaps"]) - np.array(ro["overlaps"])).max() <= 1e-10
    assert [host.blocks_from_kinds(k) for k in res["kinds"]] == [[(a, b) for a, b, _ in lay] for lay in ro["layers"]]
    sv = circ.get_statevector()
    ref_sv = O.simulate_emitted(O.emit_gates(ro["layers"], n), n)
    assert np.abs(sv - ref_sv).max() <= 1e-6
    # and without sweeps the gates themselves (north_star 1e-8 for the first four layers)
    enc.prepare_mps(mps, num_layers=L, num_sweeps=0)
    ro0 = O.prepare_mps(A, L, 0, gauge="canonical")
    g = enc.last_result["gates"].reshape(-1, 16)
    for idx, (_, _, _, _, G) in enumerate(O.flatten_layers(ro0["layers"])):
        assert np.abs(g[idx][: G.size] - G.reshape(-1)).max() <= 1e-8
