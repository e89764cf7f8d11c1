"""Host-side logic (qmprs_b200/host.py) exercised on CPU with the numpy test double of the
kernel interface (tests/fake_kernels.py): launch sequencing, reshapes, index conventions,
rank plumbing, block structure, early break.  The CUDA kernels themselves are covered by
the -m gpu tests."""
import numpy as np
import pytest

from oracle import qmprs_oracle as O
from qmprs_b200 import host
from tests.fake_kernels import FakeKernels


def gate_table(res):
    return O.flatten_layers(res["layers"])


@pytest.mark.parametrize("n,chi,L,S,seed", [(4, 4, 2, 0, 0), (6, 64, 3, 2, 1), (8, 32, 5, 3, 2), (8, 4, 4, 2, 3),
                                           (10, 512, 6, 2, 4)])
def test_pipeline_matches_oracle(n, chi, L, S, seed):
    psi = O.random_state(n, seed)
    ref = O.prepare(psi, n, chi, L, S, gauge="canonical")
    # arbitrary SVD phases (as a Jacobi SVD returns) must not matter
    out = host.prepare(FakeKernels(svd_phase_seed=seed + 10), psi, n, chi, L, S)
    assert out["n_layers"] == ref["n_layers"]
    g = out["gates"].reshape(-1, 16)
    for idx, (_, _, _, site, G) in enumerate(gate_table(ref)):
        assert np.abs(g[idx][: G.size] - G.reshape(-1)).max() < 1e-7
    assert abs(out["fidelity"] - O.circuit_fidelity(psi, ref["layers"], n)) < 1e-9
    assert host.bond_dims(out["mps"]) == O.bond_dims(ref["mps"])


def test_recompute_and_stored_sweeps_agree():
    psi = O.random_state(7, 5)
    K = FakeKernels()
    A = host.build_mps(K, K.from_host(psi), 7, 16)
    gates, kinds, _ = host.disentangle(K, A, 3, 1 - 1e-6)
    target = host.to_dense(K, A)
    g1, g2 = gates.clone(), gates.clone()
    host.optimize_layers(K, target, g1, kinds, 7, 2, stored=True)
    host.optimize_layers(K, target, g2, kinds, 7, 2, stored=False)
    assert np.abs(K.to_host(g1) - K.to_host(g2)).max() < 1e-10


def test_block_structure_and_early_break():
    h = np.array([1, 1]) / np.sqrt(2)
    z = np.array([1.0, 0.0])
    bell = np.zeros(4, dtype=complex); bell[0] = bell[3] = 1 / np.sqrt(2)
    psi = np.array([1.0 + 0j])
    for v in (bell, z, h, bell, z, h):
        psi = np.kron(psi, v)
    out = host.prepare(FakeKernels(), psi, 8, 32, 5, 0)
    assert out["n_layers"] == 1                      # early break (sequential.py:390): already a product of blocks
    assert host.blocks_from_kinds(out["kinds"][0]) == [(0, 1), (2, 2), (3, 3), (4, 5), (6, 6), (7, 7)]
    assert out["fidelity"] > 1 - 1e-12


def test_zero_overlap_and_to_dense():
    K = FakeKernels()
    psi = O.random_state(6, 8)
    A = host.build_mps(K, K.from_host(psi), 6, 64)
    dense = K.to_host(host.to_dense(K, A))
    assert np.abs(dense - psi).max() < 1e-12
    assert abs(host.zero_overlap(K, A) - np.conj(psi[0])) < 1e-12


def test_forward_and_inverse_layer_are_inverse():
    K = FakeKernels()
    psi = O.random_state(6, 9)
    A = host.build_mps(K, K.from_host(psi), 6, 64)
    B = host.copy_mps(K, A)
    gates, kinds = host.chi2_layer(K, B)
    host.apply_inverse_layer(K, B, gates, kinds)
    host.apply_inverse_layer(K, B, gates, kinds, inverse=False)
    assert np.abs(K.to_host(host.to_dense(K, B)) - psi).max() < 1e-9


def test_mirror_roundtrip():
    K = FakeKernels()
    A = host.build_mps(K, K.from_host(O.random_state(5, 1)), 5, 8)
    M = host.mirror(K, host.mirror(K, A))
    for a, b in zip(A, M):
        assert a.shape == b.shape and np.abs(K.to_host(a) - K.to_host(b)).max() == 0.0


@pytest.mark.parametrize("n,chi,L,S", [(6, 64, 4, 2), (8, 4, 4, 1), (10, 512, 8, 0)])
def test_exact_split_gives_the_same_circuit(n, chi, L, S):
    """split="exact" (no SVD when re-splitting theta) changes only the gauge of the working MPS:
    gates, layer count and fidelity equal the SVD mode."""
    psi = O.random_state(n, 21)
    a = host.prepare(FakeKernels(), psi, n, chi, L, S, split="svd")
    b = host.prepare(FakeKernels(), psi, n, chi, L, S, split="exact")
    assert a["n_layers"] == b["n_layers"] and a["kinds"] == b["kinds"]
    assert np.abs(a["gates"] - b["gates"]).max() < 1e-7
    assert abs(a["fidelity"] - b["fidelity"]) < 1e-10


def _gauge_scrambled(A, seed, scale=1.0):
    """Same state (times ``scale``), random invertible matrices inserted on every bond."""
    rng = np.random.default_rng(seed)
    B = [a.copy() for a in A]
    for i in range(len(B) - 1):
        r = B[i].shape[2]
        X = np.eye(r) + 0.3 * (rng.standard_normal((r, r)) + 1j * rng.standard_normal((r, r)))
        B[i] = np.einsum("lpr,rs->lps", B[i], X)
        B[i + 1] = np.einsum("sr,rpk->spk", np.linalg.inv(X), B[i + 1])
    B[len(B) // 2] = B[len(B) // 2] * scale
    return B


@pytest.mark.parametrize("kind", ["left", "scrambled", "scaled"])
def test_prepare_mps_any_gauge_matches_reference_preconditioning(kind):
    """sequential.py:360-376: normalize + compress('right') + canonicalize('right', normalize=True) on
    whatever gauge the caller's MPS is in.  The early-break overlaps (hence layer count and depth) and the
    gates must not depend on the input gauge (a left-canonical input gave overlaps off by 1/sqrt 2)."""
    n, chi, L, S = 7, 8, 4, 1
    psi = O.random_state(n, 21)
    A0 = O.compress_right(O.from_dense(psi, n), max_bond=chi)
    if kind == "left":
        A = O.left_canon([a.copy() for a in A0])
    elif kind == "scrambled":
        A = _gauge_scrambled(A0, 3)
    else:
        A = _gauge_scrambled(A0, 4, scale=2.5)            # not normalised either
    ref = O.prepare_mps(A, L, S, gauge="canonical")
    base = O.prepare_mps(A0, L, S, gauge="canonical")
    K = FakeKernels(svd_phase_seed=5)
    out = host.prepare(K, psi, n, chi, L, S, mps=[K.from_host(a) for a in A])
    assert out["n_layers"] == ref["n_layers"] == base["n_layers"]
    assert np.abs(np.array(out["overlaps"]) - np.array(ref["overlaps"])).max() < 1e-10
    assert np.abs(np.array(out["overlaps"]) - np.array(base["overlaps"])).max() < 1e-10
    g = out["gates"].reshape(-1, 16)
    for idx, (_, _, _, site, G) in enumerate(gate_table(ref)):
        assert np.abs(g[idx][: G.size] - G.reshape(-1)).max() < 1e-7


@pytest.mark.parametrize("schedule", ["IterDiOall", "IterDiOi"])
@pytest.mark.parametrize("n,chi,L,S,seed", [(6, 64, 3, 2, 1), (8, 32, 4, 3, 2), (8, 4, 4, 2, 3)])
def test_iterative_schedules_match_oracle(schedule, n, chi, L, S, seed):
    """the schedules the reference names as future work (notebook :459, sequential.py:410,
    428-432), composed on the host from the same stages; checked against the oracle's composition of its own."""
    psi = O.random_state(n, seed)
    ref = O.prepare(psi, n, chi, L, S, gauge="canonical", schedule=schedule)
    out = host.prepare(FakeKernels(svd_phase_seed=seed + 10), psi, n, chi, L, S, schedule=schedule)
    assert out["n_layers"] == ref["n_layers"]
    g = out["gates"].reshape(-1, 16)
    for idx, (_, _, _, site, G) in enumerate(gate_table(ref)):
        assert np.abs(g[idx][: G.size] - G.reshape(-1)).max() < 1e-7
    assert abs(out["fidelity"] - O.circuit_fidelity(psi, ref["layers"], n)) < 1e-9
    for a, b in zip(out["overlaps"], ref["overlaps"]):
        assert abs(a - b) < 1e-8


def test_iterative_schedules_without_sweeps_are_the_default_schedule():
    psi = O.random_state(7, 6)
    base = host.prepare(FakeKernels(), psi, 7, 16, 4, 0)
    for schedule in ("IterDiOall", "IterDiOi"):
        out = host.prepare(FakeKernels(), psi, 7, 16, 4, 0, schedule=schedule)
        assert out["kinds"] == base["kinds"]
        assert np.array_equal(out["gates"], base["gates"])
    with pytest.raises(ValueError):
        host.prepare(FakeKernels(), psi, 7, 16, 4, 0, schedule="IterDallOi")


def test_iterative_schedule_early_break_and_blocks():
    """A product of two-qubit blocks is disentangled by the first layer whatever the schedule."""
    bell = np.zeros(4, dtype=complex); bell[0] = bell[3] = 1 / np.sqrt(2)
    psi = np.array([1.0 + 0j])
    for v in (bell, bell, bell):
        psi = np.kron(psi, v)
    for schedule in ("IterDiOall", "IterDiOi"):
        out = host.prepare(FakeKernels(), psi, 6, 8, 4, 2, schedule=schedule)
        assert out["n_layers"] == 1 and out["fidelity"] > 1 - 1e-12


@pytest.mark.parametrize("schedule", ["IterDiOall", "IterDiOi"])
def test_iterative_schedules_on_an_mps_in_any_gauge(schedule):
    """prepare_mps(schedule=...) on a left-canonical, un-normalised MPS: the pre-conditioning of sequential.py:360-376
    comes first, as in the default schedule, and the residual of Iter DiOall is rebuilt from the PRE-CONDITIONED MPS."""
    n, chi, L, S = 7, 8, 3, 2
    psi = O.random_state(n, 22)
    A0 = O.compress_right(O.from_dense(psi, n), max_bond=chi)
    A = O.left_canon([a.copy() for a in A0])
    A[2] = 1.7 * A[2]
    ref = O.prepare_mps(A, L, S, gauge="canonical", schedule=schedule)
    K = FakeKernels(svd_phase_seed=6)
    out = host.prepare(K, psi, n, chi, L, S, mps=[K.from_host(a) for a in A], schedule=schedule)
    assert out["n_layers"] == ref["n_layers"]
    assert np.abs(np.array(out["overlaps"]) - np.array(ref["overlaps"])).max() < 1e-9
    g = out["gates"].reshape(-1, 16)
    for idx, (_, _, _, site, G) in enumerate(gate_table(ref)):
        assert np.abs(g[idx][: G.size] - G.reshape(-1)).max() < 1e-7


def _edge_state(rng, n, kind):
    if kind == "random":
        v = rng.random(2 ** n) + 1j * rng.random(2 ** n)
    elif kind == "haar":
        v = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    elif kind == "product":
        v = np.array([1.0 + 0j])
        for _ in range(n):
            v = np.kron(v, rng.normal(size=2) + 1j * rng.normal(size=2))
    elif kind == "ghz":
        v = np.zeros(2 ** n, dtype=complex); v[0] = v[-1] = 1
    elif kind == "sparse":
        v = np.zeros(2 ** n, dtype=complex)
        idx = rng.choice(2 ** n, size=min(3, 2 ** n), replace=False)
        v[idx] = rng.normal(size=idx.size) + 1j * rng.normal(size=idx.size)
    elif kind == "basis":
        v = np.zeros(2 ** n, dtype=complex); v[rng.integers(2 ** n)] = 1
    else:                                                   # product of random blocks of 1-3 qubits
        v = np.array([1.0 + 0j]); m = 0
        while m < n:
            k = min(int(rng.integers(1, 4)), n - m)
            v = np.kron(v, rng.normal(size=2 ** k) + 1j * rng.normal(size=2 ** k)); m += k
    return v / np.linalg.norm(v)


def test_edge_case_states_bonds_and_schedules_follow_the_oracle():
    """Seeded sweep over the inputs the reference's tests probe one by one (test_sequential_encoding.py:91-155:
    partially entangled and GHZ-like states; test_mps.py:49-88: small registers, tight bond dimensions) and beyond:
    2..8 qubits, bond dimension 1..64, product / basis / GHZ / sparse / block-product / dense states, every schedule.
    Layer count (early break), block structure, and fidelity must equal the oracle's; the gate matrices too wherever
    the Schmidt spectrum is non-degenerate (GHZ-like and sparse states leave the singular bases free: the reference's
    own gates depend on LAPACK there)."""
    rng = np.random.default_rng(123)
    kinds_all = ["random", "haar", "product", "ghz", "sparse", "basis", "blocks"]
    for trial in range(70):
        n = int(rng.integers(2, 9)); chi = int(rng.choice([1, 2, 3, 4, 8, 64]))
        L = int(rng.integers(1, 5)); S = int(rng.integers(0, 3))
        kind = str(rng.choice(kinds_all))
        sched = str(rng.choice(["DallOall", "DallOall", "IterDiOall", "IterDiOi"]))
        psi = _edge_state(rng, n, kind)
        ref = O.prepare(psi, n, chi, L, S, gauge="canonical", schedule=sched)
        out = host.prepare(FakeKernels(svd_phase_seed=trial), psi, n, chi, L, S, schedule=sched)
        ctx = (trial, n, chi, L, S, kind, sched)
        assert out["n_layers"] == ref["n_layers"], ctx
        flat = gate_table(ref)
        kinds = [k for kl in out["kinds"] for k in kl]
        assert kinds == [2 if G.shape[0] == 4 else 1 for (_, _, _, _, G) in flat], ctx
        assert abs(out["fidelity"] - O.circuit_fidelity(psi, ref["layers"], n)) < 1e-9, ctx
        if kind not in ("ghz", "sparse"):
            g = out["gates"].reshape(-1, 16)
            for idx, (_, _, _, _, G) in enumerate(flat):
                assert np.abs(g[idx][: G.size] - G.reshape(-1)).max() < 1e-6, ctx
