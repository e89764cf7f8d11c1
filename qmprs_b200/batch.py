"""Batches of independent target states sharded over the GPUs of one node.

BASELINE.json config 5 / SURVEY.md section 8(e): state s goes to rank ``s mod world``;
there is no communication while the states are compiled (each rank runs the single-state
path on its own B200), and one final all-gather of fixed-size records
``{gates[L][N][16], kinds[L][N], n_layers, overlap}`` over NCCL (NVLink/NVSwitch) -- about
29 KB per 12-qubit / 10-layer state.  One process per GPU (torchrun); without an initialised
process group the whole batch runs on the local device.
"""
from __future__ import annotations

import numpy as np
import torch

from qmprs_b200 import host


def shard_indices(n_states: int, rank: int, world: int) -> list[int]:
    return list(range(rank, n_states, world))


def record_len(n_sites: int, num_layers: int) -> int:
    """doubles per record: gates (re, im interleaved), kinds, n_layers, overlap (re, im)."""
    return num_layers * n_sites * 32 + num_layers * n_sites + 3


def pack_record(res: dict, n_sites: int, num_layers: int) -> np.ndarray:
    """float64 vector: gates (re,im interleaved, zero padded to num_layers), kinds, n_layers,
    <psi|circuit> as (re, im) -- the fidelity is its modulus."""
    L = res["n_layers"]
    g = np.zeros((num_layers, n_sites, 16), dtype=np.complex128)
    k = np.zeros((num_layers, n_sites), dtype=np.float64)
    g[:L] = res["gates"]
    k[:L] = np.asarray(res["kinds"], dtype=np.float64)
    ov = res.get("overlap")
    if ov is None:
        ov = (float(res["fidelity"]), 0.0)
    return np.concatenate([g.view(np.float64).reshape(-1), k.reshape(-1), [float(L), float(ov[0]), float(ov[1])]])


def unpack_record(vec: np.ndarray, n_sites: int, num_layers: int) -> dict:
    return unpack_records(np.asarray(vec).reshape(1, -1), n_sites, num_layers)[0]


def unpack_records(block: np.ndarray, n_sites: int, num_layers: int) -> list:
    """All records of a (B, record_len) float64 block at once (bulk numpy: a 4096-record block is unpacked in
    ~12 ms; the per-record Python conversions cost more than the device->host copy).  The gate arrays are views
    into ``block`` (kept alive by them), not copies.  The collector is paused while the ~10^5 small objects are
    built (its generational scans of the growing lists were 40 % of the time), and the kind tables -- identical for
    almost every record of a batch of generic states -- are converted once per distinct table."""
    import gc
    block = np.ascontiguousarray(np.asarray(block, dtype=np.float64))
    B = block.shape[0]
    ng = num_layers * n_sites * 32
    nk = num_layers * n_sites
    if B == 0:
        return []
    kinds8 = block[:, ng:ng + nk].astype(np.int8)
    same = (kinds8 == kinds8[0]).all(axis=1).tolist()
    first = kinds8[0].reshape(num_layers, n_sites).tolist()
    nl = np.rint(block[:, ng + nk]).astype(np.int64).tolist()
    ov = block[:, ng + nk + 1:ng + nk + 3]
    fid = np.hypot(ov[:, 0], ov[:, 1]).tolist()
    ovl = ov.tolist()
    gates = block[:, :ng].view(np.complex128).reshape(B, num_layers, n_sites, 16)
    was_enabled = gc.isenabled()
    gc.disable()
    try:
        out = []
        for b in range(B):
            k = nl[b]
            kinds = ([row[:] for row in first[:k]] if same[b]
                     else kinds8[b].reshape(num_layers, n_sites)[:k].tolist())
            o = ovl[b]
            out.append({"gates": gates[b, :k], "kinds": kinds, "n_layers": k, "overlap": (o[0], o[1]),
                        "fidelity": fid[b]})
    finally:
        if was_enabled:
            gc.enable()
    return out


def _to_host_pinned(t):
    """Device tensor -> numpy through page-locked memory of torch's caching host allocator: the DMA runs at link
    speed (a pageable ``.cpu()`` of the 130 MB config-5 block is staged through the driver's bounce buffer and
    page-faults its fresh destination: ~40 ms against ~5), and the block is recycled by the allocator once the
    caller has dropped every record that views it -- the returned array owns a reference to the tensor."""
    if not t.is_cuda:
        return t.numpy()
    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    h.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return h.numpy()


def prepare_state_batch(states, bond_dimension: int, num_layers: int = 1, num_sweeps: int = 0,
                        threshold: float = 1 - 1e-6, kernels=None, gather: bool = True, graph_lanes: int = 0,
                        preparer=None, return_device: bool = False):
    """Compile every row of ``states`` (B x 2^n; a numpy array, or a tensor already resident on this rank's
    GPU) and return the list of B result records (on every rank when ``gather``).  ``kernels``: kernel
    handle (default: CUDA on the local rank's device).

    Data path: the rank's shard of the states goes to HBM in one copy; with a ``preparer`` (captured CUDA
    graphs, graphs.py) every state is one graph replay whose gate record is written into the rank's device
    record block -- no host round trip per state; one ``all_gather_into_tensor`` of the fixed-size record
    blocks over NCCL; one device->host copy of the gathered records.  ``return_device`` skips that last copy
    and returns the gathered (world * per_rank, record_len) float64 tensor (rank-major, see shard_indices)."""
    import torch.distributed as dist

    on_device = hasattr(states, "data_ptr")
    if not on_device:
        states = np.asarray(states, dtype=np.complex128)
    B, dim = states.shape
    n = int(round(np.log2(dim)))
    if 2 ** n != dim:
        raise ValueError("each state must have 2^n amplitudes")
    if not isinstance(num_layers, int) or num_layers < 1:
        raise ValueError("The number of layers must be a positive integer.")
    distributed = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank() if distributed else 0
    world = dist.get_world_size() if distributed else 1
    if kernels is None:
        from qmprs_b200.kernels import get_kernels
        kernels = get_kernels()
    K = kernels
    mine = shard_indices(B, rank, world)
    rl = record_len(n, num_layers)
    per_rank = (B + world - 1) // world
    if graph_lanes or preparer is not None:
        # small states: captured CUDA graphs replayed per state on several concurrent lanes (graphs.py)
        if preparer is None:
            from qmprs_b200.graphs import GraphedPreparer
            preparer = GraphedPreparer(n, bond_dimension, num_layers, num_sweeps, threshold, lanes=graph_lanes,
                                       device=str(K.device))
        local = torch.zeros((per_rank, rl), dtype=torch.float64, device=K.device)
        preparer.run_into(states[rank::world] if world > 1 else states, local)
    else:
        local_h = np.zeros((per_rank, rl), dtype=np.float64)
        for slot, s in enumerate(mine):
            res = host.prepare(K, states[s], n, bond_dimension, num_layers, num_sweeps, threshold)
            local_h[slot] = pack_record(res, n, num_layers)
        local = torch.from_numpy(local_h).to(K.device)
    if distributed and gather and world > 1:
        # the one collective of the path: all-gather of the fixed-size record blocks
        recv = torch.empty((world * per_rank, rl), dtype=torch.float64, device=K.device)
        dist.all_gather_into_tensor(recv, local)
    else:
        recv = local
    if return_device:
        return recv
    allrec = _to_host_pinned(recv)         # the one device->host copy of the path
    if not (distributed and gather) or world == 1:
        recs = unpack_records(allrec[:len(mine)], n, num_layers)
        return recs if world == 1 else dict(zip(mine, recs))
    recs = unpack_records(allrec, n, num_layers)
    out = [None] * B
    for r in range(world):
        for slot, s in enumerate(shard_indices(B, r, world)):
            out[s] = recs[r * per_rank + slot]
    return out
