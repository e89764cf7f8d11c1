"""CUDA-graph execution of ``prepare_state`` for many small states.

For small registers (BASELINE config 5: 12 qubits, matrices <= 64x64) a state is ~10^4 tiny,
strictly dependent kernels, so launch latency and the shape read-backs dominate.  The launch
sequence only depends on data through (i) the ranks kept by the 1e-10 cutoffs, (ii) the block
structure of each layer, (iii) the early-break test, (iv) the Jacobi sweep counts.  For generic
states all four are known in advance (exact ranks, one block per layer, no early break), so the
pipeline is run SPECULATIVELY with static shapes and a fixed sweep budget, captured once into
a CUDA graph per lane, and replayed per state; every assumption is validated on the device
(``qm_expect_*``, ``qm_svd_static``) and a state whose flag comes back set is simply re-run
through the eager path.  Several lanes (graph instance + private buffers + stream) run
concurrently so that the one-CTA kernels of different states share the 148 SMs.  Lanes are streams:
they only overlap when each has its own hardware work queue (CUDA_DEVICE_MAX_CONNECTIONS, set to 32 in
``qmprs_b200/__init__.py`` unless the user chose a value).
"""
from __future__ import annotations

import os

import numpy as np
import torch

from qmprs_b200 import host
from qmprs_b200.kernels import CudaKernels, get_kernels


class _Lane:
    def __init__(self, device, n, chi, L, S, threshold, sample_state, split="svd", defer_sweeps=False):
        """``defer_sweeps``: the captured graph stops after the layer extraction and leaves ``gates`` (before any
        sweep) and the dense ``target``; the sweeps of a whole batch then run in ONE launch (run_into)."""
        self.n, self.chi, self.L, self.S, self.threshold = n, chi, L, S, threshold
        self.defer = bool(defer_sweeps)
        self.K = CudaKernels(device)                      # private workspaces
        self.stream = torch.cuda.Stream(device)
        self.psi_in = torch.empty(2 ** n, dtype=torch.complex128, device=device)
        self.pin_in = torch.empty(2 ** n, dtype=torch.complex128).pin_memory()
        self.pending = None
        K = self.K
        with torch.cuda.stream(self.stream):
            # eager warm-up on this lane: lazy initialisation, attribute calls, workspace growth
            self.psi_in.copy_(torch.from_numpy(sample_state))
            host.prepare_device(K, self.psi_in, n, chi, L, 0 if self.defer else S, threshold, split=split)
        self.stream.synchronize()
        n0 = K.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self.stream):
            K.begin_static()
            work = K.scale_copy(self.psi_in.reshape(-1, 1)).reshape(-1)
            if self.defer:
                gates, kinds, _, A = host.prepare_layers_device(K, work, n, chi, L, threshold, split=split)
                ov = None
                self.target = host.to_dense(K, A)                     # sequential.py:440 (mps.mps)
            else:
                gates, kinds, ov, _, _ = host.prepare_device(K, work, n, chi, L, S, threshold, split=split)
            K.end_static()
        self.nodes = K.launch_count() - n0
        self.gates, self.kinds, self.ov, self.mismatch = gates, kinds, ov, K.mismatch
        if len(kinds) != L:
            raise RuntimeError("static capture produced an unexpected layer count")
        self.pin_gates = torch.empty(gates.shape, dtype=gates.dtype).pin_memory()
        self.pin_ov = torch.empty(2, dtype=torch.float64).pin_memory()
        self.pin_mis = torch.empty(1, dtype=torch.int32).pin_memory()
        self.done = torch.cuda.Event()

    def submit(self, state, tag):
        self.pin_in.copy_(torch.from_numpy(np.ascontiguousarray(state)))
        with torch.cuda.stream(self.stream):
            self.psi_in.copy_(self.pin_in, non_blocking=True)
            self.mismatch.zero_()
            self.graph.replay()
            self.pin_gates.copy_(self.gates, non_blocking=True)
            self.pin_ov.copy_(self.ov, non_blocking=True)
            self.pin_mis.copy_(self.mismatch, non_blocking=True)
            self.done.record(self.stream)
        self.pending = (tag, state)

    def collect(self):
        tag, state = self.pending
        self.pending = None
        self.done.synchronize()
        if int(self.pin_mis[0]) != 0:
            return tag, None, state
        N, L = self.n, self.L
        ov = self.pin_ov.numpy()
        res = {"gates": self.pin_gates.numpy().reshape(L, N, 16).copy(), "kinds": [list(k) for k in self.kinds],
               "n_layers": L, "fidelity": float(np.hypot(ov[0], ov[1])), "n_sites": N, "overlaps": None}
        return tag, res, state


class _BatchLane:
    """One captured graph that extracts the layers of ``width`` states in LOCK STEP on one stream
    (:func:`qmprs_b200.host.prepare_layers_lockstep`).  A lane executes its kernels one after the other and the
    kernels of a small register are single CTAs, so a lane with one state per graph keeps ONE SM busy -- and the
    hardware runs at most 32 lanes side by side (work queues).  Measured on config 5 (profiles/bench_r02_c5_*.json):
    cutting the kernel nodes per state from 1814 to 691 moved the throughput by 10 %, and forking a graph into
    parallel per-state branches by nothing: the bound was the per-lane chain of single-CTA SVDs (23 of the ~28 ms of
    kernel time of a state).  In lock step EVERY step is one launch for the ``width`` states (the SVD of a split with
    grid = width, the bookkeeping kernels through their ``*_batch`` entries): 46 graph nodes per state instead of 578,
    and the layer phase of a 4096-state batch is the SM time of its single-CTA SVDs."""

    def __init__(self, device, n, chi, L, threshold, sample_state, width=8):
        self.n, self.L, self.width = n, L, int(width)
        dim, M, W = 2 ** n, L * n, int(width)
        self.K = K = CudaKernels(device)
        self.stream = torch.cuda.Stream(device)
        self.psi_in = torch.empty((W, dim), dtype=torch.complex128, device=device)
        self.gates_out = torch.empty((W, M, 16), dtype=torch.complex128, device=device)
        self.target_out = torch.empty((W, dim), dtype=torch.complex128, device=device)
        self.flags = torch.zeros(W + 1, dtype=torch.int32, device=device)
        sample = torch.from_numpy(sample_state)

        def pipeline():
            K.begin_static()
            work = self.psi_in.clone()
            gates, kinds, A = host.prepare_layers_lockstep(K, work, n, chi, L, threshold, self.flags)
            if len(kinds) != L:
                raise RuntimeError("static capture produced an unexpected layer count")
            self.gates_out.copy_(gates)
            self.target_out.copy_(host.to_dense_batch(K, A))              # sequential.py:440 (mps.mps)
            K.end_static()
            return kinds

        with torch.cuda.stream(self.stream):
            for w in range(W):
                self.psi_in[w].copy_(sample)
            pipeline()                                     # uncaptured warm-up: lazy initialisation, attribute calls
        self.stream.synchronize()
        if int(self.flags.max().item()) != 0:
            raise RuntimeError("static assumptions do not hold for the sample state")
        n0 = K.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self.stream):
            self.kinds = pipeline()
        self.nodes = (K.launch_count() - n0) / W           # kernel nodes per state
        self.done = torch.cuda.Event()


class GraphedPreparer:
    """``prepare_state`` for a stream of equally sized states through captured CUDA graphs."""

    def __init__(self, n_qubits, bond_dimension, num_layers=1, num_sweeps=0, threshold=1 - 1e-6, lanes=4,
                 device=None, split="svd", width=8):
        if not isinstance(num_layers, int) or num_layers < 1:
            raise ValueError("The number of layers must be a positive integer.")
        if device is None:
            device = f"cuda:{torch.cuda.current_device()}"
        self.device = device
        self.cfg = (int(n_qubits), int(bond_dimension), int(num_layers), int(num_sweeps), float(threshold))
        rng = np.random.default_rng(12345)
        sample = rng.random(2 ** n_qubits) + 1j * rng.random(2 ** n_qubits)
        sample /= np.linalg.norm(sample)
        self.split = split
        self._sample = sample
        self.n_lanes = int(lanes)
        self.width = max(1, int(width))      # states per graph of the batch path (advanced in lock step)
        self._full = None                    # lanes whose graph is the whole pipeline (run: one state at a time)
        self._layers = None                  # lanes whose graph stops before the sweeps (run_into: batches)
        self.eager = get_kernels(device)
        # batches of registers whose sweeps fit one SM's shared memory and whose bonds fit the small-register kernels:
        # graphs of `width` states in lock step extract the layers, then ONE k_sweeps_small launch (grid = batch)
        # sweeps every state and returns its fidelity
        nq = self.cfg[0]
        self.defer = (split == "svd" and nq <= CudaKernels.SMALL_SWEEP_MAX_SITES and self.cfg[1] >= 1
                      and nq * self.cfg[2] <= CudaKernels.SMALL_SWEEP_MAX_GATES
                      and min(2 ** (nq // 2), self.cfg[1]) <= CudaKernels.FUSED_MAX_BOND)
        self.fallbacks = 0
        self.time_phases = False             # run_into: record (layers ms, sweeps ms) of the call in self.phase_ms
        self.phase_ms = None
        self.replays = 0

    @property
    def lanes(self):
        if self._full is None:
            self._full = [_Lane(self.device, *self.cfg, self._sample, split=self.split) for _ in range(self.n_lanes)]
        return self._full

    @property
    def layer_lanes(self):
        if self._layers is None:
            n, chi, L, S, thr = self.cfg
            self._layers = [_BatchLane(self.device, n, chi, L, thr, self._sample, width=self.width)
                            for _ in range(self.n_lanes)]
        return self._layers

    @property
    def nodes_per_graph(self):
        built = self._layers if self._layers is not None else self.lanes
        return built[0].nodes

    def _finish(self, lane, out):
        tag, res, state = lane.collect()
        if res is None:                                   # an assumption failed: eager path, exact semantics
            n, chi, L, S, thr = self.cfg
            res = host.prepare(self.eager, state, n, chi, L, S, thr, split=self.split)
            res.pop("mps", None)
            self.fallbacks += 1
        out[tag] = res

    def run_into(self, states, rec_dev):
        """Batch form without a host round trip per state.  ``states``: (B, 2^n) numpy array (one pinned
        host->device copy) or a tensor already on the device; ``rec_dev``: (>= B, record_len) float64 device
        tensor that receives one record per state (layout of :func:`qmprs_b200.batch.pack_record`).
        Every state is: device copy into the lane's input, graph replay, three small device copies
        (gates, overlap, validity flag) -- all on the lane's stream.  The flags are read once at the end;
        flagged states (a static assumption failed) are re-run through the eager path and their records
        overwritten.  Returns the number of eager fallbacks of this call."""
        from qmprs_b200 import batch as qb
        n, chi, L, S, thr = self.cfg
        dev = torch.device(self.device)
        B = int(states.shape[0])
        if B == 0:
            return 0
        main = torch.cuda.current_stream(dev)
        host_states = None
        stage = None                                       # staged host->device copy: (chunk of states, events)
        if hasattr(states, "data_ptr") and states.is_cuda:
            sdev = states.contiguous()
        else:
            host_states = np.asarray(states, dtype=np.complex128)      # any row stride (a rank's shard is states[r::w])
            dim = host_states.shape[1]
            # one pinned staging buffer per preparer, grown on demand (pinning 256 MB per call costs ~0.1 s)
            pin = getattr(self, "_pin_in", None)
            if pin is None or pin.shape[0] < B or pin.shape[1] != dim:
                pin = self._pin_in = torch.empty((B, dim), dtype=torch.complex128).pin_memory()
            sdev = torch.empty((B, dim), dtype=torch.complex128, device=dev)
            # the copy goes in chunks, each with its own event: the lanes start on the first chunk while the host is
            # still staging the later ones (memcpy into pinned memory + DMA of 256 MB would otherwise sit in front of
            # the whole batch)
            chunk = max(self.width, min(B, max(1, (16 << 20) // (16 * dim))))
            chunk = ((chunk + self.width - 1) // self.width) * self.width
            stage = (chunk, {})

        def staged(upto):
            """Make sure the states [0, upto) are on their way to the device; returns the event to wait for."""
            ck, evs = stage
            last = None
            c0 = len(evs) * ck
            while c0 < min(upto, B):
                c1 = min(c0 + ck, B)
                pin[c0:c1].copy_(torch.from_numpy(host_states[c0:c1]))
                sdev[c0:c1].copy_(pin[c0:c1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(main)
                evs[c0 // ck] = ev
                c0 = c1
            return evs[(min(upto, B) - 1) // ck]

        ng, nk = L * n * 32, L * n
        # constant part of the records of the static pipeline: one block per layer, all layers used
        kinds_layer = [2] * (n - 1) + [1]
        tmpl = np.concatenate([np.tile(np.array(kinds_layer, dtype=np.float64), L), [float(L)]])
        rec_dev[:B, ng:ng + nk + 1] = torch.from_numpy(tmpl).to(dev)
        flags = torch.zeros(B, dtype=torch.int32, device=dev)
        ready = torch.cuda.Event(enable_timing=bool(os.environ.get("QM_BATCH_DEBUG")) or self.time_phases)
        ready.record(main)
        if self.defer:
            lanes = self.layer_lanes
            nl, wd = len(lanes), self.width
            gates_b = torch.empty((B, L * n, 16), dtype=torch.complex128, device=dev)
            targets_b = torch.empty((B, 2 ** n), dtype=torch.complex128, device=dev)
            ngroups = (B + wd - 1) // wd
            gflags = torch.zeros((ngroups, wd + 1), dtype=torch.int32, device=dev)
            used = lanes[:min(nl, ngroups)]
            for lane in used:
                lane.stream.wait_event(ready)
            for gi, g0 in enumerate(range(0, B, wd)):
                lane = lanes[gi % nl]
                w = min(wd, B - g0)
                if stage is not None:
                    lane.stream.wait_event(staged(g0 + w))
                with torch.cuda.stream(lane.stream):
                    lane.psi_in[:w].copy_(sdev[g0:g0 + w], non_blocking=True)
                    if w < wd:                             # ragged tail: idle slots redo the last state
                        lane.psi_in[w:].copy_(sdev[g0 + w - 1].expand(wd - w, -1), non_blocking=True)
                    lane.flags.zero_()
                    lane.graph.replay()
                    gates_b[g0:g0 + w].copy_(lane.gates_out[:w], non_blocking=True)
                    targets_b[g0:g0 + w].copy_(lane.target_out[:w], non_blocking=True)
                    gflags[gi].copy_(lane.flags, non_blocking=True)
        else:
            lanes = self.lanes
            nl = len(lanes)
            used = lanes[:min(nl, B)]
            for lane in used:
                lane.stream.wait_event(ready)
            for s in range(B):
                lane = lanes[s % nl]
                if stage is not None:
                    lane.stream.wait_event(staged(s + 1))
                with torch.cuda.stream(lane.stream):
                    lane.psi_in.copy_(sdev[s], non_blocking=True)
                    lane.mismatch.zero_()
                    lane.graph.replay()
                    rec_dev[s, :ng].copy_(lane.gates.view(torch.float64).reshape(-1), non_blocking=True)
                    rec_dev[s, ng + nk + 1:ng + nk + 3].copy_(lane.ov, non_blocking=True)
                    flags[s:s + 1].copy_(lane.mismatch, non_blocking=True)
        self.replays += B
        for lane in used:
            lane.done.record(lane.stream)
            main.wait_event(lane.done)
        debug = bool(os.environ.get("QM_BATCH_DEBUG")) or self.time_phases     # phase times of this call
        if debug:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record(main)
        if self.defer:
            # every state's sweeps + fidelity in one launch, one CTA per state (sequential.py:509-541; README.md:66)
            ov = torch.empty((B, 2), dtype=torch.float64, device=dev)
            self.eager.sweeps_small(targets_b, n, gates_b.view(B * L * n, 16), list(range(n)) * L, kinds_layer * L, S,
                                    batch=B, psis=sdev, overlaps=ov)
            rec_dev[:B, :ng] = gates_b.view(torch.float64).reshape(B, ng)
            rec_dev[:B, ng + nk + 1:ng + nk + 3] = ov
        if debug:
            ev[1].record(main)
            torch.cuda.synchronize(dev)
            self.phase_ms = (ready.elapsed_time(ev[0]), ev[0].elapsed_time(ev[1]))
            if os.environ.get("QM_BATCH_DEBUG"):
                import sys
                sys.stderr.write(f"[run_into] B={B}: layers phase ends at +{self.phase_ms[0]:.1f} ms, "
                                 f"sweeps phase {self.phase_ms[1]:.1f} ms\n")
        if self.defer:
            gf = gflags.cpu().numpy()
            per_state = (gf[:, :wd] | gf[:, wd:wd + 1]).reshape(-1)[:B]      # own flag or the group's
            bad = np.nonzero(per_state)[0]
        else:
            bad = np.nonzero(flags.cpu().numpy())[0]
        for s in bad:                                      # an assumption failed: eager path, exact semantics
            st = host_states[s] if host_states is not None else sdev[s]
            res = host.prepare(self.eager, st, n, chi, L, S, thr, split=self.split)
            rec_dev[s].copy_(torch.from_numpy(qb.pack_record(res, n, L)).to(dev))
        self.fallbacks += len(bad)
        return len(bad)

    def run(self, states):
        """Round-robin the states over the lanes.  (Driving the lanes from several host threads was
        measured: no gain -- the bound was the number of hardware work queues, not the host.)"""
        states = np.asarray(states, dtype=np.complex128)
        out = [None] * len(states)
        nl = len(self.lanes)
        for s in range(len(states)):
            lane = self.lanes[s % nl]
            if lane.pending is not None:
                self._finish(lane, out)
            lane.submit(states[s], s)
            self.replays += 1
        for lane in self.lanes:
            if lane.pending is not None:
                self._finish(lane, out)
        return out
