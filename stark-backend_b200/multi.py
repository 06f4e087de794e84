"""Multi-GPU plumbing of the prover: one process per GPU (torch.distributed), independent proofs per
rank ("replicas only"); only 32-byte commitments and timings cross ranks.
reference: OpenVM scales by proving independent segments on separate devices
(benchmarks/synthetic/README.md:23); there is no data-path collective in a single proof yet."""
import numpy as np
import torch
import torch.distributed as dist


def bind_to_gpu_numa_node(local_rank):
    """Pins the calling process to the CPUs that are local to GPU `local_rank` (NVML: nvmlDeviceGetCpuAffinity), so that
    pinned host buffers allocated afterwards are first-touched on the GPU's own NUMA node and the prover thread runs next to
    it.  With one process per GPU this removes the cross-socket hop from every trace transport (8 ranks x 1 GiB per step
    otherwise share the inter-socket links).  Returns the CPU list, or None when NVML / the affinity call is unavailable
    (reference: per-device pinned staging, cuda-common/src/pinned.rs)."""
    import os

    try:
        import pynvml

        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = local_rank
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if local_rank < len(ids) and ids[local_rank].isdigit():
                idx = int(ids[local_rank])
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


def assign_proofs(n_proofs, world, rank):
    """Round-robin assignment of independent proofs (segments / AIR groups) to ranks."""
    return [i for i in range(n_proofs) if i % world == rank]


def all_gather_commitments(root_words, device=None):
    """root_words: uint32[8] commitment of this rank -> list of every rank's commitment, in rank order."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [np.asarray(root_words, dtype=np.uint32)]
    t = torch.from_numpy(np.ascontiguousarray(root_words, dtype=np.uint32).view(np.int32)).clone()
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [o.cpu().numpy().view(np.uint32) for o in out]


def max_over_ranks(value, device=None):
    """Device-timed durations are reported as the maximum over ranks."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64)
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ---- one commitment sharded over the ranks (commit row) ------------------------------------------------
# The stacked matrix is split by columns: RS encoding is per column, so every rank encodes its own slice with no
# communication.  Leaf hashing needs whole rows, so ONE exchange follows: rank r receives, for all W columns, the
# 2^k strided row segments {q + t S : q in [r S/G, (r+1) S/G), t < 2^k} (S = query stride).  Laid out as
# [column][t][q'] that is exactly a codeword of height B H / G with query stride S / G, so the ordinary Merkle commit
# of the shard yields the global tree's node (level log2(S/G), index r); the G sub-roots are all-gathered (32 bytes
# each) and the top log2 G levels are compressed redundantly on every rank.  The root equals the single-GPU root.
# reference: MerkleTree::new / query layout, crates/stark-backend/src/prover/stacked_pcs.rs:413-485.


def column_slice(width, world, rank):
    """Contiguous stacked-column range [c0, c1) owned by `rank` (ragged when world does not divide width)."""
    return width * rank // world, width * (rank + 1) // world


def pack_codeword_slice(cw, log_rpq, world):
    """cw: (Wl, N) tensor, the rank's columns of the codeword (column-major rows).  Returns the send buffer
    (world, Wl, 2^k, S/world): for every destination rank the row segments of its queries."""
    wl, n = cw.shape
    s = n >> log_rpq
    assert s % world == 0 and s >= world, "query stride must be a multiple of the number of ranks"
    return cw.view(wl, 1 << log_rpq, world, s // world).permute(2, 0, 1, 3).contiguous()


def exchange_rows(send, width, world, rank):
    """all-to-all of the packed slices.  Returns the (width, N / world) row shard: all columns (in global order) of the
    rows of this rank's queries."""
    per_dest = send.shape[2] * send.shape[3]
    in_split = [send.shape[1] * per_dest] * world
    out_split = [(column_slice(width, world, r)[1] - column_slice(width, world, r)[0]) * per_dest for r in range(world)]
    recv = send.new_empty(sum(out_split))
    if world == 1:
        recv.copy_(send.reshape(-1))
    else:
        dist.all_to_all_single(recv, send.reshape(-1), output_split_sizes=out_split, input_split_sizes=in_split)
    return recv.view(width, per_dest)


class DeviceCommitBackend:
    """The three compute steps of the sharded commit on a B200Device (C-ABI primitives)."""

    def __init__(self, dev):
        self.dev = dev

    def stream_scope(self):
        """torch ops, NCCL / symmetric-memory barriers and the library's kernels all run on the context's stream inside this
        scope, so the steps of a sharded commit are ordered on the device and need no host synchronisation in between."""
        return torch.cuda.stream(self.dev.torch_stream())

    def rs_encode(self, trace_slice, height, wl, l_skip, log_blowup):
        from .backend import DeviceMatrix

        out = self.dev.rs_encode(DeviceMatrix(trace_slice, height, wl), l_skip, log_blowup)
        return out.buffer.view(wl, height << log_blowup)

    def merkle_layers(self, shard, log_rpq):
        """shard: (W, rows) tensor.  Returns the concatenated digest layers (device tensor) of its tree."""
        from .backend import DeviceMatrix

        w, rows = shard.shape
        return self.dev.merkle_tree(DeviceMatrix(shard.reshape(-1), rows, w), log_rpq)

    def compress_level(self, level):
        """level: (n, 8) digests -> (n / 2, 8): adjacent pairs compressed in one launch."""
        return self.dev.poseidon2_compress(level.reshape(-1).contiguous()).view(-1, 8)


class PeerExchange:
    """Row exchange of the sharded commit over NVLink peer memory instead of a collective: every rank owns a shard buffer
    in torch symmetric memory (mapped into all ranks), and one kernel of the library (`swirl_scatter_rows_to_peers`)
    stores this rank's codeword columns straight into the peers' buffers.  Barriers on the symmetric-memory signal pads
    fence the buffers before and after."""

    def __init__(self, dev, width, rows, world, rank):
        import ctypes as C

        self.dev, self.width, self.rows, self.world, self.rank = dev, width, rows, world, rank
        n = width * (rows // world)
        if world > 1:
            import torch.distributed._symmetric_memory as symm_mem

            self.buf = symm_mem.empty(n, dtype=torch.int32, device=dev.torch_device)
            self.hdl = symm_mem.rendezvous(self.buf, dist.group.WORLD)
            ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        else:
            self.buf, self.hdl = torch.empty(n, dtype=torch.int32, device=dev.torch_device), None
            ptrs = [self.buf.data_ptr()]
        self.ptrs = (C.c_void_p * world)(*ptrs)

    def exchange(self, cw, col0, log_rpq):
        """cw: (Wl, rows) tensor of this rank's codeword columns.  Returns the (width, rows / world) shard."""
        from .lib import check

        if self.hdl is not None:
            self.hdl.barrier()  # every rank is done reading its shard of the previous commitment
        self.dev._sync_torch()
        check(self.dev.lib.swirl_scatter_rows_to_peers(self.dev.ctx, cw.data_ptr(), self.rows, cw.shape[0], col0, log_rpq, self.world,
                                                       self.ptrs))
        if self.hdl is not None:
            if torch.cuda.current_stream(self.dev.torch_device).cuda_stream != self.dev.stream_ptr():
                self.dev.synchronize()  # the barrier below is enqueued on torch's stream, the scatter ran on the library's
            self.hdl.barrier()  # all peers have finished writing into this rank's shard
        return self.buf.view(self.width, self.rows // self.world)


def sharded_commit(backend, trace_slice, height, width, l_skip, log_blowup, log_rpq, world, rank, peer_exchange=None):
    """Commit to a height x width matrix (already stacked: height = 2^(l_skip + n_stack)) whose columns
    [column_slice(width, world, rank)) are in `trace_slice` (flat column-major tensor on the backend's device).
    Returns dict(root (8 words, numpy), shard (W, N/world) rows of this rank's queries, layers (local digest layers),
    sub_roots (world, 8))."""
    import contextlib

    scope = backend.stream_scope() if hasattr(backend, "stream_scope") else contextlib.nullcontext()
    with scope:
        return _sharded_commit(backend, trace_slice, height, width, l_skip, log_blowup, log_rpq, world, rank, peer_exchange)


def _sharded_commit(backend, trace_slice, height, width, l_skip, log_blowup, log_rpq, world, rank, peer_exchange):
    c0, c1 = column_slice(width, world, rank)
    cw = backend.rs_encode(trace_slice, height, c1 - c0, l_skip, log_blowup)
    if peer_exchange is not None:  # one kernel storing into the peers' shard buffers over NVLink
        shard = peer_exchange.exchange(cw, c0, log_rpq)
    else:                          # pack + all_to_all_single (also the CPU / gloo path of the tests)
        send = pack_codeword_slice(cw, log_rpq, world)
        shard = exchange_rows(send, width, world, rank)
    layers = backend.merkle_layers(shard, log_rpq)
    sub_root = layers.view(-1)[-8:].clone()
    gathered = sub_root.new_empty(world * 8)
    if world > 1:
        dist.all_gather_into_tensor(gathered, sub_root)
    else:
        gathered.copy_(sub_root)
    level = gathered.view(world, 8)
    while level.shape[0] > 1:  # the top log2(world) levels, identical on every rank
        level = backend.compress_level(level)
    as_words = lambda t: t.detach().cpu().numpy().view(np.uint32).copy()
    return dict(root=as_words(level.reshape(-1)), shard=shard, layers=layers, sub_roots=as_words(gathered).reshape(world, 8))


def sharded_commit_benchmark(dev, log_rows, cols, l_skip, log_blowup, log_rpq, world, rank, reps=5, compare_single=True):
    """Times one commitment of a 2^log_rows x cols matrix sharded over the ranks (peer-memory exchange) on the device
    (max over ranks) and, on rank 0, the same commitment on one GPU; returns a dict for the bench line."""
    from .backend import DeviceMatrix, PcsParams

    h = 1 << log_rows

    def column(c):  # column c of the common matrix, identical on every rank
        g = torch.Generator(device=dev.torch_device).manual_seed(1000 + c)
        return torch.randint(0, 0x78000001, (h,), dtype=torch.int32, device=dev.torch_device, generator=g)

    c0, c1 = column_slice(cols, world, rank)
    mine = torch.cat([column(c) for c in range(c0, c1)])
    backend, px = DeviceCommitBackend(dev), PeerExchange(dev, cols, h << log_blowup, world, rank)
    stream, times, res = dev.torch_stream(), [], None
    for _ in range(reps + 1):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a.record(stream)
        res = sharded_commit(backend, mine, h, cols, l_skip, log_blowup, log_rpq, world, rank, peer_exchange=px)
        torch.cuda.synchronize()
        b.record(stream)
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    ms = max_over_ranks(min(times[1:]), dev.torch_device)
    out = {"workload": f"one commitment of 2^{log_rows} x {cols} sharded by columns over {world} GPUs", "ms": ms,
           "cells_per_s": h * cols / (ms / 1e3), "exchange": "peer-memory scatter kernel over NVLink (symmetric memory)"}
    if compare_single and rank == 0:
        full = torch.cat([column(c) for c in range(cols)])
        single = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record(stream)
            root, pcs = dev.commit(PcsParams(l_skip, log_rows - l_skip, log_blowup, log_rpq), [DeviceMatrix(full, h, cols)])
            dev.synchronize()
            b.record(stream)
            torch.cuda.synchronize()
            single.append(a.elapsed_time(b))
            pcs.free()
        out.update(single_gpu_ms=min(single), speedup=min(single) / ms, roots_equal=bool(np.array_equal(root, res["root"])))
    return out


# -------------------------------------------------------------------------------------------------------------------
# One proof over several GPUs: the common-main commitment -- the leaf hashing is the largest phase of
# every BASELINE config -- is sharded over the ranks (column-sharded RS encode, one row exchange over NVLink, local
# fused leaf hash + subtree, 32-byte sub-roots all-gathered); rank 0 then runs the sumcheck phases on its own copy of the
# traces, and when the WHIR opening samples its query indices every rank opens the queries whose rows it holds (each
# query lives on exactly one rank) and the rows / local Merkle paths are reduced to rank 0, which appends the top
# log2(world) siblings.  The transcript is rank 0's, the proof is the one a single GPU produces, byte for byte.
# Communication per proof: the row exchange (4 N (G-1)/G^2 bytes per rank), G sub-roots, the query indices (< 1 KB) and
# the opened rows (a few MB).  The sumcheck tables are not sharded (DESIGN.md section 7 explains what that would take).
# -------------------------------------------------------------------------------------------------------------------
class ShardedProver:
    def __init__(self, dev, params, world, rank, use_peer_memory=True):
        self.dev, self.params, self.world, self.rank = dev, params, world, rank
        self.backend = DeviceCommitBackend(dev)
        self.use_peer_memory = use_peer_memory and world > 1 and dev is not None
        self._px = None
        self.timings = {}

    # ---- the commitment, sharded ---------------------------------------------------------------------------------
    def _commit(self, stacked):
        P = self.params
        H, W = stacked.height, stacked.width
        if W < self.world:  # raised on every rank alike, before any collective
            raise ValueError(f"sharded commitment: the stacked matrix has {W} columns, fewer than the {self.world} ranks")
        c0, c1 = column_slice(W, self.world, self.rank)
        full = _tensor_view(self.dev, stacked.stacked_ptr(), H * W)
        mine = full[c0 * H:c1 * H]
        rows = H << P.log_blowup
        if self.use_peer_memory:
            if self._px is None or (self._px.width, self._px.rows) != (W, rows):
                self._px = PeerExchange(self.dev, W, rows, self.world, self.rank)
        res = sharded_commit(self.backend, mine, H, W, P.l_skip, P.log_blowup, P.whir.k, self.world, self.rank,
                             peer_exchange=self._px if self.use_peer_memory else None)
        # every level of the top log2(world) levels, for the upper part of the Merkle paths
        levels = [torch.from_numpy(res["sub_roots"].view(np.int32)).to(self.dev.torch_device).view(self.world, 8)]
        while levels[-1].shape[0] > 1:
            levels.append(self.backend.compress_level(levels[-1]))
        res["top_levels"] = [l.cpu().numpy().view(np.uint32).reshape(-1, 8) for l in levels]
        return res

    # ---- openings: every rank answers for the queries it owns ----------------------------------------------------------
    def _open_local(self, res, indices, stacked):
        """rows (nq, 2^k, W) and local paths (nq, local_depth, 8) as int32 device tensors, zero for queries of other ranks."""
        P, dev = self.params, self.dev
        k = P.whir.k
        W = stacked.width
        S = (stacked.height << P.log_blowup) >> k
        s_local = S // self.world
        depth_local = s_local.bit_length() - 1
        nq = len(indices)
        rows = torch.zeros(nq * (W << k), dtype=torch.int32, device=dev.torch_device)
        paths = torch.zeros(nq * max(depth_local, 1) * 8, dtype=torch.int32, device=dev.torch_device)
        mine = [j for j, i in enumerate(indices) if i // s_local == self.rank]
        if mine:
            local = [indices[j] % s_local for j in mine]
            shard = res["shard"]
            r = dev.matrix_open_rows(shard.data_ptr(), shard.shape[1], W, s_local, k, local)  # (n, 2^k, W) numpy
            sel = torch.tensor(mine, device=dev.torch_device)
            rows.view(nq, -1)[sel] = torch.from_numpy(r.reshape(len(mine), -1).view(np.int32)).to(dev.torch_device)
            if depth_local:
                p = dev.merkle_query_proofs(res["layers"].data_ptr(), s_local, local)
                paths.view(nq, -1)[sel] = torch.from_numpy(p.reshape(len(mine), -1).view(np.int32)).to(dev.torch_device)
        return rows, paths, depth_local

    def _exchange_openings(self, res, stacked, indices):
        """Collective: rank 0 passes the indices, the others pass None.  Returns (rows, paths) on rank 0."""
        dev = self.dev
        n = torch.tensor([len(indices) if indices is not None else 0], dtype=torch.int64, device=dev.torch_device)
        if self.world > 1:
            dist.broadcast(n, 0)
        idx = torch.tensor(indices if indices is not None else [0] * int(n.item()), dtype=torch.int64, device=dev.torch_device)
        if self.world > 1:
            dist.broadcast(idx, 0)
        indices = [int(x) for x in idx.cpu().tolist()]
        rows, paths, depth_local = self._open_local(res, indices, stacked)
        if self.world > 1:  # every query has exactly one owner: a sum over ranks places each row (values < p < 2^31)
            dist.reduce(rows, 0)
            dist.reduce(paths, 0)
        if self.rank != 0:
            return None
        nq, W, k = len(indices), stacked.width, self.params.whir.k
        s_local = ((stacked.height << self.params.log_blowup) >> k) // self.world
        top = res["top_levels"]
        full_paths = np.zeros((nq, depth_local + len(top) - 1, 8), dtype=np.uint32)
        full_paths[:, :depth_local] = paths.cpu().numpy().view(np.uint32).reshape(nq, max(depth_local, 1), 8)[:, :depth_local]
        for j, i in enumerate(indices):
            node = i // s_local
            for lvl in range(len(top) - 1):
                full_paths[j, depth_local + lvl] = top[lvl][node ^ 1]
                node >>= 1
        return rows.cpu().numpy().view(np.uint32).reshape(nq, 1 << k, W), full_paths

    # ---- the proof -----------------------------------------------------------------------------------------------------
    def prove(self, vk_pre_hash, per_air_pk, per_trace):
        """Every rank passes the same arguments (the traces are resident on every rank; a rank reads only its column slice
        of the stacked matrix for the commitment, rank 0 reads everything for the sumchecks).  Returns the Proof on rank 0,
        None elsewhere.  `self.timings`: milliseconds of the sharded commitment (max over ranks) and, on rank 0, the rest."""
        from .prover import Coordinator

        dev, P = self.dev, self.params
        per_trace = sorted(per_trace, key=lambda t: (-t[1].common_main.height(), t[0]))
        stream = dev.torch_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        stacked = dev.stack(P.pcs(), [t[1].common_main for t in per_trace])
        res = self._commit(stacked)
        e1.record(stream)
        dev.synchronize()
        self.timings["sharded_commit_ms"] = max_over_ranks(e0.elapsed_time(e1), dev.torch_device)
        if self.rank != 0:
            self._exchange_openings(res, stacked, None)  # serve rank 0's single request
            stacked.free()
            return None
        stacked.attach_external(res["root"], lambda idx: self._exchange_openings(res, stacked, idx))
        import time

        t0 = time.perf_counter()
        proof = Coordinator(dev, P).prove(vk_pre_hash, per_air_pk, per_trace, precommitted=(res["root"], stacked))
        dev.synchronize()
        self.timings["rank0_rest_ms"] = 1e3 * (time.perf_counter() - t0)
        err = getattr(stacked, "_callback_error", None)
        if err is not None:
            raise err
        return proof


def _tensor_view(dev, ptr, n_words):
    """int32 torch view of `n_words` device words at `ptr` (no copy; the owner must outlive the view)."""
    class _Iface:
        pass

    o = _Iface()
    o.__cuda_array_interface__ = {"shape": (int(n_words),), "typestr": "<i4", "data": (int(ptr), False), "version": 3}
    return torch.as_tensor(o, device=dev.torch_device)
