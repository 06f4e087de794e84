"""Multi-GPU plumbing of the prover: one process per GPU (torch.distributed), independent proofs per
rank ("replicas only"); only 32-byte commitments and timings cross ranks.
reference: OpenVM scales by proving independent segments on separate devices
(benchmarks/synthetic/README.md:23); there is no data-path collective in a single proof yet."""
import numpy as np
import torch
import torch.distributed as dist


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
