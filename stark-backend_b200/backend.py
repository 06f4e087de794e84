"""Host-side mirror of the reference's prover-device interface for the commit path.

Names follow the reference (crates/stark-backend/src/prover/hal.rs:65-138,
crates/cuda-backend/src/gpu_backend.rs:59-212, cuda-backend/src/base.rs:8-12,
cuda-backend/src/stacked_pcs.rs:30-46): ``B200Device.commit(traces) -> (commitment, pcs_data)``,
``DeviceMatrix``, ``StackedPcsData`` (layout / matrix / tree), ``StackedLayout``.
PyTorch is plumbing only: it owns device buffers and the CUDA context; every computation goes
through the C ABI of libswirl_b200.so.
"""
import ctypes as C
import weakref
from dataclasses import dataclass

import numpy as np
import torch

from . import lib as _lib
from .lib import MatrixC, PcsParamsC, check

P = 0x78000001
_R = (1 << 32) % P
_RINV = pow(_R, P - 2, P)


def to_mont(x):
    """canonical -> Montgomery words (numpy, host side helper for tests / fixtures)."""
    a = np.asarray(x, dtype=np.uint64) % P
    return ((a << np.uint64(32)) % np.uint64(P)).astype(np.uint32)


def from_mont(m):
    a = np.asarray(m, dtype=np.uint64)
    return ((a * np.uint64(_RINV)) % np.uint64(P)).astype(np.uint32)


@dataclass(frozen=True)
class PcsParams:
    """The SystemParams fields the commitment depends on (reference: config.rs:52-65)."""

    l_skip: int
    n_stack: int
    log_blowup: int
    k_whir: int

    def c(self):
        return PcsParamsC(self.l_skip, self.n_stack, self.log_blowup, self.k_whir)


def _as_i32_tensor(arr):
    a = np.ascontiguousarray(arr, dtype=np.uint32)
    return torch.from_numpy(a.view(np.int32))


class DeviceMatrix:
    """Column-major base-field matrix resident in HBM (reference: DeviceMatrix, base.rs:8-12).
    ``buffer`` is a flat int32 CUDA tensor of Montgomery words, values[col*height + row]."""

    def __init__(self, buffer, height, width):
        assert buffer.is_cuda and buffer.dtype in (torch.int32, torch.uint32)
        assert buffer.numel() == height * width
        self.buffer, self._h, self._w = buffer, int(height), int(width)

    @classmethod
    def from_host(cls, values, height, width, device="cuda:0"):
        """transport_matrix_to_device (data_transporter.rs:93-106): plain H2D memcpy."""
        t = _as_i32_tensor(np.asarray(values).reshape(-1))
        return cls(t.to(device), height, width)

    def height(self):
        return self._h

    def width(self):
        return self._w

    def to_host(self):
        return self.buffer.cpu().numpy().view(np.uint32).copy()

    def ptr(self):
        return self.buffer.data_ptr()


@dataclass
class StackedSlice:
    col_idx: int
    row_idx: int
    log_height: int


@dataclass
class StackedLayout:
    """reference: prover/stacked_pcs.rs:18-42."""

    l_skip: int
    height: int
    width: int
    sorted_cols: list  # (mat_idx, col_in_mat, StackedSlice)

    @classmethod
    def new(cls, l_skip, log_stacked_height, sorted_meta):
        """StackedLayout::new (prover/stacked_pcs.rs:144-203); host-only, no GPU needed."""
        lib = _lib.load_library()
        n = len(sorted_meta)
        widths = (C.c_uint64 * max(n, 1))(*[w for w, _ in sorted_meta])
        lhs = (C.c_int32 * max(n, 1))(*[h for _, h in sorted_meta])
        ow, on = C.c_uint64(), C.c_uint64()
        check(lib.swirl_stacked_layout(l_skip, log_stacked_height, n, widths, lhs, C.byref(ow), C.byref(on), None))
        buf = (C.c_uint64 * max(5 * on.value, 1))()
        check(lib.swirl_stacked_layout(l_skip, log_stacked_height, n, widths, lhs, C.byref(ow), C.byref(on), buf))
        cols = [
            (int(buf[5 * i]), int(buf[5 * i + 1]), StackedSlice(int(buf[5 * i + 2]), int(buf[5 * i + 3]), int(buf[5 * i + 4])))
            for i in range(on.value)
        ]
        return cls(l_skip, 1 << log_stacked_height, int(ow.value), cols)


class MerkleTree:
    """View of the device-resident tree (reference: MerkleTreeGpu, merkle_tree.rs:76-88)."""

    def __init__(self, device, codeword_ptr, height, width, layers_ptr, query_stride, log_rpq):
        self._dev = device
        self.codeword_ptr, self.height, self.width = codeword_ptr, height, width
        self.layers_ptr, self._qs, self.log_rows_per_query = layers_ptr, query_stride, log_rpq

    def query_stride(self):
        return self._qs

    def rows_per_query(self):
        return 1 << self.log_rows_per_query

    def proof_depth(self):
        return self._qs.bit_length() - 1

    def digest_layers(self):
        """All layers on the host: list of (n_l, 8) uint32 arrays, layer 0 first, root last."""
        total = 2 * self._qs - 1
        flat = self._dev._d2h(self.layers_ptr, total * 8)
        out, off, n = [], 0, self._qs
        while n >= 1:
            out.append(flat[off * 8 : (off + n) * 8].reshape(n, 8))
            off += n
            n >>= 1
        return out

    def root(self):
        return self._dev._d2h(self.layers_ptr + (2 * self._qs - 2) * 32, 8)

    def backing_matrix(self):
        return self._dev._d2h(self.codeword_ptr, self.height * self.width)

    def query_merkle_proofs(self, indices):
        """batch of query_merkle_proof (stacked_pcs.rs:388-405): (num_queries, depth, 8)."""
        return self._dev.merkle_query_proofs(self.layers_ptr, self._qs, indices)

    def get_opened_rows(self, indices):
        """batch of get_opened_rows (stacked_pcs.rs:516-540): (num_queries, rows_per_query, width)."""
        return self._dev.matrix_open_rows(
            self.codeword_ptr, self.height, self.width, self._qs, self.log_rows_per_query, indices
        )


class StackedPcsData:
    """reference: StackedPcsDataGpu (cuda-backend/src/stacked_pcs.rs:30-46)."""

    def __init__(self, device, handle, params, keepalive):
        self._dev, self._h, self.params, self._keep = device, handle, params, keepalive
        device._live_pcs.add(self)  # B200Device.close() frees the handles that are still alive before the context goes
        lib = device.lib
        self.height = int(lib.swirl_pcs_stacked_height(handle))
        self.width = int(lib.swirl_pcs_stacked_width(handle))
        n = int(lib.swirl_pcs_layout(handle, None))
        buf = (C.c_uint64 * max(5 * n, 1))()
        lib.swirl_pcs_layout(handle, buf)
        cols = [
            (int(buf[5 * i]), int(buf[5 * i + 1]), StackedSlice(int(buf[5 * i + 2]), int(buf[5 * i + 3]), int(buf[5 * i + 4])))
            for i in range(n)
        ]
        self.layout = StackedLayout(params.l_skip, self.height, self.width, cols)
        self.tree = MerkleTree(
            device,
            lib.swirl_pcs_codeword(handle),
            int(lib.swirl_pcs_codeword_height(handle)),
            self.width,
            lib.swirl_pcs_layers(handle),
            int(lib.swirl_pcs_query_stride(handle)),
            params.k_whir,
        )

    def matrix(self):
        """The stacked evaluation matrix (height x width, column-major) on the host."""
        return self._dev._d2h(self._dev.lib.swirl_pcs_stacked_matrix(self._h), self.height * self.width)

    def commit(self):
        return self.external_root if getattr(self, "external_root", None) is not None else self.tree.root()

    def stacked_ptr(self):
        """Device pointer of the stacked matrix (height x width, column-major)."""
        return self._dev.lib.swirl_pcs_stacked_matrix(self._h)

    def attach_external(self, root, open_fn):
        """The commitment's codeword and tree live elsewhere (other ranks of a sharded commitment): `open_fn(indices) ->
        (rows (nq, 2^k, width) uint32, paths (nq, depth, 8) uint32)` supplies what the WHIR opening asks for."""
        dev, width, rpq = self._dev, self.width, 1 << self.params.k_whir
        depth = int(dev.lib.swirl_pcs_query_stride(self._h)).bit_length() - 1

        def thunk(_user, h_idx, nq, d_rows, d_paths):
            try:
                rows, paths = open_fn([int(h_idx[i]) for i in range(nq)])
                rows = np.ascontiguousarray(rows, dtype=np.uint32).reshape(nq, rpq, width)
                paths = np.ascontiguousarray(paths, dtype=np.uint32).reshape(nq, depth, 8)
                check(dev.lib.swirl_memcpy_h2d(dev.ctx, d_rows, rows.ctypes.data, rows.nbytes))
                check(dev.lib.swirl_memcpy_h2d(dev.ctx, d_paths, paths.ctypes.data, paths.nbytes))
                dev.synchronize()
                return 0
            except Exception as e:  # noqa: BLE001 -- an exception must not unwind through the C frames
                self._callback_error = e
                return 10001

        self._open_cb = _lib.OPEN_FN(thunk)  # kept alive as long as the handle
        root = np.ascontiguousarray(root, dtype=np.uint32)
        check(dev.lib.swirl_pcs_attach_external(self._h, root.ctypes.data, self._open_cb, None))
        self.external_root = root.copy()

    def open_rows(self, indices):
        """Opened rows of the commitment's codeword, (num_queries, rows_per_query, width): from the cached codeword or, with
        cache_rs_code_matrix = false, re-encoded by column groups (what the WHIR opening does)."""
        dev = self._dev
        idx = dev.h2d(np.asarray(indices, dtype=np.uint32))
        rpq = 1 << self.params.k_whir
        out = dev.alloc(len(indices) * rpq * self.width)
        dev._sync_torch()
        check(dev.lib.swirl_pcs_open_rows(dev.ctx, self._h, idx.data_ptr(), len(indices), out.data_ptr()))
        dev.synchronize()
        return out.cpu().numpy().view(np.uint32).reshape(len(indices), rpq, self.width)

    def free(self):
        if self._h:
            check(self._dev.lib.swirl_pcs_free(self._dev.ctx, self._h))
            self._h = None
            self._keep = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PcsTraceView:
    """A trace that lives inside a StackedPcsData produced by commit_host (the device copy made by the
    transport): same accessors as DeviceMatrix, kept alive by the PCS handle."""

    def __init__(self, pcs, ptr, height, width):
        self._pcs, self._ptr, self._h, self._w = pcs, int(ptr), int(height), int(width)

    def height(self):
        return self._h

    def width(self):
        return self._w

    def ptr(self):
        return self._ptr


class TraceTransporter:
    """Host -> device transport of trace matrices with pinned double buffering (reference: DeviceDataTransporter,
    prover/hal.rs:141-207; pinned staging in cuda-common/src/pinned.rs): `submit` starts the copy of a
    pinned host trace into one of `depth` device buffers on a dedicated copy stream and returns a ticket; `matrix`
    makes the library's stream wait for that copy and returns the DeviceMatrix.  Submitting the next proof's trace
    before proving the current one overlaps the PCIe transfer with the proof."""

    def __init__(self, dev, height, width, depth=2):
        self.dev, self.height, self.width = dev, int(height), int(width)
        self.buffers = [dev.alloc(self.height * self.width) for _ in range(depth)]
        self.stream = torch.cuda.Stream(device=dev.torch_device)
        self.next = 0
        # per buffer: event recorded on the library's stream after the last consumer of the buffer was enqueued
        self.consumed = [None] * depth
        self.outstanding = 0

    def submit(self, host_tensor):
        """Starts the copy into the next buffer.  At most `depth` tickets may be outstanding (submitted and not yet
        `retire`d); the copy waits, on the device, for the work that last read this buffer."""
        assert host_tensor.is_pinned() and host_tensor.numel() == self.height * self.width
        if self.outstanding >= len(self.buffers):
            raise RuntimeError(f"TraceTransporter: {self.outstanding} tickets outstanding with depth {len(self.buffers)}; "
                               "retire() a ticket before submitting another trace")
        slot = self.next
        buf = self.buffers[slot]
        self.next = (self.next + 1) % len(self.buffers)
        self.outstanding += 1
        ev = torch.cuda.Event()
        with torch.cuda.stream(self.stream):
            if self.consumed[slot] is not None:
                self.stream.wait_event(self.consumed[slot])  # earlier readers of this buffer on the library's stream
            buf.copy_(host_tensor, non_blocking=True)
            ev.record(self.stream)
        return buf, ev, slot

    def matrix(self, ticket):
        """The DeviceMatrix of a ticket; the library's stream waits for the copy (the host does not block).  NOTE: a
        commitment of a single full-size trace aliases this buffer as its stacked matrix (commit.cu), so the
        StackedPcsData / Proof.common_main_pcs made from it must be freed (or the proof finished) before `retire`."""
        buf, ev, _ = ticket
        self.dev.torch_stream().wait_event(ev)
        return DeviceMatrix(buf, self.height, self.width)

    def retire(self, ticket):
        """Everything that reads the ticket's buffer has been enqueued on the library's stream: the buffer may be reused by a
        later submit (which waits for that work on the device)."""
        _, _, slot = ticket
        ev = torch.cuda.Event()
        ev.record(self.dev.torch_stream())
        self.consumed[slot] = ev
        self.outstanding -= 1


class WhirConfig:
    """reference: WhirConfig / WhirRoundConfig (config.rs:172-197)."""

    def __init__(self, k, num_queries, mu_pow_bits=0, query_phase_pow_bits=0, folding_pow_bits=0):
        self.k, self.num_queries = int(k), [int(q) for q in num_queries]
        self.mu_pow_bits, self.query_phase_pow_bits, self.folding_pow_bits = mu_pow_bits, query_phase_pow_bits, folding_pow_bits

    @staticmethod
    def new(log_blowup, log_stacked_height, k, log_final_poly_len, query_phase_pow_bits, folding_pow_bits, mu_pow_bits,
            security_bits=100):
        """WhirConfig::new for the unique-decoding regime (config.rs:268-341): queries per round =
        ceil((security - query_pow) / -log2((1 + 2^-rate) / 2)), rate += k - 1 per round."""
        import math

        level = max(security_bits - query_phase_pow_bits, 0)
        rounds = -(-max(log_stacked_height - log_final_poly_len, 0) // k)
        rate, nq = log_blowup, []
        for _ in range(rounds):
            per_query = -math.log2(min(max((1.0 + 2.0 ** (-rate)) / 2.0, 5e-324), 1.0))
            nq.append(math.ceil(level / per_query))
            rate += k - 1
        return WhirConfig(k, nq, mu_pow_bits, query_phase_pow_bits, folding_pow_bits)

    def c(self):
        c = _lib.WhirConfigC()
        c.k, c.num_rounds = self.k, len(self.num_queries)
        for i, q in enumerate(self.num_queries):
            c.num_queries[i] = q
        c.mu_pow_bits, c.query_phase_pow_bits, c.folding_pow_bits = self.mu_pow_bits, self.query_phase_pow_bits, self.folding_pow_bits
        return c


class AirProvingContext:
    """One present AIR: the constraint DAG of its verifying key (SymbolicConstraintsDag,
    air_builders/symbolic/dag.rs:17-96) plus its traces (AirProvingContext, prover/types.rs:37-73).
    nodes: (n, 4) uint32 rows (op, a, b, c) in the swirl_dag_node encoding; interactions: list of
    (count_node, bus_index, [message nodes]); matrices are DeviceMatrix."""

    def __init__(self, nodes, constraint_idx, interactions, constraint_degree, need_rot, common_main, public_values=(),
                 cached_mains=(), preprocessed=None):
        self.nodes = np.ascontiguousarray(nodes, dtype=np.uint32).reshape(-1, 4)
        self.constraint_idx = np.ascontiguousarray(constraint_idx, dtype=np.uint32)
        self.interactions = [(int(c), int(b), [int(m) for m in msg]) for c, b, msg in interactions]
        self.constraint_degree, self.need_rot = int(constraint_degree), bool(need_rot)
        self.public_values = np.ascontiguousarray(public_values, dtype=np.uint32)
        self.common_main, self.cached_mains, self.preprocessed = common_main, list(cached_mains), preprocessed

    def c(self, keep):
        inter = (_lib.InteractionC * max(len(self.interactions), 1))()
        msg = []
        for i, (cnt, bus, m) in enumerate(self.interactions):
            inter[i] = _lib.InteractionC(cnt, bus, len(msg), len(m))
            msg.extend(m)
        msg = np.asarray(msg, dtype=np.uint32)
        cached = (MatrixC * max(len(self.cached_mains), 1))()
        for i, m in enumerate(self.cached_mains):
            cached[i] = MatrixC(m.ptr(), m.height(), m.width())
        prep = MatrixC(self.preprocessed.ptr(), self.preprocessed.height(), self.preprocessed.width()) if self.preprocessed else None
        keep.extend([inter, msg, cached, prep, self])
        a = _lib.AirCtxC()
        a.nodes, a.n_nodes = self.nodes.ctypes.data, len(self.nodes)
        a.constraint_idx, a.n_constraints = self.constraint_idx.ctypes.data, len(self.constraint_idx)
        a.interactions, a.n_interactions = C.addressof(inter), len(self.interactions)
        a.msg_nodes = msg.ctypes.data if len(msg) else None
        a.constraint_degree, a.need_rot = self.constraint_degree, 1 if self.need_rot else 0
        a.public_values, a.n_public_values = (self.public_values.ctypes.data if len(self.public_values) else None), len(self.public_values)
        a.common_main = MatrixC(self.common_main.ptr(), self.common_main.height(), self.common_main.width())
        a.cached_mains, a.n_cached = C.addressof(cached), len(self.cached_mains)
        a.preprocessed = C.addressof(prep) if prep is not None else None
        return a


class Transcript:
    """reference: DuplexSponge as FiatShamirTranscript (transcript/duplex_sponge.rs:16-115,
    transcript/traits.rs:11-90).  Host-resident POD state (`swirl_transcript`)."""

    def __init__(self, words18=None):
        self.lib = _lib.load_library()
        self.c = _lib.TranscriptC()
        if words18 is not None:
            self.load(words18)

    def load(self, words18):
        for i in range(16):
            self.c.state[i] = int(words18[i])
        self.c.absorb_idx, self.c.sample_idx = int(words18[16]), int(words18[17])

    def words(self):
        return np.array(list(self.c.state) + [self.c.absorb_idx, self.c.sample_idx], dtype=np.uint32)

    def observe(self, mont_words):
        a = np.ascontiguousarray(mont_words, dtype=np.uint32).reshape(-1)
        check(self.lib.swirl_transcript_observe(C.byref(self.c), a.ctypes.data, a.size))

    def sample(self, n=1):
        out = np.zeros(n, dtype=np.uint32)
        check(self.lib.swirl_transcript_sample(C.byref(self.c), out.ctypes.data, n))
        return out

    def sample_ext(self):
        return self.sample(4)

    def sample_bits(self, bits):
        v = C.c_uint32()
        check(self.lib.swirl_transcript_sample_bits(C.byref(self.c), bits, C.byref(v)))
        return int(v.value)

    def check_witness(self, bits, witness):
        ok = C.c_int()
        check(self.lib.swirl_transcript_check_witness(C.byref(self.c), bits, int(witness), C.byref(ok)))
        return bool(ok.value)


class B200Device:
    """reference: GpuDevice (cuda-backend/src/device.rs:52-110) — one CUDA device, one stream."""

    def __init__(self, device=0):
        self.lib = _lib.load_library()
        if not torch.cuda.is_available():
            raise RuntimeError("B200Device needs a CUDA device: there is no CPU fallback")
        self.device = int(device)
        self.torch_device = torch.device(f"cuda:{self.device}")
        torch.cuda.set_device(self.device)
        torch.zeros(1, device=self.torch_device)  # make sure the primary context exists
        h = C.c_void_p()
        check(self.lib.swirl_ctx_create(self.device, C.byref(h)))
        self.ctx = h
        self._live_pcs = weakref.WeakSet()

    def close(self):
        if self.ctx:
            for pcs in list(self._live_pcs):  # PCS data must not outlive the context that owns its memory
                pcs.free()
            self.lib.swirl_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- plumbing -------------------------------------------------------------------------
    def synchronize(self):
        check(self.lib.swirl_ctx_synchronize(self.ctx))

    def stream_ptr(self):
        return self.lib.swirl_ctx_stream(self.ctx)

    def torch_stream(self):
        return torch.cuda.ExternalStream(self.stream_ptr(), device=self.torch_device)

    def launch_count(self):
        return int(self.lib.swirl_ctx_launch_count(self.ctx))

    def set_cache_rs_code_matrix(self, on):
        """reference: GpuDevice::set_cache_rs_code_matrix (cuda-backend/src/device.rs:108-110)."""
        check(self.lib.swirl_ctx_set_cache_rs_code_matrix(self.ctx, 1 if on else 0))

    def set_jit(self, mode):
        """Run-time compiled per-AIR constraint kernels: 0 = interpreter only, 1 = tall traces (default), 2 = always;
        + 4 = the MLE rounds run compiled kernels as well."""
        check(self.lib.swirl_ctx_set_jit(self.ctx, int(mode)))

    def jit_stats(self):
        """{r0_compiled, r0_launches, mle_compiled, mle_launches} of this context (swirl_ctx_jit_stats)."""
        out = (C.c_uint64 * 4)()
        check(self.lib.swirl_ctx_jit_stats(self.ctx, out))
        return dict(zip(("r0_compiled", "r0_launches", "mle_compiled", "mle_launches"), (int(x) for x in out)))

    def mem_stats(self, reset_peak=False):
        """{live, peak, held, device_free} bytes of the context's scratch arena (traces handed in by the caller not counted)."""
        out = (C.c_uint64 * 4)()
        check(self.lib.swirl_ctx_mem_stats(self.ctx, 1 if reset_peak else 0, out))
        return {"live": int(out[0]), "peak": int(out[1]), "held": int(out[2]), "device_free": int(out[3])}

    def set_ntt_plan(self, max_log_radix, scratch_bytes=0):
        check(self.lib.swirl_ctx_set_ntt_plan(self.ctx, max_log_radix, scratch_bytes))

    TIMING_SLOTS = {"leaf": 0, "tree": 1, "chunk": 2, "ntt_pass": 3, "ntt_final": 4, "stack": 5, "gkr": 6, "bc_round0": 7,
                    "bc_mle": 8}

    def sync_stats(self):
        """(stream synchronisations issued by the library so far, milliseconds spent inside them)."""
        n, ms = C.c_uint64(), C.c_double()
        check(self.lib.swirl_ctx_sync_stats(self.ctx, C.byref(n), C.byref(ms)))
        return int(n.value), ms.value

    def set_round_link(self, on):
        """Sumcheck rounds through the mapped mailbox (default) or one launch + stream synchronisation per round."""
        check(self.lib.swirl_ctx_set_round_link(self.ctx, 1 if on else 0))

    def link_stats(self):
        """Round results received through the mailbox (no stream synchronisation) so far."""
        n = C.c_uint64()
        check(self.lib.swirl_ctx_link_stats(self.ctx, C.byref(n)))
        return int(n.value)

    def timing_enable(self, on=True):
        check(self.lib.swirl_ctx_timing_enable(self.ctx, 1 if on else 0))

    def timing_read(self):
        """{family: (total_ms, launches, algorithmic_bytes)} since timing_enable(True)."""
        out = {}
        for name, slot in self.TIMING_SLOTS.items():
            ms, n, b = C.c_double(), C.c_uint64(), C.c_uint64()
            check(self.lib.swirl_ctx_timing_read(self.ctx, slot, C.byref(ms), C.byref(n)))
            check(self.lib.swirl_ctx_timing_bytes(self.ctx, slot, C.byref(b)))
            out[name] = (ms.value, int(n.value), int(b.value))
        return out

    # -- host proof buffers -------------------------------------------------------------------------------------
    # Fresh multi-megabyte numpy buffers are fresh anonymous pages: the first write into each 4 KiB page faults, which costs
    # ~1 ms for the 3.4 MB WHIR proof.  Buffers handed back by Proof.release() are reused (what a pooling allocator such as
    # the reference's mimalloc does for the Rust Vecs).
    def host_buffer(self, n_words):
        pool = self.__dict__.setdefault("_host_pool", {})
        free = pool.get(int(n_words))
        if free:
            return free.pop()
        return np.zeros(int(n_words), dtype=np.uint32)

    def recycle_host_buffer(self, buf):
        """Called by Proof.release() only (never from a finaliser): the caller has given the buffer up."""
        if isinstance(buf, np.ndarray) and buf.base is None and buf.dtype == np.uint32 and buf.size >= (1 << 16):
            pool = self.__dict__.setdefault("_host_pool", {})
            lst = pool.setdefault(int(buf.size), [])
            if len(lst) < 4 and not any(b is buf for b in lst):
                lst.append(buf)

    def trim(self):
        """Return the idle scratch blocks the context keeps for the next proof to the driver."""
        check(self.lib.swirl_ctx_trim(self.ctx))

    def alloc(self, n_words):
        return torch.empty(int(n_words), dtype=torch.int32, device=self.torch_device)

    def h2d(self, arr):
        return _as_i32_tensor(np.asarray(arr).reshape(-1)).to(self.torch_device)

    def _d2h(self, ptr, n_words):
        """Device pointer -> numpy uint32 (synchronises the ctx stream first)."""
        out = np.empty(int(n_words), dtype=np.uint32)
        check(self.lib.swirl_memcpy_d2h(self.ctx, out.ctypes.data, int(ptr), int(n_words) * 4))
        return out

    def _sync_torch(self):
        # torch work on its current stream must be visible to the ctx stream (nothing to do when torch is already
        # running on the ctx stream: `with torch.cuda.stream(dev.torch_stream())`)
        cur = torch.cuda.current_stream(self.torch_device)
        if cur.cuda_stream != self.stream_ptr():
            cur.synchronize()

    # -- kernel-level primitives ------------------------------------------------------------
    def poseidon2_permute(self, states):
        """states: CUDA int32 tensor of n*16 words, permuted in place."""
        self._sync_torch()
        check(self.lib.swirl_poseidon2_permute(self.ctx, states.data_ptr(), states.numel() // 16))

    def poseidon2_compress(self, pairs):
        n = pairs.numel() // 16
        out = self.alloc(n * 8)
        self._sync_torch()
        check(self.lib.swirl_poseidon2_compress(self.ctx, pairs.data_ptr(), out.data_ptr(), n))
        return out

    def ntt_batch(self, data, log_n, cols, inverse=False):
        self._sync_torch()
        check(self.lib.swirl_ntt_batch(self.ctx, data.data_ptr(), log_n, cols, 1 if inverse else 0))

    def rs_encode(self, matrix, l_skip, log_blowup, out=None):
        """rs_code_matrix: DeviceMatrix (H x W) -> DeviceMatrix ((H << log_blowup) x W)."""
        h, w = matrix.height(), matrix.width()
        if out is None:
            out = self.alloc((h << log_blowup) * w)
        self._sync_torch()
        check(self.lib.swirl_rs_encode(self.ctx, matrix.ptr(), h, w, l_skip, log_blowup, out.data_ptr()))
        return DeviceMatrix(out, h << log_blowup, w)

    def merkle_tree(self, matrix, log_rows_per_query, out=None):
        """Digest layers (concatenated, device tensor) of the tree over `matrix`."""
        h = matrix.height()
        leaves = 1 << max(h - 1, 0).bit_length() if h > 1 else 1
        qs = leaves >> log_rows_per_query
        if out is None:
            out = self.alloc((2 * qs - 1) * 8)
        self._sync_torch()
        check(
            self.lib.swirl_merkle_tree(self.ctx, matrix.ptr(), h, matrix.width(), log_rows_per_query, out.data_ptr())
        )
        return out

    def merkle_query_proofs(self, layers_ptr, query_stride, indices):
        idx = self.h2d(np.asarray(indices, dtype=np.uint32))
        depth = query_stride.bit_length() - 1
        out = self.alloc(len(indices) * depth * 8)
        self._sync_torch()
        check(self.lib.swirl_merkle_query_proofs(self.ctx, layers_ptr, query_stride, idx.data_ptr(), len(indices), out.data_ptr()))
        self.synchronize()
        return out.cpu().numpy().view(np.uint32).reshape(len(indices), depth, 8)

    def matrix_open_rows(self, matrix_ptr, height, width, query_stride, log_rpq, indices):
        idx = self.h2d(np.asarray(indices, dtype=np.uint32))
        out = self.alloc(len(indices) * (width << log_rpq))
        self._sync_torch()
        check(
            self.lib.swirl_matrix_open_rows(
                self.ctx, matrix_ptr, height, width, query_stride, log_rpq, idx.data_ptr(), len(indices), out.data_ptr()
            )
        )
        self.synchronize()
        return out.cpu().numpy().view(np.uint32).reshape(len(indices), 1 << log_rpq, width)

    def sponge_grind(self, state18, bits, min_w=0, max_w=P):
        """Smallest canonical PoW witness for the 18-word sponge state, or None."""
        st = (C.c_uint32 * 18)(*[int(x) for x in state18])
        w = C.c_uint32()
        check(self.lib.swirl_sponge_grind(self.ctx, st, bits, min_w, max_w, C.byref(w)))
        return None if w.value == 0xFFFFFFFF else int(w.value)

    def transcript_grind(self, ts, bits):
        w = C.c_uint32()
        check(self.lib.swirl_transcript_grind(self.ctx, C.byref(ts.c), bits, C.byref(w)))
        return int(w.value)

    # -- LogUp-GKR (fractional_sumcheck, fractional_sumcheck_gkr.rs:60-213) ---------------------------
    def fold_mle(self, table, r, out=None):
        """table: CUDA int32 tensor of 2n EF (flat, column-major matrix of even height); r: 4 Montgomery words.
        Returns the n folded EF: out[j] = t[2j] + (t[2j+1] - t[2j]) r (fold_mle_evals, sumcheck.rs:395-414)."""
        n_out = table.numel() // 8
        out = out if out is not None else self.alloc(n_out * 4)
        rr = np.ascontiguousarray(r, dtype=np.uint32)
        self._sync_torch()
        check(self.lib.swirl_fold_mle(self.ctx, table.data_ptr(), out.data_ptr(), n_out, rr.ctypes.data))
        return out

    def gkr_fractional_sumcheck(self, ts, leaves, log_n, assert_zero=True, n_stored=None, pad_q=None):
        """leaves: CUDA int32 tensor of 2^log_n Frac<EF> (8 words each) — or only the first n_stored of
        them when the rest is the constant fraction (0, pad_q).  Returns dict(frac_sum,
        claims[log_n,16], polys[log_n(log_n-1)/2,12], xi[log_n,4]) of Montgomery words."""
        n_polys = log_n * (log_n - 1) // 2
        frac_sum = np.zeros(8, np.uint32)
        claims = np.zeros((log_n, 16), np.uint32)
        polys = np.zeros((max(n_polys, 1), 12), np.uint32)
        xi = np.zeros((log_n, 4), np.uint32)
        self._sync_torch()
        if n_stored is None:
            check(self.lib.swirl_gkr_fractional_sumcheck(self.ctx, C.byref(ts.c), leaves.data_ptr(), log_n, 1 if assert_zero else 0,
                                                         frac_sum.ctypes.data, claims.ctypes.data, polys.ctypes.data, xi.ctypes.data))
        else:
            pq = np.ascontiguousarray(pad_q, dtype=np.uint32)
            check(self.lib.swirl_gkr_fractional_sumcheck_padded(self.ctx, C.byref(ts.c), leaves.data_ptr(), int(n_stored), pq.ctypes.data,
                                                                log_n, 1 if assert_zero else 0, frac_sum.ctypes.data, claims.ctypes.data,
                                                                polys.ctypes.data, xi.ctypes.data))
        return dict(frac_sum=frac_sum, claims=claims, polys=polys[:n_polys], xi=xi)

    # -- MultiRapProver::prove_rap_constraints (prove_zerocheck_and_logup, logup_zerocheck/mod.rs:40-438) --
    def prove_batch_constraints(self, ts, l_skip, max_constraint_degree, logup_pow_bits, airs):
        """airs: AirProvingContext list sorted by descending height.  Returns (flat proof words, r)."""
        keep = []
        arr = (_lib.AirCtxC * len(airs))(*[a.c(keep) for a in airs])
        n = int(self.lib.swirl_batch_constraints_proof_words(l_skip, max_constraint_degree, arr, len(airs)))
        if n == 0:
            raise _lib.SwirlError(10001, "invalid batch constraint inputs")
        proof = np.zeros(n, dtype=np.uint32)
        n_max = max(max(a.common_main.height().bit_length() - 1 - l_skip for a in airs), 0)
        r = np.zeros((n_max + 1, 4), dtype=np.uint32)
        self._sync_torch()
        check(self.lib.swirl_prove_batch_constraints(self.ctx, C.byref(ts.c), l_skip, max_constraint_degree, logup_pow_bits,
                                                     arr, len(airs), proof.ctypes.data, n, r.ctypes.data))
        return proof, r

    # -- stacked opening reduction (prove_stacked_opening_reduction, prover/stacked_reduction.rs:67-127) --
    def stacked_reduction(self, ts, pcs_list, need_rot_per_commit, r):
        """need_rot_per_commit[c][mat] -> bool.  r: (>= 1 + n_max, 4) Montgomery words.
        Returns dict(univariate_round_coeffs, sumcheck_round_polys, stacking_openings (list per commit), u, flat)."""
        handles = (C.c_void_p * len(pcs_list))(*[d._h for d in pcs_list])
        n = int(self.lib.swirl_stacked_reduction_proof_words(handles, len(pcs_list)))
        proof = np.zeros(n, dtype=np.uint32)
        rots = [np.asarray([1 if b else 0 for b in rr], dtype=np.uint8) for rr in need_rot_per_commit]
        rot_ptrs = (C.c_void_p * len(rots))(*[a.ctypes.data for a in rots])
        r = np.ascontiguousarray(r, dtype=np.uint32)
        p0 = pcs_list[0].params
        u = np.zeros((p0.n_stack + 1, 4), dtype=np.uint32)
        self._sync_torch()
        check(self.lib.swirl_stacked_reduction(self.ctx, C.byref(ts.c), handles, len(pcs_list), rot_ptrs, r.ctypes.data,
                                               r.size // 4, proof.ctypes.data, n, u.ctypes.data))
        n0 = (2 * ((1 << p0.l_skip) - 1) + 1) * 4
        n1 = n0 + p0.n_stack * 8
        openings, off = [], n1
        for d in pcs_list:
            openings.append(proof[off : off + d.width * 4].reshape(d.width, 4))
            off += d.width * 4
        return dict(univariate_round_coeffs=proof[:n0].reshape(-1, 4), sumcheck_round_polys=proof[n0:n1].reshape(-1, 2, 4),
                    stacking_openings=openings, u=u, flat=proof)

    # -- OpeningProver::prove_openings (cpu_backend.rs:139-220) -------------------------------------------
    def prove_openings(self, ts, whir_cfg, pcs_list, need_rot_per_commit, r):
        """Returns (flat StackingProof words, flat WhirProof words)."""
        handles = (C.c_void_p * len(pcs_list))(*[d._h for d in pcs_list])
        cc, pc = whir_cfg.c(), pcs_list[0].params.c()
        widths = np.array([d.width for d in pcs_list], dtype=np.uint64)
        n_st = int(self.lib.swirl_stacked_reduction_proof_words(handles, len(pcs_list)))
        n_wh = int(self.lib.swirl_whir_proof_words(C.byref(pc), C.byref(cc), len(pcs_list), widths.ctypes.data))
        if n_wh == 0:
            raise _lib.SwirlError(10001, "invalid WHIR configuration")
        st, wh = self.host_buffer(n_st), self.host_buffer(n_wh)  # every word is written by the library
        rots = [np.asarray([1 if b else 0 for b in rr], dtype=np.uint8) for rr in need_rot_per_commit]
        rot_ptrs = (C.c_void_p * len(rots))(*[a.ctypes.data for a in rots])
        r = np.ascontiguousarray(r, dtype=np.uint32)
        self._sync_torch()
        check(self.lib.swirl_prove_openings(self.ctx, C.byref(ts.c), C.byref(cc), handles, len(pcs_list), rot_ptrs, r.ctypes.data,
                                            r.size // 4, st.ctypes.data, n_st, wh.ctypes.data, n_wh))
        return st, wh

    # -- WHIR opening (prove_whir_opening, prover/whir.rs:78-341) -----------------------------------
    def whir_open(self, ts, cfg, params, pcs_list, u):
        """pcs_list: StackedPcsData (common main first).  u: (l_skip+n_stack, 4) Montgomery words.
        Returns the flat WhirProof words (layout in include/swirl_b200.h)."""
        cc, pc = cfg.c(), params.c()
        widths = np.array([d.width for d in pcs_list], dtype=np.uint64)
        n = int(self.lib.swirl_whir_proof_words(C.byref(pc), C.byref(cc), len(pcs_list), widths.ctypes.data))
        if n == 0:
            raise _lib.SwirlError(10001, "invalid WHIR configuration")
        proof = np.zeros(n, dtype=np.uint32)
        handles = (C.c_void_p * len(pcs_list))(*[d._h for d in pcs_list])
        u = np.ascontiguousarray(u, dtype=np.uint32)
        self._sync_torch()
        check(self.lib.swirl_whir_open(self.ctx, C.byref(ts.c), C.byref(cc), handles, len(pcs_list), u.ctypes.data,
                                       proof.ctypes.data, n))
        return proof

    def stack(self, params, traces):
        """Layout + stacked matrix only (first half of commit): the handle has no codeword or tree until
        `attach_external` (sharded commitments, multi.py)."""
        n = len(traces)
        arr = (MatrixC * max(n, 1))()
        for i, t in enumerate(traces):
            arr[i] = MatrixC(t.ptr(), t.height(), t.width())
        h = C.c_void_p()
        pc = params.c()
        self._sync_torch()
        check(self.lib.swirl_stack(self.ctx, C.byref(pc), arr, n, C.byref(h)))
        return StackedPcsData(self, h, params, list(traces))

    # -- TraceCommitter::commit ---------------------------------------------------------------
    def commit(self, params, traces):
        """stacked_commit over device-resident, height-sorted traces.
        Returns (root: uint32[8] Montgomery words, StackedPcsData)."""
        n = len(traces)
        arr = (MatrixC * max(n, 1))()
        for i, t in enumerate(traces):
            arr[i] = MatrixC(t.ptr(), t.height(), t.width())
        root = np.zeros(8, dtype=np.uint32)
        h = C.c_void_p()
        pc = params.c()
        self._sync_torch()
        check(self.lib.swirl_commit(self.ctx, C.byref(pc), arr, n, root.ctypes.data, C.byref(h)))
        return root, StackedPcsData(self, h, params, list(traces))

    def commit_host(self, params, host_traces):
        """Same through host buffers: host_traces = [(uint32 ndarray col-major flat, height, width)].
        The H2D transport happens inside the call (the e2e path bench.py times)."""
        n = len(host_traces)
        arr = (MatrixC * max(n, 1))()
        keep = []
        for i, (vals, hh, ww) in enumerate(host_traces):
            if isinstance(vals, torch.Tensor):
                ptr = vals.data_ptr()
            else:
                vals = np.ascontiguousarray(vals, dtype=np.uint32)
                ptr = vals.ctypes.data
            keep.append(vals)
            arr[i] = MatrixC(ptr, hh, ww)
        root = np.zeros(8, dtype=np.uint32)
        h = C.c_void_p()
        pc = params.c()
        check(self.lib.swirl_commit_host(self.ctx, C.byref(pc), arr, n, root.ctypes.data, C.byref(h)))
        return root, StackedPcsData(self, h, params, None)
