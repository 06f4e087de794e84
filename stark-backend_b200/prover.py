"""Host-side mirror of the reference prover driver over the C ABI.

reference: Coordinator::prove (crates/stark-backend/src/prover/mod.rs:104-198) driving a
ProverDevice = TraceCommitter + MultiRapProver + OpeningProver (prover/hal.rs:65-138); model
implementation crates/cuda-backend/src/gpu_backend.rs:44-212.  Every compute step is a call into
libswirl_b200.so; this file only sequences them and feeds the transcript, as the Rust Coordinator does.
"""
import ctypes as C

import numpy as np

from . import lib as _lib
from .backend import PcsParams, Transcript, WhirConfig
from .lib import check


class SystemParams:
    """reference: SystemParams (config.rs:36-50)."""

    def __init__(self, l_skip, n_stack, log_blowup, whir, logup_pow_bits, max_constraint_degree):
        self.l_skip, self.n_stack, self.log_blowup = l_skip, n_stack, log_blowup
        self.whir, self.logup_pow_bits, self.max_constraint_degree = whir, logup_pow_bits, max_constraint_degree

    def pcs(self):
        return PcsParams(self.l_skip, self.n_stack, self.log_blowup, self.whir.k)


class CommittedTraceData:
    """reference: CommittedTraceData (prover/types.rs:61-68): commitment + trace + PcsData."""

    def __init__(self, commitment, trace, data):
        self.commitment, self.trace, self.data = commitment, trace, data


class AirProvingKey:
    """The parts of DeviceStarkProvingKey the driver reads (prover/types.rs:90-110)."""

    def __init__(self, is_required=True, preprocessed_data=None):
        self.is_required, self.preprocessed_data = is_required, preprocessed_data


class Proof:
    """reference: Proof (proof.rs:20-60), field elements as Montgomery words, sub-proofs flat in the
    layouts documented in include/swirl_b200.h."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def release(self):
        """Hand the large host proof buffers back to the device's pool for the next proof (what a pooling allocator does for
        the reference's Vecs).  EXPLICIT: the caller promises not to read this proof's sections afterwards; a proof that is
        merely dropped keeps its buffers until the garbage collector frees them.  Also usable as a context manager."""
        dev = self.__dict__.pop("_device", None)
        if dev is not None:
            for name in ("whir_proof", "stacking_proof", "constraints_proof"):
                buf = self.__dict__.pop(name, None)
                if buf is not None:
                    dev.recycle_host_buffer(buf)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.release()
        return False

    def words(self):
        return np.concatenate([self.common_main_commit, self.constraints_proof, self.stacking_proof, self.whir_proof])

    def encode(self):
        """Bytes of `Proof::encode_to_vec()` (proof.rs:226-257, codec.rs): canonical LE u32 words with the
        reference's length prefixes."""
        from . import codec

        return codec.encode_proof(self.shape, self.common_main_commit, self.constraints_proof, self.stacking_proof,
                                  self.whir_proof)


class Coordinator:
    def __init__(self, device, params, transcript=None, plan_memory=False):
        """plan_memory: before committing, model the proof's device memory (memory.py, after the reference's
        memory_metering.rs) and switch the device to cache_rs_code_matrix = false when keeping the codeword would not
        fit the free HBM (the reference's default for its GPU prover, cuda-backend/src/device.rs:113-121)."""
        self.device, self.params = device, params
        self.transcript = transcript if transcript is not None else Transcript()
        self.plan_memory = plan_memory
        self.memory_estimate = None

    def _plan(self, airs):
        from . import memory as M

        dev, P = self.device, self.params
        counts = M.ProvingMemoryCounts.from_airs(airs, P.l_skip)
        H = 1 << (P.l_skip + P.n_stack)
        alias = len(airs) == 1 and airs[0].common_main.height() == H
        cfg = M.ProvingMemoryConfig(P.l_skip, P.l_skip + P.n_stack, P.log_blowup, P.whir.k, True, alias)
        st = dev.mem_stats()
        free = st["device_free"] + (st["held"] - st["live"])  # idle arena blocks are reusable
        cache, est = M.choose_cache_rs_code_matrix(cfg, counts, free)
        dev.set_cache_rs_code_matrix(cache)
        self.memory_estimate = {"cache_rs_code_matrix": cache, "free_bytes": free, **est.__dict__}

    def prove_host(self, vk_pre_hash, per_air_pk, air, host_trace, height, width):
        """Single-AIR proof with the common main trace in HOST memory: the H2D transport runs inside
        swirl_commit_host, pipelined with the RS encoding and leaf hashing by column groups; the later
        phases read the device copy the commitment keeps (DeviceDataTransporter + Coordinator::prove)."""
        from .backend import AirProvingContext, PcsTraceView

        dev, P = self.device, self.params
        root, common = dev.commit_host(P.pcs(), [(host_trace, height, width)])
        view = PcsTraceView(common, dev.lib.swirl_pcs_stacked_matrix(common._h), height, width)
        ctx = AirProvingContext(air.nodes, air.constraint_idx, air.interactions, air.constraint_degree, air.need_rot, view,
                                air.public_values)
        return self.prove(vk_pre_hash, per_air_pk, [(0, ctx, [])], precommitted=(root, common))

    def prove(self, vk_pre_hash, per_air_pk, per_trace, precommitted=None):
        """per_air_pk: list of AirProvingKey indexed by air id.  per_trace: list of
        (air_id, AirProvingContext, [CommittedTraceData cached mains]) — sorted here as
        ProvingContext::into_sorted does (height descending, then air id; prover/types.rs:144-148)."""
        import os
        import time

        dev, ts, P = self.device, self.transcript, self.params
        trace_on = bool(os.environ.get("SWIRL_PHASE_TIMING"))
        marks = [("start", time.perf_counter())]

        def mark(name):
            if trace_on:
                dev.synchronize()
                marks.append((name, time.perf_counter()))

        ts.observe(vk_pre_hash)
        per_trace = sorted(per_trace, key=lambda t: (-t[1].common_main.height(), t[0]))
        if self.plan_memory and precommitted is None:
            self._plan([t[1] for t in per_trace])
        if precommitted is None:
            root, common = dev.commit(P.pcs(), [t[1].common_main for t in per_trace])
        else:
            root, common = precommitted
        ts.observe(root)
        present = {t[0]: t for t in per_trace}
        for air_id, pk in enumerate(per_air_pk):
            t = present.get(air_id)
            if not pk.is_required:
                ts.observe(np.array([0x0FFFFFFE if t is not None else 0], dtype=np.uint32))
            if t is not None:
                if pk.preprocessed_data is not None:
                    ts.observe(pk.preprocessed_data.commitment)
                else:
                    ts.observe(np.array([_to_mont(t[1].common_main.height().bit_length() - 1)], dtype=np.uint32))
                for cd in t[2]:
                    ts.observe(cd.commitment)
                ts.observe(t[1].public_values)
        airs = [t[1] for t in per_trace]
        mark("commit")
        constraints_proof, r = dev.prove_batch_constraints(ts, P.l_skip, P.max_constraint_degree, P.logup_pow_bits, airs)
        # prove_openings (cpu_backend.rs:139-220)
        pcs_list, need_rot = [common], [[a.need_rot for a in airs]]
        for air_id, a, cached in per_trace:
            pk = per_air_pk[air_id]
            for cd in ([pk.preprocessed_data] if pk.preprocessed_data is not None else []) + list(cached):
                pcs_list.append(cd.data)
                need_rot.append([a.need_rot])
        mark("batch_constraints")
        stacking, whir = dev.prove_openings(ts, P.whir, pcs_list, need_rot, r)
        mark("openings")
        from .codec import AirShape, ProofShape

        lifted = lambda a: max(a.common_main.height(), 1 << P.l_skip)
        total_interactions = sum(len(a.interactions) * lifted(a) for a in airs)
        log_h = lambda a: a.common_main.height().bit_length() - 1
        shape = ProofShape(
            l_skip=P.l_skip, n_stack=P.n_stack, log_blowup=P.log_blowup, max_constraint_degree=P.max_constraint_degree,
            k_whir=P.whir.k, num_queries=list(P.whir.num_queries),
            airs=[AirShape([a.common_main.width()] + ([a.preprocessed.width()] if a.preprocessed is not None else [])
                           + [m.width() for m in a.cached_mains], a.need_rot) for a in airs],
            gkr_layers=total_interactions.bit_length() if total_interactions else 0,  # calculate_n_logup, lib.rs:82-93
            n_max=max(max(log_h(a) for a in airs) - P.l_skip, 0), commit_widths=[d.width for d in pcs_list],
            trace_vdata=[(log_h(present[i][1]), [cd.commitment for cd in present[i][2]]) if i in present else None
                         for i in range(len(per_air_pk))],
            public_values=[present[i][1].public_values if i in present else np.zeros(0, np.uint32)
                           for i in range(len(per_air_pk))])
        self.phase_ms = {marks[i][0]: 1e3 * (marks[i][1] - marks[i - 1][1]) for i in range(1, len(marks))}
        return Proof(common_main_commit=root, constraints_proof=constraints_proof, stacking_proof=stacking, whir_proof=whir, shape=shape, _device=dev,
                     r=r, public_values=[a.public_values for a in airs], common_main_pcs=common,
                     log_heights=[a.common_main.height().bit_length() - 1 for a in airs])


def _to_mont(x):
    return int((int(x) % 0x78000001) * (1 << 32) % 0x78000001)
