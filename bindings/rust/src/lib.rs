//! `openvm-b200-backend`: `ProverBackend` / `ProverDevice` / `DeviceDataTransporter` / `StarkEngine` over
//! `libswirl_b200.so` — the shim of SURVEY.md §8(f)-1, modelled on the reference's own GPU backend
//! (crates/cuda-backend/src/{gpu_backend.rs:44-212, data_transporter.rs:35-106, engine.rs:24-83}).
//!
//! The three phase traits map one to one onto the phase-level entry points of include/swirl_b200.h.  Proof parts come
//! back as flat Montgomery words in the field order of proof.rs and are rebuilt into the reference structs in
//! `proof_parts`, so `Proof::encode_to_vec()`, the reference verifier and `backend_test_suite!` are used unchanged
//! (tests/backend_suite.rs).
//!
//! NOT COMPILED in the repository this library was developed in (no Rust toolchain in that image, Plonky3 crates not
//! vendored).  The same call sequence, buffer sizes and section offsets are exercised from Python
//! (stark-backend_b200/{backend,prover,codec}.py) and checked bit for bit against the CPU oracle and the reference's
//! own CUDA kernels; tests/test_abi.py keeps `ffi.rs` in sync with the header.
pub mod ffi;
pub mod proof_parts;

use std::{ffi::CStr, sync::Arc};

use ffi::*;
use openvm_stark_backend::{
    air_builders::symbolic::{
        symbolic_variable::Entry, SymbolicConstraintsDag, SymbolicExpressionNode,
    },
    keygen::types::MultiStarkProvingKey,
    proof::{BatchConstraintProof, GkrProof, StackingProof, WhirProof},
    prover::{
        stacked_pcs::StackedPcsData, AirProvingContext, ColMajorMatrix, CommittedTraceData, Coordinator,
        DeviceDataTransporter, DeviceMultiStarkProvingKey, DeviceStarkProvingKey, MatrixDimensions, MultiRapProver,
        OpeningProver, ProverBackend, ProverDevice, ProvingContext, TraceCommitter,
    },
    FiatShamirTranscript, StarkEngine, SystemParams,
};
use openvm_stark_sdk::config::baby_bear_poseidon2::{BabyBearPoseidon2Config as SC, Digest, EF, F};
use p3_field::{PrimeCharacteristicRing, PrimeField32};

// ---------------------------------------------------------------------------------------------------------------
// errors (hal.rs:68-74: one Error that every sub-error converts into)
// ---------------------------------------------------------------------------------------------------------------
#[derive(Debug, thiserror::Error)]
pub enum B200Error {
    /// 1..999: cudaError_t of the failing CUDA call (cuda-common/src/error.rs:53-60)
    #[error("CUDA error {0}: {1}")]
    Cuda(i32, String),
    /// SWIRL_ERR_NONZERO_ROOT_SUM = LogupZerocheckError::NonZeroRootSum (fractional_sumcheck_gkr.rs:88-91)
    #[error("LogUp numerator sum is not zero")]
    NonZeroRootSum,
    #[error("swirl_b200 error {0}: {1}")]
    Other(i32, String),
}

pub fn check(rc: i32) -> Result<(), B200Error> {
    if rc == 0 {
        return Ok(());
    }
    let msg = unsafe { CStr::from_ptr(swirl_last_error()) }.to_string_lossy().into_owned();
    Err(match rc {
        1..=999 => B200Error::Cuda(rc, msg),
        10005 => B200Error::NonZeroRootSum,
        _ => B200Error::Other(rc, msg),
    })
}

// ---------------------------------------------------------------------------------------------------------------
// field <-> word conversions.  A p3 `BabyBear` is one u32 in Montgomery form (x * 2^32 mod p): the same bytes the
// library reads and writes (data_transporter.rs:93-106 memcpys `Vec<BabyBear>` as is).
// ---------------------------------------------------------------------------------------------------------------
#[inline]
pub fn f_words(v: &[F]) -> &[u32] {
    // SAFETY: BabyBear = MontyField31 is #[repr(transparent)] over u32
    unsafe { std::slice::from_raw_parts(v.as_ptr() as *const u32, v.len()) }
}
#[inline]
pub fn f_from_word(w: u32) -> F {
    // SAFETY: as above; every word the library returns is canonical (< p)
    unsafe { std::mem::transmute::<u32, F>(w) }
}
#[inline]
pub fn f_to_word(x: F) -> u32 {
    unsafe { std::mem::transmute::<F, u32>(x) }
}

// ---------------------------------------------------------------------------------------------------------------
// device context shared by every handle (ADVICE: PCS data must not outlive the context that owns its memory)
// ---------------------------------------------------------------------------------------------------------------
pub struct Ctx(pub *mut SwirlCtx);
unsafe impl Send for Ctx {}
unsafe impl Sync for Ctx {}
impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { swirl_ctx_destroy(self.0) };
    }
}

/// DeviceMatrix of this backend (cuda-backend/src/base.rs:8-12): column-major Montgomery words in HBM.
#[derive(Clone)]
pub struct B200Matrix {
    ctx: Arc<Ctx>,
    buf: Arc<DeviceBuf>,
    height: usize,
    width: usize,
}
struct DeviceBuf {
    ctx: Arc<Ctx>,
    ptr: *mut u32,
}
unsafe impl Send for DeviceBuf {}
unsafe impl Sync for DeviceBuf {}
impl Drop for DeviceBuf {
    fn drop(&mut self) {
        unsafe { swirl_free(self.ctx.0, self.ptr as *mut _) };
    }
}
impl MatrixDimensions for B200Matrix {
    fn height(&self) -> usize {
        self.height
    }
    fn width(&self) -> usize {
        self.width
    }
}
impl B200Matrix {
    pub fn pod(&self) -> SwirlMatrix {
        SwirlMatrix { data: self.buf.ptr as *const u32, height: self.height as u64, width: self.width as u64 }
    }
}

/// StackedPcsData of this backend (cuda-backend/src/stacked_pcs.rs:30-46): layout, stacked matrix, codeword (when cached)
/// and digest layers stay on the device behind the opaque handle.
pub struct B200PcsData {
    ctx: Arc<Ctx>,
    pub pcs: *mut SwirlPcs,
    pub commitment: Digest,
    /// keeps the device traces the commitment aliases alive (a single full-height trace is its own stacked matrix)
    _traces: Vec<B200Matrix>,
}
unsafe impl Send for B200PcsData {}
unsafe impl Sync for B200PcsData {}
impl Drop for B200PcsData {
    fn drop(&mut self) {
        unsafe { swirl_pcs_free(self.ctx.0, self.pcs) };
    }
}
impl B200PcsData {
    pub fn stacked_width(&self) -> usize {
        unsafe { swirl_pcs_stacked_width(self.pcs) as usize }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// per-AIR prover data: the vk's SymbolicConstraintsDag flattened into the PODs of the C ABI once per proving key
// (reference: AirDataGpu built in transport_pk_to_device, cuda-backend/src/data_transporter.rs:52-58)
// ---------------------------------------------------------------------------------------------------------------
pub struct B200AirData {
    pub nodes: Vec<SwirlDagNode>,
    pub constraint_idx: Vec<u32>,
    pub interactions: Vec<SwirlInteraction>,
    pub msg_nodes: Vec<u32>,
}

impl B200AirData {
    /// air_builders/symbolic/dag.rs:17-96 -> include/swirl_b200.h SWIRL_NODE_*.
    pub fn new(dag: &SymbolicConstraintsDag<F>) -> Self {
        let nodes = dag
            .constraints
            .nodes
            .iter()
            .map(|n| match n {
                SymbolicExpressionNode::Variable(v) => match v.entry {
                    Entry::Preprocessed { offset } => SwirlDagNode { op: SWIRL_NODE_VAR_PREP, a: v.index as u32, b: offset as u32, c: 0 },
                    Entry::Main { part_index, offset } => {
                        SwirlDagNode { op: SWIRL_NODE_VAR_MAIN, a: v.index as u32, b: offset as u32, c: part_index as u32 }
                    }
                    Entry::Public => SwirlDagNode { op: SWIRL_NODE_VAR_PUBLIC, a: v.index as u32, b: 0, c: 0 },
                    Entry::Challenge => unreachable!("SWIRL AIRs have no challenge phase"),
                },
                SymbolicExpressionNode::IsFirstRow => SwirlDagNode { op: SWIRL_NODE_IS_FIRST, a: 0, b: 0, c: 0 },
                SymbolicExpressionNode::IsLastRow => SwirlDagNode { op: SWIRL_NODE_IS_LAST, a: 0, b: 0, c: 0 },
                SymbolicExpressionNode::IsTransition => SwirlDagNode { op: SWIRL_NODE_IS_TRANSITION, a: 0, b: 0, c: 0 },
                SymbolicExpressionNode::Constant(c) => SwirlDagNode { op: SWIRL_NODE_CONST, a: f_to_word(*c), b: 0, c: 0 },
                SymbolicExpressionNode::Add { left_idx, right_idx, .. } => {
                    SwirlDagNode { op: SWIRL_NODE_ADD, a: *left_idx as u32, b: *right_idx as u32, c: 0 }
                }
                SymbolicExpressionNode::Sub { left_idx, right_idx, .. } => {
                    SwirlDagNode { op: SWIRL_NODE_SUB, a: *left_idx as u32, b: *right_idx as u32, c: 0 }
                }
                SymbolicExpressionNode::Neg { idx, .. } => SwirlDagNode { op: SWIRL_NODE_NEG, a: *idx as u32, b: 0, c: 0 },
                SymbolicExpressionNode::Mul { left_idx, right_idx, .. } => {
                    SwirlDagNode { op: SWIRL_NODE_MUL, a: *left_idx as u32, b: *right_idx as u32, c: 0 }
                }
            })
            .collect();
        let mut msg_nodes = Vec::new();
        let interactions = dag
            .interactions
            .iter()
            .map(|it| {
                let off = msg_nodes.len() as u32;
                msg_nodes.extend(it.message.iter().map(|&i| i as u32));
                SwirlInteraction { count_node: it.count as u32, bus_index: it.bus_index as u32, msg_offset: off, msg_len: it.message.len() as u32 }
            })
            .collect();
        Self { nodes, constraint_idx: dag.constraints.constraint_idx.iter().map(|&i| i as u32).collect(), interactions, msg_nodes }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// the backend (gpu_backend.rs:44-63)
// ---------------------------------------------------------------------------------------------------------------
#[derive(Clone, Copy, Default)]
pub struct B200Backend;

impl ProverBackend for B200Backend {
    const CHALLENGE_EXT_DEGREE: u8 = 4;
    type Val = F;
    type Challenge = EF;
    type Commitment = Digest;
    type Matrix = B200Matrix;
    type PcsData = B200PcsData;
    type OtherAirData = B200AirData;

    fn constraint_eval_buffer_size(pk: &DeviceStarkProvingKey<Self>) -> usize {
        // value slots of the compiled constraint program are sized inside the library; the DAG size bounds them
        pk.other_data.nodes.len()
    }
}

/// Fiat-Shamir transcript = the reference's DuplexSponge state as the POD the library shares with its kernels
/// (`DeviceSpongeState`, cuda-backend/cuda/src/sponge.cu:13-17).  observe / sample run the library's host sponge, so host
/// and device agree by construction; `grind` runs on the GPU and returns the smallest witness.
#[derive(Clone)]
pub struct B200Transcript {
    pub state: SwirlTranscript,
    ctx: Option<Arc<Ctx>>,
}
impl Default for B200Transcript {
    fn default() -> Self {
        Self { state: SwirlTranscript { state: [0; 16], absorb_idx: 0, sample_idx: 0 }, ctx: None }
    }
}
impl FiatShamirTranscript<SC> for B200Transcript {
    fn observe(&mut self, value: F) {
        let w = f_to_word(value);
        check(unsafe { swirl_transcript_observe(&mut self.state, &w, 1) }).expect("observe");
    }
    fn sample(&mut self) -> F {
        let mut w = 0u32;
        check(unsafe { swirl_transcript_sample(&mut self.state, &mut w, 1) }).expect("sample");
        f_from_word(w)
    }
    fn observe_commit(&mut self, digest: Digest) {
        check(unsafe { swirl_transcript_observe(&mut self.state, f_words(&digest).as_ptr(), 8) }).expect("observe_commit");
    }
    fn grind(&mut self, bits: usize) -> F {
        match &self.ctx {
            Some(ctx) => {
                let mut w = 0u32; // canonical
                check(unsafe { swirl_transcript_grind(ctx.0, &mut self.state, bits as i32, &mut w) }).expect("grind");
                F::from_u32(w)
            }
            None => {
                // no device attached (verifier side): the trait's brute force
                let witness = (0..F::ORDER_U32).map(F::from_u32).find(|w| self.clone().check_witness(bits, *w)).expect("PoW");
                assert!(self.check_witness(bits, witness));
                witness
            }
        }
    }
}

/// GpuDevice of this backend (cuda-backend/src/device.rs:52-110): one CUDA device, one non-blocking stream, the
/// SystemParams it proves for.
#[derive(Clone)]
pub struct B200Device {
    pub ctx: Arc<Ctx>,
    pub params: SystemParams,
}

impl B200Device {
    pub fn new(device: i32, params: SystemParams) -> Result<Self, B200Error> {
        let mut ctx = std::ptr::null_mut();
        check(unsafe { swirl_ctx_create(device, &mut ctx) })?;
        Ok(Self { ctx: Arc::new(Ctx(ctx)), params })
    }
    /// GpuDevice::set_cache_rs_code_matrix (device.rs:108-110)
    pub fn set_cache_rs_code_matrix(&self, on: bool) -> Result<(), B200Error> {
        check(unsafe { swirl_ctx_set_cache_rs_code_matrix(self.ctx.0, on as i32) })
    }
    /// Run-time compiled per-AIR round-0 kernels: 0 = interpreter only, 1 = tall traces (default), 2 = always.
    pub fn set_jit(&self, mode: i32) -> Result<(), B200Error> {
        check(unsafe { swirl_ctx_set_jit(self.ctx.0, mode) })
    }
    /// Sumcheck rounds through the mapped mailbox (default) or one launch + stream synchronisation per round, the
    /// reference's pattern (logup_zerocheck/fractional.rs:649-).  Same proof either way; the library turns it off by
    /// itself under a profiler or sanitizer.
    pub fn set_round_link(&self, on: bool) -> Result<(), B200Error> {
        check(unsafe { swirl_ctx_set_round_link(self.ctx.0, on as i32) })
    }
    /// (stream synchronisations, rounds received through the mailbox) since the context was created.
    pub fn sync_stats(&self) -> Result<(u64, u64), B200Error> {
        let (mut syncs, mut ms, mut links) = (0u64, 0f64, 0u64);
        check(unsafe { swirl_ctx_sync_stats(self.ctx.0, &mut syncs, &mut ms) })?;
        check(unsafe { swirl_ctx_link_stats(self.ctx.0, &mut links) })?;
        Ok((syncs, links))
    }
    fn pcs_params(&self) -> SwirlPcsParams {
        SwirlPcsParams {
            l_skip: self.params.l_skip as i32,
            n_stack: self.params.n_stack as i32,
            log_blowup: self.params.log_blowup as i32,
            k_whir: self.params.k_whir() as i32,
        }
    }
    fn whir_config(&self) -> SwirlWhirConfig {
        let w = &self.params.whir;
        let mut num_queries = [0i32; 32];
        for (i, r) in w.rounds.iter().enumerate() {
            num_queries[i] = r.num_queries as i32;
        }
        SwirlWhirConfig {
            k: w.k as i32,
            num_rounds: w.rounds.len() as i32,
            num_queries,
            mu_pow_bits: w.mu_pow_bits as i32,
            query_phase_pow_bits: w.query_phase_pow_bits as i32,
            folding_pow_bits: w.folding_pow_bits as i32,
        }
    }
    fn alloc(&self, words: usize) -> Result<Arc<DeviceBuf>, B200Error> {
        let mut p = std::ptr::null_mut();
        check(unsafe { swirl_malloc(self.ctx.0, words * 4, &mut p) })?;
        Ok(Arc::new(DeviceBuf { ctx: self.ctx.clone(), ptr: p as *mut u32 }))
    }
}

impl TraceCommitter<B200Backend> for B200Device {
    type Error = B200Error;

    /// hal.rs:84-87; gpu_backend.rs:65-88.  Borrows the device matrices; the returned PcsData owns the codeword and tree.
    fn commit(&self, traces: &[&B200Matrix]) -> Result<(Digest, B200PcsData), B200Error> {
        let pods: Vec<SwirlMatrix> = traces.iter().map(|t| t.pod()).collect();
        let (mut root, mut pcs) = ([0u32; 8], std::ptr::null_mut());
        check(unsafe { swirl_commit(self.ctx.0, &self.pcs_params(), pods.as_ptr(), pods.len(), root.as_mut_ptr(), &mut pcs) })?;
        let commitment: Digest = root.map(f_from_word);
        Ok((commitment, B200PcsData { ctx: self.ctx.clone(), pcs, commitment, _traces: traces.iter().map(|t| (*t).clone()).collect() }))
    }
}

/// One present AIR as the C ABI wants it; `keep` owns the arrays the PODs point into.
struct AirPods {
    ctxs: Vec<SwirlAirCtx>,
    _cached: Vec<Vec<SwirlMatrix>>,
    _prep: Vec<Box<SwirlMatrix>>,
    _pvs: Vec<Vec<u32>>,
}

fn air_pods(mpk: &DeviceMultiStarkProvingKey<B200Backend>, ctx: &ProvingContext<B200Backend>) -> AirPods {
    let n = ctx.per_trace.len();
    let (mut ctxs, mut cached, mut prep, mut pvs) = (Vec::with_capacity(n), Vec::with_capacity(n), Vec::new(), Vec::with_capacity(n));
    // the Coordinator has sorted per_trace by (height desc, air id) already (prover/types.rs:144-148)
    for (air_id, air) in &ctx.per_trace {
        let pk = &mpk.per_air[*air_id];
        let d = &pk.other_data;
        cached.push(air.cached_mains.iter().map(|c| c.trace.pod()).collect::<Vec<_>>());
        pvs.push(f_words(&air.public_values).to_vec());
        let preprocessed = match &pk.preprocessed_data {
            Some(p) => {
                prep.push(Box::new(p.trace.pod()));
                &**prep.last().unwrap() as *const SwirlMatrix
            }
            None => std::ptr::null(),
        };
        let (c, pv) = (cached.last().unwrap(), pvs.last().unwrap());
        ctxs.push(SwirlAirCtx {
            nodes: d.nodes.as_ptr(),
            n_nodes: d.nodes.len() as u64,
            constraint_idx: d.constraint_idx.as_ptr(),
            n_constraints: d.constraint_idx.len() as u64,
            interactions: d.interactions.as_ptr(),
            n_interactions: d.interactions.len() as u64,
            msg_nodes: d.msg_nodes.as_ptr(),
            constraint_degree: pk.vk.max_constraint_degree as u32,
            need_rot: pk.vk.params.need_rot as u32,
            public_values: pv.as_ptr(),
            n_public_values: pv.len() as u64,
            common_main: air.common_main.pod(),
            cached_mains: c.as_ptr(),
            n_cached: c.len() as u64,
            preprocessed,
        });
    }
    AirPods { ctxs, _cached: cached, _prep: prep, _pvs: pvs }
}

impl MultiRapProver<B200Backend, B200Transcript> for B200Device {
    type PartialProof = (GkrProof<SC>, BatchConstraintProof<SC>);
    type Artifacts = Vec<EF>;
    type Error = B200Error;

    /// hal.rs:94-112; gpu_backend.rs:104-143 -> swirl_prove_batch_constraints (LogUp-GKR + batch constraint sumcheck).
    fn prove_rap_constraints(
        &self,
        transcript: &mut B200Transcript,
        mpk: &DeviceMultiStarkProvingKey<B200Backend>,
        ctx: &ProvingContext<B200Backend>,
        _common_main_pcs_data: &B200PcsData,
    ) -> Result<(Self::PartialProof, Vec<EF>), B200Error> {
        transcript.ctx = Some(self.ctx.clone());
        let pods = air_pods(mpk, ctx);
        let (l_skip, d) = (self.params.l_skip as i32, mpk.max_constraint_degree as i32);
        let words = unsafe { swirl_batch_constraints_proof_words(l_skip, d, pods.ctxs.as_ptr(), pods.ctxs.len()) };
        let n_max = ctx.per_trace.iter().map(|(_, a)| a.common_main.height().trailing_zeros() as usize).max().unwrap_or(0)
            .saturating_sub(self.params.l_skip);
        let (mut flat, mut r) = (vec![0u32; words], vec![0u32; 4 * (n_max + 1)]);
        check(unsafe {
            swirl_prove_batch_constraints(self.ctx.0, &mut transcript.state, l_skip, d, self.params.logup.pow_bits as i32,
                                          pods.ctxs.as_ptr(), pods.ctxs.len(), flat.as_mut_ptr(), words, r.as_mut_ptr())
        })?;
        let shape = proof_parts::BatchShape::new(&self.params, mpk, ctx);
        Ok((proof_parts::split_gkr_and_batch(&flat, &shape), proof_parts::ef_vec(&r)))
    }
}

impl OpeningProver<B200Backend, B200Transcript> for B200Device {
    type OpeningProof = (StackingProof<SC>, WhirProof<SC>);
    type OpeningPoints = Vec<EF>;
    type Error = B200Error;

    /// hal.rs:118-138; gpu_backend.rs:145-211 -> swirl_prove_openings (stacked reduction, u_cube, WHIR).
    fn prove_openings(
        &self,
        transcript: &mut B200Transcript,
        mpk: &DeviceMultiStarkProvingKey<B200Backend>,
        ctx: ProvingContext<B200Backend>,
        common_main_pcs_data: B200PcsData,
        r: Vec<EF>,
    ) -> Result<Self::OpeningProof, B200Error> {
        transcript.ctx = Some(self.ctx.clone());
        // commitment order of StackedReductionProver::new (stacked_reduction.rs:36-50): common main, then per trace the
        // preprocessed commitment (if any) and the cached mains
        let mut pcs: Vec<*const SwirlPcs> = vec![common_main_pcs_data.pcs as *const _];
        let mut need_rot: Vec<Vec<u8>> = vec![ctx.per_trace.iter().map(|(i, _)| mpk.per_air[*i].vk.params.need_rot as u8).collect()];
        let mut keep: Vec<Arc<B200PcsData>> = Vec::new();
        for (air_id, air) in &ctx.per_trace {
            let pk = &mpk.per_air[*air_id];
            for cd in pk.preprocessed_data.iter().chain(air.cached_mains.iter()) {
                pcs.push(cd.data.pcs as *const _);
                need_rot.push(vec![pk.vk.params.need_rot as u8]);
                keep.push(cd.data.clone());
            }
        }
        let widths: Vec<u64> = pcs.iter().map(|p| unsafe { swirl_pcs_stacked_width(*p) }).collect();
        let rot_ptrs: Vec<*const u8> = need_rot.iter().map(|v| v.as_ptr()).collect();
        let whir = self.whir_config();
        let n_st = unsafe { swirl_stacked_reduction_proof_words(pcs.as_ptr(), pcs.len()) };
        let n_wh = unsafe { swirl_whir_proof_words(&self.pcs_params(), &whir, pcs.len(), widths.as_ptr()) };
        let (mut st, mut wh) = (vec![0u32; n_st], vec![0u32; n_wh]);
        let r_words: Vec<u32> = r.iter().flat_map(proof_parts::ef_words).collect();
        check(unsafe {
            swirl_prove_openings(self.ctx.0, &mut transcript.state, &whir, pcs.as_ptr(), pcs.len(), rot_ptrs.as_ptr(),
                                 r_words.as_ptr(), r.len(), st.as_mut_ptr(), n_st, wh.as_mut_ptr(), n_wh)
        })?;
        drop(keep);
        drop(common_main_pcs_data); // "owned by the function and may be mutated" (hal.rs:124-127): freed here
        let widths: Vec<usize> = widths.iter().map(|&w| w as usize).collect();
        Ok((proof_parts::split_stacking(&st, &self.params, &widths), proof_parts::split_whir(&wh, &self.params, &widths)))
    }
}

impl ProverDevice<B200Backend, B200Transcript> for B200Device {
    type Error = B200Error;
    type DeviceCtx = Arc<Ctx>;
    fn device_ctx(&self) -> &Arc<Ctx> {
        &self.ctx
    }
}

// ---------------------------------------------------------------------------------------------------------------
// DeviceDataTransporter (hal.rs:141-207; data_transporter.rs:35-106)
// ---------------------------------------------------------------------------------------------------------------
impl DeviceDataTransporter<SC, B200Backend> for B200Device {
    fn transport_pk_to_device(&self, mpk: &MultiStarkProvingKey<SC>) -> DeviceMultiStarkProvingKey<B200Backend> {
        let per_air = mpk
            .per_air
            .iter()
            .map(|pk| DeviceStarkProvingKey {
                air_name: pk.air_name.clone(),
                vk: pk.vk.clone(),
                preprocessed_data: pk.preprocessed_data.as_ref().map(|d| {
                    // a preprocessed trace is committed on its own: transport the matrix and commit it again on the
                    // device (the commitment equals the keygen's, which the backend test-suite asserts)
                    let trace = self.transport_matrix_to_device(&d.trace);
                    let (commitment, data) = self.commit(&[&trace]).expect("preprocessed commit");
                    CommittedTraceData { commitment, trace, data: Arc::new(data) }
                }),
                other_data: B200AirData::new(&pk.vk.symbolic_constraints),
            })
            .collect();
        unsafe { swirl_ctx_synchronize(self.ctx.0) };
        DeviceMultiStarkProvingKey::new(per_air, mpk.trace_height_constraints.clone(), mpk.max_constraint_degree,
                                        mpk.params.clone(), mpk.vk_pre_hash)
    }

    fn transport_matrix_to_device(&self, matrix: &ColMajorMatrix<F>) -> B200Matrix {
        let words = f_words(&matrix.values);
        let buf = self.alloc(words.len().max(1)).expect("device allocation");
        check(unsafe { swirl_memcpy_h2d(self.ctx.0, buf.ptr as *mut _, words.as_ptr() as *const _, words.len() * 4) }).expect("H2D");
        B200Matrix { ctx: self.ctx.clone(), buf, height: matrix.height(), width: matrix.width() }
    }

    fn transport_pcs_data_to_device(&self, pcs_data: &StackedPcsData<F, Digest>) -> B200PcsData {
        // host PcsData (cached trace committed by the CPU backend): re-commit its stacked matrix on the device; the
        // Merkle root is a function of the matrix alone, so the commitment is unchanged
        let m = self.transport_matrix_to_device(&pcs_data.matrix);
        let (root, data) = self.commit(&[&m]).expect("cached commit");
        debug_assert_eq!(root, pcs_data.commit());
        data
    }

    fn transport_matrix_from_device_to_host(&self, matrix: &B200Matrix) -> ColMajorMatrix<F> {
        let mut words = vec![0u32; matrix.height * matrix.width];
        check(unsafe { swirl_memcpy_d2h(self.ctx.0, words.as_mut_ptr() as *mut _, matrix.buf.ptr as *const _, words.len() * 4) }).expect("D2H");
        ColMajorMatrix::new(words.into_iter().map(f_from_word).collect(), matrix.width)
    }
}

// ---------------------------------------------------------------------------------------------------------------
// StarkEngine (engine.rs:40-110)
// ---------------------------------------------------------------------------------------------------------------
pub struct B200Engine {
    device: B200Device,
    config: SC,
}

impl StarkEngine for B200Engine {
    type SC = SC;
    type PB = B200Backend;
    type PD = B200Device;
    type TS = B200Transcript;

    fn new(params: SystemParams) -> Self {
        Self { device: B200Device::new(0, params.clone()).expect("no CUDA device: libswirl_b200 has no CPU fallback"), config: SC::default_from_params(params) }
    }
    fn config(&self) -> &SC {
        &self.config
    }
    fn device(&self) -> &B200Device {
        &self.device
    }
    fn initial_transcript(&self) -> B200Transcript {
        B200Transcript { ctx: Some(self.device.ctx.clone()), ..Default::default() }
    }
    fn prover_from_transcript(&self, transcript: B200Transcript) -> Coordinator<SC, B200Backend, B200Device, B200Transcript> {
        Coordinator::new(B200Backend, self.device.clone(), transcript)
    }
}

#[allow(dead_code)]
fn _assert_air_ctx_is_unused(_: &AirProvingContext<B200Backend>) {}
