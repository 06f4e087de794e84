//! `ProverBackend` / `ProverDevice` over `libswirl_b200.so` — the shim of SURVEY.md §8(f)-1, modelled on
//! crates/cuda-backend/src/gpu_backend.rs:44-212.  The three phase traits map one to one onto the phase-level entry
//! points of include/swirl_b200.h; proofs come back as flat Montgomery words in the field order of proof.rs and are
//! rebuilt into the reference structs here, so `Proof::encode_to_vec()` and the reference verifier are used unchanged.
//! Untested in this repository (no Rust toolchain in the development image); the same call sequence is exercised from
//! Python in stark-backend_b200/{backend,prover}.py and checked bit for bit against the CPU oracle.
pub mod ffi;

use std::ffi::CStr;

use ffi::*;

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

/// Owns a `swirl_ctx` bound to the caller's non-blocking stream (GpuDeviceCtx, cuda-common/src/stream.rs:132-151).
pub struct B200Device {
    pub ctx: *mut SwirlCtx,
    pub pcs_params: SwirlPcsParams,
}
unsafe impl Send for B200Device {}

impl B200Device {
    pub fn new(device: i32, stream: *mut std::ffi::c_void, pcs_params: SwirlPcsParams) -> Result<Self, B200Error> {
        let mut ctx = std::ptr::null_mut();
        check(unsafe { swirl_ctx_create_on_stream(device, stream, &mut ctx) })?;
        Ok(Self { ctx, pcs_params })
    }

    /// TraceCommitter::commit (hal.rs:84-87): borrows the device matrices, returns the commitment and the PcsData handle.
    pub fn commit(&self, traces: &[SwirlMatrix]) -> Result<([u32; 8], B200PcsData), B200Error> {
        let (mut root, mut pcs) = ([0u32; 8], std::ptr::null_mut());
        check(unsafe { swirl_commit(self.ctx, &self.pcs_params, traces.as_ptr(), traces.len(), root.as_mut_ptr(), &mut pcs) })?;
        Ok((root, B200PcsData { ctx: self.ctx, pcs }))
    }

    /// MultiRapProver::prove_rap_constraints (hal.rs:94-112): flat GkrProof + BatchConstraintProof words and the point r.
    pub fn prove_rap_constraints(&self, ts: &mut SwirlTranscript, l_skip: i32, max_constraint_degree: i32, logup_pow_bits: i32,
                                 airs: &[SwirlAirCtx], n_max: usize) -> Result<(Vec<u32>, Vec<u32>), B200Error> {
        let words = unsafe { swirl_batch_constraints_proof_words(l_skip, max_constraint_degree, airs.as_ptr(), airs.len()) };
        let (mut flat, mut r) = (vec![0u32; words], vec![0u32; 4 * (n_max + 1)]);
        check(unsafe {
            swirl_prove_batch_constraints(self.ctx, ts, l_skip, max_constraint_degree, logup_pow_bits, airs.as_ptr(), airs.len(),
                                          flat.as_mut_ptr(), words, r.as_mut_ptr())
        })?;
        Ok((flat, r))
    }

    /// OpeningProver::prove_openings (hal.rs:118-138): flat StackingProof and WhirProof words.
    pub fn prove_openings(&self, ts: &mut SwirlTranscript, whir: &SwirlWhirConfig, pcs: &[*const SwirlPcs], widths: &[u64],
                          need_rot: &[*const u8], r: &[u32]) -> Result<(Vec<u32>, Vec<u32>), B200Error> {
        let n_st = unsafe { swirl_stacked_reduction_proof_words(pcs.as_ptr(), pcs.len()) };
        let n_wh = unsafe { swirl_whir_proof_words(&self.pcs_params, whir, pcs.len(), widths.as_ptr()) };
        let (mut st, mut wh) = (vec![0u32; n_st], vec![0u32; n_wh]);
        check(unsafe {
            swirl_prove_openings(self.ctx, ts, whir, pcs.as_ptr(), pcs.len(), need_rot.as_ptr(), r.as_ptr(), r.len() / 4,
                                 st.as_mut_ptr(), n_st, wh.as_mut_ptr(), n_wh)
        })?;
        Ok((st, wh))
    }
}

impl Drop for B200Device {
    fn drop(&mut self) {
        unsafe { swirl_ctx_destroy(self.ctx) };
    }
}

/// StackedPcsData of this backend (cuda-backend/src/stacked_pcs.rs:30-46): codeword + digest layers stay on the device.
pub struct B200PcsData {
    ctx: *mut SwirlCtx,
    pub pcs: *mut SwirlPcs,
}
impl Drop for B200PcsData {
    fn drop(&mut self) {
        unsafe { swirl_pcs_free(self.ctx, self.pcs) };
    }
}

// impl ProverBackend for B200Backend { type Val = BabyBear; type Challenge = BinomialExtensionField<BabyBear, 4>;
//     type Commitment = [BabyBear; 8]; type Matrix = DeviceMatrix<BabyBear>; type PcsData = B200PcsData; ... }
// impl TraceCommitter / MultiRapProver / OpeningProver for B200Device: the three methods above, with
// `split_gkr_and_batch(&flat, shapes)` etc. rebuilding proof.rs structs from the documented section offsets
// (stark-backend_b200/codec.py is the executable description of those offsets).
