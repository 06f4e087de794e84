"""Device-memory plan of one `Coordinator.prove` call.

Reference: crates/stark-backend/src/memory_metering.rs (ProvingMemoryCounts / ProvingMemoryConfig / ProvingMemoryEstimate:
a closed-form model of the prover's peak, used to meter segments and to choose cache_rs_code_matrix) and
crates/cuda-backend/src/device.rs:113-121 (the GPU prover's cache switches).  The *shape* of the model is the
reference's; the per-phase formulas are this library's own buffers (csrc/{commit,gkr,batch,stacked_reduction,whir}.cu),
since the layout in HBM is not the reference's:

  main            4 B per main-trace cell: the caller's matrices (+ a stacked copy unless one trace fills the matrix)
  rs_code_matrix  4 * B * H * W when cached, else 0; digest layers 64 * B*H / 2^k always
  commit          scratch while committing: NTT inter-pass scratch (+ 32-column codeword window and 64 B of sponge
                  state per codeword row when the codeword is streamed)
  gkr             LogUp leaves 32 B, stored tree prefix ~32 B, two ping-pong SoA tables 16 + 8 B per leaf
  batch           folded EF matrices: 16 B per main cell / 2^l_skip, + half of it for the ping-pong buffer, + selectors
  stacking        the same for the stacked matrix (q evals) + eq/k_rot tables
  whir            ~88 B per stacked row + the 32-column codeword window when the codeword is recomputed

`choose_cache_rs_code_matrix` is the planner: keep the codeword only if the modelled peak fits the free HBM.
`tests/test_memory_plan.py` checks the model against the arena's measured high-water mark (swirl_ctx_mem_stats).
"""
from dataclasses import dataclass

STREAM_GROUP = 32  # csrc/commit.cu
OVERHEAD = 256 << 20  # fixed scratch (reduction partials, twiddles, descriptors): memory_metering.rs uses the same constant


@dataclass
class ProvingMemoryCounts:
    main_cells_with_rot: int = 0     # sum over AIRs of lifted_height * width, AIRs that open next-row rotations
    main_cells_without_rot: int = 0
    interaction_cells: int = 0       # sum over AIRs of lifted_height * num_interactions (LogUp leaves before padding)

    @property
    def main_cells(self):
        return self.main_cells_with_rot + self.main_cells_without_rot

    @classmethod
    def from_airs(cls, airs, l_skip):
        c = cls()
        for a in airs:
            h = max(a.common_main.height(), 1 << l_skip)
            w = a.common_main.width() + sum(m.width() for m in a.cached_mains) + (a.preprocessed.width() if a.preprocessed is not None else 0)
            if a.need_rot:
                c.main_cells_with_rot += h * w
            else:
                c.main_cells_without_rot += h * w
            c.interaction_cells += h * len(a.interactions)
        return c


@dataclass
class ProvingMemoryConfig:
    l_skip: int
    log_stacked_height: int
    log_blowup: int
    k_whir: int
    cache_rs_code_matrix: bool = True
    stacked_aliases_trace: bool = False  # one trace that fills its stacked columns is used in place
    ntt_scratch_bytes: int = 4 << 30


@dataclass
class ProvingMemoryEstimate:
    total: int
    main: int
    stacked_matrix: int
    rs_code_matrix: int
    digest_layers: int
    commit: int
    gkr: int
    batch_constraint: int
    stacking: int
    whir: int
    secondary_peak: int


def estimate(cfg: ProvingMemoryConfig, counts: ProvingMemoryCounts, include_main=True) -> ProvingMemoryEstimate:
    H = 1 << cfg.log_stacked_height
    B = 1 << cfg.log_blowup
    cells = counts.main_cells
    W = -(-cells // H)
    main = 4 * cells if include_main else 0
    stacked = 0 if cfg.stacked_aliases_trace else 4 * H * W
    codeword = 4 * B * H * W
    layers = 64 * (B * H >> cfg.k_whir)
    ntt = min(cfg.ntt_scratch_bytes, codeword)
    if cfg.cache_rs_code_matrix or W <= STREAM_GROUP:
        commit = ntt
        rs = codeword if cfg.cache_rs_code_matrix else 0
        commit_extra = 0 if cfg.cache_rs_code_matrix else codeword  # a narrow matrix is encoded in one piece, then dropped
    else:
        window = 4 * B * H * STREAM_GROUP
        commit = min(ntt, window) + window + 64 * B * H
        rs, commit_extra = 0, 0
    leaves = counts.interaction_cells
    gkr = 88 * leaves
    ef = 16 * (cells >> cfg.l_skip)  # one EF per cell of the folded matrices
    rot = 16 * (counts.main_cells_with_rot >> cfg.l_skip)
    batch = ef + rot + (ef + rot) // 2 + 12 * cells // max(W, 1) + 32 * leaves  # + selectors; the leaves live until GKR ends
    stacking = 24 * (H * W >> cfg.l_skip) + 64 * (H >> cfg.l_skip)
    whir = 88 * H + (0 if cfg.cache_rs_code_matrix else 4 * B * H * min(W, STREAM_GROUP) + min(ntt, 4 * B * H * min(W, STREAM_GROUP)))
    secondary = max(commit + commit_extra, gkr + 32 * leaves, batch, stacking, whir) + OVERHEAD
    total = main + stacked + rs + layers + secondary
    return ProvingMemoryEstimate(total, main, stacked, rs, layers, commit + commit_extra, gkr, batch, stacking, whir, secondary)


def choose_cache_rs_code_matrix(cfg: ProvingMemoryConfig, counts: ProvingMemoryCounts, free_bytes: int, main_resident=True,
                                headroom=0.9):
    """The planner: cache the codeword iff the modelled peak (traces already resident are not counted again) fits into
    `headroom` of the free HBM; otherwise stream and recompute.  Returns (cache?, estimate used)."""
    for cache in (True, False):
        c = ProvingMemoryConfig(**{**cfg.__dict__, "cache_rs_code_matrix": cache})
        e = estimate(c, counts, include_main=not main_resident)
        if e.total <= headroom * free_bytes or not cache:
            return cache, e
