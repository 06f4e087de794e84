// Internal definition of the opaque `swirl_pcs` handle (include/swirl_b200.h): the stacked PCS
// data the commit phase produces and the opening phases consume.
// Reference: StackedPcsDataGpu, cuda-backend/src/stacked_pcs.rs:30-46 (layout + stacked matrix +
// MerkleTreeGpu with its backing codeword and digest layers).
#pragma once
#include <vector>

#include "common.cuh"

namespace swirl {

struct LayoutCol {
    uint64_t mat_idx, col_in_mat, col_idx, row_idx;
    int log_height;
};

struct Layout {
    int l_skip = 0;
    uint64_t height = 0, width = 0;
    std::vector<LayoutCol> cols;
};

// StackedLayout::new (prover/stacked_pcs.rs:144-203); `sorted` = (width, log_height) by descending height
int make_layout(int l_skip, int log_stacked_height, size_t n, const uint64_t* widths, const int32_t* log_heights,
                Layout* out);

}  // namespace swirl

struct swirl_pcs {
    swirl_pcs_params params{};
    swirl::Layout layout;
    uint64_t codeword_height = 0, query_stride = 0;
    const uint32_t* stacked = nullptr;  // device; owned iff owns_stacked
    bool owns_stacked = false;
    uint32_t* codeword = nullptr;  // device, owned
    uint32_t* layers = nullptr;    // device, owned
    std::vector<uint32_t*> owned_traces;  // device copies made by swirl_commit_host
    // external tree (swirl_pcs_attach_external): codeword and digest layers live elsewhere (other GPUs of a sharded
    // commitment); the WHIR opening asks the callback for the opened rows and Merkle paths of its query indices
    swirl_open_fn open_fn = nullptr;
    void* open_user = nullptr;
    uint32_t ext_root[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

