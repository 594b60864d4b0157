// Launcher table of the compiled jet specs (one jet_inst.cu object each).
#pragma once
#include "jet_common.cuh"
#include "jet_forward_tc.cuh"
#include "jet_backward_tc.cuh"
#include "jet_pair_tc.cuh"

// (NF, SMASK) list. SMASK bits enumerate pairs (0,0),(0,1),..,(1,1),.. over the NF first-order channels.
#define TDE_SPEC_LIST(X) \
    X(0, 0)              \
    X(1, 0)              \
    X(1, 1)              \
    X(2, 0)              \
    X(2, 1)              \
    X(2, 5)              \
    X(2, 7)              \
    X(3, 0)              \
    X(3, 9)              \
    X(3, 63)

namespace tde {
typedef cudaError_t (*JetLauncher)(const JetParams&, int grid, int smem, cudaStream_t);
typedef cudaError_t (*TcLauncher)(const TcParams&, int grid, int smem, cudaStream_t);
typedef cudaError_t (*TcBwdLauncher)(const TcBwdParams&, int grid, int smem, cudaStream_t);
typedef cudaError_t (*TcDwLauncher)(const TcDwParams&, int grid, int smem, cudaStream_t);
#define TDE_DECL(nf, sm)                                                                  \
    cudaError_t launch_forward_##nf##_##sm(const JetParams&, int, int, cudaStream_t);     \
    cudaError_t launch_backward_##nf##_##sm(const JetParams&, int, int, cudaStream_t);     \
    cudaError_t launch_forward_tc_##nf##_##sm(const TcParams&, int, int, cudaStream_t);   \
    cudaError_t launch_backward_tc_##nf##_##sm(const TcBwdParams&, int, int, cudaStream_t); \
    cudaError_t launch_dw_tc_##nf##_##sm(const TcDwParams&, int, int, cudaStream_t);        \
    int pair_max_clusters_##nf##_##sm(int smem);
TDE_SPEC_LIST(TDE_DECL)
#undef TDE_DECL
}  // namespace tde
