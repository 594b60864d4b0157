// One translation unit per compiled jet spec: built with -DTDE_NF=<n> -DTDE_SMASK=<mask>.
#include "jet_forward.cuh"
#include "jet_backward.cuh"
#include "jet_forward_tc.cuh"
#include "jet_backward_tc.cuh"
#include "jet_pair_tc.cuh"
#include "jet_launch.h"

#define TDE_CAT_(a, b, c, d) a##b##_##d
#define TDE_CAT(a, b, c, d) TDE_CAT_(a, b, c, d)

namespace tde {

// single-CTA tensor-core kernels exist where the merged N = 2 x PTS J MMA and the accumulator sets the scheme needs fit
template <int NF, int SMASK, int W, bool OK = tc_shape_ok(W / 128, tc_points(W, JetSpecT<NF, SMASK>::J) * JetSpecT<NF, SMASK>::J)>
struct TcKernels {
    static void (*fwd())(TcParams) { return jet_forward_tc_kernel<NF, SMASK, W>; }
    static void (*bwd())(TcBwdParams) { return jet_backward_tc_kernel<NF, SMASK, W>; }
};
template <int NF, int SMASK, int W>
struct TcKernels<NF, SMASK, W, false> {
    static void (*fwd())(TcParams) { return nullptr; }
    static void (*bwd())(TcBwdParams) { return nullptr; }
};
// CTA-pair kernels exist for jet specs with 96 J <= 512 TMEM columns only
template <int NF, int SMASK, bool OK = (JetSpecT<NF, SMASK>::J <= 5)>
struct PairKernels {
    static void (*fwd())(TcParams) { return jet_forward_pair_kernel<NF, SMASK>; }
    static void (*bwd())(TcBwdParams) { return jet_backward_pair_kernel<NF, SMASK>; }
};
template <int NF, int SMASK>
struct PairKernels<NF, SMASK, false> {
    static void (*fwd())(TcParams) { return nullptr; }
    static void (*bwd())(TcBwdParams) { return nullptr; }
};

cudaError_t TDE_CAT(launch_forward_, TDE_NF, _, TDE_SMASK)(const JetParams& p, int grid, int smem, cudaStream_t st) {
    auto k = jet_forward_kernel<TDE_NF, TDE_SMASK>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    k<<<grid, kThreads, smem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t TDE_CAT(launch_backward_, TDE_NF, _, TDE_SMASK)(const JetParams& p, int grid, int smem, cudaStream_t st) {
    auto k = jet_backward_kernel<TDE_NF, TDE_SMASK>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    k<<<grid, kThreads, smem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t TDE_CAT(launch_forward_tc_, TDE_NF, _, TDE_SMASK)(const TcParams& p, int grid, int smem, cudaStream_t st) {
    void (*k)(TcParams) = p.W == 512 ? TcKernels<TDE_NF, TDE_SMASK, 512>::fwd()
                          : p.W == 256 ? TcKernels<TDE_NF, TDE_SMASK, 256>::fwd() : TcKernels<TDE_NF, TDE_SMASK, 128>::fwd();
    int threads = tc_threads(p.W, tc_points(p.W, JetSpecT<TDE_NF, TDE_SMASK>::J)), cluster = kTcCluster;
    if (p.pair) {
        k = PairKernels<TDE_NF, TDE_SMASK>::fwd();
        threads = kPairThreads;
        cluster = kPairCluster;
        if (!k) return cudaErrorInvalidValue;
    }
    if (!k) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k, p);
}

cudaError_t TDE_CAT(launch_backward_tc_, TDE_NF, _, TDE_SMASK)(const TcBwdParams& p, int grid, int smem, cudaStream_t st) {
    void (*k)(TcBwdParams) = p.W == 512 ? TcKernels<TDE_NF, TDE_SMASK, 512>::bwd()
                             : p.W == 256 ? TcKernels<TDE_NF, TDE_SMASK, 256>::bwd() : TcKernels<TDE_NF, TDE_SMASK, 128>::bwd();
    int threads = tc_threads(p.W, tc_points(p.W, JetSpecT<TDE_NF, TDE_SMASK>::J)), cluster = kTcCluster;
    if (p.pair) {
        k = PairKernels<TDE_NF, TDE_SMASK>::bwd();
        threads = kPairThreads;
        cluster = kPairCluster;
        if (!k) return cudaErrorInvalidValue;
    }
    if (!k) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k, p);
}

// weight-gradient GEMM (recomputes Z from the stash, so it is compiled per jet spec); J <= 6: its raw ring holds J stages
template <int NF, int SMASK, bool OK = (JetSpecT<NF, SMASK>::J <= 6)>
struct DwKernel {
    static void (*get())(TcDwParams) { return jet_dw_tc_kernel<NF, SMASK>; }
};
template <int NF, int SMASK>
struct DwKernel<NF, SMASK, false> {
    static void (*get())(TcDwParams) { return nullptr; }
};
cudaError_t TDE_CAT(launch_dw_tc_, TDE_NF, _, TDE_SMASK)(const TcDwParams& p, int grid, int smem, cudaStream_t st) {
    void (*k)(TcDwParams) = DwKernel<TDE_NF, TDE_SMASK>::get();
    if (!k) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    k<<<grid, kDwThreads, smem, st>>>(p);
    return cudaGetLastError();
}

// how many clusters of the pair kernels can be resident at once (a cluster of 4 must fit inside one GPC, so it can be fewer
// than SMs / 4); the persistent grid must not exceed it.  Returns 0 when the spec has no pair kernel, -1 on error.
int TDE_CAT(pair_max_clusters_, TDE_NF, _, TDE_SMASK)(int smem) {
    void (*kf)(TcParams) = PairKernels<TDE_NF, TDE_SMASK>::fwd();
    void (*kb)(TcBwdParams) = PairKernels<TDE_NF, TDE_SMASK>::bwd();
    if (!kf || !kb) return 0;
    if (cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(kPairCluster * 64);
    cfg.blockDim = dim3(kPairThreads);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kPairCluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int nf = 0, nb = 0;
    if (cudaOccupancyMaxActiveClusters(&nf, kf, &cfg) != cudaSuccess) return -1;
    if (cudaOccupancyMaxActiveClusters(&nb, kb, &cfg) != cudaSuccess) return -1;
    return nf < nb ? nf : nb;
}

}  // namespace tde
