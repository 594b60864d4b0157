// One translation unit per compiled jet spec: built with -DTDE_NF=<n> -DTDE_SMASK=<mask>.
#include "jet_forward.cuh"
#include "jet_backward.cuh"
#include "jet_forward_tc.cuh"
#include "jet_backward_tc.cuh"
#include "jet_launch.h"

#define TDE_CAT_(a, b, c, d) a##b##_##d
#define TDE_CAT(a, b, c, d) TDE_CAT_(a, b, c, d)

namespace tde {

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
    auto k = p.W == 256 ? jet_forward_tc_kernel<TDE_NF, TDE_SMASK, 256> : jet_forward_tc_kernel<TDE_NF, TDE_SMASK, 128>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(tc_threads(p.W));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kTcCluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k, p);
}

cudaError_t TDE_CAT(launch_backward_tc_, TDE_NF, _, TDE_SMASK)(const TcBwdParams& p, int grid, int smem, cudaStream_t st) {
    auto k = p.W == 256 ? jet_backward_tc_kernel<TDE_NF, TDE_SMASK, 256> : jet_backward_tc_kernel<TDE_NF, TDE_SMASK, 128>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(tc_threads(p.W));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kTcCluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k, p);
}

}  // namespace tde
