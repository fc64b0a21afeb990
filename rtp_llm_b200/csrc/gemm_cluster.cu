// One-kernel-per-GEMM path (wo_gemm.cuh: cluster split-K merge through DSMEM). Used for FP16 weights (lm_head, the FP16
// model config) and for batches above 64; weight-only INT8/INT4 decode GEMMs go through the persistent stream-K kernel
// (decode_program.cuh) unless B200_GEMM_CLUSTERED=1.
#include "internal.h"
#include "wo_gemm.cuh"

using namespace b200;
using namespace b200_host;

namespace {

template <int FMT, typename T, int BPAD>
int launch_one(const CUtensorMap& xmap, const CUtensorMap& wmap, const GemmParams& p, int n_tiles, cudaStream_t st) {
    auto kern = wo_gemm_kernel<FMT, T, BPAD, 0>;
    constexpr int smem = gemm_smem_bytes(FMT, BPAD, 0);
    static bool configured[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 16) return fail(B200_EINVAL, "device ordinal %d out of range", dev);
    if (!configured[dev]) {
        CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured[dev] = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(n_tiles, p.nsplit, 1);
    cfg.blockDim = dim3(gemm_threads(0), 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (p.use_pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (p.nsplit > 1 && p.cluster_reduce) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 1;
        attr[na].val.clusterDim.y = (unsigned)p.nsplit;
        attr[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, xmap, wmap, p));
    return launched("wo_gemm_kernel");
}

template <int FMT, typename T>
int by_bpad(int bpad, const CUtensorMap& xmap, const CUtensorMap& wmap, const GemmParams& p, int n_tiles, cudaStream_t st) {
    switch (bpad) {
        case 16: return launch_one<FMT, T, 16>(xmap, wmap, p, n_tiles, st);
        case 32: return launch_one<FMT, T, 32>(xmap, wmap, p, n_tiles, st);
        case 64: return launch_one<FMT, T, 64>(xmap, wmap, p, n_tiles, st);
        default: return launch_one<FMT, T, 128>(xmap, wmap, p, n_tiles, st);
    }
}
template <typename T>
int by_fmt(int fmt, int bpad, const CUtensorMap& xmap, const CUtensorMap& wmap, const GemmParams& p, int n_tiles, cudaStream_t st) {
    switch (fmt) {
        case B200_FMT_F16: return by_bpad<kFmtF16, T>(bpad, xmap, wmap, p, n_tiles, st);
        case B200_FMT_INT8: return by_bpad<kFmtInt8, T>(bpad, xmap, wmap, p, n_tiles, st);
        case B200_FMT_INT8G: return by_bpad<kFmtInt8G, T>(bpad, xmap, wmap, p, n_tiles, st);
        default: return by_bpad<kFmtInt4, T>(bpad, xmap, wmap, p, n_tiles, st);
    }
}

}  // namespace

namespace b200_host {
int launch_cluster_gemm(int fmt, bool bf16, int bpad, const CUtensorMap& xmap, const CUtensorMap& wmap, const GemmParams& p,
                        int n_tiles, cudaStream_t st) {
    if (bf16) return by_fmt<__nv_bfloat16>(fmt, bpad, xmap, wmap, p, n_tiles, st);
    return by_fmt<__half>(fmt, bpad, xmap, wmap, p, n_tiles, st);
}
}  // namespace b200_host
