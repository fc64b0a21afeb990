// C-ABI entry points of libb200_decode.so (declared in include/b200_decode_ops.h).
// Host side only: argument checks, tensor-map encoding, launch-shape heuristics, launches. No torch, no allocation.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#include "../../include/b200_decode_ops.h"
#include "aux_kernels.cuh"
#include "paged_decode_attn.cuh"
#include "peer_allreduce.cuh"
#include "wo_gemm.cuh"

using namespace b200;

namespace {

thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define ARG_CHECK(cond, ...)                                \
    do {                                                    \
        if (!(cond)) return fail(B200_EINVAL, __VA_ARGS__); \
    } while (0)
#define CUDA_CHECK(expr)                                                                               \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess) return fail(B200_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)

int launched(const char* what) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(B200_ECUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return B200_OK;
}

std::atomic<int> g_pdl{0};

// Launch through cudaLaunchKernelEx so a kernel can carry the programmatic-dependent-launch attribute: it may then be
// scheduled while its predecessor drains (every kernel of the decode chain executes griddepcontrol.wait before it
// touches dependent data, so this only overlaps launch latency, prologues and -- for the GEMM -- the weight prefetch).
template <typename... KArgs, typename... Args>
cudaError_t launch_ex(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

int num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

// ---- cuTensorMapEncodeTiled through the runtime (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

int make_map(CUtensorMap* map, bool bf16, int rank, const void* base, const cuuint64_t* dims,
             const cuuint64_t* strides_bytes /*rank-1*/, const cuuint32_t* box) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(B200_ECUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank,
                    const_cast<void*>(base), dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(B200_ECUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return B200_OK;
}

size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- attention launch shape
constexpr int kAttnMaxSplit = 64;
void attn_split(int units, int max_tiles, int* nsplit, int* tiles_per_split) {
    int forced = env_int("B200_ATTN_TILES_PER_SPLIT", 0);
    int best_c = max_tiles;
    double best = 1e30;
    const int sms = num_sms();
    // Model fitted to measurements on B200 (profiles/r01_kernel_bench.txt, tools/tp8_shapes.py; within ~10 % for 32..256
    // (sequence, kv-head) units): a CTA has ~3 us of fixed time, streams its 32 KB tiles at min(60 GB/s -- what its 96 KB ring
    // sustains --, HBM / resident CTAs), CTAs beyond 2*SMs are back-filled (fractional waves), a split adds ~3 us of merge.
    const double slots = 2.0 * sms, hbm = 6.2e12, r_cta = 60e9, fixed = 3e-6, merge = 3e-6, tile_bytes = 32768.0;
    for (int c = 1; c <= max_tiles; ++c) {
        const int ns = (max_tiles + c - 1) / c;
        if (ns > kAttnMaxSplit) continue;
        const double ctas = (double)units * ns;
        const double waves = std::max(1.0, ctas / slots);   // beyond one wave CTAs are back-filled, not lock-stepped
        const double resident = ctas < slots ? ctas : slots;
        const double rate = std::min(r_cta, hbm / resident);
        const double cost = waves * (fixed + c * tile_bytes / rate) + (ns > 1 ? merge : 0.0);
        if (cost < best * (1.0 - 1e-6) || (std::fabs(cost - best) <= best * 1e-6 && c > best_c)) {
            best = cost;
            best_c = c;
        }
    }
    if (forced > 0) best_c = forced < max_tiles ? forced : max_tiles;
    int ns = (max_tiles + best_c - 1) / best_c;
    if (ns > kAttnMaxSplit) {
        ns = kAttnMaxSplit;
        best_c = (max_tiles + ns - 1) / ns;
        ns = (max_tiles + best_c - 1) / best_c;
    }
    *nsplit = ns;
    *tiles_per_split = best_c;
}

// ---- GEMM launch shape
int gemm_bpad(int B) { return B <= 16 ? 16 : (B <= 32 ? 32 : (B <= 64 ? 64 : 128)); }
void gemm_split(int n_tiles, int k_blocks, int max_split, int* nsplit, int* kb_per_split) {
    // Measured (profiles/r01_gemm_trace.txt): a CTA spends ~0.4 us per 128-deep k-block plus ~1.5-2 us of fixed time, two
    // CTAs are resident per SM, and the split-K merge costs ~1 us. So: fill one wave of 2*SMs slots, then shorten chains.
    const int forced = env_int("B200_GEMM_SPLITK", 0);
    const long slots = 2L * num_sms();
    int best_s = 1;
    double best = 1e30;
    const int smax = k_blocks < max_split ? k_blocks : max_split;
    const bool cluster = max_split <= 8;   // cluster mode: the nsplit CTAs of a tile must be co-scheduled inside one GPC
    for (int s = 1; s <= smax; ++s) {
        if (s > 1 && n_tiles >= slots) break;   // already more than a wave of tiles: splitting only adds merge traffic
        if (cluster && s > 1 && (s & (s - 1))) continue;                       // powers of two pack GPCs without strays
        if (cluster && s > 1 && n_tiles > 8 * (32 / s)) continue;              // >= 16 SMs (32 slots) per GPC, 8 GPCs
        const int kbp = (k_blocks + s - 1) / s;
        const int se = (k_blocks + kbp - 1) / kbp;
        if (se != s) continue;
        const long ctas = (long)n_tiles * se;
        const long waves = (ctas + slots - 1) / slots;
        const double cost = (double)waves * (kbp + 4.0) + (se > 1 ? 2.0 : 0.0);
        if (cost < best - 1e-9) {
            best = cost;
            best_s = se;
        }
    }
    if (forced > 0) best_s = forced < smax ? forced : smax;
    int kbp = (k_blocks + best_s - 1) / best_s;
    *kb_per_split = kbp;
    *nsplit = (k_blocks + kbp - 1) / kbp;
}
constexpr size_t kGemmSemBytes = 16384;

template <int FMT, typename T, int BPAD, int VAR>
int launch_gemm_var(const CUtensorMap& xmap, const CUtensorMap& wmap, const GemmParams& p, int n_tiles, cudaStream_t st) {
    auto kern = wo_gemm_kernel<FMT, T, BPAD, VAR>;
    constexpr int smem = gemm_smem_bytes(FMT, BPAD, VAR);
    static bool configured = false;
    if (!configured) {
        CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(n_tiles, p.nsplit, 1);
    cfg.blockDim = dim3(gemm_threads(VAR), 1, 1);
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

template <int FMT, typename T, int BPAD>
int launch_gemm(const CUtensorMap& xmap, const CUtensorMap& wmap, const GemmParams& p, int n_tiles, cudaStream_t st) {
    return launch_gemm_var<FMT, T, BPAD, 0>(xmap, wmap, p, n_tiles, st);
}

template <int FMT, typename T>
int dispatch_gemm_bpad(int bpad, const CUtensorMap& xmap, const CUtensorMap& wmap, const GemmParams& p, int n_tiles,
                       cudaStream_t st) {
    switch (bpad) {
        case 16: return launch_gemm<FMT, T, 16>(xmap, wmap, p, n_tiles, st);
        case 32: return launch_gemm<FMT, T, 32>(xmap, wmap, p, n_tiles, st);
        case 64: return launch_gemm<FMT, T, 64>(xmap, wmap, p, n_tiles, st);
        default: return launch_gemm<FMT, T, 128>(xmap, wmap, p, n_tiles, st);
    }
}
template <typename T>
int dispatch_gemm_fmt(int fmt, int bpad, const CUtensorMap& xmap, const CUtensorMap& wmap, const GemmParams& p,
                      int n_tiles, cudaStream_t st) {
    switch (fmt) {
        case B200_FMT_F16: return dispatch_gemm_bpad<kFmtF16, T>(bpad, xmap, wmap, p, n_tiles, st);
        case B200_FMT_INT8: return dispatch_gemm_bpad<kFmtInt8, T>(bpad, xmap, wmap, p, n_tiles, st);
        default: return dispatch_gemm_bpad<kFmtInt4, T>(bpad, xmap, wmap, p, n_tiles, st);
    }
}

}  // namespace

extern "C" {

const char* b200_last_error(void) { return g_err.c_str(); }

int b200_set_pdl(int enable) {
    g_pdl.store(enable ? 1 : 0);
    return B200_OK;
}

uint64_t b200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int b200_device_check(int device) {
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return fail(B200_EUNSUPPORTED, "no CUDA device %d: %s", device, cudaGetErrorString(e));
    if (prop.major != 10) return fail(B200_EUNSUPPORTED, "device %d is sm_%d%d, this library is sm_100a only", device,
                                      prop.major, prop.minor);
    if (!encode_fn()) return fail(B200_EUNSUPPORTED, "driver lacks cuTensorMapEncodeTiled");
    return B200_OK;
}

// ------------------------------------------------------------------------------------------------ launch-shape introspection
int b200_plan_attn_split(int units, int max_seq_len, int* nsplit, int* tiles_per_split) {
    ARG_CHECK(units > 0 && max_seq_len > 0 && nsplit && tiles_per_split, "plan_attn_split: bad argument");
    attn_split(units, (max_seq_len + kAttnTile - 1) / kAttnTile, nsplit, tiles_per_split);
    return B200_OK;
}

int b200_plan_gemm_split(int K, int N, int* nsplit, int* k_blocks_per_split) {
    ARG_CHECK(K > 0 && K % kGemmBK == 0 && N > 0 && nsplit && k_blocks_per_split, "plan_gemm_split: bad argument");
    gemm_split((N + kGemmTileN - 1) / kGemmTileN, K / kGemmBK, env_int("B200_GEMM_CLUSTER", 1) ? 8 : 16, nsplit, k_blocks_per_split);
    return B200_OK;
}

// ------------------------------------------------------------------------------------------------ indexing
int b200_convert_block_table(int32_t* page_list, const int32_t* block_ids, int batch, int max_blocks, void* stream) {
    ARG_CHECK(page_list && block_ids, "convert_block_table: null pointer");
    ARG_CHECK(batch >= 0 && max_blocks >= 0, "convert_block_table: negative size");
    if (batch == 0 || max_blocks == 0) return B200_OK;
    const int total = batch * max_blocks;
    const int threads = 256, blocks = (total + threads - 1) / threads;
    convert_block_table_kernel<<<blocks < 1024 ? blocks : 1024, threads, 0, (cudaStream_t)stream>>>(page_list, block_ids,
                                                                                                  batch, max_blocks);
    return launched("convert_block_table_kernel");
}

int b200_paged_attn_plan(const int32_t* input_lengths, const int32_t* sequence_lengths, const int32_t* prefix_lengths,
                         const int32_t* block_ids, int batch, int max_blocks, int tokens_per_block,
                         int32_t* paged_kv_last_page_len, int32_t* decode_page_indptr, int32_t* page_indice,
                         int32_t* batch_indice, int32_t* positions, void* stream) {
    if (batch == 0) return B200_OK;
    ARG_CHECK(batch > 0 && batch <= 1024, "paged_attn_plan: batch %d exceeds the single-CTA limit 1024", batch);
    ARG_CHECK(prefix_lengths || sequence_lengths, "paged_attn_plan: need either prefix_lengths or sequence_lengths");
    ARG_CHECK(!prefix_lengths || input_lengths, "paged_attn_plan: prefill mode needs input_lengths");
    ARG_CHECK(tokens_per_block > 0, "paged_attn_plan: tokens_per_block must be positive");
    ARG_CHECK(paged_kv_last_page_len && decode_page_indptr && batch_indice && positions, "paged_attn_plan: null output");
    ARG_CHECK(!block_ids || page_indice, "paged_attn_plan: block_ids given but page_indice is null");
    const int threads = (batch + 31) / 32 * 32;
    paged_attn_plan_kernel<<<1, threads, 0, (cudaStream_t)stream>>>(input_lengths, sequence_lengths, prefix_lengths,
                                                                    block_ids, batch, max_blocks, tokens_per_block,
                                                                    paged_kv_last_page_len, decode_page_indptr,
                                                                    page_indice, batch_indice, positions);
    return launched("paged_attn_plan_kernel");
}

// ------------------------------------------------------------------------------------------------ attention
size_t b200_paged_decode_attn_workspace_bytes(size_t batch, size_t head_num, size_t kv_head_num, size_t max_seq_len) {
    size_t tiles = (max_seq_len + kAttnTile - 1) / kAttnTile;
    size_t ns = tiles < (size_t)kAttnMaxSplit ? (tiles ? tiles : 1) : (size_t)kAttnMaxSplit;
    size_t sem = round_up(batch * kv_head_num * sizeof(int), 256);
    size_t ml = round_up(batch * head_num * ns * 2 * sizeof(float), 256);
    size_t o = batch * head_num * ns * kAttnD * sizeof(float);
    return sem + ml + o;
}

int b200_paged_decode_attn(const void* q, int is_bf16, void* out, size_t head_num, size_t kv_head_num, size_t head_dim,
                           size_t batch, size_t max_blocks_per_seq, size_t max_seq_len, size_t page_size,
                           const void* kv_pool, const int32_t* page_list, const uint32_t* sequence_lengths,
                           float q_scale, void* workspace, size_t workspace_bytes, void* stream) {
    if (batch == 0) return B200_OK;
    ARG_CHECK(q && out && kv_pool && page_list && sequence_lengths, "paged_decode_attn: null pointer");
    ARG_CHECK(head_dim == (size_t)kAttnD, "paged_decode_attn: head_dim %zu unsupported (128 only)", head_dim);
    ARG_CHECK(kv_head_num > 0 && head_num % kv_head_num == 0, "paged_decode_attn: head_num %zu not a multiple of kv_head_num %zu",
              head_num, kv_head_num);
    const int group = (int)(head_num / kv_head_num);
    ARG_CHECK(group >= 1 && group <= 16, "paged_decode_attn: group size %d unsupported (1..16)", group);
    ARG_CHECK(page_size == 16 || page_size == 32 || page_size == 64 || page_size == 128,
              "paged_decode_attn: page_size %zu unsupported (16/32/64/128)", page_size);
    ARG_CHECK(max_seq_len >= 1, "paged_decode_attn: max_seq_len must be >= 1");
    ARG_CHECK(max_seq_len <= max_blocks_per_seq * page_size, "paged_decode_attn: max_seq_len %zu exceeds page table capacity %zu",
              max_seq_len, max_blocks_per_seq * page_size);
    ARG_CHECK(batch * kv_head_num <= 65535, "paged_decode_attn: batch*kv_heads %zu exceeds grid limit", batch * kv_head_num);
    ARG_CHECK(((uintptr_t)kv_pool & 15) == 0 && ((uintptr_t)q & 3) == 0, "paged_decode_attn: misaligned pointer");

    AttnParams p{};
    p.q = q;
    p.out = out;
    p.page_list = page_list;
    p.seq_lens = reinterpret_cast<const int32_t*>(sequence_lengths);
    p.B = (int)batch;
    p.Hq = (int)head_num;
    p.Hkv = (int)kv_head_num;
    p.group = group;
    p.M = (int)max_blocks_per_seq;
    p.T = (int)page_size;
    p.log2T = page_size == 16 ? 4 : page_size == 32 ? 5 : page_size == 64 ? 6 : 7;
    p.box_h = page_size < (size_t)kAttnTile ? (int)page_size : kAttnTile;
    p.boxes_per_tile = kAttnTile / p.box_h;
    p.scale_log2 = q_scale / std::sqrt((float)head_dim) * 1.4426950408889634f;
    const int max_tiles = (int)((max_seq_len + kAttnTile - 1) / kAttnTile);
    attn_split((int)(batch * kv_head_num), max_tiles, &p.nsplit, &p.tiles_per_split);

    // workspace carve-up: [sem][ml][o]
    const size_t sem_b = round_up(batch * kv_head_num * sizeof(int), 256);
    const size_t ml_b = round_up(batch * head_num * (size_t)p.nsplit * 2 * sizeof(float), 256);
    const size_t o_b = batch * head_num * (size_t)p.nsplit * kAttnD * sizeof(float);
    if (p.nsplit > 1) {
        ARG_CHECK(workspace && workspace_bytes >= sem_b + ml_b + o_b,
                  "paged_decode_attn: workspace too small (%zu < %zu)", workspace_bytes, sem_b + ml_b + o_b);
        p.sem = reinterpret_cast<int*>(workspace);
        p.ws_ml = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + sem_b);
        p.ws_o = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + sem_b + ml_b);
    }

    CUtensorMap map;
    {
        const cuuint64_t dims[4] = {(cuuint64_t)kAttnD, (cuuint64_t)page_size, (cuuint64_t)kv_head_num, (cuuint64_t)1 << 31};
        const cuuint64_t strides[3] = {(cuuint64_t)kAttnD * 2, (cuuint64_t)page_size * kAttnD * 2,
                                       (cuuint64_t)kv_head_num * page_size * kAttnD * 2};
        const cuuint32_t box[4] = {64, (cuuint32_t)p.box_h, 1, 1};
        int rc = make_map(&map, is_bf16 != 0, 4, kv_pool, dims, strides, box);
        if (rc) return rc;
    }
    const dim3 grid(p.nsplit, (unsigned)(batch * kv_head_num), 1);
    static bool configured[2] = {false, false};
    if (is_bf16) {
        if (!configured[1]) {
            CUDA_CHECK(cudaFuncSetAttribute(paged_decode_attn_kernel<__nv_bfloat16>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmemBytes));
            configured[1] = true;
        }
        CUDA_CHECK(launch_ex(paged_decode_attn_kernel<__nv_bfloat16>, grid, dim3(kAttnThreads), kAttnSmemBytes,
                             (cudaStream_t)stream, g_pdl.load() != 0, map, p));
    } else {
        if (!configured[0]) {
            CUDA_CHECK(cudaFuncSetAttribute(paged_decode_attn_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            kAttnSmemBytes));
            configured[0] = true;
        }
        CUDA_CHECK(launch_ex(paged_decode_attn_kernel<__half>, grid, dim3(kAttnThreads), kAttnSmemBytes,
                             (cudaStream_t)stream, g_pdl.load() != 0, map, p));
    }
    return launched("paged_decode_attn_kernel");
}

// ------------------------------------------------------------------------------------------------ GEMM
size_t b200_wo_gemm_packed_bytes(int fmt, int K, int N) {
    if (K <= 0 || N <= 0 || K % kGemmBK) return 0;
    const size_t n_tiles = (N + kGemmTileN - 1) / kGemmTileN, kb = K / kGemmBK;
    if (fmt == B200_FMT_INT4) return n_tiles * kb * kW4BlockBytes;
    if (fmt == B200_FMT_INT8) return n_tiles * kb * kW8BlockBytes;
    return (size_t)N * K * 2;
}

int b200_pack_w4(const uint8_t* q_packed, const void* scales, const void* zeros_x_scales, int K, int N, int group,
                 void* blob, void* stream) {
    ARG_CHECK(q_packed && scales && zeros_x_scales && blob, "pack_w4: null pointer");
    ARG_CHECK(group == kGemmBK, "pack_w4: group size %d unsupported (128 only)", group);
    ARG_CHECK(K > 0 && K % kGemmBK == 0, "pack_w4: K=%d must be a positive multiple of 128", K);
    ARG_CHECK(N > 0 && N % 2 == 0, "pack_w4: N=%d must be positive and even", N);
    const size_t words = b200_wo_gemm_packed_bytes(B200_FMT_INT4, K, N) / 4;
    const int threads = 256;
    const size_t blocks = (words + threads - 1) / threads;
    pack_w4_kernel<<<(unsigned)(blocks < 65535 ? blocks : 65535), threads, 0, (cudaStream_t)stream>>>(
        q_packed, reinterpret_cast<const uint16_t*>(scales), reinterpret_cast<const uint16_t*>(zeros_x_scales), K, N,
        reinterpret_cast<uint8_t*>(blob));
    return launched("pack_w4_kernel");
}

int b200_pack_w8(const int8_t* q, int K, int N, void* blob, void* stream) {
    ARG_CHECK(q && blob, "pack_w8: null pointer");
    ARG_CHECK(K > 0 && K % kGemmBK == 0, "pack_w8: K=%d must be a positive multiple of 128", K);
    ARG_CHECK(N > 0, "pack_w8: N must be positive");
    const size_t bytes = b200_wo_gemm_packed_bytes(B200_FMT_INT8, K, N);
    const int threads = 256;
    const size_t blocks = (bytes + threads - 1) / threads;
    pack_w8_kernel<<<(unsigned)(blocks < 65535 ? blocks : 65535), threads, 0, (cudaStream_t)stream>>>(
        q, K, N, reinterpret_cast<uint8_t*>(blob));
    return launched("pack_w8_kernel");
}

size_t b200_wo_gemm_workspace_bytes(int max_batch, int N, int K) {
    if (max_batch <= 0 || N <= 0 || K <= 0 || K % kGemmBK) return 0;
    const int n_tiles = (N + kGemmTileN - 1) / kGemmTileN;
    int ns, kbp;
    gemm_split(n_tiles, K / kGemmBK, 16, &ns, &kbp);
    // sized for the largest split the heuristic (or the env override) may pick for any batch <= max_batch
    const size_t part = ns > 1 ? (size_t)ns * n_tiles * gemm_bpad(max_batch) * kGemmTileN * sizeof(float) : 0;
    return kGemmSemBytes + part;
}

int b200_wo_gemm(int fmt, int is_bf16, const void* x, int B, int K, int N, const void* w, const void* col_scale,
                 const void* bias, void* y, void* workspace, size_t workspace_bytes, int flags, void* stream) {
    if (B == 0) return B200_OK;
    ARG_CHECK(fmt == B200_FMT_F16 || fmt == B200_FMT_INT8 || fmt == B200_FMT_INT4, "wo_gemm: unknown weight format %d", fmt);
    ARG_CHECK(x && w && y, "wo_gemm: null pointer");
    ARG_CHECK(B > 0 && B <= 128, "wo_gemm: batch %d unsupported (1..128 per call)", B);
    ARG_CHECK(K > 0 && K % kGemmBK == 0, "wo_gemm: K=%d must be a positive multiple of 128", K);
    ARG_CHECK(N > 0, "wo_gemm: N must be positive");
    ARG_CHECK(fmt != B200_FMT_INT8 || col_scale, "wo_gemm: INT8 needs col_scale");
    ARG_CHECK((((uintptr_t)x | (uintptr_t)w | (uintptr_t)y) & 15) == 0, "wo_gemm: x / w / y must be 16-byte aligned");
    const int n_tiles = (N + kGemmTileN - 1) / kGemmTileN;
    ARG_CHECK((size_t)n_tiles * sizeof(int) <= kGemmSemBytes, "wo_gemm: N=%d too large", N);
    const int bpad = gemm_bpad(B);

    GemmParams p{};
    p.w_blob = reinterpret_cast<const uint8_t*>(w);
    p.col_scale = col_scale;
    p.bias = bias;
    p.y = y;
    p.B = B;
    p.N = N;
    p.K = K;
    p.k_blocks = K / kGemmBK;
    p.use_pdl = ((flags & B200_GEMM_PDL) || g_pdl.load()) ? 1 : 0;
    p.silu_mul = (flags & B200_GEMM_SILU_MUL) ? 1 : 0;
    ARG_CHECK(!p.silu_mul || N % 128 == 0, "wo_gemm: SILU_MUL needs N %% 128 == 0 (gate/up interleaved in 64-feature halves)");
    p.dbg = env_int("B200_GEMM_DBG", 0);
#ifdef B200_GEMM_DEV
    {
        const char* tr = getenv("B200_GEMM_TRACE_PTR");   // developer timeline buffer (device pointer)
        p.trace = (tr && *tr) ? reinterpret_cast<long long*>(strtoull(tr, nullptr, 0)) : nullptr;
    }
    p.dbg = env_int("B200_GEMM_DBG", 0);
#endif
    p.cluster_reduce = env_int("B200_GEMM_CLUSTER", 1) ? 1 : 0;   // split-K merge through DSMEM (cluster <= 8) vs global semaphores
    gemm_split(n_tiles, p.k_blocks, p.cluster_reduce ? 8 : 16, &p.nsplit, &p.kb_per_split);
    if (p.silu_mul && !p.cluster_reduce) {
        p.nsplit = 1;   // the fused activation is implemented for the direct and the cluster-merge epilogues
        p.kb_per_split = p.k_blocks;
    }
    if (p.nsplit > 1 && !p.cluster_reduce) {
        const size_t tile_bytes = (size_t)n_tiles * bpad * kGemmTileN * sizeof(float);
        if (!workspace || workspace_bytes < kGemmSemBytes + tile_bytes * 2) {
            p.nsplit = 1;  // no room for partials: fall back to one CTA per n-tile (still correct, just less parallel)
            p.kb_per_split = p.k_blocks;
        } else {
            const size_t fit = (workspace_bytes - kGemmSemBytes) / tile_bytes;
            if ((size_t)p.nsplit > fit) {
                p.kb_per_split = (p.k_blocks + (int)fit - 1) / (int)fit;
                p.nsplit = (p.k_blocks + p.kb_per_split - 1) / p.kb_per_split;
            }
            p.sem = reinterpret_cast<int*>(workspace);
            p.ws = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + kGemmSemBytes);
        }
    }

    CUtensorMap xmap, wmap;
    {
        const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)B};
        const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
        const cuuint32_t box[2] = {64, (cuuint32_t)bpad};
        int rc = make_map(&xmap, is_bf16 != 0, 2, x, dims, strides, box);
        if (rc) return rc;
    }
    if (fmt == B200_FMT_F16) {
        const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
        const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
        const cuuint32_t box[2] = {64, (cuuint32_t)kGemmTileN};
        int rc = make_map(&wmap, is_bf16 != 0, 2, w, dims, strides, box);
        if (rc) return rc;
    } else {
        wmap = xmap;  // unused by the kernel
    }
    if (is_bf16) return dispatch_gemm_fmt<__nv_bfloat16>(fmt, bpad, xmap, wmap, p, n_tiles, (cudaStream_t)stream);
    return dispatch_gemm_fmt<__half>(fmt, bpad, xmap, wmap, p, n_tiles, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------ glue ops
int b200_add_rmsnorm(const void* x, void* residual, const void* gamma, void* y, int is_bf16, int rows, int hidden,
                     float eps, void* stream) {
    if (rows == 0) return B200_OK;
    ARG_CHECK(x && gamma && y, "add_rmsnorm: null pointer");
    ARG_CHECK(hidden > 0 && hidden % 8 == 0 && hidden <= 12288, "add_rmsnorm: hidden=%d must be a multiple of 8, <= 12288", hidden);
    const int threads = hidden / 8 >= 512 ? 512 : (hidden / 8 >= 256 ? 256 : 128);
    const size_t smem = (size_t)hidden * sizeof(float);
    const bool pdl = g_pdl.load() != 0;
    if (is_bf16)
        CUDA_CHECK(launch_ex(add_rmsnorm_kernel<__nv_bfloat16>, dim3(rows), dim3(threads), smem, (cudaStream_t)stream, pdl,
                             (const __nv_bfloat16*)x, (__nv_bfloat16*)residual, (const __nv_bfloat16*)gamma,
                             (__nv_bfloat16*)y, hidden, eps));
    else
        CUDA_CHECK(launch_ex(add_rmsnorm_kernel<__half>, dim3(rows), dim3(threads), smem, (cudaStream_t)stream, pdl,
                             (const __half*)x, (__half*)residual, (const __half*)gamma, (__half*)y, hidden, eps));
    return launched("add_rmsnorm_kernel");
}

int b200_silu_and_mul(const void* gate_up, void* y, int is_bf16, int rows, int inter, void* stream) {
    if (rows == 0) return B200_OK;
    ARG_CHECK(gate_up && y, "silu_and_mul: null pointer");
    ARG_CHECK(inter > 0 && inter % 2 == 0, "silu_and_mul: inter=%d must be positive and even", inter);
    const size_t total = (size_t)rows * inter / 2;
    const int threads = 256;
    const size_t blocks = (total + threads - 1) / threads;
    const unsigned g = (unsigned)(blocks < 4 * 148 * 8 ? blocks : 4 * 148 * 8);
    const bool pdl = g_pdl.load() != 0;
    if (is_bf16)
        CUDA_CHECK(launch_ex(silu_and_mul_kernel<__nv_bfloat16>, dim3(g), dim3(threads), 0, (cudaStream_t)stream, pdl,
                             (const __nv_bfloat16*)gate_up, (__nv_bfloat16*)y, rows, inter));
    else
        CUDA_CHECK(launch_ex(silu_and_mul_kernel<__half>, dim3(g), dim3(threads), 0, (cudaStream_t)stream, pdl,
                             (const __half*)gate_up, (__half*)y, rows, inter));
    return launched("silu_and_mul_kernel");
}

int b200_rope_append(const void* qkv, void* q_out, void* kv_pool, const int32_t* page_list,
                     const int32_t* sequence_lengths, int is_bf16, int batch, int head_num, int kv_head_num,
                     int head_dim, int max_blocks_per_seq, int page_size, float rope_base, void* stream) {
    if (batch == 0) return B200_OK;
    ARG_CHECK(qkv && q_out && kv_pool && page_list && sequence_lengths, "rope_append: null pointer");
    ARG_CHECK(head_dim > 0 && head_dim % 2 == 0 && head_dim <= 512, "rope_append: head_dim %d unsupported", head_dim);
    ARG_CHECK(rope_base > 1.f, "rope_append: rope_base must be > 1");
    const dim3 grid(batch, head_num + 2 * kv_head_num);
    const float l2b = std::log2(rope_base);
    const bool pdl = g_pdl.load() != 0;
    if (is_bf16)
        CUDA_CHECK(launch_ex(rope_append_kernel<__nv_bfloat16>, grid, dim3(head_dim / 2), 0, (cudaStream_t)stream, pdl,
                             (const __nv_bfloat16*)qkv, (__nv_bfloat16*)q_out, (__nv_bfloat16*)kv_pool, page_list,
                             sequence_lengths, head_num, kv_head_num, head_dim, (int)max_blocks_per_seq, page_size, l2b));
    else
        CUDA_CHECK(launch_ex(rope_append_kernel<__half>, grid, dim3(head_dim / 2), 0, (cudaStream_t)stream, pdl,
                             (const __half*)qkv, (__half*)q_out, (__half*)kv_pool, page_list, sequence_lengths, head_num,
                             kv_head_num, head_dim, (int)max_blocks_per_seq, page_size, l2b));
    return launched("rope_append_kernel");
}

int b200_embedding(const int32_t* ids, const void* table, void* out, int is_bf16, int rows, int hidden, void* stream) {
    (void)is_bf16;
    if (rows == 0) return B200_OK;
    ARG_CHECK(ids && table && out, "embedding: null pointer");
    ARG_CHECK(hidden > 0 && hidden % 8 == 0, "embedding: hidden=%d must be a multiple of 8", hidden);
    embedding_kernel<__half><<<rows, 128, 0, (cudaStream_t)stream>>>(ids, (const __half*)table, (__half*)out, hidden);
    return launched("embedding_kernel");
}

int b200_argmax(const void* logits, int dtype, int rows, int vocab, int32_t* out, void* stream) {
    if (rows == 0) return B200_OK;
    ARG_CHECK(logits && out, "argmax: null pointer");
    ARG_CHECK(vocab > 0, "argmax: vocab must be positive");
    ARG_CHECK(dtype >= 0 && dtype <= 2, "argmax: dtype %d unknown (0 fp16, 1 bf16, 2 fp32)", dtype);
    if (dtype == 0)
        argmax_kernel<__half><<<rows, 1024, 0, (cudaStream_t)stream>>>((const __half*)logits, vocab, out);
    else if (dtype == 1)
        argmax_kernel<__nv_bfloat16><<<rows, 1024, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)logits, vocab, out);
    else
        argmax_kernel<float><<<rows, 1024, 0, (cudaStream_t)stream>>>((const float*)logits, vocab, out);
    return launched("argmax_kernel");
}

// ------------------------------------------------------------------------------------------------ peer all-reduce
size_t b200_peer_ar_region_bytes(size_t max_message_bytes) {
    // [2 parities][2 areas (one-shot / reduce-scatter, all-gather)][W sources][LL slot = 2 x message] + per-CTA epochs
    return 2 * 2 * (size_t)kArMaxWorld * 2 * round_up(max_message_bytes, 256) + kArMaxCtas * sizeof(uint32_t) + 256;
}

int b200_peer_alloc(size_t bytes, void** ptr, void* ipc_handle_out /*64 bytes*/) {
    ARG_CHECK(ptr && ipc_handle_out && bytes > 0, "peer_alloc: bad argument");
    void* d = nullptr;
    CUDA_CHECK(cudaMalloc(&d, bytes));
    CUDA_CHECK(cudaMemset(d, 0, bytes));
    CUDA_CHECK(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    CUDA_CHECK(cudaIpcGetMemHandle(&h, d));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(ipc_handle_out, &h, sizeof(h));
    *ptr = d;
    return B200_OK;
}

int b200_peer_open(const void* ipc_handle /*64 bytes*/, void** ptr) {
    ARG_CHECK(ipc_handle && ptr, "peer_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, sizeof(h));
    void* d = nullptr;
    CUDA_CHECK(cudaIpcOpenMemHandle(&d, h, cudaIpcMemLazyEnablePeerAccess));
    *ptr = d;
    return B200_OK;
}

int b200_peer_allreduce(const void* in, void* out, size_t bytes, int is_bf16, void* const* regions, size_t max_message_bytes,
                        int call_parity, int rank, int world, void* stream) {
    ARG_CHECK(in && out && regions, "peer_allreduce: null pointer");
    ARG_CHECK(world >= 2 && world <= kArMaxWorld && rank >= 0 && rank < world, "peer_allreduce: bad rank/world %d/%d", rank, world);
    ARG_CHECK(bytes > 0 && bytes % 16 == 0 && bytes <= max_message_bytes, "peer_allreduce: message of %zu bytes unsupported", bytes);
    ARG_CHECK((((uintptr_t)in | (uintptr_t)out) & 15) == 0, "peer_allreduce: in/out must be 16-byte aligned");
    const size_t src_stride = 2 * round_up(max_message_bytes, 256);          // LL doubles the bytes
    const size_t area_stride = (size_t)kArMaxWorld * src_stride;
    const size_t parity_stride = 2 * area_stride;
    PeerArParams p{};
    p.in = in;
    p.out = out;
    for (int r = 0; r < world; ++r) {
        uint8_t* base = reinterpret_cast<uint8_t*>(regions[r]);
        ARG_CHECK(base, "peer_allreduce: region %d is null", r);
        p.slots[r] = base + (size_t)(call_parity & 1) * parity_stride;
        p.slots2[r] = p.slots[r] + area_stride;
    }
    p.epoch = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(regions[rank]) + 2 * parity_stride);
    p.src_stride = src_stride;
    p.n16 = (int)(bytes / 16);
    p.rank = rank;
    p.world = world;
    int ctas = (p.n16 + 2 * kArThreads - 1) / (2 * kArThreads);   // ~2 chunks (32 B of payload) per thread
    if (ctas < 1) ctas = 1;
    ctas = env_int("B200_AR_CTAS", ctas);
    if (ctas > kArMaxCtas) ctas = kArMaxCtas;
    const bool pdl = g_pdl.load() != 0;
    // W >= 3: reduce-scatter + all-gather (2 hops, (W-1)/W of the volume twice); W == 2: one hop. Env override for tests.
    const int two_shot = env_int("B200_AR_TWOSHOT", world >= 3 ? 1 : 0) && (p.n16 % world == 0);
    if (two_shot) {
        int c2 = (p.n16 / world + kArThreads - 1) / kArThreads;
        if (c2 < 1) c2 = 1;
        if (c2 > kArMaxCtas) c2 = kArMaxCtas;
        c2 = env_int("B200_AR_CTAS", c2);
        if (c2 > kArMaxCtas) c2 = kArMaxCtas;
        if (is_bf16)
            CUDA_CHECK(launch_ex(peer_allreduce_twoshot_kernel<__nv_bfloat16>, dim3(c2), dim3(kArThreads), 0, (cudaStream_t)stream, pdl, p));
        else
            CUDA_CHECK(launch_ex(peer_allreduce_twoshot_kernel<__half>, dim3(c2), dim3(kArThreads), 0, (cudaStream_t)stream, pdl, p));
        return launched("peer_allreduce_twoshot_kernel");
    }
    if (is_bf16)
        CUDA_CHECK(launch_ex(peer_allreduce_kernel<__nv_bfloat16>, dim3(ctas), dim3(kArThreads), 0, (cudaStream_t)stream, pdl, p));
    else
        CUDA_CHECK(launch_ex(peer_allreduce_kernel<__half>, dim3(ctas), dim3(kArThreads), 0, (cudaStream_t)stream, pdl, p));
    return launched("peer_allreduce_kernel");
}

// ------------------------------------------------------------------------------------------------ GPU-side checkers
int b200_ref_paged_decode_attn(const void* q, int is_bf16, void* out, int head_num, int kv_head_num, int head_dim,
                               int batch, int max_blocks_per_seq, int page_size, const void* kv_pool,
                               const int32_t* page_list, const int32_t* sequence_lengths, float q_scale, void* stream) {
    if (batch == 0) return B200_OK;
    ARG_CHECK(head_dim % 32 == 0 && head_dim <= 256, "ref attn: head_dim %d unsupported", head_dim);
    const dim3 grid(batch, head_num);
    const float scale = q_scale / std::sqrt((float)head_dim);
    if (is_bf16)
        ref_paged_decode_attn_kernel<__nv_bfloat16><<<grid, 32, 0, (cudaStream_t)stream>>>(
            (const __nv_bfloat16*)q, (__nv_bfloat16*)out, (const __nv_bfloat16*)kv_pool, page_list, sequence_lengths,
            head_num, kv_head_num, head_dim, max_blocks_per_seq, page_size, scale);
    else
        ref_paged_decode_attn_kernel<__half><<<grid, 32, 0, (cudaStream_t)stream>>>(
            (const __half*)q, (__half*)out, (const __half*)kv_pool, page_list, sequence_lengths, head_num, kv_head_num,
            head_dim, max_blocks_per_seq, page_size, scale);
    return launched("ref_paged_decode_attn_kernel");
}

int b200_ref_dequant_gemm(int fmt, int is_bf16, const void* x, int B, int K, int N, const void* w, const void* scales,
                          const void* zeros_x_scales, int group, const void* bias, void* y, void* stream) {
    if (B == 0) return B200_OK;
    const dim3 grid((N + 127) / 128, B);
    if (is_bf16)
        ref_dequant_gemm_kernel<__nv_bfloat16><<<grid, 128, 0, (cudaStream_t)stream>>>(
            (const __nv_bfloat16*)x, B, K, N, fmt, w, (const __nv_bfloat16*)scales, (const __nv_bfloat16*)zeros_x_scales,
            group, (const __nv_bfloat16*)bias, (__nv_bfloat16*)y);
    else
        ref_dequant_gemm_kernel<__half><<<grid, 128, 0, (cudaStream_t)stream>>>(
            (const __half*)x, B, K, N, fmt, w, (const __half*)scales, (const __half*)zeros_x_scales, group,
            (const __half*)bias, (__half*)y);
    return launched("ref_dequant_gemm_kernel");
}

}  // extern "C"
