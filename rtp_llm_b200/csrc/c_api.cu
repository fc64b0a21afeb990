// C-ABI entry points of libb200_decode.so (declared in include/b200_decode_ops.h).
// Host side only: argument checks, tensor-map encoding, launch-shape heuristics, launches, and the recorder that turns a
// sequence of op calls into a decode program (b200_program_*). No torch, no allocation on the op path.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <vector>

#include "aux_kernels.cuh"
#include "decode_program.cuh"
#include "internal.h"
#include "paged_decode_attn.cuh"
#include "peer_allreduce.cuh"
#include "sampling.cuh"
#include "wo_gemm.cuh"

using namespace b200;
using namespace b200_host;

namespace b200_host {

thread_local std::string g_err;
thread_local unsigned long long* g_ftrace_next = nullptr;
static std::atomic<uint64_t> g_launches{0};

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

int launched(const char* what) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(B200_ECUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return B200_OK;
}

int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

int num_sms() {
    static int sms[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 16) return 148;
    if (!sms[dev]) {
        cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
        if (sms[dev] <= 0) sms[dev] = 148;
    }
    return sms[dev];
}

int seg_dispatch_f16(int bpad, int qfmt, const ProgOp* op0, const ProgOp* d_ops, int nops, unsigned* gbar, int grid, bool pdl,
                     unsigned long long* trace, cudaStream_t st, int* grid_out);
int seg_dispatch_bf16(int bpad, int qfmt, const ProgOp* op0, const ProgOp* d_ops, int nops, unsigned* gbar, int grid, bool pdl,
                      unsigned long long* trace, cudaStream_t st, int* grid_out);

int segment_grid(bool bf16, int bpad, int qfmt, int* grid_out) {
    return bf16 ? seg_dispatch_bf16(bpad, qfmt, nullptr, nullptr, 0, nullptr, 0, false, nullptr, nullptr, grid_out)
                : seg_dispatch_f16(bpad, qfmt, nullptr, nullptr, 0, nullptr, 0, false, nullptr, nullptr, grid_out);
}
int launch_segment(bool bf16, int bpad, int qfmt, const ProgOp* op0, const ProgOp* d_ops, int nops, unsigned* gbar, int grid,
                   bool pdl, unsigned long long* trace, cudaStream_t st) {
    return bf16 ? seg_dispatch_bf16(bpad, qfmt, op0, d_ops, nops, gbar, grid, pdl, trace, st, nullptr)
                : seg_dispatch_f16(bpad, qfmt, op0, d_ops, nops, gbar, grid, pdl, trace, st, nullptr);
}

}  // namespace b200_host

namespace {

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
void attn_split(int units, int max_tiles, int* nsplit, int* tiles_per_split, int head_dim = 128) {
    int forced = env_int("B200_ATTN_TILES_PER_SPLIT", 0);
    int best_c = max_tiles;
    double best = 1e30;
    const int sms = num_sms();
    // Model fitted to measurements on B200 (profiles/r01_kernel_bench.txt, tools/tp8_shapes.py; within ~10 % for 32..256
    // (sequence, kv-head) units): a CTA has ~3 us of fixed time, streams its 32 KB tiles at min(60 GB/s -- what its 96 KB ring
    // sustains --, HBM / resident CTAs), CTAs beyond 2*SMs are back-filled (fractional waves), a split adds ~3 us of merge.
    const double slots = 2.0 * sms, hbm = 6.2e12, r_cta = 60e9, fixed = 3e-6, merge = 3e-6, tile_bytes = 32768.0 * head_dim / 128.0;
    const bool cluster_ok = env_int("B200_ATTN_CLUSTER", 1) != 0;
    for (int c = 1; c <= max_tiles; ++c) {
        const int ns = (max_tiles + c - 1) / c;
        if (ns > kAttnMaxSplit) continue;
        const double ctas = (double)units * ns;
        const double waves = std::max(1.0, ctas / slots);   // beyond one wave CTAs are back-filled, not lock-stepped
        const double resident = ctas < slots ? ctas : slots;
        const double rate = std::min(r_cta, hbm / resident);
        // merging the splits: <= 8 splits form a cluster and merge through DSMEM (~1.5 us); more go through the L2 workspace and
        // a last-arriver loop whose cost grows with the split count (measured 26 us for 16 splits x 8 tiles, r02)
        const double merge_cost = ns == 1 ? 0.0 : (ns <= 8 && cluster_ok ? 1.5e-6 : merge + 0.55e-6 * ns);
        const double cost = waves * (fixed + c * tile_bytes / rate) + merge_cost;
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
// one-kernel-per-GEMM path (gemm_cluster.cu): split-K so that one wave of 2*SMs CTAs is filled
void gemm_split(int n_tiles, int k_blocks, int max_split, int* nsplit, int* kb_per_split) {
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
constexpr size_t kGemmSemBytes = 16384;      // [n_tiles] tile semaphores ... [last 64 bytes] grid-barrier words of a stand-alone call

// stream-K plan of the persistent kernel: the n_tiles*k_blocks work list is cut into runs of `per` k-blocks, one per CTA
struct SkPlan {
    int per, used, max_contrib, aligned;
};
int sk_max_contrib(int n_tiles, int k_blocks, int per) {
    int mc = 1;
    for (int t = 0; t < n_tiles; ++t) {
        const long a = (long)t * k_blocks, b = a + k_blocks - 1;
        const int c = (int)(b / per - a / per) + 1;
        if (c > mc) mc = c;
    }
    return mc;
}
// Default plan = UNIFORM split-K: the run length is the smallest divisor of k_blocks that still gives every CTA at most one
// run (or whole tiles when there are more tiles than CTAs), so every CTA owns exactly one segment and the S = k_blocks / per
// contributors of a tile (S a power of two) merge their partials together. B200_SK_STREAMK=1 selects ragged stream-K runs
// (equal k-block counts per CTA, tiles finished by their lowest contributor) for experiments.
SkPlan sk_plan(int n_tiles, int k_blocks, int grid, int bpad, size_t ws_partial_bytes) {
    const long total = (long)n_tiles * k_blocks;
    const size_t slot = (size_t)bpad * kGemmTileN * sizeof(float);
    int want = (int)((total + grid - 1) / grid);
    SkPlan p{};
    if (!env_int("B200_SK_STREAMK", 0)) {
        int per = 0;
        if (want >= k_blocks) {
            per = (want + k_blocks - 1) / k_blocks * k_blocks;          // whole tiles, possibly several per CTA
        } else {
            for (int d = want; d <= k_blocks; ++d) {
                const int S = k_blocks / d;
                if (k_blocks % d == 0 && (S & (S - 1)) == 0 && S <= 16 && (size_t)n_tiles * S * slot <= ws_partial_bytes) {
                    per = d;
                    break;
                }
            }
            if (!per) per = k_blocks;
        }
        p.per = per;
        p.used = (int)((total + per - 1) / per);
        p.max_contrib = per >= k_blocks ? 1 : k_blocks / per;
        p.aligned = 1;
        return p;
    }
    int per = want;
    const int min_run = env_int("B200_SK_MIN_RUN", 2);
    if (per < min_run) per = min_run < k_blocks ? min_run : k_blocks;
    int mc = sk_max_contrib(n_tiles, k_blocks, per);
    if (mc > 1 && (size_t)n_tiles * mc * slot > ws_partial_bytes) {
        per = (per + k_blocks - 1) / k_blocks * k_blocks;   // no room for partials: whole tiles per CTA (always valid)
        mc = 1;
    }
    p.per = per;
    p.used = (int)((total + per - 1) / per);
    p.max_contrib = mc;
    p.aligned = 0;
    return p;
}
size_t sk_ws_bytes(int n_tiles, int k_blocks, int grid, int bpad) {
    SkPlan p = sk_plan(n_tiles, k_blocks, grid, bpad, (size_t)-1);
    size_t need = p.max_contrib > 1 ? (size_t)n_tiles * p.max_contrib * bpad * kGemmTileN * sizeof(float) : 0;
    // the ragged stream-K plan (developer switch) never needs more than 3 + k_blocks / per slots per tile: cover it too
    const long total = (long)n_tiles * k_blocks;
    int per = (int)((total + grid - 1) / grid);
    if (per < 2) per = 2;
    const size_t ragged = (size_t)n_tiles * (size_t)(sk_max_contrib(n_tiles, k_blocks, per < k_blocks ? per : k_blocks)) * bpad * kGemmTileN * sizeof(float);
    return need > ragged ? need : ragged;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ decode program recorder
// A program is recorded from ordinary op calls, like a CUDA-graph capture (the reference captures its decode step with
// cudaStreamBeginCapture in cpp/cuda_graph/cuda_graph_runner.cc): between b200_program_begin and b200_program_end the
// calling thread's op calls are appended to the program instead of being launched. Consecutive ops the persistent kernel
// implements (weight-only GEMMs, norms, rope + append, embedding, block-table conversion) become ONE launch; every other
// op (attention, FP16 GEMMs, argmax, all-reduce) is replayed as its own launch in between.
struct b200_program {
    struct Item {
        bool segment = false;
        int first = 0, count = 0;          // ops[first, first+count)
        int bf16 = -1, bpad = -1, qfmt = -1;
        std::function<int(void*)> fn;      // non-segment: replays the recorded call on a stream
    };
    std::vector<ProgOp> ops;
    std::vector<Item> items;
    ProgOp* d_ops = nullptr;
    unsigned* d_bar = nullptr;
    unsigned long long* trace = nullptr;
    bool finalized = false;
    bool fuse = true;
};

namespace {

thread_local b200_program* g_rec = nullptr;

int rec_segment_op(const ProgOp& op, int bf16, int bpad, int qfmt) {
    b200_program* P = g_rec;
    auto compat = [](int a, int b) { return a < 0 || b < 0 || a == b; };
    if (!P->items.empty()) {
        b200_program::Item& it = P->items.back();
        if (it.segment && compat(it.bf16, bf16) && compat(it.bpad, bpad) && compat(it.qfmt, qfmt)) {
            if (it.bf16 < 0) it.bf16 = bf16;
            if (it.bpad < 0) it.bpad = bpad;
            if (it.qfmt < 0) it.qfmt = qfmt;
            P->ops.push_back(op);
            ++it.count;
            return B200_OK;
        }
    }
    b200_program::Item it;
    it.segment = true;
    it.first = (int)P->ops.size();
    it.count = 1;
    it.bf16 = bf16;
    it.bpad = bpad;
    it.qfmt = qfmt;
    P->ops.push_back(op);
    P->items.push_back(std::move(it));
    return B200_OK;
}

int rec_call(std::function<int(void*)> fn) {
    b200_program::Item it;
    it.fn = std::move(fn);
    g_rec->items.push_back(std::move(it));
    return B200_OK;
}
// developer switch B200_PROGRAM_FUSE_MASK: bit (1 << op type) must be set for an op type to be fused (default: all)
bool recording_fused(int op_type) { return g_rec && g_rec->fuse && ((env_int("B200_PROGRAM_FUSE_MASK", -1) >> op_type) & 1); }

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
    if (recording_fused(kOpBlockTable)) {
        ProgOp op{};
        op.type = kOpBlockTable;
        op.t.page_list = page_list;
        op.t.block_ids = block_ids;
        op.t.batch = batch;
        op.t.max_blocks = max_blocks;
        return rec_segment_op(op, -1, -1, -1);
    }
    if (g_rec) return rec_call([=](void* st) { return b200_convert_block_table(page_list, block_ids, batch, max_blocks, st); });
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
    if (g_rec)
        return rec_call([=](void* st) {
            return b200_paged_attn_plan(input_lengths, sequence_lengths, prefix_lengths, block_ids, batch, max_blocks,
                                        tokens_per_block, paged_kv_last_page_len, decode_page_indptr, page_indice, batch_indice,
                                        positions, st);
        });
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
    size_t o = batch * head_num * ns * 256 * sizeof(float);      // sized for the largest head_dim (256)
    return sem + ml + o;
}

static int attn_impl(const void* q, const void* qkv, float rope_base, int q_len, int is_bf16, void* out, size_t head_num, size_t kv_head_num,
                     size_t head_dim, size_t batch, size_t max_blocks_per_seq, size_t max_seq_len, size_t page_size, const void* kv_pool,
                     const int32_t* page_list, const uint32_t* sequence_lengths, float q_scale, void* workspace, size_t workspace_bytes,
                     void* stream);

int b200_paged_decode_attn(const void* q, int is_bf16, void* out, size_t head_num, size_t kv_head_num, size_t head_dim,
                           size_t batch, size_t max_blocks_per_seq, size_t max_seq_len, size_t page_size,
                           const void* kv_pool, const int32_t* page_list, const uint32_t* sequence_lengths,
                           float q_scale, void* workspace, size_t workspace_bytes, void* stream) {
    ARG_CHECK(q || batch == 0, "paged_decode_attn: null pointer");
    return attn_impl(q, nullptr, 0.f, 1, is_bf16, out, head_num, kv_head_num, head_dim, batch, max_blocks_per_seq, max_seq_len, page_size,
                     kv_pool, page_list, sequence_lengths, q_scale, workspace, workspace_bytes, stream);
}

int b200_paged_decode_attn_multi(const void* q, int is_bf16, void* out, size_t head_num, size_t kv_head_num, size_t head_dim,
                                 size_t batch, size_t q_len, size_t max_blocks_per_seq, size_t max_seq_len, size_t page_size,
                                 const void* kv_pool, const int32_t* page_list, const uint32_t* sequence_lengths, float q_scale,
                                 void* workspace, size_t workspace_bytes, void* stream) {
    ARG_CHECK(q || batch == 0, "paged_decode_attn_multi: null pointer");
    ARG_CHECK(q_len >= 1 && q_len <= 16, "paged_decode_attn_multi: q_len %zu unsupported", q_len);
    return attn_impl(q, nullptr, 0.f, (int)q_len, is_bf16, out, head_num, kv_head_num, head_dim, batch, max_blocks_per_seq, max_seq_len,
                     page_size, kv_pool, page_list, sequence_lengths, q_scale, workspace, workspace_bytes, stream);
}

int b200_paged_decode_attn_rope(const void* qkv, int is_bf16, void* out, size_t head_num, size_t kv_head_num, size_t head_dim,
                                size_t batch, size_t max_blocks_per_seq, size_t max_seq_len, size_t page_size, void* kv_pool,
                                const int32_t* page_list, const uint32_t* sequence_lengths, float q_scale, float rope_base,
                                void* workspace, size_t workspace_bytes, void* stream) {
    ARG_CHECK(qkv || batch == 0, "paged_decode_attn_rope: null pointer");
    ARG_CHECK(rope_base > 1.f, "paged_decode_attn_rope: rope_base must be > 1");
    return attn_impl(nullptr, qkv, rope_base, 1, is_bf16, out, head_num, kv_head_num, head_dim, batch, max_blocks_per_seq, max_seq_len,
                     page_size, kv_pool, page_list, sequence_lengths, q_scale, workspace, workspace_bytes, stream);
}

static int attn_impl(const void* q, const void* qkv, float rope_base, int q_len, int is_bf16, void* out, size_t head_num, size_t kv_head_num,
                     size_t head_dim, size_t batch, size_t max_blocks_per_seq, size_t max_seq_len, size_t page_size, const void* kv_pool,
                     const int32_t* page_list, const uint32_t* sequence_lengths, float q_scale, void* workspace, size_t workspace_bytes,
                     void* stream) {
    if (batch == 0) return B200_OK;
    ARG_CHECK((q || qkv) && out && kv_pool && page_list && sequence_lengths, "paged_decode_attn: null pointer");
    ARG_CHECK(head_dim == 64 || head_dim == 128 || head_dim == 256, "paged_decode_attn: head_dim %zu unsupported (64 / 128 / 256)", head_dim);
    ARG_CHECK(!qkv || head_dim == 128, "paged_decode_attn_rope: the fused rope is built for head_dim 128 (got %zu)", head_dim);
    ARG_CHECK(kv_head_num > 0 && head_num % kv_head_num == 0, "paged_decode_attn: head_num %zu not a multiple of kv_head_num %zu",
              head_num, kv_head_num);
    const int group = (int)(head_num / kv_head_num);
    ARG_CHECK(group >= 1 && group <= 16, "paged_decode_attn: group size %d unsupported (1..16)", group);
    ARG_CHECK(group * q_len <= 16, "paged_decode_attn: group %d x q_len %d exceeds the 16 rows of one MMA tile", group, q_len);
    ARG_CHECK(page_size == 16 || page_size == 32 || page_size == 64 || page_size == 128,
              "paged_decode_attn: page_size %zu unsupported (16/32/64/128)", page_size);
    ARG_CHECK(max_seq_len >= 1, "paged_decode_attn: max_seq_len must be >= 1");
    ARG_CHECK(max_seq_len <= max_blocks_per_seq * page_size, "paged_decode_attn: max_seq_len %zu exceeds page table capacity %zu",
              max_seq_len, max_blocks_per_seq * page_size);
    ARG_CHECK(batch * kv_head_num <= 65535, "paged_decode_attn: batch*kv_heads %zu exceeds grid limit", batch * kv_head_num);
    ARG_CHECK(((uintptr_t)kv_pool & 15) == 0 && (((uintptr_t)q | (uintptr_t)qkv) & 3) == 0, "paged_decode_attn: misaligned pointer");
    if (g_rec)
        return rec_call([=](void* st) {
            return attn_impl(q, qkv, rope_base, q_len, is_bf16, out, head_num, kv_head_num, head_dim, batch, max_blocks_per_seq, max_seq_len,
                             page_size, kv_pool, page_list, sequence_lengths, q_scale, workspace, workspace_bytes, st);
        });

    AttnParams p{};
    p.q = q;
    p.qkv = qkv;
    p.kv_pool_rw = const_cast<void*>(kv_pool);
    p.log2_base = qkv ? std::log2(rope_base) : 0.f;
    p.out = out;
    p.page_list = page_list;
    p.seq_lens = reinterpret_cast<const int32_t*>(sequence_lengths);
    p.B = (int)batch;
    p.Hq = (int)head_num;
    p.Hkv = (int)kv_head_num;
    p.group = group;
    p.q_len = q_len;
    p.M = (int)max_blocks_per_seq;
    p.T = (int)page_size;
    p.log2T = page_size == 16 ? 4 : page_size == 32 ? 5 : page_size == 64 ? 6 : 7;
    p.box_h = page_size < (size_t)kAttnTile ? (int)page_size : kAttnTile;
    p.boxes_per_tile = kAttnTile / p.box_h;
    p.scale_log2 = q_scale / std::sqrt((float)head_dim) * 1.4426950408889634f;
    const int max_tiles = (int)((max_seq_len + kAttnTile - 1) / kAttnTile);
    attn_split((int)(batch * kv_head_num), max_tiles, &p.nsplit, &p.tiles_per_split, (int)head_dim);

    // workspace carve-up: [sem][ml][o]
    const size_t sem_b = round_up(batch * kv_head_num * sizeof(int), 256);
    const size_t ml_b = round_up(batch * head_num * q_len * (size_t)p.nsplit * 2 * sizeof(float), 256);
    const size_t o_b = batch * head_num * q_len * (size_t)p.nsplit * head_dim * sizeof(float);
    if (p.nsplit > 8 || (p.nsplit > 1 && !env_int("B200_ATTN_CLUSTER", 1))) {
        ARG_CHECK(workspace && workspace_bytes >= sem_b + ml_b + o_b,
                  "paged_decode_attn: workspace too small (%zu < %zu)", workspace_bytes, sem_b + ml_b + o_b);
        p.sem = reinterpret_cast<int*>(workspace);
        p.ws_ml = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + sem_b);
        p.ws_o = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + sem_b + ml_b);
    }

    CUtensorMap map;
    {
        const cuuint64_t dims[4] = {(cuuint64_t)head_dim, (cuuint64_t)page_size, (cuuint64_t)kv_head_num, (cuuint64_t)1 << 31};
        const cuuint64_t strides[3] = {(cuuint64_t)head_dim * 2, (cuuint64_t)page_size * head_dim * 2,
                                       (cuuint64_t)kv_head_num * page_size * head_dim * 2};
        const cuuint32_t box[4] = {64, (cuuint32_t)p.box_h, 1, 1};
        int rc = make_map(&map, is_bf16 != 0, 4, kv_pool, dims, strides, box);
        if (rc) return rc;
    }
    const dim3 grid(p.nsplit, (unsigned)(batch * kv_head_num), 1);
    p.cluster_merge = (p.nsplit >= 2 && p.nsplit <= 8 && env_int("B200_ATTN_CLUSTER", 1)) ? 1 : 0;
    static bool configured_dev[16][6] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    ARG_CHECK(dev >= 0 && dev < 16, "device ordinal %d out of range", dev);
    auto launch = [&](auto kern, bool& configured, int smem_bytes) -> cudaError_t {
        if (!configured) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
            if (e != cudaSuccess) return e;
            configured = true;
        }
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = grid;
        cfg.blockDim = dim3(kAttnThreads);
        cfg.dynamicSmemBytes = smem_bytes;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[2];
        int na = 0;
        if (g_pdl.load() != 0) {
            attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[na].val.programmaticStreamSerializationAllowed = 1;
            ++na;
        }
        if (p.cluster_merge) {
            attr[na].id = cudaLaunchAttributeClusterDimension;
            attr[na].val.clusterDim.x = (unsigned)p.nsplit;
            attr[na].val.clusterDim.y = 1;
            attr[na].val.clusterDim.z = 1;
            ++na;
        }
        cfg.attrs = attr;
        cfg.numAttrs = na;
        return cudaLaunchKernelEx(&cfg, kern, map, p);
    };
    const int di = head_dim == 64 ? 0 : (head_dim == 128 ? 1 : 2);
    bool& configured = configured_dev[dev][di * 2 + (is_bf16 ? 1 : 0)];
    if (is_bf16) {
        if (di == 0) CUDA_CHECK(launch(paged_decode_attn_kernel<__nv_bfloat16, 64>, configured, attn_smem_bytes(64)));
        else if (di == 1) CUDA_CHECK(launch(paged_decode_attn_kernel<__nv_bfloat16, 128>, configured, attn_smem_bytes(128)));
        else CUDA_CHECK(launch(paged_decode_attn_kernel<__nv_bfloat16, 256>, configured, attn_smem_bytes(256)));
    } else {
        if (di == 0) CUDA_CHECK(launch(paged_decode_attn_kernel<__half, 64>, configured, attn_smem_bytes(64)));
        else if (di == 1) CUDA_CHECK(launch(paged_decode_attn_kernel<__half, 128>, configured, attn_smem_bytes(128)));
        else CUDA_CHECK(launch(paged_decode_attn_kernel<__half, 256>, configured, attn_smem_bytes(256)));
    }
    return launched("paged_decode_attn_kernel");
}

// ------------------------------------------------------------------------------------------------ GEMM
size_t b200_wo_gemm_packed_bytes(int fmt, int K, int N) {
    if (K <= 0 || N <= 0 || K % kGemmBK) return 0;
    const size_t n_tiles = (N + kGemmTileN - 1) / kGemmTileN, kb = K / kGemmBK;
    if (fmt == B200_FMT_INT4) return n_tiles * kb * kW4BlockBytes;
    if (fmt == B200_FMT_INT8) return n_tiles * kb * kW8BlockBytes;
    if (fmt == B200_FMT_INT8G) return n_tiles * kb * kW8GBlockBytes;
    if (fmt != B200_FMT_F16) return 0;
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

int b200_pack_w8g(const int8_t* q, const void* scales, const void* zeros_x_scales, int K, int N, int group, void* blob,
                  void* stream) {
    ARG_CHECK(q && scales && zeros_x_scales && blob, "pack_w8g: null pointer");
    ARG_CHECK(group == kGemmBK, "pack_w8g: group size %d unsupported (128 only)", group);
    ARG_CHECK(K > 0 && K % kGemmBK == 0, "pack_w8g: K=%d must be a positive multiple of 128", K);
    ARG_CHECK(N > 0, "pack_w8g: N must be positive");
    const size_t bytes = b200_wo_gemm_packed_bytes(B200_FMT_INT8G, K, N);
    const int threads = 256;
    const size_t blocks = (bytes + threads - 1) / threads;
    pack_w8g_kernel<<<(unsigned)(blocks < 65535 ? blocks : 65535), threads, 0, (cudaStream_t)stream>>>(
        q, reinterpret_cast<const uint16_t*>(scales), reinterpret_cast<const uint16_t*>(zeros_x_scales), K, N,
        reinterpret_cast<uint8_t*>(blob));
    return launched("pack_w8g_kernel");
}

size_t b200_wo_gemm_workspace_bytes(int max_batch, int N, int K) {
    if (max_batch <= 0 || N <= 0 || K <= 0 || K % kGemmBK) return 0;
    const int n_tiles = (N + kGemmTileN - 1) / kGemmTileN;
    int ns, kbp;
    gemm_split(n_tiles, K / kGemmBK, 16, &ns, &kbp);
    // sized for the largest split either GEMM path may pick for any batch <= max_batch
    const int bpad = gemm_bpad(max_batch);
    const size_t part = ns > 1 ? (size_t)ns * n_tiles * bpad * kGemmTileN * sizeof(float) : 0;
    const size_t sk = sk_ws_bytes(n_tiles, K / kGemmBK, 2 * num_sms(), bpad);
    return kGemmSemBytes + (part > sk ? part : sk);
}

namespace {
struct RsDesc {
    void* regions[kArMaxWorld];
    size_t max_message_bytes;
    int rank, world;
};
int wo_gemm_impl(int fmt, int is_bf16, const void* x, int B, int K, int N, const void* w, const void* col_scale,
                 const void* bias, void* y, void* workspace, size_t workspace_bytes, int flags, const RsDesc* rs, void* stream);
int fill_ar_params(PeerArParams& p, void* const* regions, size_t max_message_bytes, int rank, int world);
}  // namespace

int b200_wo_gemm(int fmt, int is_bf16, const void* x, int B, int K, int N, const void* w, const void* col_scale,
                 const void* bias, void* y, void* workspace, size_t workspace_bytes, int flags, void* stream) {
    return wo_gemm_impl(fmt, is_bf16, x, B, K, N, w, col_scale, bias, y, workspace, workspace_bytes, flags, nullptr, stream);
}

int b200_wo_gemm_rs(int fmt, int is_bf16, const void* x, int B, int K, int N, const void* w, const void* col_scale,
                    const void* bias, void* y, void* workspace, size_t workspace_bytes, int flags, void* const* regions,
                    size_t max_message_bytes, int rank, int world, void* stream) {
    ARG_CHECK(regions, "wo_gemm_rs: null regions");
    ARG_CHECK(world >= 2 && world <= kArMaxWorld && rank >= 0 && rank < world, "wo_gemm_rs: bad rank/world %d/%d", rank, world);
    ARG_CHECK(N % 128 == 0 && N % (8 * world) == 0 && N <= 8192, "wo_gemm_rs: N=%d must be a multiple of 128 and of 8*world, <= 8192", N);
    ARG_CHECK((size_t)B * N * 2 <= max_message_bytes, "wo_gemm_rs: message of %zu bytes exceeds the region", (size_t)B * N * 2);
    ARG_CHECK(!(flags & B200_GEMM_SILU_MUL), "wo_gemm_rs: SILU_MUL cannot be combined with the reduce-scatter push");
    RsDesc d{};
    for (int r = 0; r < world; ++r) {
        ARG_CHECK(regions[r], "wo_gemm_rs: region %d is null", r);
        d.regions[r] = regions[r];
    }
    d.max_message_bytes = max_message_bytes;
    d.rank = rank;
    d.world = world;
    return wo_gemm_impl(fmt, is_bf16, x, B, K, N, w, col_scale, bias, y, workspace, workspace_bytes, flags, &d, stream);
}

namespace {
int wo_gemm_impl(int fmt, int is_bf16, const void* x, int B, int K, int N, const void* w, const void* col_scale,
                 const void* bias, void* y, void* workspace, size_t workspace_bytes, int flags, const RsDesc* rs, void* stream) {
    if (B == 0) return B200_OK;
    ARG_CHECK(fmt == B200_FMT_F16 || fmt == B200_FMT_INT8 || fmt == B200_FMT_INT4 || fmt == B200_FMT_INT8G,
              "wo_gemm: unknown weight format %d", fmt);
    ARG_CHECK(x && w && y, "wo_gemm: null pointer");
    ARG_CHECK(B > 0 && B <= 128, "wo_gemm: batch %d unsupported (1..128 per call)", B);
    ARG_CHECK(K > 0 && K % kGemmBK == 0, "wo_gemm: K=%d must be a positive multiple of 128", K);
    ARG_CHECK(N > 0, "wo_gemm: N must be positive");
    ARG_CHECK(fmt != B200_FMT_INT8 || col_scale, "wo_gemm: INT8 needs col_scale");
    ARG_CHECK((((uintptr_t)x | (uintptr_t)w | (uintptr_t)y) & 15) == 0, "wo_gemm: x / w / y must be 16-byte aligned");
    const int n_tiles = (N + kGemmTileN - 1) / kGemmTileN;
    ARG_CHECK(n_tiles <= 2032, "wo_gemm: N=%d too large", N);   // arrivals [0, 2048) + departures [2048, 4080) + 64 bytes of barrier words
    const int bpad = gemm_bpad(B);
    const bool silu_mul = (flags & B200_GEMM_SILU_MUL) != 0;
    ARG_CHECK(!silu_mul || N % 128 == 0, "wo_gemm: SILU_MUL needs N %% 128 == 0 (gate/up interleaved per 128-feature tile)");
    const bool use_pdl = (flags & B200_GEMM_PDL) || g_pdl.load();

    CUtensorMap xmap;
    {
        const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)B};
        const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
        const cuuint32_t box[2] = {64, (cuuint32_t)bpad};
        int rc = make_map(&xmap, is_bf16 != 0, 2, x, dims, strides, box);
        if (rc) return rc;
    }

    // ---- weight-only INT8 / INT4 at decode batches: inside a decode program the GEMM is an op of the persistent kernel
    // (uniform split-K merged through L2). A stand-alone call uses the cluster kernel below (split-K merged through DSMEM):
    // measured on B200 at B=32 (profiles/r02_gemm_paths.txt) the cluster kernel is faster for every Llama-3-8B shape;
    // B200_GEMM_PERSISTENT=1 routes stand-alone calls through the persistent kernel too (experiments / tests).
    const bool seg_capable = (fmt == B200_FMT_INT8 || fmt == B200_FMT_INT4) && bpad <= 64;   // INT8G / FP16: cluster kernel only
    const bool streamk = seg_capable && !rs && (recording_fused(kOpGemm) || env_int("B200_GEMM_PERSISTENT", 0));
    if (streamk) {
        ARG_CHECK(workspace && workspace_bytes >= kGemmSemBytes, "wo_gemm: workspace of at least %zu bytes required", kGemmSemBytes);
        int grid = 0;
        int rc = segment_grid(is_bf16 != 0, bpad, fmt, &grid);
        if (rc) return rc;
        const SkPlan plan = sk_plan(n_tiles, K / kGemmBK, grid, bpad, workspace_bytes - kGemmSemBytes);
        ProgOp op{};
        op.xmap = xmap;
        op.type = kOpGemm;
        op.fmt = fmt;
        SkGemmParams& g = op.g;
        g.w_blob = reinterpret_cast<const uint8_t*>(w);
        g.col_scale = col_scale;
        g.bias = bias;
        g.y = y;
        g.B = B;
        g.N = N;
        g.K = K;
        g.k_blocks = K / kGemmBK;
        g.n_tiles = n_tiles;
        g.total_kb = n_tiles * g.k_blocks;
        g.per_cta = plan.per;
        g.max_contrib = plan.max_contrib;
        g.aligned = plan.aligned;
        g.sem = reinterpret_cast<int*>(workspace);
        g.ws = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + kGemmSemBytes);
        g.silu_mul = silu_mul ? 1 : 0;
        if (recording_fused(kOpGemm)) return rec_segment_op(op, is_bf16 ? 1 : 0, bpad, fmt);
        if (g_rec)
            return rec_call([=](void* st) {
                return b200_wo_gemm(fmt, is_bf16, x, B, K, N, w, col_scale, bias, y, workspace, workspace_bytes, flags, st);
            });
        unsigned* gbar = reinterpret_cast<unsigned*>(reinterpret_cast<uint8_t*>(workspace) + kGemmSemBytes - 64);
        return launch_segment(is_bf16 != 0, bpad, fmt, &op, nullptr, 1, gbar, plan.used < grid ? plan.used : grid, use_pdl,
                              nullptr, (cudaStream_t)stream);
    }
    if (g_rec) {
        if (rs) {
            const RsDesc d = *rs;
            return rec_call([=](void* st) {
                return wo_gemm_impl(fmt, is_bf16, x, B, K, N, w, col_scale, bias, y, workspace, workspace_bytes, flags, &d, st);
            });
        }
        return rec_call([=](void* st) {
            return b200_wo_gemm(fmt, is_bf16, x, B, K, N, w, col_scale, bias, y, workspace, workspace_bytes, flags, st);
        });
    }

    // ---- FP16 weights / batches above 64: one kernel per GEMM, cluster split-K (gemm_cluster.cu)
    GemmParams p{};
    p.w_blob = reinterpret_cast<const uint8_t*>(w);
    p.col_scale = col_scale;
    p.bias = bias;
    p.y = y;
    p.B = B;
    p.N = N;
    p.K = K;
    p.k_blocks = K / kGemmBK;
    p.use_pdl = use_pdl ? 1 : 0;
    p.silu_mul = silu_mul ? 1 : 0;
    p.dbg = 0;
#ifdef B200_GEMM_DEV
    {
        const char* tr = getenv("B200_GEMM_TRACE_PTR");   // developer timeline buffer (device pointer)
        p.trace = (tr && *tr) ? reinterpret_cast<long long*>(strtoull(tr, nullptr, 0)) : nullptr;
    }
    p.dbg = env_int("B200_GEMM_DBG", 0);
#endif
    p.cluster_reduce = (rs || env_int("B200_GEMM_CLUSTER", 1)) ? 1 : 0;   // split-K merge through DSMEM (cluster <= 8) vs global semaphores

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
    CUtensorMap wmap;
    if (fmt == B200_FMT_F16) {
        const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
        const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
        const cuuint32_t box[2] = {64, (cuuint32_t)kGemmTileN};
        int rc = make_map(&wmap, is_bf16 != 0, 2, w, dims, strides, box);
        if (rc) return rc;
    } else {
        wmap = xmap;  // unused by the kernel
    }
    if (rs) {
        GemmParamsRs prs{};
        static_cast<GemmParams&>(prs) = p;
        PeerArParams ap{};
        int rc = fill_ar_params(ap, rs->regions, rs->max_message_bytes, rs->rank, rs->world);
        if (rc) return rc;
        for (int r = 0; r < rs->world; ++r) prs.rs.region[r] = ap.region[r];
        prs.rs.epoch = ap.epoch;
        prs.rs.src_stride = ap.src_stride;
        prs.rs.parity_stride = ap.parity_stride;
        prs.rs.rank = rs->rank;
        prs.rs.world = rs->world;
        return launch_cluster_gemm_rs(fmt, is_bf16 != 0, bpad, xmap, wmap, prs, n_tiles, (cudaStream_t)stream);
    }
    return launch_cluster_gemm(fmt, is_bf16 != 0, bpad, xmap, wmap, p, n_tiles, (cudaStream_t)stream);
}
}  // namespace

// ------------------------------------------------------------------------------------------------ glue ops
int b200_add_rmsnorm(const void* x, void* residual, const void* gamma, void* y, int is_bf16, int rows, int hidden,
                     float eps, void* stream) {
    if (rows == 0) return B200_OK;
    ARG_CHECK(x && gamma && y, "add_rmsnorm: null pointer");
    ARG_CHECK(hidden > 0 && hidden % 8 == 0 && hidden <= 12288, "add_rmsnorm: hidden=%d must be a multiple of 8, <= 12288", hidden);
    if (recording_fused(kOpNorm) && hidden <= kSegNormMaxHidden) {
        ProgOp op{};
        op.type = kOpNorm;
        op.n.x = x;
        op.n.residual = residual;
        op.n.gamma = gamma;
        op.n.y = y;
        op.n.rows = rows;
        op.n.hidden = hidden;
        op.n.eps = eps;
        return rec_segment_op(op, is_bf16 ? 1 : 0, -1, -1);
    }
    if (g_rec) return rec_call([=](void* st) { return b200_add_rmsnorm(x, residual, gamma, y, is_bf16, rows, hidden, eps, st); });
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

int b200_qk_rmsnorm(void* qkv, const void* q_gamma, const void* k_gamma, const void* q_bias, const void* k_bias, int is_bf16,
                    int rows, int head_num, int kv_head_num, int head_dim, float eps, void* stream) {
    if (rows == 0) return B200_OK;
    ARG_CHECK(qkv && q_gamma && k_gamma, "qk_rmsnorm: null pointer");
    ARG_CHECK((q_bias == nullptr) == (k_bias == nullptr), "qk_rmsnorm: give both biases or none");
    ARG_CHECK(head_dim > 0 && head_dim % 64 == 0, "qk_rmsnorm: head_dim %d must be a multiple of 64 (fused_qk_rmsnorm.cu:104-106)", head_dim);
    ARG_CHECK(rows > 0 && head_num > 0 && kv_head_num > 0, "qk_rmsnorm: bad shape");
    if (g_rec)
        return rec_call([=](void* st) {
            return b200_qk_rmsnorm(qkv, q_gamma, k_gamma, q_bias, k_bias, is_bf16, rows, head_num, kv_head_num, head_dim, eps, st);
        });
    const int units = rows * (head_num + kv_head_num), wpb = 8;
    const dim3 grid((units + wpb - 1) / wpb);
    const bool pdl = g_pdl.load() != 0;
    if (is_bf16)
        CUDA_CHECK(launch_ex(qk_rmsnorm_kernel<__nv_bfloat16>, grid, dim3(wpb * 32), 0, (cudaStream_t)stream, pdl, (__nv_bfloat16*)qkv,
                             (const __nv_bfloat16*)q_gamma, (const __nv_bfloat16*)k_gamma, (const __nv_bfloat16*)q_bias,
                             (const __nv_bfloat16*)k_bias, rows, head_num, kv_head_num, head_dim, eps));
    else
        CUDA_CHECK(launch_ex(qk_rmsnorm_kernel<__half>, grid, dim3(wpb * 32), 0, (cudaStream_t)stream, pdl, (__half*)qkv, (const __half*)q_gamma,
                             (const __half*)k_gamma, (const __half*)q_bias, (const __half*)k_bias, rows, head_num, kv_head_num, head_dim, eps));
    return launched("qk_rmsnorm_kernel");
}

int b200_silu_and_mul(const void* gate_up, void* y, int is_bf16, int rows, int inter, void* stream) {
    if (rows == 0) return B200_OK;
    ARG_CHECK(gate_up && y, "silu_and_mul: null pointer");
    ARG_CHECK(inter > 0 && inter % 2 == 0, "silu_and_mul: inter=%d must be positive and even", inter);
    if (g_rec) return rec_call([=](void* st) { return b200_silu_and_mul(gate_up, y, is_bf16, rows, inter, st); });
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
    const float l2b = std::log2(rope_base);
    if (recording_fused(kOpRope)) {
        ProgOp op{};
        op.type = kOpRope;
        op.r.qkv = qkv;
        op.r.q_out = q_out;
        op.r.kv_pool = kv_pool;
        op.r.page_list = page_list;
        op.r.seq_lens = sequence_lengths;
        op.r.B = batch;
        op.r.head_num = head_num;
        op.r.kv_head_num = kv_head_num;
        op.r.head_dim = head_dim;
        op.r.max_blocks = max_blocks_per_seq;
        op.r.page_size = page_size;
        op.r.log2_base = l2b;
        return rec_segment_op(op, is_bf16 ? 1 : 0, -1, -1);
    }
    if (g_rec)
        return rec_call([=](void* st) {
            return b200_rope_append(qkv, q_out, kv_pool, page_list, sequence_lengths, is_bf16, batch, head_num, kv_head_num,
                                    head_dim, max_blocks_per_seq, page_size, rope_base, st);
        });
    const dim3 grid(batch, head_num + 2 * kv_head_num);
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

int b200_rope_append_ex(const void* qkv, const void* qkv_bias, void* q_out, void* kv_pool, const int32_t* page_list,
                        const int32_t* sequence_lengths, const int32_t* position_ids, const float* cos_sin_cache,
                        int cache_positions, const b200_rope_config* cfg, int use_logn_attn, int is_bf16, int batch, int head_num,
                        int kv_head_num, int head_dim, int max_blocks_per_seq, int page_size, void* stream) {
    if (batch == 0) return B200_OK;
    ARG_CHECK(qkv && q_out && kv_pool && page_list && sequence_lengths && cfg, "rope_append_ex: null pointer");
    ARG_CHECK(head_dim > 0 && head_dim % 2 == 0 && head_dim <= 512, "rope_append_ex: head_dim %d unsupported", head_dim);
    ARG_CHECK(cfg->style == 0 || (cfg->style >= 1 && cfg->style <= 7 && cfg->style != 2), "rope_append_ex: rope style %d unsupported (Glm2 is out of scope)", cfg->style);
    ARG_CHECK(cfg->dim > 0 && cfg->dim % 2 == 0 && cfg->dim <= head_dim, "rope_append_ex: rotary dim %d must be even and <= head_dim", cfg->dim);
    ARG_CHECK(cfg->style == 0 || cfg->base > 1.f, "rope_append_ex: rope base must be > 1");
    ARG_CHECK(!cos_sin_cache || cache_positions > 0, "rope_append_ex: a cos/sin cache needs its number of positions");
    ARG_CHECK(!use_logn_attn || cfg->max_pos > 1, "rope_append_ex: logn attention needs max_pos > 1");
    const b200_rope_config cc = *cfg;
    if (g_rec)
        return rec_call([=](void* st) {
            return b200_rope_append_ex(qkv, qkv_bias, q_out, kv_pool, page_list, sequence_lengths, position_ids, cos_sin_cache,
                                       cache_positions, &cc, use_logn_attn, is_bf16, batch, head_num, kv_head_num, head_dim,
                                       max_blocks_per_seq, page_size, st);
        });
    RopeCfg rc{cc.style, cc.dim, cc.base, cc.scale, cc.factor1, cc.factor2, cc.max_pos, cc.extrapolation_factor, cc.mscale};
    const dim3 grid(batch, head_num + 2 * kv_head_num);
    const bool pdl = g_pdl.load() != 0;
    const float2* cache = reinterpret_cast<const float2*>(cos_sin_cache);
    if (is_bf16)
        CUDA_CHECK(launch_ex(rope_append_ex_kernel<__nv_bfloat16>, grid, dim3(head_dim / 2), 0, (cudaStream_t)stream, pdl,
                             (const __nv_bfloat16*)qkv, (const __nv_bfloat16*)qkv_bias, (__nv_bfloat16*)q_out, (__nv_bfloat16*)kv_pool,
                             page_list, sequence_lengths, position_ids, cache, cache_positions, rc, use_logn_attn, head_num,
                             kv_head_num, head_dim, max_blocks_per_seq, page_size));
    else
        CUDA_CHECK(launch_ex(rope_append_ex_kernel<__half>, grid, dim3(head_dim / 2), 0, (cudaStream_t)stream, pdl, (const __half*)qkv,
                             (const __half*)qkv_bias, (__half*)q_out, (__half*)kv_pool, page_list, sequence_lengths, position_ids,
                             cache, cache_positions, rc, use_logn_attn, head_num, kv_head_num, head_dim, max_blocks_per_seq,
                             page_size));
    return launched("rope_append_ex_kernel");
}

int b200_embedding(const int32_t* ids, const void* table, void* out, int is_bf16, int rows, int hidden, void* stream) {
    (void)is_bf16;
    if (rows == 0) return B200_OK;
    ARG_CHECK(ids && table && out, "embedding: null pointer");
    ARG_CHECK(hidden > 0 && hidden % 8 == 0, "embedding: hidden=%d must be a multiple of 8", hidden);
    if (recording_fused(kOpEmbed)) {
        ProgOp op{};
        op.type = kOpEmbed;
        op.e.ids = ids;
        op.e.table = table;
        op.e.out = out;
        op.e.rows = rows;
        op.e.hidden = hidden;
        return rec_segment_op(op, -1, -1, -1);
    }
    if (g_rec) return rec_call([=](void* st) { return b200_embedding(ids, table, out, is_bf16, rows, hidden, st); });
    embedding_kernel<__half><<<rows, 128, 0, (cudaStream_t)stream>>>(ids, (const __half*)table, (__half*)out, hidden);
    return launched("embedding_kernel");
}

int b200_argmax(const void* logits, int dtype, int rows, int vocab, int32_t* out, void* stream) {
    if (rows == 0) return B200_OK;
    ARG_CHECK(logits && out, "argmax: null pointer");
    ARG_CHECK(vocab > 0, "argmax: vocab must be positive");
    ARG_CHECK(dtype >= 0 && dtype <= 2, "argmax: dtype %d unknown (0 fp16, 1 bf16, 2 fp32)", dtype);
    if (g_rec) return rec_call([=](void* st) { return b200_argmax(logits, dtype, rows, vocab, out, st); });
    if (dtype == 0)
        argmax_kernel<__half><<<rows, 1024, 0, (cudaStream_t)stream>>>((const __half*)logits, vocab, out);
    else if (dtype == 1)
        argmax_kernel<__nv_bfloat16><<<rows, 1024, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)logits, vocab, out);
    else
        argmax_kernel<float><<<rows, 1024, 0, (cudaStream_t)stream>>>((const float*)logits, vocab, out);
    return launched("argmax_kernel");
}

int b200_sample(float* logits, int rows, int vocab, const int32_t* history, const int32_t* hist_len, int hist_stride,
                int32_t* count_ws, const float* temperature, const float* repetition, const float* presence,
                const float* frequency, const int32_t* top_k, const float* top_p, const float* uniform, const uint8_t* process,
                int32_t* token_out, float* token_prob_out, float* probs_out, void* stream) {
    if (rows == 0) return B200_OK;
    ARG_CHECK(logits && top_k && top_p && uniform && token_out, "sample: null pointer");
    ARG_CHECK(rows > 0 && rows <= 65535 && vocab > 0, "sample: bad shape [%d, %d]", rows, vocab);
    ARG_CHECK(!history || (hist_len && hist_stride > 0), "sample: history needs hist_len and a positive stride");
    ARG_CHECK(!(repetition || presence || frequency) || (history && count_ws), "sample: penalties need history and count_ws");
    if (g_rec)
        return rec_call([=](void* st) {
            return b200_sample(logits, rows, vocab, history, hist_len, hist_stride, count_ws, temperature, repetition, presence,
                               frequency, top_k, top_p, uniform, process, token_out, token_prob_out, probs_out, st);
        });
    SampleParams p{};
    p.logits = logits;
    p.probs_out = probs_out;
    p.history = history;
    p.hist_len = hist_len;
    p.count_ws = count_ws;
    p.temperature = temperature;
    p.repetition = repetition;
    p.presence = presence;
    p.frequency = frequency;
    p.top_k = top_k;
    p.top_p = top_p;
    p.uniform = uniform;
    p.process = process;
    p.token_out = token_out;
    p.token_prob_out = token_prob_out;
    p.rows = rows;
    p.vocab = vocab;
    p.hist_stride = hist_stride;
    CUDA_CHECK(launch_ex(sample_kernel, dim3(rows), dim3(kSampleThreads), 0, (cudaStream_t)stream, g_pdl.load() != 0, p));
    return launched("sample_kernel");
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

namespace {
int fill_ar_params(PeerArParams& p, void* const* regions, size_t max_message_bytes, int rank, int world) {
    const size_t src_stride = 2 * round_up(max_message_bytes, 256);          // LL doubles the bytes
    const size_t area_stride = (size_t)kArMaxWorld * src_stride;
    const size_t parity_stride = 2 * area_stride;
    for (int r = 0; r < world; ++r) {
        ARG_CHECK(regions[r], "peer collective: region %d is null", r);
        p.region[r] = reinterpret_cast<uint8_t*>(regions[r]);
    }
    p.epoch = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(regions[rank]) + 2 * parity_stride);
    p.src_stride = src_stride;
    p.area_stride = area_stride;
    p.parity_stride = parity_stride;
    p.rank = rank;
    p.world = world;
    return B200_OK;
}
}  // namespace

int b200_peer_allreduce(const void* in, void* out, size_t bytes, int is_bf16, void* const* regions, size_t max_message_bytes,
                        int call_parity, int rank, int world, void* stream) {
    (void)call_parity;   // since r02 the slot parity comes from the device-side call counter (graph replays of any length are safe)
    ARG_CHECK(in && out && regions, "peer_allreduce: null pointer");
    ARG_CHECK(world >= 2 && world <= kArMaxWorld && rank >= 0 && rank < world, "peer_allreduce: bad rank/world %d/%d", rank, world);
    ARG_CHECK(bytes > 0 && bytes % 16 == 0 && bytes <= max_message_bytes, "peer_allreduce: message of %zu bytes unsupported", bytes);
    ARG_CHECK((((uintptr_t)in | (uintptr_t)out) & 15) == 0, "peer_allreduce: in/out must be 16-byte aligned");
    if (g_rec) {
        std::vector<void*> regs(regions, regions + world);
        return rec_call([=](void* st) {
            return b200_peer_allreduce(in, out, bytes, is_bf16, regs.data(), max_message_bytes, 0, rank, world, st);
        });
    }
    PeerArParams p{};
    p.in = in;
    p.out = out;
    int rc = fill_ar_params(p, regions, max_message_bytes, rank, world);
    if (rc) return rc;
    p.n16 = (int)(bytes / 16);
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

namespace {
int allreduce_norm_impl(const void* in, void* residual, const void* gamma, void* y, int is_bf16, int rows, int hidden,
                        float eps, void* const* regions, size_t max_message_bytes, int rank, int world, int pushed, void* stream);
}
int b200_peer_allreduce_norm(const void* in, void* residual, const void* gamma, void* y, int is_bf16, int rows, int hidden,
                             float eps, void* const* regions, size_t max_message_bytes, int rank, int world, void* stream) {
    return allreduce_norm_impl(in, residual, gamma, y, is_bf16, rows, hidden, eps, regions, max_message_bytes, rank, world, 0, stream);
}
int b200_peer_gather_norm(const void* in, void* residual, const void* gamma, void* y, int is_bf16, int rows, int hidden,
                          float eps, void* const* regions, size_t max_message_bytes, int rank, int world, void* stream) {
    return allreduce_norm_impl(in, residual, gamma, y, is_bf16, rows, hidden, eps, regions, max_message_bytes, rank, world, 1, stream);
}
namespace {
int allreduce_norm_impl(const void* in, void* residual, const void* gamma, void* y, int is_bf16, int rows, int hidden,
                        float eps, void* const* regions, size_t max_message_bytes, int rank, int world, int pushed, void* stream) {
    ARG_CHECK(in && residual && gamma && y && regions, "peer_allreduce_norm: null pointer");
    ARG_CHECK(world >= 2 && world <= kArMaxWorld && rank >= 0 && rank < world, "peer_allreduce_norm: bad rank/world %d/%d", rank, world);
    ARG_CHECK(rows > 0 && rows <= 65535, "peer_allreduce_norm: rows %d unsupported", rows);
    ARG_CHECK(hidden > 0 && hidden % (8 * world) == 0 && hidden <= 8192, "peer_allreduce_norm: hidden %d must be a multiple of 8*world and <= 8192", hidden);
    ARG_CHECK((size_t)rows * hidden * 2 <= max_message_bytes, "peer_allreduce_norm: message of %zu bytes exceeds the region", (size_t)rows * hidden * 2);
    ARG_CHECK((((uintptr_t)in | (uintptr_t)residual | (uintptr_t)gamma | (uintptr_t)y) & 15) == 0, "peer_allreduce_norm: pointers must be 16-byte aligned");
    if (g_rec) {
        std::vector<void*> regs(regions, regions + world);
        return rec_call([=](void* st) {
            return allreduce_norm_impl(in, residual, gamma, y, is_bf16, rows, hidden, eps, regs.data(), max_message_bytes, rank, world, pushed, st);
        });
    }
    PeerArParams p{};
    p.in = in;
    int rc = fill_ar_params(p, regions, max_message_bytes, rank, world);
    if (rc) return rc;
    p.residual = residual;
    p.gamma = gamma;
    p.y = y;
    p.rows = rows;
    p.hidden = hidden;
    p.eps = eps;
    p.pushed = pushed;
    const size_t smem = (size_t)hidden * sizeof(float) + (size_t)(hidden / 8 / world) * 16;
    const bool pdl = g_pdl.load() != 0;
    if (is_bf16)
        CUDA_CHECK(launch_ex(peer_allreduce_norm_kernel<__nv_bfloat16>, dim3(rows), dim3(kArThreads), smem, (cudaStream_t)stream, pdl, p));
    else
        CUDA_CHECK(launch_ex(peer_allreduce_norm_kernel<__half>, dim3(rows), dim3(kArThreads), smem, (cudaStream_t)stream, pdl, p));
    return launched("peer_allreduce_norm_kernel");
}
}  // namespace

int b200_peer_argmax(const void* logits, int is_bf16, int rows, int vocab_local, int vocab_total, int32_t* out,
                     void* const* regions, size_t max_message_bytes, int rank, int world, void* stream) {
    ARG_CHECK(logits && out && regions, "peer_argmax: null pointer");
    ARG_CHECK(world >= 2 && world <= kArMaxWorld && rank >= 0 && rank < world, "peer_argmax: bad rank/world %d/%d", rank, world);
    ARG_CHECK(rows > 0 && (size_t)rows * 16 <= max_message_bytes && vocab_local > 0 && vocab_total > 0, "peer_argmax: bad shape");
    if (g_rec) {
        std::vector<void*> regs(regions, regions + world);
        return rec_call([=](void* st) {
            return b200_peer_argmax(logits, is_bf16, rows, vocab_local, vocab_total, out, regs.data(), max_message_bytes, rank, world, st);
        });
    }
    PeerArParams p{};
    int rc = fill_ar_params(p, regions, max_message_bytes, rank, world);
    if (rc) return rc;
    p.logits = logits;
    p.token_out = out;
    p.rows = rows;
    p.vocab_local = vocab_local;
    p.vocab_total = vocab_total;
    const bool pdl = g_pdl.load() != 0;
    if (is_bf16)
        CUDA_CHECK(launch_ex(peer_argmax_kernel<__nv_bfloat16>, dim3(rows), dim3(kArThreads), 0, (cudaStream_t)stream, pdl, p));
    else
        CUDA_CHECK(launch_ex(peer_argmax_kernel<__half>, dim3(rows), dim3(kArThreads), 0, (cudaStream_t)stream, pdl, p));
    return launched("peer_argmax_kernel");
}

// ------------------------------------------------------------------------------------------------ decode programs
int b200_program_create(b200_program** out) {
    ARG_CHECK(out, "program_create: null pointer");
    *out = new b200_program();
    (*out)->fuse = env_int("B200_PROGRAM_FUSE", 1) != 0;
    return B200_OK;
}

int b200_program_begin(b200_program* p) {
    ARG_CHECK(p && !p->finalized, "program_begin: null or already finalised program");
    ARG_CHECK(!g_rec, "program_begin: this thread is already recording a program");
    g_rec = p;
    return B200_OK;
}

int b200_program_end(b200_program* p) {
    ARG_CHECK(p && g_rec == p, "program_end: this thread is not recording this program");
    g_rec = nullptr;
    if (!p->ops.empty()) {
        CUDA_CHECK(cudaMalloc(&p->d_ops, p->ops.size() * sizeof(ProgOp)));
        CUDA_CHECK(cudaMemcpy(p->d_ops, p->ops.data(), p->ops.size() * sizeof(ProgOp), cudaMemcpyHostToDevice));
    }
    const size_t bar_bytes = (p->items.size() + 1) * 256;
    CUDA_CHECK(cudaMalloc(&p->d_bar, bar_bytes));
    CUDA_CHECK(cudaMemset(p->d_bar, 0, bar_bytes));
    CUDA_CHECK(cudaDeviceSynchronize());
    p->finalized = true;
    return B200_OK;
}

int b200_program_set_trace(b200_program* p, void* device_buffer) {
    ARG_CHECK(p, "program_set_trace: null program");
    p->trace = reinterpret_cast<unsigned long long*>(device_buffer);
    return B200_OK;
}

int b200_program_num_ops(const b200_program* p) { return p ? (int)p->ops.size() : 0; }
int b200_program_num_launches(const b200_program* p) { return p ? (int)p->items.size() : 0; }

int b200_program_launch(b200_program* p, void* stream) {
    ARG_CHECK(p && p->finalized, "program_launch: program not finalised (b200_program_end)");
    ARG_CHECK(!g_rec, "program_launch: cannot launch while recording");
    const bool pdl = g_pdl.load() != 0;
    for (size_t i = 0; i < p->items.size(); ++i) {
        const b200_program::Item& it = p->items[i];
        int rc;
        if (it.segment) {
            const bool bf16 = it.bf16 > 0;
            const int bpad = it.bpad > 0 ? it.bpad : 32, qfmt = it.qfmt > 0 ? it.qfmt : B200_FMT_INT4;
            int grid = 0;
            rc = segment_grid(bf16, bpad, qfmt, &grid);
            if (rc) return rc;
            unsigned long long* tr = p->trace ? p->trace + (size_t)it.first * grid * 2 : nullptr;
            // fine timeline (CTA 0): after the coarse region, [op][kFtSlots][kFtBlocks]
            g_ftrace_next = p->trace ? p->trace + p->ops.size() * (size_t)grid * 2 + (size_t)it.first * kFtSlots * kFtBlocks : nullptr;
            rc = launch_segment(bf16, bpad, qfmt, nullptr, p->d_ops + it.first, it.count, p->d_bar + i * 64, grid, pdl, tr,
                                (cudaStream_t)stream);
            g_ftrace_next = nullptr;
        } else {
            rc = it.fn(stream);
        }
        if (rc) return rc;
    }
    return B200_OK;
}

int b200_program_destroy(b200_program* p) {
    if (!p) return B200_OK;
    if (g_rec == p) g_rec = nullptr;
    if (p->d_ops) cudaFree(p->d_ops);
    if (p->d_bar) cudaFree(p->d_bar);
    delete p;
    return B200_OK;
}

}  // extern "C"
