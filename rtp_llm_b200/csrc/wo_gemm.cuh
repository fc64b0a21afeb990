// Weight-only (INT4 group-128 / INT8 per-column / INT8 group-128 / FP16) x FP16|BF16 GEMM for decode batches, sm_100a.
//
//   Y[b][n] = sum_k X[b][k] * W'[k][n] (+ bias[n]),   W' per /root/reference/rtp_llm/device/device_impl.py:183-222,242-300
//   (SURVEY.md section 8 a8-a10).  Replaces the cutlass fpA_intB "mixed gemm" the reference loader still prepares weights
//   for (device_impl.py:392-479) but whose kernel is absent from the snapshot; plugs in behind LinearBase.forward
//   (/root/reference/rtp_llm/models_py/modules/factory/linear/linear_base.py:81).
//
// B200 design (swap-AB, weights stationary on the MMA M side):
//   * one CTA = 128 output features x a run of 128-deep k-blocks (split-K across blockIdx.y);
//   * two TMA warps: weights were re-laid out at load time into per-(n-tile,k-block) contiguous blobs (packed nibbles +
//                the group's scales and zero*scale), so ONE cp.async.bulk stages a whole block (weight ring); the
//                activation block [B x 128] comes through a 128B-swizzled tensor map, OOB rows zero-filled (activation ring);
//   * dequant warps: dequantise IN REGISTERS (lop3 magic-number int->fp16, exact; hfma2 with the group scale / zero*scale)
//                and store the fp16 operand straight INTO TENSOR MEMORY (tcgen05.st) -- the A operand of
//   * MMA warp : one elected thread issues tcgen05.mma (M=128 features, N=batch pad, K=16) with A from TMEM and
//                B (activations, K-major SW128) from shared memory, fp32 accumulators in TMEM;
//   * epilogue : the dequant warps read the accumulators (tcgen05.ld); the split-K CTAs of a tile form a thread-block cluster
//                and merge their fp32 partials through distributed shared memory in fixed order (deterministic); bias /
//                per-column scale / optional fused SiLU(gate)*up applied; coalesced stores.
//   The FP16-weight variant feeds A from shared memory (TMA tensor map, SW128) with the same pipeline.
#pragma once
#include <type_traits>

#include "ptx.cuh"

namespace b200 {

enum : int { kFmtF16 = 0, kFmtInt8 = 1, kFmtInt4 = 2, kFmtInt8G = 3 };   // Int8G: 8-bit GPTQ/AWQ, group 128 (device_impl.py:256-258)

constexpr int kGemmBK = 128;       // k elements per pipeline stage (= one INT4 quantisation group)
constexpr int kGemmTileN = 128;    // output features per CTA (MMA M)
constexpr int kW4BlockBytes = 8192 + 256 + 256;
constexpr int kW8BlockBytes = 16384;
constexpr int kW8GBlockBytes = 16384 + 256 + 256;   // + the group's scales and zero*scale, like the INT4 blob
constexpr int kW16BlockBytes = 32768;

__host__ __device__ constexpr int gemm_w_bytes(int fmt) {
    return fmt == kFmtInt4 ? kW4BlockBytes : (fmt == kFmtInt8 ? kW8BlockBytes : (fmt == kFmtInt8G ? kW8GBlockBytes : kW16BlockBytes));
}
// Shared-memory rings.  The weight ring (HBM stream) and the activation ring (L2 hits) are SEPARATE: a weight stage is
// released as soon as the dequant warps hold its nibbles in registers, so its lifetime is one TMA latency, not the
// whole dequant -> TMEM -> MMA chain, and the bytes in flight per SM (Little's law: ~44 GB/s * latency) stay high.
// 8 dequant warps = two groups of four taking alternate k-blocks (see the dequant section of the kernel).
__host__ __device__ constexpr int gemm_ndq_warps(int var) { (void)var; return 8; }
__host__ __device__ constexpr int gemm_threads(int var) { return (gemm_ndq_warps(var) + 3) * 32; }
__host__ __device__ constexpr int gemm_x_stage_bytes(int bpad) { return bpad * 256; }
#ifndef B200_GEMM_XS
#define B200_GEMM_XS 4
#endif
__host__ __device__ constexpr int gemm_x_stages(int bpad) { return bpad <= 32 ? B200_GEMM_XS : 3; }
constexpr int kGemmSmemMisc = 1024 /*align*/ + 1024 /*barriers*/;
__host__ __device__ constexpr int gemm_w_stages_for(int fmt, int bpad, int budget) {
    int n = (budget - kGemmSmemMisc - gemm_x_stages(bpad) * gemm_x_stage_bytes(bpad)) / gemm_w_bytes(fmt);
    return n > 12 ? 12 : n;
}
// two CTAs per SM when at least 4 weight stages fit in half an SM's shared memory
__host__ __device__ constexpr int gemm_min_ctas(int fmt, int bpad, int var) {
    (void)var;
    return gemm_w_stages_for(fmt, bpad, 106 * 1024) >= 4 ? 2 : 1;
}
__host__ __device__ constexpr int gemm_w_stages(int fmt, int bpad, int var) {
    return gemm_min_ctas(fmt, bpad, var) == 2 ? gemm_w_stages_for(fmt, bpad, 106 * 1024)
                                              : gemm_w_stages_for(fmt, bpad, 216 * 1024);
}
// Accumulator tiles the 8 k-steps of a block rotate over (the epilogue sums them). Measured (tools/gemm_trace.py): one
// tile is enough -- back-to-back tcgen05.mma into the same accumulator pipeline fine; kept as a knob.
__host__ __device__ constexpr int gemm_nacc(int bpad) { return bpad <= 64 ? 1 : 2; }
__host__ __device__ constexpr int gemm_a_stages(int bpad, int var) {
    (void)var;
    return 3;
}
__host__ __device__ constexpr int gemm_tmem_cols(int fmt, int bpad, int var) {
    int need = bpad * gemm_nacc(bpad) + (fmt == kFmtF16 ? 0 : gemm_a_stages(bpad, var) * 64);
    int c = 32;
    while (c < need) c *= 2;
    return c;
}
__host__ __device__ constexpr int gemm_smem_bytes(int fmt, int bpad, int var) {
    return gemm_x_stages(bpad) * gemm_x_stage_bytes(bpad) + gemm_w_stages(fmt, bpad, var) * gemm_w_bytes(fmt) + kGemmSmemMisc;
}

// Row-parallel GEMM under tensor parallelism (b200_wo_gemm_rs): the epilogue pushes every output element that another rank
// owns straight into that rank's reduce-scatter slot over NVLink, as LL words {two elements, epoch} in the layout
// peer_allreduce_norm_kernel polls (peer_allreduce.cuh: area 0, slot [src = rank][row][chunk of the owner's slice]); the
// columns this rank owns are stored to y as usual. world == 0: off.
struct RsPush {
    uint8_t* region[8];                              // region[r]: base of rank r's peer-visible region
    const uint32_t* epoch;                           // device-side call counter of the communicator (read, not advanced)
    unsigned long long src_stride, parity_stride;
    int rank, world;
};

struct GemmParams {
    const uint8_t* w_blob;   // int4 / int8: pre-tiled blobs [n_tiles][k_blocks][block bytes]; f16: unused (tensor map)
    const void* col_scale;   // int8: per-column scale [N] (T); else null
    const void* bias;        // [N] (T) or null
    void* y;                 // [B][N] (T)
    int B, N, K;
    int k_blocks;            // K / 128
    int kb_per_split, nsplit;
    float* ws;               // [nsplit][n_tiles][bpad][128] fp32 split-K partials
    int* sem;                // [n_tiles], zero on entry / exit
    int use_pdl;
    int silu_mul;            // 1: tile rows 2r / 2r+1 are gate / up feature tile*64+r; y[b][tile*64+r] = silu(g)*u
    int cluster_reduce;      // 1: the nsplit CTAs of a tile form a cluster (1,nsplit,1) and merge through DSMEM
    long long* trace;        // developer timeline (tools/gemm_trace.py); null in production
    int dbg;                 // developer experiments (env B200_GEMM_DBG): 1 no math, 2 no TMEM store, 4 no MMA; 0 in production
};

// The VAR == 1 kernels (gemm_cluster_rs.cu) take the descriptor of the reduce-scatter push after the common parameters; the
// plain kernels keep the smaller struct (a larger parameter block cost them a spilled register).
struct GemmParamsRs : GemmParams {
    RsPush rs;
};
template <int VAR>
struct GemmParamsOf {
    using type = GemmParams;
};
template <>
struct GemmParamsOf<1> {
    using type = GemmParamsRs;
};

// ---- int4 -> fp16/bf16 pairs.  Nibble p of a word holds u = q_s + 8; p<4 <-> k = 2p, p>=4 <-> k = 2(p-4)+1, so each
// (w >> 4i) & 0x000f000f yields the (even k, odd k) pair of TMEM column i directly.
template <typename T>
struct Dequant4;
template <>
struct Dequant4<__half> {
    // out[i] = q_s * s + zs for columns i = 0..3 of one 32-bit word
    static __device__ __forceinline__ void word(uint32_t w, __half2 s2, __half2 zs2, uint32_t* out) {
        const uint32_t MAGIC = 0x64006400u;                 // 1024.0 | 1024.0
        const __half2 k1032 = __half2half2(__ushort_as_half(0x6408));   // 1032
        const __half2 k16th = __half2half2(__ushort_as_half(0x2C00));   // 1/16
        const __half2 k72 = __half2half2(__ushort_as_half(0xD480));     // -72
        const uint32_t w2 = w >> 8;
        uint32_t t0, t1, t2, t3;
        asm("lop3.b32 %0, %1, 0x000f000f, %2, 0xea;" : "=r"(t0) : "r"(w), "r"(MAGIC));
        asm("lop3.b32 %0, %1, 0x00f000f0, %2, 0xea;" : "=r"(t1) : "r"(w), "r"(MAGIC));
        asm("lop3.b32 %0, %1, 0x000f000f, %2, 0xea;" : "=r"(t2) : "r"(w2), "r"(MAGIC));
        asm("lop3.b32 %0, %1, 0x00f000f0, %2, 0xea;" : "=r"(t3) : "r"(w2), "r"(MAGIC));
        __half2 h0 = __hsub2(*reinterpret_cast<__half2*>(&t0), k1032);          // u - 8, exact
        __half2 h1 = __hfma2(*reinterpret_cast<__half2*>(&t1), k16th, k72);     // (1024+16u)/16 - 72 = u - 8, exact
        __half2 h2 = __hsub2(*reinterpret_cast<__half2*>(&t2), k1032);
        __half2 h3 = __hfma2(*reinterpret_cast<__half2*>(&t3), k16th, k72);
        h0 = __hfma2(h0, s2, zs2);
        h1 = __hfma2(h1, s2, zs2);
        h2 = __hfma2(h2, s2, zs2);
        h3 = __hfma2(h3, s2, zs2);
        out[0] = *reinterpret_cast<uint32_t*>(&h0);
        out[1] = *reinterpret_cast<uint32_t*>(&h1);
        out[2] = *reinterpret_cast<uint32_t*>(&h2);
        out[3] = *reinterpret_cast<uint32_t*>(&h3);
    }
};
template <>
struct Dequant4<__nv_bfloat16> {
    static __device__ __forceinline__ void word(uint32_t w, __nv_bfloat162 s2, __nv_bfloat162 zs2, uint32_t* out) {
        const uint32_t MAGIC = 0x43004300u;  // 128.0 | 128.0 (ulp 1)
        const __nv_bfloat162 k136 = __bfloat162bfloat162(__ushort_as_bfloat16(0x4308));  // 136
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t t;
            asm("lop3.b32 %0, %1, 0x000f000f, %2, 0xea;" : "=r"(t) : "r"(w >> (4 * i)), "r"(MAGIC));
            __nv_bfloat162 h = __hsub2(*reinterpret_cast<__nv_bfloat162*>(&t), k136);  // u - 8, exact
            h = __hfma2(h, s2, zs2);
            out[i] = *reinterpret_cast<uint32_t*>(&h);
        }
    }
};

// ---- int8 (stored as u = q + 128) -> exact fp16/bf16 pairs; the per-column scale is applied in the epilogue.
template <typename T>
struct Dequant8;
template <>
struct Dequant8<__half> {
    static __device__ __forceinline__ void word(uint32_t w, uint32_t* out) {  // 4 bytes -> 2 columns
        const __half2 k1152 = __half2half2(__ushort_as_half(0x6480));  // 1024 + 128
        uint32_t t0 = __byte_perm(w, 0x64646464u, 0x4140);
        uint32_t t1 = __byte_perm(w, 0x64646464u, 0x4342);
        __half2 h0 = __hsub2(*reinterpret_cast<__half2*>(&t0), k1152);
        __half2 h1 = __hsub2(*reinterpret_cast<__half2*>(&t1), k1152);
        out[0] = *reinterpret_cast<uint32_t*>(&h0);
        out[1] = *reinterpret_cast<uint32_t*>(&h1);
    }
};
template <>
struct Dequant8<__nv_bfloat16> {
    static __device__ __forceinline__ void word(uint32_t w, uint32_t* out) {
        // bf16 has 8 significant bits: 1024+u is not representable; go through fp32 magic 2^23 instead.
        // float(0x4B000000 | u) = 8388608 + u  -> minus (8388608 + 128) = q, exact; then round to bf16 (exact, |q| <= 128).
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            float f0 = __uint_as_float(0x4B000000u | ((w >> (16 * i)) & 0xFFu)) - 8388736.f;
            float f1 = __uint_as_float(0x4B000000u | ((w >> (16 * i + 8)) & 0xFFu)) - 8388736.f;
            out[i] = pack2<__nv_bfloat16>(f0, f1);
        }
    }
};

// One thread's share of the reduce-scatter push: the thread owns output column n (fixed) of every batch row it visits.
struct RsLane {
    uint8_t* remote;         // slot address of (row 0, the word holding columns n, n+1) in the owner's region; null: this rank owns n
    uint32_t row_bytes;      // bytes between consecutive batch rows in that slot
    uint32_t epoch;
};
__device__ __forceinline__ RsLane rs_lane(const GemmParamsRs& p, int n) {
    RsLane l;
    l.epoch = *reinterpret_cast<const volatile uint32_t*>(p.rs.epoch) + 1u;   // the call the follow-up gather+norm kernel will run as
    const int Cs = p.N / 8 / p.rs.world, c = n >> 3, owner = c / Cs;
    l.row_bytes = (uint32_t)Cs * 32u;
    l.remote = owner == p.rs.rank ? nullptr
                                  : p.rs.region[owner] + (size_t)(l.epoch & 1u) * p.rs.parity_stride + (size_t)p.rs.rank * p.rs.src_stride +
                                        (size_t)(c - owner * Cs) * 32 + ((n & 7) >> 1) * 8;
    return l;
}
// v = this lane's value for (row b, column n); lanes 2i / 2i+1 hold adjacent columns: the even lane emits the pair.
// Must be called by all 32 lanes of the warp with the same b.
template <typename T>
__device__ __forceinline__ void rs_emit(const GemmParams& p, const RsLane& l, int b, int n, float v, int lane) {
    const float hi = __shfl_down_sync(0xffffffffu, v, 1);
    if (lane & 1) return;
    const uint32_t data = pack2<T>(v, hi);
    if (l.remote == nullptr) {
        *reinterpret_cast<uint32_t*>(reinterpret_cast<T*>(p.y) + (size_t)b * p.N + n) = data;
    } else {
        asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(l.remote + (size_t)b * l.row_bytes), "r"(data), "r"(l.epoch) : "memory");
    }
}

template <typename T>
struct Pair;
template <>
struct Pair<__half> {
    using type = __half2;
    static __device__ __forceinline__ __half2 bcast(uint16_t bits) { return __half2half2(__ushort_as_half(bits)); }
};
template <>
struct Pair<__nv_bfloat16> {
    using type = __nv_bfloat162;
    static __device__ __forceinline__ __nv_bfloat162 bcast(uint16_t bits) {
        return __bfloat162bfloat162(__ushort_as_bfloat16(bits));
    }
};

// developer timeline (only in -DB200_GEMM_DEV builds, see tools/gemm_trace.py): TRACE(slot, it) stores clock64() for CTA
// (0,0); slots: 0 w-issue, 1 x-issue, 2 dq wfull, 3 dq math done, 4 dq aempty seen, 5 dq afull arrive (for it-1),
// 6 mma operands ready, 7 mma issued
#ifdef B200_GEMM_DEV
#define B200_TRACE(slot, it)                                                                     \
    do {                                                                                         \
        if (p.trace && blockIdx.x == 0 && blockIdx.y == 0 && (it) < 64) p.trace[(slot) * 64 + (it)] = clock64(); \
    } while (0)
#define B200_DBG(bit) (p.dbg & (bit))
#define B200_TRACING (p.trace != nullptr)
#else
#define B200_TRACE(slot, it) do { } while (0)
#define B200_DBG(bit) 0
#define B200_TRACING false
#endif

template <int FMT, typename T, int BPAD, int VAR>
__global__ void __launch_bounds__(gemm_threads(VAR), gemm_min_ctas(FMT, BPAD, VAR))
wo_gemm_kernel(const __grid_constant__ CUtensorMap x_map, const __grid_constant__ CUtensorMap w_map,
               const typename GemmParamsOf<VAR>::type p) {
    constexpr int WS = gemm_w_stages(FMT, BPAD, VAR);      // weight ring depth
    constexpr int XS = gemm_x_stages(BPAD);                // activation ring depth
    constexpr int NDQ_WARPS = gemm_ndq_warps(VAR);
    constexpr int NDQ_THREADS = NDQ_WARPS * 32;
    constexpr int W_WTMA = NDQ_WARPS, W_XTMA = NDQ_WARPS + 1, W_MMA = NDQ_WARPS + 2;
    constexpr int X_BYTES = gemm_x_stage_bytes(BPAD);
    constexpr int W_BYTES = gemm_w_bytes(FMT);
    constexpr int A_STAGES = gemm_a_stages(BPAD, VAR);
    constexpr int TMEM_COLS = gemm_tmem_cols(FMT, BPAD, VAR);
    constexpr int NACC = gemm_nacc(BPAD);
    constexpr bool kBf16 = std::is_same<T, __nv_bfloat16>::value;
    constexpr uint32_t IDESC = make_idesc_f16(kGemmTileN, BPAD, kBf16);
    static_assert(WS >= 3, "weight ring too shallow");

    // 1024-byte alignment is required by the 128B-swizzled tiles. The array is DECLARED aligned (the kernel has no static
    // shared memory, so the dynamic segment starts at the aligned window base) and checked once; keeping the pointer's
    // shared-space provenance lets nvcc emit LDS/STS instead of generic accesses with 64-bit address math.
    extern __shared__ __align__(1024) uint8_t smem[];
    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) __trap();
    uint8_t* xring = smem;                                 // XS stages of [2 boxes][BPAD rows][128 B], SW128
    uint8_t* wring = smem + XS * X_BYTES;                  // WS stages of one weight block each
    uint64_t* wfull = reinterpret_cast<uint64_t*>(wring + WS * W_BYTES);   // W_BYTES is a multiple of 512
    uint64_t* wempty = wfull + WS;
    uint64_t* xfull = wempty + WS;
    uint64_t* xempty = xfull + XS;
    uint64_t* afull = xempty + XS;                         // TMEM A buffer written
#ifdef B200_GEMM_HALF_PUBLISH   // experiment (measured: no gain): publish the two 64-k halves of a TMEM buffer separately
    uint64_t* aempty = afull + 2 * A_STAGES;               // afull[a] = first 64 k of buffer a, afull[A_STAGES + a] = second 64 k
#else
    uint64_t* aempty = afull + A_STAGES;                   // TMEM A buffer consumed by the MMA
#endif
    uint64_t* dfull = aempty + A_STAGES;                   // accumulator complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dfull + 1);
    int* s_flag = reinterpret_cast<int*>(tmem_slot + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x, split = blockIdx.y;
    if (B200_TRACING && threadIdx.x == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        atomicMin(reinterpret_cast<unsigned long long*>(&p.trace[8 * 64 + 14]), gt);   // earliest CTA start (ns)
        {
            const int cta = blockIdx.y * gridDim.x + blockIdx.x;
            if (cta < 1024) {
                uint32_t smid;
                asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                p.trace[8 * 64 + 16 + cta * 3 + 0] = smid;
                p.trace[8 * 64 + 16 + cta * 3 + 1] = (long long)gt;
            }
        }
        if ((blockIdx.x == 0 || blockIdx.x == gridDim.x - 1) && blockIdx.y == 0) {
            const int o = blockIdx.x == 0 ? 0 : 4;
            p.trace[8 * 64 + o + 0] = clock64();
            p.trace[8 * 64 + o + 2] = (long long)gt;
        }
    }
    const int kb0 = split * p.kb_per_split;
    const int kb1 = min(kb0 + p.kb_per_split, p.k_blocks);
    const int nkb = kb1 - kb0;  // >= 1 by construction of the grid

    if (threadIdx.x == 0) {
        for (int s = 0; s < WS; ++s) {
            mbar_init(&wfull[s], 1);
            mbar_init(&wempty[s], FMT == kFmtF16 ? 1 : 4);
        }
        for (int s = 0; s < XS; ++s) {
            mbar_init(&xfull[s], 1);
            mbar_init(&xempty[s], 1);
        }
        for (int a = 0; a < A_STAGES; ++a) {
            mbar_init(&afull[a], 4);
#ifdef B200_GEMM_HALF_PUBLISH
            mbar_init(&afull[A_STAGES + a], 4);
#endif
            mbar_init(&aempty[a], 1);
        }
        mbar_init(dfull, 1);
        fence_mbar_init();
    }
    if (warp == W_MMA) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_d = tmem_base;                // NACC accumulator tiles of BPAD fp32 columns, lane = output feature
    const uint32_t tmem_a = tmem_base + NACC * BPAD;  // then A_STAGES buffers of 64 columns: the fp16 A operand

    if (warp == W_WTMA) {
        // ------------------------------------------------------------------ weight producer (HBM stream)
        // Weights never depend on the previous kernel: under programmatic dependent launch this warp starts
        // streaming at once; only the activation producer waits for the upstream grid.
        if (elect_one()) {
            if (FMT == kFmtF16) tma_prefetch_desc(&w_map);
            const uint8_t* wsrc = p.w_blob + ((size_t)tile * p.k_blocks + kb0) * (size_t)W_BYTES;
            int s = 0;
            uint32_t ph = 0;
            for (int it = 0; it < nkb; ++it) {
                uint8_t* stage = wring + s * W_BYTES;
                mbar_wait(&wempty[s], ph ^ 1);
                B200_TRACE(0, it);
                mbar_arrive_expect_tx(&wfull[s], W_BYTES);
                if (FMT == kFmtF16) {
                    const int k0 = (kb0 + it) * kGemmBK;
                    tma_load_2d(stage, &w_map, k0, tile * kGemmTileN, &wfull[s]);
                    tma_load_2d(stage + 16384, &w_map, k0 + 64, tile * kGemmTileN, &wfull[s]);
                } else {
                    tma_bulk_load(stage, wsrc + (size_t)it * W_BYTES, W_BYTES, &wfull[s]);
                }
                if (++s == WS) {
                    s = 0;
                    ph ^= 1;
                }
            }
        }
    } else if (warp == W_XTMA) {
        // ------------------------------------------------------------------ activation producer (L2 hits)
        if (elect_one()) {
            tma_prefetch_desc(&x_map);
            if (p.use_pdl) pdl_wait();
            int s = 0;
            uint32_t ph = 0;
            for (int it = 0; it < nkb; ++it) {
                uint8_t* stage = xring + s * X_BYTES;
                const int k0 = (kb0 + it) * kGemmBK;
                mbar_wait(&xempty[s], ph ^ 1);
                B200_TRACE(1, it);
                mbar_arrive_expect_tx(&xfull[s], X_BYTES);
                tma_load_2d(stage, &x_map, k0, 0, &xfull[s]);
                tma_load_2d(stage + BPAD * 128, &x_map, k0 + 64, 0, &xfull[s]);
                if (++s == XS) {
                    s = 0;
                    ph ^= 1;
                }
            }
        }
    } else if (warp == W_MMA) {
        // ------------------------------------------------------------------ MMA issuer (one elected thread)
        // Warp-uniform control flow + elect.sync: ptxas then knows a single thread issues the tcgen05 instructions and
        // moves their operands to uniform registers directly (with `if (lane == 0)` it emits an ELECT/R2UR/branch
        // "waterfall" loop around EVERY UTCHMMA, ~100 cycles each -- measured with tools/gemm_trace.py).
        if (p.use_pdl && lane == 0) pdl_launch_dependents();
        int sw = 0, sx = 0, a = 0;
        uint32_t phw = 0, phx = 0, aph = 0;
        for (int it = 0; it < nkb; ++it) {
            mbar_wait(&xfull[sx], phx);
            if (FMT == kFmtF16) mbar_wait(&wfull[sw], phw);
            else mbar_wait(&afull[a], aph);
            int sx_n = sx + 1, sw_n = sw + 1, a_n = a + 1;
            uint32_t phx_n = phx, phw_n = phw, aph_n = aph;
            if (sx_n == XS) { sx_n = 0; phx_n ^= 1; }
            if (sw_n == WS) { sw_n = 0; phw_n ^= 1; }
            if (a_n == A_STAGES) { a_n = 0; aph_n ^= 1; }
            tc_fence_after();
            if (elect_one()) {
                B200_TRACE(6, it);
                const uint32_t xs = smem_u32(xring + sx * X_BYTES);
                const uint32_t wsm = smem_u32(wring + sw * W_BYTES);
#pragma unroll
                for (int j = 0; j < kGemmBK / 16; ++j) {
                    if (B200_DBG(4)) break;
#ifdef B200_GEMM_HALF_PUBLISH
                    if (FMT != kFmtF16 && j == 4) {        // the second 64 k of the block are published separately
                        mbar_wait(&afull[A_STAGES + a], aph);
                        tc_fence_after();
                    }
#endif
                    const uint64_t bdesc = make_smem_desc_sw128(xs + (j >> 2) * (BPAD * 128) + (j & 3) * 32);
                    const uint32_t acc = (it > 0 || j >= NACC) ? 1u : 0u;
                    const uint32_t dcol = tmem_d + (j % NACC) * BPAD;
                    if (FMT == kFmtF16) {
                        const uint64_t adesc = make_smem_desc_sw128(wsm + (j >> 2) * 16384 + (j & 3) * 32);
                        umma_ss_f16(dcol, adesc, bdesc, IDESC, acc);
                    } else {
                        umma_ts_f16(dcol, tmem_a + a * 64 + j * 8, bdesc, IDESC, acc);
                    }
                }
                umma_commit(&xempty[sx]);                       // activation stage consumed
                if (FMT == kFmtF16) umma_commit(&wempty[sw]);   // f16 weights are read by the MMA itself
                else umma_commit(&aempty[a]);                   // TMEM A buffer reusable
                if (it == nkb - 1) umma_commit(dfull);          // accumulator final
                B200_TRACE(7, it);
            }
            __syncwarp();
            sx = sx_n; phx = phx_n;
            sw = sw_n; phw = phw_n;
            a = a_n; aph = aph_n;
        }
    } else {
        // ------------------------------------------------------------------ dequant warps (TMEM A producers)
        // Two groups of four warps take ALTERNATE k-blocks (two blocks in flight per CTA); inside a group warp q owns TMEM
        // lane quarter q, i.e. 32 feature rows x the whole 128-deep block, processed as two 64-k halves. One barrier round
        // (weights landed / TMEM buffer free / buffer published) therefore covers 4096 weights per warp: the round's
        // latency (~100 cycles per mbarrier operation even when the phase is already complete) is what bounds this loop.
        const int quarter = warp & 3;     // TMEM lane quarter this warp may touch (hardware rule: warp id % 4)
        const int grp = warp >> 2;        // dequant group 0/1
        const int row = quarter * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        if (FMT != kFmtF16) {
            constexpr int G = 2;
            int s = grp % WS, a = grp % A_STAGES;
            uint32_t ph = 0, aph = 0;
            for (int it = grp; it < nkb; it += G) {
                mbar_wait(&wfull[s], ph);
                if (threadIdx.x == 0) B200_TRACE(2, it);
                const uint32_t wb = smem_u32(wring) + s * W_BYTES;   // 32-bit shared address of this stage
                typename Pair<T>::type s2, zs2;
                if (FMT == kFmtInt4 || FMT == kFmtInt8G) {
                    constexpr int Q_BYTES = FMT == kFmtInt4 ? 8192 : 16384;
                    s2 = Pair<T>::bcast((uint16_t)lds_u16(wb + Q_BYTES + row * 2));
                    zs2 = Pair<T>::bcast((uint16_t)lds_u16(wb + Q_BYTES + 256 + row * 2));
                }
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t regs[2][16];
                    if (FMT == kFmtInt4) {
                        uint4 v[2];
#pragma unroll
                        for (int cc = 0; cc < 2; ++cc) v[cc] = lds_v4(wb + (half * 2 + cc) * 2048 + row * 16);
                        if (half == 1) {   // every byte of the stage this warp needs is in registers: release it
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&wempty[s]);
                        }
#pragma unroll
                        for (int cc = 0; cc < 2; ++cc) {
                            if (B200_DBG(1)) {
#pragma unroll
                                for (int q = 0; q < 16; ++q) regs[cc][q] = (q & 1) ? v[cc].x ^ v[cc].w : v[cc].y ^ v[cc].z;
                                continue;
                            }
                            Dequant4<T>::word(v[cc].x, s2, zs2, &regs[cc][0]);
                            Dequant4<T>::word(v[cc].y, s2, zs2, &regs[cc][4]);
                            Dequant4<T>::word(v[cc].z, s2, zs2, &regs[cc][8]);
                            Dequant4<T>::word(v[cc].w, s2, zs2, &regs[cc][12]);
                        }
                    } else {
                        uint4 v[2][2];
#pragma unroll
                        for (int cc = 0; cc < 2; ++cc)
#pragma unroll
                            for (int h = 0; h < 2; ++h) v[cc][h] = lds_v4(wb + ((half * 2 + cc) * 2 + h) * 2048 + row * 16);
                        if (half == 1) {
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&wempty[s]);
                        }
#pragma unroll
                        for (int cc = 0; cc < 2; ++cc)
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                Dequant8<T>::word(v[cc][h].x, &regs[cc][h * 8 + 0]);
                                Dequant8<T>::word(v[cc][h].y, &regs[cc][h * 8 + 2]);
                                Dequant8<T>::word(v[cc][h].z, &regs[cc][h * 8 + 4]);
                                Dequant8<T>::word(v[cc][h].w, &regs[cc][h * 8 + 6]);
                            }
                        if (FMT == kFmtInt8G) {   // group-wise 8-bit: W' = q_s * s + zero*scale, one rounding (as INT4)
#pragma unroll
                            for (int cc = 0; cc < 2; ++cc)
#pragma unroll
                                for (int q = 0; q < 16; ++q) {
                                    typename Pair<T>::type h = *reinterpret_cast<typename Pair<T>::type*>(&regs[cc][q]);
                                    h = __hfma2(h, s2, zs2);
                                    regs[cc][q] = *reinterpret_cast<uint32_t*>(&h);
                                }
                        }
                    }
                    if (half == 0) {
                        if (threadIdx.x == 0) B200_TRACE(3, it);
                        mbar_wait(&aempty[a], aph ^ 1);
                        if (threadIdx.x == 0) B200_TRACE(4, it);
                        tc_fence_after();
                    }
                    const uint32_t dst = tmem_a + lane_addr + a * 64 + half * 32;
                    if (B200_DBG(2)) {
                        if (regs[0][0] == 0x12345678u && regs[1][15] == 0x9abcdef0u) s_flag[1] = 1;  // keep the math alive
                    } else {
                        tmem_st_32x32b_x16(dst, regs[0]);
                        tmem_st_32x32b_x16(dst + 16, regs[1]);
                    }
#ifdef B200_GEMM_HALF_PUBLISH
                    // publish each 64-k half on its own barrier: the MMA starts on half 0 while half 1 is being dequantised
                    tmem_wait_st();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&afull[half * A_STAGES + a]);
#endif
                }
#ifndef B200_GEMM_HALF_PUBLISH
                // publish the block at once (measured: deferring the publish behind the next block's math is slower --
                // the dequant -> MMA -> release chain, not the TMEM store latency, bounds the loop)
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&afull[a]);
#endif
                if (threadIdx.x == 0) B200_TRACE(5, it);
                s += G;
                if (s >= WS) {
                    s -= WS;
                    ph ^= 1;
                }
                a += G;
                if (a >= A_STAGES) {
                    a -= A_STAGES;
                    aph ^= 1;
                }
            }
        }

        // ------------------------------------------------------------------ epilogue
#define B200_TRACE_E(i)                                                                                          \
    do {                                                                                                         \
        if (B200_TRACING && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) p.trace[8 * 64 + 8 + (i)] = clock64(); \
    } while (0)
        B200_TRACE_E(0);
        mbar_wait(dfull, 0);
        tc_fence_after();
        B200_TRACE_E(1);
        // warps with the same lane quarter share the batch columns in slices of 16
        constexpr int SLICES = BPAD / 16;                   // 16-column slices of the accumulator
        constexpr int KCS = NDQ_WARPS / 4;                  // warps per lane quarter
        const int kce = warp >> 2;                          // this warp's index among them
        const int n = tile * kGemmTileN + row;
        const bool n_ok = n < p.N;
        const int n_tiles = gridDim.x;
        T* yp = reinterpret_cast<T*>(p.y);
        float cscale = 1.f, bias = 0.f;
        if (n_ok) {
            if (FMT == kFmtInt8) cscale = to_f32<T>(reinterpret_cast<const T*>(p.col_scale)[n]);
            if (p.bias) bias = to_f32<T>(reinterpret_cast<const T*>(p.bias)[n]);
        }
        float* fin = reinterpret_cast<float*>(wring);       // [bpad][128] fp32 staging for the fused SiLU*mul (weight ring is idle)
        // 16 batch columns of this thread's feature row, summed over the NACC accumulator tiles
        auto load_acc = [&](int sl, float (&v)[16]) {
            uint32_t r[NACC][16];
#pragma unroll
            for (int q = 0; q < NACC; ++q) tmem_ld_32x32b_x16(tmem_d + lane_addr + q * BPAD + sl * 16, r[q]);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float a = __uint_as_float(r[0][j]);
#pragma unroll
                for (int q = 1; q < NACC; ++q) a += __uint_as_float(r[q][j]);
                v[j] = a;
            }
        };
        if (p.nsplit == 1) {
            RsLane rsl{};
            if constexpr (VAR == 1) rsl = rs_lane(p, n);
            for (int sl = kce; sl < SLICES; sl += KCS) {
                float v[16];
                load_acc(sl, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int b = sl * 16 + j;
                    if (p.silu_mul) fin[b * kGemmTileN + row] = fmaf(v[j], cscale, bias);
                    else if constexpr (VAR == 1) {
                        if (b < p.B) rs_emit<T>(p, rsl, b, n, fmaf(v[j], cscale, bias), lane);    // N % 128 == 0 in this mode
                    } else if (n_ok && b < p.B) yp[(size_t)b * p.N + n] = from_f32<T>(fmaf(v[j], cscale, bias));
                }
            }
        } else if (p.cluster_reduce) {
            // split-K merge through distributed shared memory: park the fp32 partial tile [bpad][128] in this CTA's smem
            // (the activation ring is idle once the accumulator is final); the merge itself follows the cluster barrier below.
            float* red = reinterpret_cast<float*>(xring);
            for (int sl = kce; sl < SLICES; sl += KCS) {
                float v[16];
                load_acc(sl, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) red[(sl * 16 + j) * kGemmTileN + row] = v[j];
            }
        } else {
            // write the fp32 partial tile [bpad][128] (coalesced along n), then the last CTA of this n-tile reduces
            float* wsp = p.ws + ((size_t)split * n_tiles + tile) * (size_t)(BPAD * kGemmTileN);
            for (int sl = kce; sl < SLICES; sl += KCS) {
                float v[16];
                load_acc(sl, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int b = sl * 16 + j;
                    if (b < p.B) __stcg(&wsp[(size_t)b * kGemmTileN + row], v[j]);
                }
            }
            B200_TRACE_E(2);
            __threadfence();
            asm volatile("bar.sync 1, %0;" ::"n"(NDQ_THREADS) : "memory");
            B200_TRACE_E(3);
            if (threadIdx.x == 0) {
                const int prev = atomicAdd(&p.sem[tile], 1);
                *s_flag = (prev == p.nsplit - 1);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(NDQ_THREADS) : "memory");
            B200_TRACE_E(4);
            if (*s_flag) {
                __threadfence();
                for (int sl = kce; sl < SLICES; sl += KCS) {
                    // all loads of a slice are issued before the first use (the partials sit in L2)
                    float acc[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
                    for (int sp = 0; sp < p.nsplit; ++sp) {
                        const float* src = p.ws + ((size_t)sp * n_tiles + tile) * (size_t)(BPAD * kGemmTileN) + row;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int b = sl * 16 + j;
                            if (b < p.B) acc[j] += __ldcg(src + (size_t)b * kGemmTileN);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int b = sl * 16 + j;
                        if (n_ok && b < p.B) yp[(size_t)b * p.N + n] = from_f32<T>(fmaf(acc[j], cscale, bias));
                    }
                }
                if (threadIdx.x == 0) p.sem[tile] = 0;
            }
        }
        B200_TRACE_E(5);
        tc_fence_before();
    }

    if (p.nsplit > 1 && p.cluster_reduce) {
        // every thread of every CTA of the cluster (1, nsplit, 1) passes both barriers
        cluster_sync_all();
        if (warp < NDQ_WARPS) {
            const int S = p.nsplit;
            const int rank = (int)cluster_ctarank();          // == blockIdx.y
            const int row = threadIdx.x & (kGemmTileN - 1);
            const int lane_grp = threadIdx.x >> 7;             // NDQ_THREADS / 128 column lanes
            const int n = tile * kGemmTileN + row;
            const bool n_ok = n < p.N;
            float cscale = 1.f, bias = 0.f;
            if (n_ok) {
                if (FMT == kFmtInt8) cscale = to_f32<T>(reinterpret_cast<const T*>(p.col_scale)[n]);
                if (p.bias) bias = to_f32<T>(reinterpret_cast<const T*>(p.bias)[n]);
            }
            const float* red = reinterpret_cast<const float*>(xring);
            T* yp = reinterpret_cast<T*>(p.y);
            RsLane rsl{};
            if constexpr (VAR == 1) rsl = rs_lane(p, n);
            // CTA `rank` owns batch columns rank, rank+S, ...; fixed summation order sp = 0..S-1 (deterministic)
            for (int b = rank + lane_grp * S; b < p.B; b += S * (NDQ_THREADS / kGemmTileN)) {
                const uint32_t laddr = smem_u32(red + b * kGemmTileN + row);
                float part[8];
#pragma unroll
                for (int sp = 0; sp < 8; ++sp) part[sp] = sp < S ? dsmem_ld_f32(dsmem_addr(laddr, sp)) : 0.f;
                float acc = 0.f;
#pragma unroll
                for (int sp = 0; sp < 8; ++sp) acc += part[sp];
                if (p.silu_mul) reinterpret_cast<float*>(wring)[b * kGemmTileN + row] = fmaf(acc, cscale, bias);
                else if constexpr (VAR == 1) rs_emit<T>(p, rsl, b, n, fmaf(acc, cscale, bias), lane);
                else if (n_ok) yp[(size_t)b * p.N + n] = from_f32<T>(fmaf(acc, cscale, bias));
            }
        }
        cluster_sync_all();   // peers may still be reading this CTA's partial
    }
    if (p.silu_mul) {
        // fused SiLU(gate) * up (replaces rtp_llm_ops.silu_and_mul after the w13 GEMM, modules/hybrid/dense_mlp.py:95-106)
        __syncthreads();
        if (warp < NDQ_WARPS) {
            const float* fin = reinterpret_cast<const float*>(wring);
            T* yp = reinterpret_cast<T*>(p.y);
            const int n_out = p.N / 2;
            const int S = (p.nsplit > 1 && p.cluster_reduce) ? p.nsplit : 1;
            const int rank = S > 1 ? (int)cluster_ctarank() : 0;
            for (int idx = threadIdx.x; idx < BPAD * 64; idx += NDQ_THREADS) {
                const int b = idx >> 6, r = idx & 63;
                if (b >= p.B || (b % S) != rank) continue;      // with a cluster merge each CTA owns the columns b % S == rank
                const int j = tile * 64 + r;
                if (j >= n_out) continue;
                const float g = fin[b * kGemmTileN + 2 * r], u = fin[b * kGemmTileN + 2 * r + 1];   // rows 2r / 2r+1 = gate / up pair
                yp[(size_t)b * n_out + j] = from_f32<T>(__fdividef(g, 1.f + __expf(-g)) * u);   // fast divide: <= 2 ulp in fp32, below the rounding of the result
            }
        }
    }
    __syncthreads();
    if (B200_TRACING && threadIdx.x == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        atomicMax(reinterpret_cast<unsigned long long*>(&p.trace[8 * 64 + 15]), gt);   // latest CTA end (ns)
        {
            const int cta = blockIdx.y * gridDim.x + blockIdx.x;
            if (cta < 1024) p.trace[8 * 64 + 16 + cta * 3 + 2] = (long long)gt;
        }
        if ((blockIdx.x == 0 || blockIdx.x == gridDim.x - 1) && blockIdx.y == 0) {
            const int o = blockIdx.x == 0 ? 0 : 4;
            p.trace[8 * 64 + o + 1] = clock64();
            p.trace[8 * 64 + o + 3] = (long long)gt;
        }
    }
    if (warp == W_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ---- load-time re-layout kernels (replace device_impl.py:392-479 preprocess_weights_for_mixed_gemm): reference
// (un-permuted) tensors -> per-(n-tile, k-block) blobs consumed above.
// int4: q_packed [K][N/2] (byte = hi nibble col 2j+1, lo nibble col 2j, two's complement q_s), scales/zs [K/128][N] (16-bit).
static __global__ void pack_w4_kernel(const uint8_t* __restrict__ q_packed, const uint16_t* __restrict__ scales,
                               const uint16_t* __restrict__ zs, int K, int N, uint8_t* __restrict__ blob) {
    const int k_blocks = K / kGemmBK;
    const int n_tiles = (N + kGemmTileN - 1) / kGemmTileN;
    const size_t total_words = (size_t)n_tiles * k_blocks * (kW4BlockBytes / 4);
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total_words;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int word_in_block = (int)(idx % (kW4BlockBytes / 4));
        const size_t blk = idx / (kW4BlockBytes / 4);
        const int kb = (int)(blk % k_blocks), tile = (int)(blk / k_blocks);
        uint32_t out = 0;
        if (word_in_block < 2048) {
            const int c = word_in_block / 512, r = (word_in_block % 512) / 4, j = word_in_block % 4;
            const int n = tile * kGemmTileN + r;
            if (n < N) {
#pragma unroll
                for (int pos = 0; pos < 8; ++pos) {
                    const int kl = c * 32 + j * 8 + (pos < 4 ? 2 * pos : 2 * (pos - 4) + 1);
                    const uint8_t byte = q_packed[(size_t)(kb * kGemmBK + kl) * (N / 2) + n / 2];
                    const uint32_t nib = (n & 1) ? (byte >> 4) : (byte & 0xF);
                    const uint32_t u = (nib + 8) & 0xF;  // two's complement q_s -> q_s + 8
                    out |= u << (4 * pos);
                }
            } else {
                out = 0x88888888u;  // q_s = 0 for padded features
            }
        } else {
            const int w = word_in_block - 2048;          // 0..127: 64 words of scales then 64 words of zs
            const uint16_t* src = (w < 64) ? scales : zs;
            const int r0 = (w % 64) * 2;
            uint32_t lo = 0, hi = 0;
            const int n0 = tile * kGemmTileN + r0;
            if (n0 < N) lo = src[(size_t)kb * N + n0];
            if (n0 + 1 < N) hi = src[(size_t)kb * N + n0 + 1];
            out = lo | (hi << 16);
        }
        reinterpret_cast<uint32_t*>(blob)[idx] = out;
    }
}
// int8: q [K][N] int8 -> blobs of u = q + 128
static __global__ void pack_w8_kernel(const int8_t* __restrict__ q, int K, int N, uint8_t* __restrict__ blob) {
    const int k_blocks = K / kGemmBK;
    const int n_tiles = (N + kGemmTileN - 1) / kGemmTileN;
    const size_t total = (size_t)n_tiles * k_blocks * kW8BlockBytes;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int byte_in_block = (int)(idx % kW8BlockBytes);
        const size_t blk = idx / kW8BlockBytes;
        const int kb = (int)(blk % k_blocks), tile = (int)(blk / k_blocks);
        const int c = byte_in_block / 2048, r = (byte_in_block % 2048) / 16, i = byte_in_block % 16;
        const int n = tile * kGemmTileN + r, k = kb * kGemmBK + c * 16 + i;
        blob[idx] = (n < N) ? (uint8_t)((int)q[(size_t)k * N + n] + 128) : (uint8_t)128;
    }
}

// int8 group-wise: q [K][N] int8 (q_s = q_u - 128), scales / zs [K/128][N] (16-bit) -> blobs of u = q_s + 128 + the group's scales, zs
static __global__ void pack_w8g_kernel(const int8_t* __restrict__ q, const uint16_t* __restrict__ scales,
                                       const uint16_t* __restrict__ zs, int K, int N, uint8_t* __restrict__ blob) {
    const int k_blocks = K / kGemmBK;
    const int n_tiles = (N + kGemmTileN - 1) / kGemmTileN;
    const size_t total = (size_t)n_tiles * k_blocks * kW8GBlockBytes;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int byte_in_block = (int)(idx % kW8GBlockBytes);
        const size_t blk = idx / kW8GBlockBytes;
        const int kb = (int)(blk % k_blocks), tile = (int)(blk / k_blocks);
        if (byte_in_block < 16384) {
            const int c = byte_in_block / 2048, r = (byte_in_block % 2048) / 16, i = byte_in_block % 16;
            const int n = tile * kGemmTileN + r, k = kb * kGemmBK + c * 16 + i;
            blob[idx] = (n < N) ? (uint8_t)((int)q[(size_t)k * N + n] + 128) : (uint8_t)128;
        } else {
            const int w = byte_in_block - 16384;         // 256 bytes of scales then 256 bytes of zs
            const uint16_t* src = (w < 256) ? scales : zs;
            const int r = (w % 256) / 2, n = tile * kGemmTileN + r;
            const uint16_t v = n < N ? src[(size_t)kb * N + n] : (uint16_t)0;
            blob[idx] = (uint8_t)((w & 1) ? (v >> 8) : (v & 0xFF));
        }
    }
}

}  // namespace b200
