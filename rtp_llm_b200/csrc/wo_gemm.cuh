// Weight-only (INT4 group-128 / INT8 per-column / FP16) x FP16|BF16 GEMM for decode batches, sm_100a.
//
//   Y[b][n] = sum_k X[b][k] * W'[k][n] (+ bias[n]),   W' per /root/reference/rtp_llm/device/device_impl.py:183-222,242-300
//   (SURVEY.md section 8 a8-a10).  Replaces the cutlass fpA_intB "mixed gemm" the reference loader still prepares weights
//   for (device_impl.py:392-479) but whose kernel is absent from the snapshot; plugs in behind LinearBase.forward
//   (/root/reference/rtp_llm/models_py/modules/factory/linear/linear_base.py:81).
//
// B200 design (swap-AB, weights stationary on the MMA M side):
//   * one CTA = 128 output features x a run of 128-deep k-blocks (split-K across blockIdx.y);
//   * warp 0   : TMA producer. Weights were re-laid out at load time into per-(n-tile,k-block) contiguous blobs
//                (packed nibbles + the group's scales and zero*scale), so one cp.async.bulk stages a whole block;
//                the activation block [B x 128] comes through a 128B-swizzled tensor map (OOB rows zero-filled);
//   * warps 2-9: dequantise IN REGISTERS (lop3 magic-number int->fp16, exact; hfma2 with the group scale / zero*scale)
//                and store the fp16 operand straight INTO TENSOR MEMORY (tcgen05.st) -- the A operand of
//   * warp 1   : one elected thread issues tcgen05.mma (M=128 features, N=batch pad, K=16) with A from TMEM and
//                B (activations, K-major SW128) from shared memory, fp32 accumulators in TMEM;
//   * epilogue : warps 2-9 read the accumulators (tcgen05.ld), split-K partials are merged by the last-arriving CTA
//                (fixed order -> deterministic), bias / per-column scale applied, coalesced stores.
//   The FP16-weight variant feeds A from shared memory (TMA tensor map, SW128) with the same pipeline.
#pragma once
#include <type_traits>

#include "ptx.cuh"

namespace b200 {

enum : int { kFmtF16 = 0, kFmtInt8 = 1, kFmtInt4 = 2 };

constexpr int kGemmBK = 128;       // k elements per pipeline stage (= one INT4 quantisation group)
constexpr int kGemmTileN = 128;    // output features per CTA (MMA M)
constexpr int kGemmThreads = 320;  // warp0 TMA, warp1 MMA, warps 2..9 dequant + epilogue
constexpr int kW4BlockBytes = 8192 + 256 + 256;
constexpr int kW8BlockBytes = 16384;
constexpr int kW16BlockBytes = 32768;

__host__ __device__ constexpr int gemm_w_bytes(int fmt) {
    return fmt == kFmtInt4 ? kW4BlockBytes : (fmt == kFmtInt8 ? kW8BlockBytes : kW16BlockBytes);
}
__host__ __device__ constexpr int gemm_stage_bytes(int fmt, int bpad) {
    return ((bpad * 256 + gemm_w_bytes(fmt)) + 1023) / 1024 * 1024;
}
__host__ __device__ constexpr int gemm_stages(int fmt, int bpad) {
    // keep two CTAs per SM where the stage is small; >= 3 stages always
    return fmt == kFmtF16 ? (bpad <= 64 ? 4 : 3) : (bpad <= 32 ? 6 : (bpad <= 64 ? 4 : 3));
}
__host__ __device__ constexpr int gemm_a_stages(int bpad) { return bpad <= 64 ? 3 : 2; }
__host__ __device__ constexpr int gemm_tmem_cols(int fmt, int bpad) {
    int need = bpad + (fmt == kFmtF16 ? 0 : gemm_a_stages(bpad) * 64);
    int c = 32;
    while (c < need) c *= 2;
    return c;
}
__host__ __device__ constexpr int gemm_smem_bytes(int fmt, int bpad) {
    return gemm_stages(fmt, bpad) * gemm_stage_bytes(fmt, bpad) + 1024 /*align*/ + 512 /*barriers*/;
}

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

template <int FMT, typename T, int BPAD>
__global__ void __launch_bounds__(kGemmThreads, (gemm_smem_bytes(FMT, BPAD) <= 110 * 1024) ? 2 : 1)
wo_gemm_kernel(const __grid_constant__ CUtensorMap x_map, const __grid_constant__ CUtensorMap w_map, const GemmParams p) {
    constexpr int STAGES = gemm_stages(FMT, BPAD);
    constexpr int STAGE_BYTES = gemm_stage_bytes(FMT, BPAD);
    constexpr int X_BYTES = BPAD * 256;
    constexpr int W_BYTES = gemm_w_bytes(FMT);
    constexpr int A_STAGES = gemm_a_stages(BPAD);
    constexpr int TMEM_COLS = gemm_tmem_cols(FMT, BPAD);
    constexpr int NDQ_WARPS = 8;
    constexpr bool kBf16 = std::is_same<T, __nv_bfloat16>::value;
    constexpr uint32_t IDESC = make_idesc_f16(kGemmTileN, BPAD, kBf16);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);  // TMA landed (W + X)
    uint64_t* empty_bar = full_bar + STAGES;                                       // smem stage free again
    uint64_t* afull_bar = empty_bar + STAGES;                                      // TMEM A buffer written
    uint64_t* aempty_bar = afull_bar + A_STAGES;                                   // TMEM A buffer consumed by the MMA
    uint64_t* dfull_bar = aempty_bar + A_STAGES;                                   // accumulator complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dfull_bar + 1);
    int* s_flag = reinterpret_cast<int*>(tmem_slot + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x, split = blockIdx.y;
    const int kb0 = split * p.kb_per_split;
    const int kb1 = min(kb0 + p.kb_per_split, p.k_blocks);
    const int nkb = kb1 - kb0;  // >= 1 by construction of the grid

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], FMT == kFmtF16 ? 1 : NDQ_WARPS + 1);
        }
        for (int a = 0; a < A_STAGES; ++a) {
            mbar_init(&afull_bar[a], NDQ_WARPS);
            mbar_init(&aempty_bar[a], 1);
        }
        mbar_init(dfull_bar, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_d = tmem_base;            // columns [0, BPAD): fp32 accumulators, lane = output feature
    const uint32_t tmem_a = tmem_base + BPAD;     // columns [BPAD, BPAD + A_STAGES*64): fp16 A operand buffers

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            tma_prefetch_desc(&x_map);
            if (FMT == kFmtF16) tma_prefetch_desc(&w_map);
            const uint8_t* wsrc = p.w_blob + ((size_t)tile * p.k_blocks + kb0) * (size_t)W_BYTES;
            // Weights never depend on the previous kernel: with programmatic dependent launch the first ring of
            // weight blocks is requested BEFORE waiting on the upstream grid; only the activations wait.
            const int pre = p.use_pdl ? min(nkb, STAGES) : 0;
            for (int it = 0; it < pre; ++it) {
                uint8_t* stage = smem + it * STAGE_BYTES;
                mbar_arrive_expect_tx(&full_bar[it], X_BYTES + W_BYTES);
                if (FMT == kFmtF16) {
                    tma_load_2d(stage + X_BYTES, &w_map, (kb0 + it) * kGemmBK, tile * kGemmTileN, &full_bar[it]);
                    tma_load_2d(stage + X_BYTES + 16384, &w_map, (kb0 + it) * kGemmBK + 64, tile * kGemmTileN,
                                &full_bar[it]);
                } else {
                    tma_bulk_load(stage + X_BYTES, wsrc + (size_t)it * W_BYTES, W_BYTES, &full_bar[it]);
                }
            }
            if (p.use_pdl) pdl_wait();
            for (int it = 0; it < nkb; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                uint8_t* stage = smem + s * STAGE_BYTES;
                const int k0 = (kb0 + it) * kGemmBK;
                if (it >= pre) {
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    mbar_arrive_expect_tx(&full_bar[s], X_BYTES + W_BYTES);
                    if (FMT == kFmtF16) {
                        tma_load_2d(stage + X_BYTES, &w_map, k0, tile * kGemmTileN, &full_bar[s]);
                        tma_load_2d(stage + X_BYTES + 16384, &w_map, k0 + 64, tile * kGemmTileN, &full_bar[s]);
                    } else {
                        tma_bulk_load(stage + X_BYTES, wsrc + (size_t)it * W_BYTES, W_BYTES, &full_bar[s]);
                    }
                }
                tma_load_2d(stage, &x_map, k0, 0, &full_bar[s]);
                tma_load_2d(stage + BPAD * 128, &x_map, k0 + 64, 0, &full_bar[s]);
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (one elected thread)
        if (p.use_pdl && lane == 0) pdl_launch_dependents();
        for (int it = 0; it < nkb; ++it) {
            const int s = it % STAGES;
            const uint32_t ph = (it / STAGES) & 1;
            const int a = it % A_STAGES;
            const uint32_t aph = (it / A_STAGES) & 1;
            mbar_wait(&full_bar[s], ph);
            if (FMT != kFmtF16) mbar_wait(&afull_bar[a], aph);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t xs = smem_u32(smem + s * STAGE_BYTES);
#pragma unroll
                for (int j = 0; j < kGemmBK / 16; ++j) {
                    const uint64_t bdesc = make_smem_desc_sw128(xs + (j >> 2) * (BPAD * 128) + (j & 3) * 32);
                    const uint32_t acc = (it > 0 || j > 0) ? 1u : 0u;
                    if (FMT == kFmtF16) {
                        const uint64_t adesc = make_smem_desc_sw128(xs + X_BYTES + (j >> 2) * 16384 + (j & 3) * 32);
                        umma_ss_f16(tmem_d, adesc, bdesc, IDESC, acc);
                    } else {
                        umma_ts_f16(tmem_d, tmem_a + a * 64 + j * 8, bdesc, IDESC, acc);
                    }
                }
                umma_commit(&empty_bar[s]);                      // X (and f16 W) of this stage consumed
                if (FMT != kFmtF16) umma_commit(&aempty_bar[a]); // TMEM A buffer reusable
                if (it == nkb - 1) umma_commit(dfull_bar);       // accumulator final
            }
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------------ dequant warps (TMEM A producers)
        const int dq = warp - 2;          // 0..7
        const int quarter = warp & 3;     // TMEM lane quarter this warp may touch
        const int sub = dq >> 2;          // which half of the k-block / of the batch columns
        const int row = quarter * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        if (FMT != kFmtF16) {
            for (int it = 0; it < nkb; ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                const int a = it % A_STAGES;
                const uint32_t aph = (it / A_STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                const uint8_t* wb = smem + s * STAGE_BYTES + X_BYTES;
                uint32_t regs[2][16];
                if (FMT == kFmtInt4) {
                    const uint16_t* sc = reinterpret_cast<const uint16_t*>(wb + 8192);
                    const typename Pair<T>::type s2 = Pair<T>::bcast(sc[row]);
                    const typename Pair<T>::type zs2 = Pair<T>::bcast(sc[128 + row]);
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc) {
                        const int c = sub * 2 + cc;  // 32-k chunk
                        const uint4 v = *reinterpret_cast<const uint4*>(wb + c * 2048 + row * 16);
                        Dequant4<T>::word(v.x, s2, zs2, &regs[cc][0]);
                        Dequant4<T>::word(v.y, s2, zs2, &regs[cc][4]);
                        Dequant4<T>::word(v.z, s2, zs2, &regs[cc][8]);
                        Dequant4<T>::word(v.w, s2, zs2, &regs[cc][12]);
                    }
                } else {
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int c = sub * 4 + cc * 2 + h;  // 16-k chunk
                            const uint4 v = *reinterpret_cast<const uint4*>(wb + c * 2048 + row * 16);
                            Dequant8<T>::word(v.x, &regs[cc][h * 8 + 0]);
                            Dequant8<T>::word(v.y, &regs[cc][h * 8 + 2]);
                            Dequant8<T>::word(v.z, &regs[cc][h * 8 + 4]);
                            Dequant8<T>::word(v.w, &regs[cc][h * 8 + 6]);
                        }
                    }
                }
                // the packed weights are in registers: this warp is done with the smem stage
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[s]);
                mbar_wait(&aempty_bar[a], aph ^ 1);
                tc_fence_after();
                const uint32_t dst = tmem_a + lane_addr + a * 64 + sub * 32;
                tmem_st_32x32b_x16(dst, regs[0]);
                tmem_st_32x32b_x16(dst + 16, regs[1]);
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&afull_bar[a]);
            }
        }

        // ------------------------------------------------------------------ epilogue
        mbar_wait(dfull_bar, 0);
        tc_fence_after();
        constexpr int COLS_PER_SUB = BPAD >= 32 ? BPAD / 2 : BPAD;  // BPAD=16: sub 0 takes all 16 columns
        const bool active = (BPAD >= 32) || sub == 0;
        const int col0 = (BPAD >= 32) ? sub * COLS_PER_SUB : 0;
        const int n = tile * kGemmTileN + row;
        const bool n_ok = n < p.N;
        const int n_tiles = gridDim.x;
        T* yp = reinterpret_cast<T*>(p.y);
        float cscale = 1.f, bias = 0.f;
        if (n_ok) {
            if (FMT == kFmtInt8) cscale = to_f32<T>(reinterpret_cast<const T*>(p.col_scale)[n]);
            if (p.bias) bias = to_f32<T>(reinterpret_cast<const T*>(p.bias)[n]);
        }
        if (p.nsplit == 1) {
            if (active) {
#pragma unroll
                for (int cb = 0; cb < COLS_PER_SUB; cb += 16) {
                    uint32_t v[16];
                    tmem_ld_32x32b_x16(tmem_d + lane_addr + col0 + cb, v);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int b = col0 + cb + j;
                        if (n_ok && b < p.B) yp[(size_t)b * p.N + n] = from_f32<T>(fmaf(__uint_as_float(v[j]), cscale, bias));
                    }
                }
            }
        } else {
            // write the fp32 partial tile [bpad][128] (coalesced along n), then the last CTA of this n-tile reduces
            float* wsp = p.ws + ((size_t)split * n_tiles + tile) * (size_t)(BPAD * kGemmTileN);
            if (active) {
#pragma unroll
                for (int cb = 0; cb < COLS_PER_SUB; cb += 16) {
                    uint32_t v[16];
                    tmem_ld_32x32b_x16(tmem_d + lane_addr + col0 + cb, v);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int b = col0 + cb + j;
                        if (b < p.B) __stcg(&wsp[(size_t)b * kGemmTileN + row], __uint_as_float(v[j]));
                    }
                }
            }
            __threadfence();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (threadIdx.x == 64) {
                const int prev = atomicAdd(&p.sem[tile], 1);
                *s_flag = (prev == p.nsplit - 1);
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (*s_flag) {
                __threadfence();
                if (active) {
                    for (int cb = 0; cb < COLS_PER_SUB; ++cb) {
                        const int b = col0 + cb;
                        if (b >= p.B) break;
                        float acc = 0.f;
                        for (int sp = 0; sp < p.nsplit; ++sp)
                            acc += __ldcg(&p.ws[((size_t)sp * n_tiles + tile) * (size_t)(BPAD * kGemmTileN) +
                                                (size_t)b * kGemmTileN + row]);
                        if (n_ok) yp[(size_t)b * p.N + n] = from_f32<T>(fmaf(acc, cscale, bias));
                    }
                }
                if (threadIdx.x == 64) p.sem[tile] = 0;
            }
        }
        tc_fence_before();
    }

    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ---- load-time re-layout kernels (replace device_impl.py:392-479 preprocess_weights_for_mixed_gemm): reference
// (un-permuted) tensors -> per-(n-tile, k-block) blobs consumed above.
// int4: q_packed [K][N/2] (byte = hi nibble col 2j+1, lo nibble col 2j, two's complement q_s), scales/zs [K/128][N] (16-bit).
__global__ void pack_w4_kernel(const uint8_t* __restrict__ q_packed, const uint16_t* __restrict__ scales,
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
__global__ void pack_w8_kernel(const int8_t* __restrict__ q, int K, int N, uint8_t* __restrict__ blob) {
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

}  // namespace b200
