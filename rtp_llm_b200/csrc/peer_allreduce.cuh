// One-shot SUM all-reduce of a small [rows][hidden] fp16/bf16 tensor over NVLink peer memory, for the TP all-reduce after
// the row-parallel GEMMs (reference: rtp_llm/models_py/distributed/collective_torch.py:694-722; its fast path is torch
// symmetric memory one-/two-shot + multimem, symm_mem.py:126-185). 256 KiB messages are pure latency, so the exchange is
// PUSH-based with the flag travelling inside the data ("LL" style): every rank writes its contribution, as 8-byte
// {data32, epoch32} words, straight into a per-source slot of every peer's region, then polls its OWN memory until all
// sources carry the current epoch and sums them in rank order 0..W-1 (same order everywhere -> bit-identical results on all
// ranks). One NVLink traversal, no separate barrier, no remote reads.
// Slot reuse: two parity slots alternate between consecutive calls; a peer can only start call k+2 after it completed
// call k+1, which needed this rank's call-k+1 data, which this rank sends only after finishing call k -- so the slot of
// call k is no longer being read when call k+2 overwrites it.
#pragma once
#include "ptx.cuh"

namespace b200 {

constexpr int kArMaxWorld = 8;
constexpr int kArMaxCtas = 64;
constexpr int kArThreads = 256;

struct PeerArParams {
    const void* in;                 // local contribution [n16 * 16 bytes]
    void* out;                      // local result (may alias `in`)
    uint8_t* region[kArMaxWorld];   // region[r]: base of rank r's peer-visible region: [2 parities][2 areas][W sources][LL slot]
    uint32_t* epoch;                // local {call counter, finished-CTA ticket}: advanced in-kernel (graph-replay safe)
    size_t src_stride;              // bytes between the per-source slots
    size_t area_stride;             // bytes between the reduce-scatter / all-gather areas
    size_t parity_stride;           // bytes between the two parity copies
    int n16;                        // number of 16-byte payload chunks
    int rank, world;
    // fused all-reduce + residual add + RMSNorm (peer_allreduce_norm_kernel)
    void* residual;                 // [rows][hidden], updated in place
    const void* gamma;              // [hidden]
    void* y;                        // [rows][hidden] normalised output
    int rows, hidden;
    float eps;
    int pushed;                     // 1: the producing GEMM already pushed the reduce-scatter words (b200_wo_gemm_rs): skip phase 1
    // vocab-parallel greedy sampling (peer_argmax_kernel)
    const void* logits;             // [rows][vocab_local]
    int32_t* token_out;             // [rows]
    int vocab_local, vocab_total;
};
// The slot parity of a call is derived from the DEVICE-side call counter (same value on every rank), so consecutive calls
// alternate slots whatever the host does: replaying a CUDA graph that holds an odd number of all-reduces stays correct.
__device__ __forceinline__ uint8_t* ar_area(const PeerArParams& p, int r, uint32_t epoch, int area) {
    return p.region[r] + (size_t)(epoch & 1u) * p.parity_stride + (size_t)area * p.area_stride;
}

__device__ __forceinline__ void st_ll(uint8_t* p, uint32_t data, uint32_t flag) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(data), "r"(flag) : "memory");
}
// two LL words in one 16-byte store (each 8-byte half carries its own flag, so a split delivery is still safe)
__device__ __forceinline__ void st_ll2(uint8_t* p, uint32_t d0, uint32_t d1, uint32_t flag) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %2};" ::"l"(p), "r"(d0), "r"(flag), "r"(d1) : "memory");
}
__device__ __forceinline__ void push_chunk(uint8_t* dst, const uint4& v, uint32_t epoch) {
    st_ll2(dst, v.x, v.y, epoch);
    st_ll2(dst + 16, v.z, v.w, epoch);
}
// Liveness guard: a peer that never arrives (died, or issued a different call sequence) makes this rank trap after a few
// seconds of polling instead of hanging the GPU (NCCL relies on its host watchdog for the same purpose).
__device__ __forceinline__ uint4 poll_chunk(const uint8_t* q, uint32_t epoch) {
    uint4 lo, hi;
    uint32_t spins = 0;
    do {
        if (++spins > (1u << 24)) __trap();
        asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w) : "l"(q) : "memory");
        asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w) : "l"(q + 16) : "memory");
    } while (lo.y != epoch || lo.w != epoch || hi.y != epoch || hi.w != epoch);
    return make_uint4(lo.x, lo.z, hi.x, hi.z);
}
__device__ __forceinline__ uint4 ld_ll2(const uint8_t* p) {   // two LL words: {d0, f0, d1, f1}
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

template <typename T>
__device__ __forceinline__ void acc2(float& a0, float& a1, uint32_t w) {
    if (std::is_same<T, __half>::value) {
        float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
        a0 += f.x;
        a1 += f.y;
    } else {
        float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
        a0 += f.x;
        a1 += f.y;
    }
}

// shared tail: the last CTA to finish advances the call counter
__device__ __forceinline__ void ar_finish(const PeerArParams& p, uint32_t epoch, int nctas) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(p.epoch + 1, 1u) == (uint32_t)nctas - 1) {
            p.epoch[1] = 0;
            __threadfence();
            *reinterpret_cast<volatile uint32_t*>(p.epoch) = epoch;
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kArThreads) peer_allreduce_kernel(const PeerArParams p) {
    pdl_launch_dependents();
    pdl_wait();
    const int cta = blockIdx.x, nctas = gridDim.x;
    const int per = (p.n16 + nctas - 1) / nctas;
    const int c0 = cta * per, c1 = min(c0 + per, p.n16);
    const uint32_t epoch = *reinterpret_cast<volatile uint32_t*>(p.epoch) + 1;   // same value in every CTA and on every rank
    const uint4* src = reinterpret_cast<const uint4*>(p.in);

    // 1) push: my chunk range goes into slot [src = rank] of every peer, each 32-bit word tagged with the epoch
    for (int i = c0 + threadIdx.x; i < c1; i += kArThreads) {
        const uint4 v = src[i];
#pragma unroll
        for (int r = 0; r < kArMaxWorld; ++r) {
            if (r < p.world && r != p.rank) {
                push_chunk(ar_area(p, r, epoch, 0) + (size_t)p.rank * p.src_stride + (size_t)i * 32, v, epoch);
            }
        }
    }
    // 2) pull from my own memory: wait for every source's words of this epoch, sum in rank order
    uint4* dst = reinterpret_cast<uint4*>(p.out);
    for (int i = c0 + threadIdx.x; i < c1; i += kArThreads) {
        const uint4 mine = src[i];
        float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int r = 0; r < kArMaxWorld; ++r) {
            if (r >= p.world) continue;
            uint32_t w0, w1, w2, w3;
            if (r == p.rank) {
                w0 = mine.x; w1 = mine.y; w2 = mine.z; w3 = mine.w;
            } else {
                const uint4 g = poll_chunk(ar_area(p, p.rank, epoch, 0) + (size_t)r * p.src_stride + (size_t)i * 32, epoch);
                w0 = g.x; w1 = g.y; w2 = g.z; w3 = g.w;
            }
            acc2<T>(a[0], a[1], w0);
            acc2<T>(a[2], a[3], w1);
            acc2<T>(a[4], a[5], w2);
            acc2<T>(a[6], a[7], w3);
        }
        uint4 o;
        o.x = pack2<T>(a[0], a[1]);
        o.y = pack2<T>(a[2], a[3]);
        o.z = pack2<T>(a[4], a[5]);
        o.w = pack2<T>(a[6], a[7]);
        dst[i] = o;
    }
    // the last CTA to finish advances the call counter (every CTA of this launch has read it by then; the next launch reads
    // it only after griddepcontrol.wait / stream order)
    ar_finish(p, epoch, nctas);
}

// Two-shot variant for W >= 3 (the one-shot push moves (W-1) x the message per rank; at 8 ranks that is 3.6 MB of 8-byte
// flagged words per all-reduce): reduce-scatter + all-gather, both push-LL. Rank r owns slice r (n16/W chunks): phase 1
// pushes slice r of the local data to rank r; phase 2 sums slice `rank` over all sources in rank order and pushes the
// reduced slice to every peer; phase 3 collects the other reduced slices. A thread handles the same chunk offset j in all
// three phases, so no CTA-level synchronisation is needed; every slice is reduced by exactly one rank (identical bits
// everywhere).
template <typename T>
__global__ void __launch_bounds__(kArThreads) peer_allreduce_twoshot_kernel(const PeerArParams p) {
    pdl_launch_dependents();
    pdl_wait();
    const int cta = blockIdx.x, nctas = gridDim.x;
    const int ns = p.n16 / p.world;                        // chunks per slice (host guarantees divisibility)
    const int per = (ns + nctas - 1) / nctas;
    const int j0 = cta * per, j1 = min(j0 + per, ns);
    const uint32_t epoch = *reinterpret_cast<volatile uint32_t*>(p.epoch) + 1;   // same value in every CTA and on every rank
    const uint4* src = reinterpret_cast<const uint4*>(p.in);
    uint4* dst = reinterpret_cast<uint4*>(p.out);
    // 1) scatter: slice r of my data -> rank r's reduce-scatter slot [src = rank]
    for (int j = j0 + threadIdx.x; j < j1; j += kArThreads) {
#pragma unroll
        for (int r = 0; r < kArMaxWorld; ++r)
            if (r < p.world && r != p.rank)
                push_chunk(ar_area(p, r, epoch, 0) + (size_t)p.rank * p.src_stride + (size_t)j * 32, src[(size_t)r * ns + j], epoch);
    }
    // 2) reduce my slice in rank order, store it locally and push it to every peer's all-gather slot [src = rank]
    for (int j = j0 + threadIdx.x; j < j1; j += kArThreads) {
        float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int s = 0; s < kArMaxWorld; ++s) {
            if (s >= p.world) continue;
            const uint4 g = (s == p.rank) ? src[(size_t)p.rank * ns + j]
                                          : poll_chunk(ar_area(p, p.rank, epoch, 0) + (size_t)s * p.src_stride + (size_t)j * 32, epoch);
            acc2<T>(a[0], a[1], g.x);
            acc2<T>(a[2], a[3], g.y);
            acc2<T>(a[4], a[5], g.z);
            acc2<T>(a[6], a[7], g.w);
        }
        uint4 o;
        o.x = pack2<T>(a[0], a[1]);
        o.y = pack2<T>(a[2], a[3]);
        o.z = pack2<T>(a[4], a[5]);
        o.w = pack2<T>(a[6], a[7]);
        dst[(size_t)p.rank * ns + j] = o;
#pragma unroll
        for (int r = 0; r < kArMaxWorld; ++r)
            if (r < p.world && r != p.rank) push_chunk(ar_area(p, r, epoch, 1) + (size_t)p.rank * p.src_stride + (size_t)j * 32, o, epoch);
    }
    // 3) gather the slices the other ranks reduced
    for (int j = j0 + threadIdx.x; j < j1; j += kArThreads) {
#pragma unroll
        for (int s = 0; s < kArMaxWorld; ++s)
            if (s < p.world && s != p.rank)
                dst[(size_t)s * ns + j] = poll_chunk(ar_area(p, p.rank, epoch, 1) + (size_t)s * p.src_stride + (size_t)j * 32, epoch);
    }
    // the last CTA to finish advances the call counter (every CTA of this launch has read it by then; the next launch reads
    // it only after griddepcontrol.wait / stream order)
    ar_finish(p, epoch, nctas);
}

// Fused TP exchange of a row-parallel GEMM output: SUM all-reduce (two-shot, push-LL) + residual add + RMSNorm in ONE kernel
// (reference sequence: all_reduce(t, Group.TP) collective_torch.py:694-722, then fused_add_rmsnorm RegisterBaseBindings.hpp:54;
// call sites hybrid/causal_attention.py:91-92, dense_mlp.py:104-105 followed by the next layer norm). Numerics are those of
// the unfused sequence: the reduced value is rounded to T (what all_reduce stores), then x = float(sum) + float(residual)
// stays unrounded for the variance and the output, the stored residual is rounded.
// One CTA per row; a row has C = hidden/8 16-byte chunks, rank r owns the chunks [r*C/W, (r+1)*C/W) of every row.
template <typename T>
__global__ void __launch_bounds__(kArThreads) peer_allreduce_norm_kernel(const PeerArParams p) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ __align__(16) uint8_t ar_smem[];
    float* s_x = reinterpret_cast<float*>(ar_smem);                    // [hidden] fp32 row
    uint4* s_own = reinterpret_cast<uint4*>(s_x + p.hidden);           // [C / W] my reduced slice of this row
    __shared__ float s_red[kArThreads / 32];
    const int row = blockIdx.x, C = p.hidden / 8, Cs = C / p.world;
    const uint32_t epoch = *reinterpret_cast<volatile uint32_t*>(p.epoch) + 1;
    const uint4* src = reinterpret_cast<const uint4*>(p.in) + (size_t)row * C;
    // 1) reduce-scatter: chunk c of my row goes to its owner's slot [src = rank] -- unless the GEMM that produced the row
    //    already pushed it from its epilogue (b200_wo_gemm_rs: same slot layout, same epoch, same rounded bits)
    if (!p.pushed) {
        for (int c = threadIdx.x; c < C; c += kArThreads) {
            const int owner = c / Cs;
            if (owner != p.rank)
                push_chunk(ar_area(p, owner, epoch, 0) + (size_t)p.rank * p.src_stride + ((size_t)row * Cs + (c - owner * Cs)) * 32, src[c], epoch);
        }
    }
    // 2) reduce my slice in rank order, round to T, keep it and push it to every peer's all-gather slot [src = rank]
    for (int cs = threadIdx.x; cs < Cs; cs += kArThreads) {
        float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int s = 0; s < kArMaxWorld; ++s) {
            if (s >= p.world) continue;
            const uint4 g = (s == p.rank) ? src[p.rank * Cs + cs]
                                          : poll_chunk(ar_area(p, p.rank, epoch, 0) + (size_t)s * p.src_stride + ((size_t)row * Cs + cs) * 32, epoch);
            acc2<T>(a[0], a[1], g.x);
            acc2<T>(a[2], a[3], g.y);
            acc2<T>(a[4], a[5], g.z);
            acc2<T>(a[6], a[7], g.w);
        }
        uint4 o;
        o.x = pack2<T>(a[0], a[1]);
        o.y = pack2<T>(a[2], a[3]);
        o.z = pack2<T>(a[4], a[5]);
        o.w = pack2<T>(a[6], a[7]);
        s_own[cs] = o;
#pragma unroll
        for (int r = 0; r < kArMaxWorld; ++r)
            if (r < p.world && r != p.rank)
                push_chunk(ar_area(p, r, epoch, 1) + (size_t)p.rank * p.src_stride + ((size_t)row * Cs + cs) * 32, o, epoch);
    }
    __syncthreads();
    // 3) gather the reduced row, add the residual, RMSNorm
    uint4* res = reinterpret_cast<uint4*>(p.residual) + (size_t)row * C;
    float ss = 0.f;
    for (int c = threadIdx.x; c < C; c += kArThreads) {
        const int owner = c / Cs;
        const uint4 v = owner == p.rank ? s_own[c - owner * Cs]
                                        : poll_chunk(ar_area(p, p.rank, epoch, 1) + (size_t)owner * p.src_stride + ((size_t)row * Cs + (c - owner * Cs)) * 32, epoch);
        const uint4 r4 = res[c];
        const uint32_t vw[4] = {v.x, v.y, v.z, v.w}, rw[4] = {r4.x, r4.y, r4.z, r4.w};
        uint32_t ow[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
            acc2<T>(a0, a1, vw[j]);
            acc2<T>(b0, b1, rw[j]);
            const float x0 = a0 + b0, x1 = a1 + b1;
            s_x[c * 8 + 2 * j] = x0;
            s_x[c * 8 + 2 * j + 1] = x1;
            ss = fmaf(x0, x0, fmaf(x1, x1, ss));
            ow[j] = pack2<T>(x0, x1);
        }
        res[c] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < kArThreads / 32; ++w) tot += s_red[w];
    const float inv = rsqrtf(tot / (float)p.hidden + p.eps);
    const uint4* gv = reinterpret_cast<const uint4*>(p.gamma);
    uint4* yv = reinterpret_cast<uint4*>(p.y) + (size_t)row * C;
    for (int c = threadIdx.x; c < C; c += kArThreads) {
        const uint4 g4 = gv[c];
        const uint32_t gw[4] = {g4.x, g4.y, g4.z, g4.w};
        uint32_t ow[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float g0 = 0.f, g1 = 0.f;
            acc2<T>(g0, g1, gw[j]);
            ow[j] = pack2<T>(s_x[c * 8 + 2 * j] * inv * g0, s_x[c * 8 + 2 * j + 1] * inv * g1);
        }
        yv[c] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
    ar_finish(p, epoch, gridDim.x);
}

// Vocab-parallel greedy sampling: every rank takes the argmax of its logits shard, the (value, global index) pairs travel as
// LL words to every peer, and every rank picks the overall winner (largest value, lowest index on ties -- torch.argmax
// semantics over the concatenated vocabulary, padded columns >= vocab_total excluded). Replaces the logits all-gather +
// argmax of the reference (cpp/models/PyWrappedModel.cc:915-936,1001-1052, CudaSampleOp.cc:453) for top_k == 1.
template <typename T>
__global__ void __launch_bounds__(kArThreads) peer_argmax_kernel(const PeerArParams p) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float s_v[kArThreads / 32];
    __shared__ int s_i[kArThreads / 32];
    const int row = blockIdx.x;
    const uint32_t epoch = *reinterpret_cast<volatile uint32_t*>(p.epoch) + 1;
    const T* lg = reinterpret_cast<const T*>(p.logits) + (size_t)row * p.vocab_local;
    const int base = p.rank * p.vocab_local;
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < p.vocab_local; i += kArThreads) {
        const int gi = base + i;
        if (gi >= p.vocab_total) break;
        const float v = to_f32<T>(lg[i]);
        if (v > bv || (v == bv && gi < bi)) {
            bv = v;
            bi = gi;
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) {
            bv = ov;
            bi = oi;
        }
    }
    if ((threadIdx.x & 31) == 0) {
        s_v[threadIdx.x >> 5] = bv;
        s_i[threadIdx.x >> 5] = bi;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        bv = threadIdx.x < kArThreads / 32 ? s_v[threadIdx.x] : -INFINITY;
        bi = threadIdx.x < kArThreads / 32 ? s_i[threadIdx.x] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) {
                bv = ov;
                bi = oi;
            }
        }
        // lanes 0..W-1: lane r pushes my pair to rank r; then lane s polls source s
        const int lane = threadIdx.x;
        if (lane < p.world && lane != p.rank)
            st_ll2(ar_area(p, lane, epoch, 0) + (size_t)p.rank * p.src_stride + (size_t)row * 16, __float_as_uint(bv), (uint32_t)bi, epoch);
        float cv = -INFINITY;
        int ci = 0x7fffffff;
        if (lane < p.world) {
            if (lane == p.rank) {
                cv = bv;
                ci = bi;
            } else {
                const uint8_t* q = ar_area(p, p.rank, epoch, 0) + (size_t)lane * p.src_stride + (size_t)row * 16;
                uint4 w;
                uint32_t spins = 0;
                do {
                    if (++spins > (1u << 24)) __trap();
                    w = ld_ll2(q);
                } while (w.y != epoch || w.w != epoch);
                cv = __uint_as_float(w.x);
                ci = (int)w.z;
            }
        }
#pragma unroll
        for (int o = 4; o; o >>= 1) {      // world <= 8
            const float ov = __shfl_xor_sync(0xffffffffu, cv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, ci, o);
            if (ov > cv || (ov == cv && oi < ci)) {
                cv = ov;
                ci = oi;
            }
        }
        if (lane == 0) p.token_out[row] = ci;
    }
    ar_finish(p, epoch, gridDim.x);
}

}  // namespace b200
