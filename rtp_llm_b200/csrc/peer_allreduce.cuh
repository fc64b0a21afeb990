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
    uint8_t* slots[kArMaxWorld];    // slots[r]: base of rank r's receive area for THIS call parity: [src W][n16_max * 32 bytes]
    uint8_t* slots2[kArMaxWorld];   // second receive area (two-shot: the all-gather of the reduced slices)
    uint32_t* epoch;                // local {call counter, finished-CTA ticket}: advanced in-kernel (graph-replay safe)
    size_t src_stride;              // bytes between the per-source slots
    int n16;                        // number of 16-byte payload chunks
    int rank, world;
};

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
__device__ __forceinline__ uint4 poll_chunk(const uint8_t* q, uint32_t epoch) {
    uint4 lo, hi;
    do {
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
                push_chunk(p.slots[r] + (size_t)p.rank * p.src_stride + (size_t)i * 32, v, epoch);
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
                const uint4 g = poll_chunk(p.slots[p.rank] + (size_t)r * p.src_stride + (size_t)i * 32, epoch);
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
                push_chunk(p.slots[r] + (size_t)p.rank * p.src_stride + (size_t)j * 32, src[(size_t)r * ns + j], epoch);
    }
    // 2) reduce my slice in rank order, store it locally and push it to every peer's all-gather slot [src = rank]
    for (int j = j0 + threadIdx.x; j < j1; j += kArThreads) {
        float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int s = 0; s < kArMaxWorld; ++s) {
            if (s >= p.world) continue;
            const uint4 g = (s == p.rank) ? src[(size_t)p.rank * ns + j]
                                          : poll_chunk(p.slots[p.rank] + (size_t)s * p.src_stride + (size_t)j * 32, epoch);
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
            if (r < p.world && r != p.rank) push_chunk(p.slots2[r] + (size_t)p.rank * p.src_stride + (size_t)j * 32, o, epoch);
    }
    // 3) gather the slices the other ranks reduced
    for (int j = j0 + threadIdx.x; j < j1; j += kArThreads) {
#pragma unroll
        for (int s = 0; s < kArMaxWorld; ++s)
            if (s < p.world && s != p.rank)
                dst[(size_t)s * ns + j] = poll_chunk(p.slots2[p.rank] + (size_t)s * p.src_stride + (size_t)j * 32, epoch);
    }
    // the last CTA to finish advances the call counter (every CTA of this launch has read it by then; the next launch reads
    // it only after griddepcontrol.wait / stream order)
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

}  // namespace b200
