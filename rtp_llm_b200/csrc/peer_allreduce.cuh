// One-shot SUM all-reduce of a small [rows][hidden] fp16/bf16 tensor over NVLink peer memory, for the TP all-reduce after
// the row-parallel GEMMs (reference: rtp_llm/models_py/distributed/collective_torch.py:694-722; its fast path is torch
// symmetric memory one-/two-shot + multimem, symm_mem.py:126-185). 256 KiB messages are latency-bound: every rank copies
// its contribution into its own peer-visible slot, raises one flag per (peer, CTA), then reads all W slots over NVLink
// and sums them in rank order 0..W-1 -- the same order on every rank, so all ranks produce bit-identical results.
// No second barrier: two slots alternate between consecutive calls, and a peer can only raise its flag for call k+1
// after it has finished reading call k (stream order), so slot (k & 1) is free again when call k+2 writes it.
#pragma once
#include "ptx.cuh"

namespace b200 {

constexpr int kArMaxWorld = 8;
constexpr int kArMaxCtas = 64;
constexpr int kArThreads = 256;

struct PeerArParams {
    const void* in;                 // local contribution [n16 * 16 bytes]
    void* out;                      // local result (may alias `in`)
    uint8_t* slot[kArMaxWorld];     // peer-visible data slot of every rank for THIS call (rank r's own entry is local memory)
    uint32_t* flags[kArMaxWorld];   // flags[r] = rank r's flag array [kArMaxWorld][kArMaxCtas] (we write ours into row `rank`)
    uint32_t* epoch;                // local [kArMaxCtas] call counters, one per CTA (graph-replay safe: advanced in-kernel)
    int n16;                        // number of 16-byte chunks
    int rank, world;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// peer lines may sit stale in the local L1 from two calls ago (same slot): volatile loads always go to the owner
__device__ __forceinline__ uint4 ld_volatile_v4(const uint4* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

template <typename T>
__device__ __forceinline__ void acc8(float (&a)[8], const uint4& v) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (sizeof(T) == 2 && std::is_same<T, __half>::value) {
            float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
            a[2 * i] += f.x;
            a[2 * i + 1] += f.y;
        } else {
            float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
            a[2 * i] += f.x;
            a[2 * i + 1] += f.y;
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
    const uint32_t epoch = p.epoch[cta] + 1;

    // 1) publish this CTA's chunk range in the local peer-visible slot
    const uint4* src = reinterpret_cast<const uint4*>(p.in);
    uint4* mine = reinterpret_cast<uint4*>(p.slot[p.rank]);
    for (int i = c0 + threadIdx.x; i < c1; i += kArThreads) mine[i] = src[i];
    __syncthreads();
    // 2) one flag per (peer, CTA): "rank `rank`, CTA `cta` has published call `epoch`". The release store of a thread that
    //    passed the CTA barrier is cumulative over the other threads' stores -- no per-thread system fence needed.
    if (threadIdx.x < p.world) st_release_sys(p.flags[threadIdx.x] + p.rank * kArMaxCtas + cta, epoch);
    // 3) wait until every rank's matching CTA has published
    if (threadIdx.x < p.world) {
        const uint32_t* f = p.flags[p.rank] + threadIdx.x * kArMaxCtas + cta;
        while ((int32_t)(ld_acquire_sys(f) - epoch) < 0) {
        }
    }
    __syncthreads();
    // 4) sum the W slots in rank order (identical on every rank)
    uint4* dst = reinterpret_cast<uint4*>(p.out);
    for (int i = c0 + threadIdx.x; i < c1; i += kArThreads) {
        uint4 v[kArMaxWorld];
#pragma unroll
        for (int r = 0; r < kArMaxWorld; ++r)
            if (r < p.world) v[r] = ld_volatile_v4(reinterpret_cast<const uint4*>(p.slot[r]) + i);
        float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int r = 0; r < kArMaxWorld; ++r)
            if (r < p.world) acc8<T>(a, v[r]);
        uint4 o;
        o.x = pack2<T>(a[0], a[1]);
        o.y = pack2<T>(a[2], a[3]);
        o.z = pack2<T>(a[4], a[5]);
        o.w = pack2<T>(a[6], a[7]);
        dst[i] = o;
    }
    if (threadIdx.x == 0) p.epoch[cta] = epoch;
}

}  // namespace b200
