// Paged GQA decode attention for sm_100a.
//
// Replaces (same math, same cache layout, same page-list format):
//   runXqa                      /root/reference/rtp_llm/models_py/bindings/cuda/ops/CudaXqa.h:65-84 (+CudaXqa.cc:124-221)
//   XQA sm_90a kernel           /root/reference/3rdparty/xqa/mha_sm90.cu (cannot run on B200)
//   FlashInferTRTLLMDecodeOp    /root/reference/rtp_llm/models_py/modules/factory/attention/cuda_impl/trtllm_gen.py:503-546
//
// Work decomposition (B200-first, not a port of XQA's warpgroup pipeline):
//   grid = (nsplit, B*Hkv); one CTA owns (sequence b, kv head, a contiguous run of 64-token tiles).
//   Warp 4 lane 0 = TMA producer: reads the K/V page ids of each tile from the page list and issues 4-D tensor-map
//   loads (128-byte swizzle, one box per page and per 64-channel half) into a 3-stage shared-memory ring.
//   Warps 0-3 = consumers: each takes 16 of the tile's 64 tokens, S = Q.K^T and O += P.V on the legacy tensor path
//   (ldmatrix + mma.sync m16n8k16; the whole GQA group is packed into the MMA M dimension so every K/V byte is read
//   from HBM exactly once), online softmax with quad-level shuffles, private running (m, l, O) per warp, merged
//   through shared memory at the end.  Sequence splits are merged by the last-arriving CTA (semaphore), XQA-style.
//   The kernel is HBM-bound: per CTA 96 KB of K/V in flight, two CTAs per SM.
#pragma once
#include "ptx.cuh"

namespace b200 {

constexpr int kAttnTile = 64;          // tokens per pipeline stage
constexpr int kAttnStages = 3;
constexpr int kAttnD = 128;            // head dim
constexpr int kAttnConsumerWarps = 4;
constexpr int kAttnThreads = (kAttnConsumerWarps + 1) * 32;
// per head dim D (64 / 128 / 256): a stage holds the D/64 64-channel halves of K, then those of V, 8 KB each (64 tokens x 128 B)
__host__ __device__ constexpr int attn_stage_bytes(int D) { return 2 * (D / 64) * kAttnTile * 128; }
__host__ __device__ constexpr int attn_smem_bytes(int D) {
    return kAttnStages * attn_stage_bytes(D) + 1024 /*align*/ + 256 /*barriers etc*/ + 512 /*rope cos/sin table*/;
}
constexpr int kAttnSmemBytes = attn_smem_bytes(kAttnD);

struct AttnParams {
    const void* q;            // [B][Hq][D]
    void* out;                // [B][Hq*D]
    const int32_t* page_list; // [B][1][2][M]
    const int32_t* seq_lens;  // [B] tokens already cached (the new token sits at this index)
    int B, Hq, Hkv, group, M;
    int q_len;                // query tokens per sequence (1 = plain decode; > 1 = verification of speculative tokens, XQA's
                              // max_q_len / trtllm-gen's q_len_per_req): q [B][q_len][Hq][D], query j attends to positions
                              // 0 .. seq_lens[b] + j, its GQA rows sit at MMA rows j * group + g (group * q_len <= 16)
    int T, log2T;             // tokens per page
    int box_h, boxes_per_tile;
    int nsplit, tiles_per_split;
    float scale_log2;         // q_scale * D^-1/2 * log2(e)
    float* ws_o;              // [B*Hkv][nsplit][group][D]
    float* ws_ml;             // [B*Hkv][nsplit][group][2]
    int* sem;                 // [B*Hkv], zero on entry, zero on exit
    // fused RoPE + KV append (XQA's USE_INPUT_KV + ROPE_STYLE, 3rdparty/xqa/mha.h:82-86): when qkv != null the kernel reads the
    // un-rotated q | k | v of the new token from the qkv GEMM output [B][(Hq+2Hkv)*D], rotates q on load, and the CTA that owns
    // the tile of position seq_lens[b] rotates k and appends k, v to the paged cache before streaming it. Same arithmetic as
    // rope_append_kernel (RopeStyle::Base, NeoX pairing) -> bit-identical cache contents and attention output.
    const void* qkv;
    void* kv_pool_rw;
    float log2_base;
    int cluster_merge;        // 1: the nsplit (<= 8) CTAs of a (sequence, kv head) form a thread-block cluster (nsplit,1,1) and merge
                              //    their partial (m, l, O) through distributed shared memory; no workspace, no semaphore
};

template <typename T, int D = kAttnD>
__global__ void __launch_bounds__(kAttnThreads, D <= 128 ? 2 : 1)
paged_decode_attn_kernel(const __grid_constant__ CUtensorMap kv_map, const AttnParams p) {
    constexpr int NH = D / 64;                       // 64-channel halves per K / V row
    constexpr int STAGE = attn_stage_bytes(D);
    constexpr int OROW = D + 8;                      // floats; +8 banks per row -> conflict-free float2 stores
    static_assert(D == 64 || D == 128 || D == 256, "head_dim 64 / 128 / 256");
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment for the 128B-swizzled boxes
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kAttnStages * STAGE);
    uint64_t* empty_bar = full_bar + kAttnStages;
    int* s_flag = reinterpret_cast<int*>(empty_bar + kAttnStages);
    float2* cs_tab = reinterpret_cast<float2*>(smem + kAttnStages * STAGE + 256);   // fused rope: (cos, sin) of the 64 channel pairs

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int split = blockIdx.x, bh = blockIdx.y;
    const int b = bh / p.Hkv, kvh = bh % p.Hkv;
    pdl_launch_dependents();   // let the o-projection GEMM start prefetching its weights
    pdl_wait();                // q and the appended K/V come from the previous kernel

    // the host sized the split for max_seq_len: a longer sequence is clamped to that bound (never more tiles than the
    // splits cover, so the last-arriver count below always matches the CTAs that really contribute)
    const int len = min(p.seq_lens[b] + p.q_len, p.nsplit * p.tiles_per_split * kAttnTile);
    const int rows = p.group * p.q_len;
    const int ntiles = (len + kAttnTile - 1) / kAttnTile;
    const int t0 = split * p.tiles_per_split;
    const int t1 = min(t0 + p.tiles_per_split, ntiles);
    const bool empty = t0 >= t1;                 // this split holds no tokens of this sequence (uniform for the CTA)
    if (empty && !p.cluster_merge) return;
    const int nact = (ntiles + p.tiles_per_split - 1) / p.tiles_per_split;  // CTAs that contribute to (b, kvh)

    if (threadIdx.x == 0) {
        for (int s = 0; s < kAttnStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kAttnConsumerWarps);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == kAttnConsumerWarps) {
        // ------------------------------------------------------------------ TMA producer
        // Tile issue is split in two so that, with the fused rope, the append of the new token (all 32 lanes) overlaps the first
        // loads in flight: lane 0 first fills the ring with tiles that cannot contain the new token, the warp appends, then
        // lane 0 streams the rest (the token's tile is the sequence's last tile).
        const int32_t* kpl = p.page_list + (size_t)(b * 2 + 0) * p.M;
        const int32_t* vpl = p.page_list + (size_t)(b * 2 + 1) * p.M;
        auto issue = [&](int i) {
            const int it = i - t0, s = it % kAttnStages;
            const uint32_t ph = (it / kAttnStages) & 1;
            // page ids first (global loads overlap the wait below)
            int kpage[4], vpage[4], inpage[4];
            int nbox = 0;
#pragma unroll
            for (int bx = 0; bx < 4; ++bx) {
                const int tok = i * kAttnTile + bx * p.box_h;
                if (bx < p.boxes_per_tile && tok < len) {
                    const int pidx = tok >> p.log2T;
                    kpage[bx] = kpl[pidx];
                    vpage[bx] = vpl[pidx];
                    inpage[bx] = tok & (p.T - 1);
                    nbox = bx + 1;
                }
            }
            mbar_wait(&empty_bar[s], ph ^ 1);
            uint8_t* stage = smem + s * STAGE;
            mbar_arrive_expect_tx(&full_bar[s], (uint32_t)(nbox * 2 * NH * p.box_h * 128));
#pragma unroll
            for (int bx = 0; bx < 4; ++bx) {
                if (bx < nbox) {
                    const int row_off = bx * p.box_h * 128;
#pragma unroll
                    for (int h = 0; h < NH; ++h) {
                        tma_load_4d(stage + h * 8192 + row_off, &kv_map, 64 * h, inpage[bx], kvh, kpage[bx], &full_bar[s]);
                        tma_load_4d(stage + (NH + h) * 8192 + row_off, &kv_map, 64 * h, inpage[bx], kvh, vpage[bx], &full_bar[s]);
                    }
                }
            }
        };
        int first_end = t1;                               // tiles [t0, first_end) are issued before the append
        bool own = false;
        int pos = 0;
        if (D == 128 && p.qkv != nullptr) {   // fused rope is built for head_dim 128 (the host rejects other sizes)
            pos = p.seq_lens[b];
            own = pos < len && (pos / kAttnTile) >= t0 && (pos / kAttnTile) < t1;      // this CTA streams the new token's tile
            if (own) first_end = min(min(t0 + kAttnStages, t1), pos / kAttnTile);
        }
        if (lane == 0) {
            tma_prefetch_desc(&kv_map);
            for (int i = t0; i < first_end; ++i) issue(i);
        }
        __syncwarp();
        if (own) {
            const T* row = reinterpret_cast<const T*>(p.qkv) + (size_t)b * (p.Hq + 2 * p.Hkv) * kAttnD;
            const T* ksrc = row + (size_t)(p.Hq + kvh) * kAttnD;
            const T* vsrc = row + (size_t)(p.Hq + p.Hkv + kvh) * kAttnD;
            const size_t page_elems = (size_t)p.Hkv * p.T * kAttnD;
            const int32_t kpage = kpl[pos >> p.log2T], vpage = vpl[pos >> p.log2T];
            T* pool = reinterpret_cast<T*>(p.kv_pool_rw);
            T* kdst = pool + (size_t)kpage * page_elems + ((size_t)kvh * p.T + (pos & (p.T - 1))) * kAttnD;
            T* vdst = pool + (size_t)vpage * page_elems + ((size_t)kvh * p.T + (pos & (p.T - 1))) * kAttnD;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int i = lane + 32 * j;                        // channel pair (i, i + 64)
                const float x0 = to_f32<T>(ksrc[i]), x1 = to_f32<T>(ksrc[i + 64]);
                const float inv_freq = exp2f(-2.0f * (float)i / (float)kAttnD * p.log2_base);
                float sn, cs;
                sincosf((float)pos * inv_freq, &sn, &cs);
                float r0, r1;
                rope_rotate(x0, x1, cs, sn, r0, r1);
                kdst[i] = from_f32<T>(r0);
                kdst[i + 64] = from_f32<T>(r1);
                vdst[i] = vsrc[i];
                vdst[i + 64] = vsrc[i + 64];
            }
            __threadfence();              // every lane's stores are performed before lane 0 (after the warp barrier) ...
            fence_proxy_async_global();   // ... issues the TMA loads of the token's tile, which read them through the async proxy
        }
        __syncwarp();
        if (own && lane == 0) fence_proxy_async_global();
        if (lane == 0) {
            for (int i = first_end; i < t1; ++i) issue(i);
        }
        if (!p.cluster_merge) return;  // the producer warp takes no part in the merge (named barrier 1 counts 128 threads)
        cluster_sync_all();            // ... but every thread of the cluster passes the two cluster barriers of the DSMEM merge
        cluster_sync_all();
        return;
    }

    // ---------------------------------------------------------------------- consumers
    const int qrow = lane >> 2;          // fragment row (and row + 8)
    const int qcol = (lane & 3) * 2;     // fragment column pair
    // Q fragments for the whole head dim: 8 k-steps x 4 regs. Rows >= group are zero.
    uint32_t qf[D / 16][4];
    bool q_loaded = false;
    if constexpr (D == 128) {
      if (p.qkv) {
        q_loaded = true;
        // un-rotated q straight from the qkv GEMM output; the NeoX partner of column c < 64 is column c + 64 = the same
        // register index four k-steps later, so the rotation stays inside a thread
        const T* qbase = reinterpret_cast<const T*>(p.qkv) + ((size_t)b * (p.Hq + 2 * p.Hkv) + (size_t)kvh * p.group) * kAttnD;
        const bool v0 = qrow < p.group, v1 = (qrow + 8) < p.group;
        const int pos = p.seq_lens[b];
        if (threadIdx.x < 64) {                              // one sincosf per channel pair and CTA, shared through smem
            const float inv_freq = exp2f(-2.0f * (float)threadIdx.x / (float)kAttnD * p.log2_base);
            float sn, cs;
            sincosf((float)pos * inv_freq, &sn, &cs);
            cs_tab[threadIdx.x] = make_float2(cs, sn);
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int hc = 0; hc < 2; ++hc) {                 // column groups c and c + 8 of the k-step
                const int c = kk * 16 + qcol + hc * 8;
                const float2 t0c = cs_tab[c], t1c = cs_tab[c + 1];
                const float cs[2] = {t0c.x, t1c.x}, sn[2] = {t0c.y, t1c.y};
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {             // fragment rows qrow and qrow + 8
                    const bool ok = rr ? v1 : v0;
                    uint32_t lo = 0u, hi = 0u;
                    if (ok) {
                        const T* src = qbase + (size_t)(qrow + 8 * rr) * kAttnD + c;
                        const float a0 = to_f32<T>(src[0]), a1 = to_f32<T>(src[1]);          // x0 (columns c, c+1)
                        const float b0 = to_f32<T>(src[64]), b1 = to_f32<T>(src[65]);        // x1 (columns c+64, c+65)
                        float l0r, h0r, l1r, h1r;
                        rope_rotate(a0, b0, cs[0], sn[0], l0r, h0r);
                        rope_rotate(a1, b1, cs[1], sn[1], l1r, h1r);
                        lo = pack2<T>(l0r, l1r);
                        hi = pack2<T>(h0r, h1r);
                    }
                    qf[kk][hc * 2 + rr] = lo;
                    qf[kk + 4][hc * 2 + rr] = hi;
                }
            }
        }
      }
    }
    if (!q_loaded) {
        // MMA row R = j * group + g  <->  query token j, head kvh * group + g
        const bool v0 = qrow < rows, v1 = (qrow + 8) < rows;
        const T* qall = reinterpret_cast<const T*>(p.q);
        const T* q0 = qall + (((size_t)b * p.q_len + qrow / p.group) * p.Hq + (size_t)kvh * p.group + qrow % p.group) * D;
        const T* q1 = qall + (((size_t)b * p.q_len + (qrow + 8) / p.group) * p.Hq + (size_t)kvh * p.group + (qrow + 8) % p.group) * D;
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
            const int c = kk * 16 + qcol;
            qf[kk][0] = v0 ? *reinterpret_cast<const uint32_t*>(q0 + c) : 0u;
            qf[kk][1] = v1 ? *reinterpret_cast<const uint32_t*>(q1 + c) : 0u;
            qf[kk][2] = v0 ? *reinterpret_cast<const uint32_t*>(q0 + c + 8) : 0u;
            qf[kk][3] = v1 ? *reinterpret_cast<const uint32_t*>(q1 + c + 8) : 0u;
        }
    }

    float o[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY;  // running max (raw score units) for rows qrow / qrow+8
    float l0 = 0.f, l1 = 0.f;              // per-thread partial row sums (quad-reduced at the end)
    const float sl2 = p.scale_log2;
    const int off0 = p.q_len - 1 - min(qrow / p.group, p.q_len - 1), off1 = p.q_len - 1 - min((qrow + 8) / p.group, p.q_len - 1);

    // ldmatrix lane geometry
    const int mi = lane >> 3, mr = lane & 7;
    const int k_tokrow = warp * 16 + (mi >> 1) * 8 + mr;  // K: matrices (tok0-7,c),(tok0-7,c+1),(tok8-15,c),(tok8-15,c+1)
    const int k_chunk_add = mi & 1;
    const int v_tokrow = warp * 16 + (mi & 1) * 8 + mr;   // V (trans): (tok0-7,c),(tok8-15,c),(tok0-7,c+1),(tok8-15,c+1)
    const int v_chunk_add = mi >> 1;

    for (int i = t0; i < t1; ++i) {
        const int it = i - t0, s = it % kAttnStages;
        const uint32_t ph = (it / kAttnStages) & 1;
        mbar_wait(&full_bar[s], ph);
        uint8_t* stage = smem + s * STAGE;
        const uint32_t k_base = smem_u32(stage), v_base = k_base + NH * 8192;

        const int tok0 = i * kAttnTile + warp * 16;
        const int valid = len - tok0;  // tokens of this warp's slice that exist (may be <= 0)
        if (valid < 16) {
            // Rows past the end of the sequence hold stale / never-written bytes: P is 0 there but 0 * NaN = NaN,
            // so the V rows are zeroed before use (K garbage is masked on the scores below).
            const int first = valid < 0 ? 0 : valid;
            for (int r = first; r < 16; ++r) {
                uint8_t* row = stage + NH * 8192 + (warp * 16 + r) * 128;
#pragma unroll
                for (int h = 0; h < NH; ++h) reinterpret_cast<uint32_t*>(row + h * 8192)[lane] = 0u;
            }
            __syncwarp();
        }

        // ---- S = Q K^T for 16 heads x 16 tokens
        float sc[2][4];
#pragma unroll
        for (int j = 0; j < 2; ++j) sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk) {
            const int half = kk >> 2, c = (kk & 3) * 2 + k_chunk_add;
            uint32_t kb[4];
            ldmatrix_x4(kb, k_base + half * 8192 + k_tokrow * 128 + ((c ^ (k_tokrow & 7)) << 4));
            mma_m16n8k16<T>(sc[0], qf[kk], kb[0], kb[1]);
            mma_m16n8k16<T>(sc[1], qf[kk], kb[2], kb[3]);
        }

        // ---- mask + online softmax (rows qrow and qrow+8)
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                // query j of a q_len-token step sees q_len - 1 - j fewer tokens than the last one (causal among the new tokens)
                const bool ok0 = (j * 8 + qcol + e) < valid - off0, ok1 = (j * 8 + qcol + e) < valid - off1;
                sc[j][e] = ok0 ? sc[j][e] : -INFINITY;
                sc[j][2 + e] = ok1 ? sc[j][2 + e] : -INFINITY;
                mx0 = fmaxf(mx0, sc[j][e]);
                mx1 = fmaxf(mx1, sc[j][2 + e]);
            }
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
        // a fully masked slice keeps m = -inf; use 0 as the subtrahend so exp2(-inf - 0) = 0 (no NaN)
        const float ms0 = (mn0 == -INFINITY) ? 0.f : mn0 * sl2, ms1 = (mn1 == -INFINITY) ? 0.f : mn1 * sl2;
        const float a0 = fast_exp2(m0 * sl2 - ms0), a1 = fast_exp2(m1 * sl2 - ms1);  // m = -inf -> 0
        m0 = mn0;
        m1 = mn1;
        float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                sc[j][e] = fast_exp2(fmaf(sc[j][e], sl2, -ms0));
                sc[j][2 + e] = fast_exp2(fmaf(sc[j][2 + e], sl2, -ms1));
                ps0 += sc[j][e];
                ps1 += sc[j][2 + e];
            }
        }
        l0 = fmaf(l0, a0, ps0);
        l1 = fmaf(l1, a1, ps1);
        if (a0 != 1.f || a1 != 1.f) {
#pragma unroll
            for (int n = 0; n < D / 8; ++n) {
                o[n][0] *= a0;
                o[n][1] *= a0;
                o[n][2] *= a1;
                o[n][3] *= a1;
            }
        }
        // P (rounded to the cache element type, 3rdparty/xqa/ref.py:80) as the A fragment of the second product
        uint32_t pf[4];
        pf[0] = pack2<T>(sc[0][0], sc[0][1]);
        pf[1] = pack2<T>(sc[0][2], sc[0][3]);
        pf[2] = pack2<T>(sc[1][0], sc[1][1]);
        pf[3] = pack2<T>(sc[1][2], sc[1][3]);

        // ---- O += P V
#pragma unroll
        for (int n2 = 0; n2 < D / 16; ++n2) {
            const int half = n2 >> 2, c = (n2 & 3) * 2 + v_chunk_add;
            uint32_t vb[4];
            ldmatrix_x4_trans(vb, v_base + half * 8192 + v_tokrow * 128 + ((c ^ (v_tokrow & 7)) << 4));
            mma_m16n8k16<T>(o[2 * n2], pf, vb[0], vb[1]);
            mma_m16n8k16<T>(o[2 * n2 + 1], pf, vb[2], vb[3]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);
    }

    // quad-reduce the row sums
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);

    // ---------------------------------------------------------------------- merge the 4 warps through smem
    asm volatile("bar.sync 1, 128;" ::: "memory");  // everyone is done reading the K/V ring
    float* o_s = reinterpret_cast<float*>(smem);                              // [4][16][OROW]
    float* m_s = o_s + kAttnConsumerWarps * 16 * OROW;                        // [4][16]
    float* l_s = m_s + kAttnConsumerWarps * 16;                               // [4][16]
    {
        float* ow = o_s + (size_t)warp * 16 * OROW;
#pragma unroll
        for (int n = 0; n < D / 8; ++n) {
            *reinterpret_cast<float2*>(ow + qrow * OROW + n * 8 + qcol) = make_float2(o[n][0], o[n][1]);
            *reinterpret_cast<float2*>(ow + (qrow + 8) * OROW + n * 8 + qcol) = make_float2(o[n][2], o[n][3]);
        }
        if ((lane & 3) == 0) {
            m_s[warp * 16 + qrow] = m0;
            m_s[warp * 16 + qrow + 8] = m1;
            l_s[warp * 16 + qrow] = l0;
            l_s[warp * 16 + qrow + 8] = l1;
        }
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");

    // thread t finishes the output channels t, t + 128, ... < D
    constexpr int DPT = (D + 127) / 128;
    const int d0 = threadIdx.x;
    const size_t ws_base = ((size_t)bh * p.nsplit + split) * rows;
    // output row of MMA row r: token r / group, head kvh * group + r % group of out [B][q_len][Hq * D]
    auto out_row = [&](int r) {
        return reinterpret_cast<T*>(p.out) + (((size_t)b * p.q_len + r / p.group) * p.Hq + (size_t)kvh * p.group + r % p.group) * D;
    };
    // per-CTA partial of the cluster merge: [16 rows][D] fp32 + (m, l) per row, behind the per-warp staging area
    float* part_o = l_s + kAttnConsumerWarps * 16;
    float* part_ml = part_o + 16 * D;
    for (int r = 0; r < rows; ++r) {
        float m = -INFINITY;
#pragma unroll
        for (int w = 0; w < kAttnConsumerWarps; ++w) m = fmaxf(m, m_s[w * 16 + r]);
        float f[kAttnConsumerWarps], l = 0.f;
#pragma unroll
        for (int w = 0; w < kAttnConsumerWarps; ++w) {
            f[w] = (m == -INFINITY) ? 0.f : fast_exp2((m_s[w * 16 + r] - m) * sl2);  // -inf -> 0
            l = fmaf(f[w], l_s[w * 16 + r], l);
        }
        if (d0 == 0) {
            if (p.cluster_merge) {
                part_ml[r * 2 + 0] = m;
                part_ml[r * 2 + 1] = l;
            } else if (nact > 1) {
                p.ws_ml[(ws_base + r) * 2 + 0] = m;
                p.ws_ml[(ws_base + r) * 2 + 1] = l;
            }
        }
#pragma unroll
        for (int j = 0; j < DPT; ++j) {
            const int d = d0 + 128 * j;
            if (d >= D) break;
            float acc = 0.f;
#pragma unroll
            for (int w = 0; w < kAttnConsumerWarps; ++w) acc = fmaf(f[w], o_s[((size_t)w * 16 + r) * OROW + d], acc);
            if (p.cluster_merge) part_o[r * D + d] = acc;
            else if (nact == 1) out_row(r)[d] = from_f32<T>(acc / l);
            else p.ws_o[(ws_base + r) * D + d] = acc;
        }
    }
    if (p.cluster_merge) {
        // ---------------------------------------------------------------------- cross-CTA merge through DSMEM
        // CTA `rank` of the cluster (= split) finishes the GQA rows r = rank, rank + nsplit, ...: every row is reduced by
        // exactly one CTA in fixed split order 0..nsplit-1 (deterministic); empty splits carry m = -inf, l = 0.
        cluster_sync_all();
        const int S = p.nsplit;
        const uint32_t o_addr = smem_u32(part_o), ml_addr = smem_u32(part_ml);
        for (int r = (int)cluster_ctarank(); r < rows; r += S) {
            float mm[8], ll[8];
            float m = -INFINITY;
#pragma unroll
            for (int sp = 0; sp < 8; ++sp) {
                if (sp < S) {
                    mm[sp] = dsmem_ld_f32(dsmem_addr(ml_addr + (r * 2 + 0) * 4, sp));
                    ll[sp] = dsmem_ld_f32(dsmem_addr(ml_addr + (r * 2 + 1) * 4, sp));
                    m = fmaxf(m, mm[sp]);
                }
            }
            float l = 0.f;
#pragma unroll
            for (int sp = 0; sp < 8; ++sp) {
                if (sp < S) {
                    mm[sp] = fast_exp2((mm[sp] - m) * sl2);      // -inf -> 0 (m is finite: the sequence has >= 1 token)
                    l = fmaf(mm[sp], ll[sp], l);
                }
            }
#pragma unroll
            for (int j = 0; j < DPT; ++j) {
                const int d = d0 + 128 * j;
                if (d >= D) break;
                float oo[8];
#pragma unroll
                for (int sp = 0; sp < 8; ++sp)
                    if (sp < S) oo[sp] = dsmem_ld_f32(dsmem_addr(o_addr + (r * D + d) * 4, sp));
                float acc = 0.f;
#pragma unroll
                for (int sp = 0; sp < 8; ++sp)
                    if (sp < S) acc = fmaf(mm[sp], oo[sp], acc);
                out_row(r)[d] = from_f32<T>(acc / l);
            }
        }
        cluster_sync_all();   // peers may still be reading this CTA's partial
        return;
    }
    if (nact == 1) return;

    // ---------------------------------------------------------------------- cross-CTA merge: last arriver reduces
    __threadfence();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (threadIdx.x == 0) {
        const int prev = atomicAdd(&p.sem[bh], 1);
        *s_flag = (prev == nact - 1);
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (!*s_flag) return;
    __threadfence();
    for (int r = 0; r < rows; ++r) {
        float m = -INFINITY;
        for (int sp = 0; sp < nact; ++sp)
            m = fmaxf(m, __ldcg(&p.ws_ml[(((size_t)bh * p.nsplit + sp) * rows + r) * 2 + 0]));
#pragma unroll
        for (int j = 0; j < DPT; ++j) {
            const int d = d0 + 128 * j;
            if (d >= D) break;
            float acc = 0.f, l = 0.f;
            for (int sp = 0; sp < nact; ++sp) {
                const size_t row = ((size_t)bh * p.nsplit + sp) * rows + r;
                const float fsp = fast_exp2((__ldcg(&p.ws_ml[row * 2 + 0]) - m) * sl2);
                acc = fmaf(fsp, __ldcg(&p.ws_o[row * D + d]), acc);
                l = fmaf(fsp, __ldcg(&p.ws_ml[row * 2 + 1]), l);
            }
            out_row(r)[d] = from_f32<T>(acc / l);
        }
    }
    if (threadIdx.x == 0) p.sem[bh] = 0;  // self-reset for the next launch
}

}  // namespace b200
