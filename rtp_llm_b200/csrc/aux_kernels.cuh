// Small HBM-bound kernels either side of the two hot kernels: page-table indexing, the glue ops of a decode layer
//  and greedy sampling.
#pragma once
#include "ptx.cuh"

namespace b200 {

// ---------------------------------------------------------------------------------------------- indexing
// block table [B][M] -> page list [B][2][M] (K = 2*id, V = 2*id+1); bit-exact restatement of the behaviour of
// /root/reference/rtp_llm/models_py/bindings/common/kernels/kv_cache_kernels.cu:49-63.
static __global__ void convert_block_table_kernel(int32_t* __restrict__ page_list, const int32_t* __restrict__ block_ids,
                                           int batch, int max_blocks) {
    const int total = batch * max_blocks;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int b = idx / max_blocks, j = idx - b * max_blocks;
        const int32_t id = block_ids[idx];
        page_list[(size_t)(2 * b) * max_blocks + j] = id * 2;
        page_list[(size_t)(2 * b + 1) * max_blocks + j] = id * 2 + 1;
    }
}

// decode-mode flashinfer plan metadata (mha_paged_attn_plan.cu:28-97, decode branch + prefill branch).
// One CTA; a warp-shuffle scan replaces the reference's thread-0 serial scan.
static __global__ void paged_attn_plan_kernel(const int32_t* __restrict__ input_lengths,
                                       const int32_t* __restrict__ sequence_lengths,
                                       const int32_t* __restrict__ prefix_lengths,
                                       const int32_t* __restrict__ block_ids, int batch, int max_blocks,
                                       int tokens_per_block, int32_t* __restrict__ last_page_len,
                                       int32_t* __restrict__ page_indptr, int32_t* __restrict__ page_indice,
                                       int32_t* __restrict__ batch_indice, int32_t* __restrict__ positions) {
    __shared__ int32_t s_wp[32], s_wt[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int input_len = 0, prefix_len = 0, seq_len = 0, pages = 0;
    if (tid < batch) {
        if (prefix_lengths) {
            input_len = input_lengths[tid];
            prefix_len = prefix_lengths[tid];
            seq_len = input_len + prefix_len;
        } else {
            input_len = 1;
            seq_len = sequence_lengths[tid] + 1;
        }
        pages = (seq_len + tokens_per_block - 1) / tokens_per_block;
        last_page_len[tid] = (seq_len - 1) % tokens_per_block + 1;
    }
    // inclusive warp scans, then scan of warp totals
    int ip = pages, itk = input_len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int vp = __shfl_up_sync(0xffffffffu, ip, o), vt = __shfl_up_sync(0xffffffffu, itk, o);
        if (lane >= o) {
            ip += vp;
            itk += vt;
        }
    }
    if (lane == 31) {
        s_wp[wid] = ip;
        s_wt[wid] = itk;
    }
    __syncthreads();
    if (wid == 0) {
        int wp = s_wp[lane], wt = s_wt[lane];
        const int nw = (blockDim.x + 31) >> 5;
        if (lane >= nw) wp = wt = 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int vp = __shfl_up_sync(0xffffffffu, wp, o), vt = __shfl_up_sync(0xffffffffu, wt, o);
            if (lane >= o) {
                wp += vp;
                wt += vt;
            }
        }
        s_wp[lane] = wp;
        s_wt[lane] = wt;
    }
    __syncthreads();
    const int pbase = (wid ? s_wp[wid - 1] : 0), tbase = (wid ? s_wt[wid - 1] : 0);
    const int p_end = pbase + ip, t_end = tbase + itk;
    if (tid == 0) page_indptr[0] = 0;
    if (tid < batch) {
        page_indptr[tid + 1] = p_end;
        const int p_start = p_end - pages, t_start = t_end - input_len;
        if (prefix_lengths) {
            for (int j = 0; j < input_len; ++j) {
                batch_indice[t_start + j] = tid;
                positions[t_start + j] = j + prefix_len;
            }
        } else {
            batch_indice[t_start] = tid;
            positions[t_start] = sequence_lengths[tid];
        }
        if (block_ids)
            for (int j = 0; j < pages; ++j) page_indice[p_start + j] = block_ids[(size_t)tid * max_blocks + j];
    }
}

// ---------------------------------------------------------------------------------------------- glue ops
template <typename T>
__device__ __forceinline__ float2 unpack2(uint32_t v);
template <>
__device__ __forceinline__ float2 unpack2<__half>(uint32_t v) { return __half22float2(*reinterpret_cast<__half2*>(&v)); }
template <>
__device__ __forceinline__ float2 unpack2<__nv_bfloat16>(uint32_t v) {
    return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&v));
}

// y = rmsnorm(x (+ residual)) * gamma; residual (if given) is updated in place with x + residual (rounded to T), while the
// UNROUNDED fp32 sum feeds both the variance and the output -- the numerics of the op the reference binds
// (RegisterBaseBindings.hpp:45-60 -> flashinfer FusedAddRMSNormKernel). One CTA per row, 16-byte vector loads, hidden % 8 == 0.
template <typename T>
__global__ void add_rmsnorm_kernel(const T* __restrict__ x, T* __restrict__ residual, const T* __restrict__ gamma,
                                   T* __restrict__ y, int hidden, float eps) {
    pdl_launch_dependents();  // the next kernel (a GEMM) may start streaming its weights now
    pdl_wait();               // our inputs come from the previous kernel
    extern __shared__ float s_row[];  // hidden floats
    __shared__ float s_red[32];
    const int row = blockIdx.x;
    const uint4* xv = reinterpret_cast<const uint4*>(x + (size_t)row * hidden);
    uint4* rv = residual ? reinterpret_cast<uint4*>(residual + (size_t)row * hidden) : nullptr;
    float ss = 0.f;
    for (int i = threadIdx.x; i < hidden / 8; i += blockDim.x) {
        uint4 a = xv[i];
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
        uint32_t rw[4] = {0u, 0u, 0u, 0u}, ow[4];
        if (rv) {
            uint4 r = rv[i];
            rw[0] = r.x; rw[1] = r.y; rw[2] = r.z; rw[3] = r.w;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 fa = unpack2<T>(aw[j]), fr = unpack2<T>(rw[j]);
            const float x0 = fa.x + fr.x, x1 = fa.y + fr.y;
            s_row[i * 8 + j * 2] = x0;
            s_row[i * 8 + j * 2 + 1] = x1;
            ss = fmaf(x0, x0, fmaf(x1, x1, ss));
            ow[j] = pack2<T>(x0, x1);
        }
        if (rv) rv[i] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? s_red[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) s_red[0] = rsqrtf(v / hidden + eps);
    }
    __syncthreads();
    const float inv = s_red[0];
    const uint4* gv = reinterpret_cast<const uint4*>(gamma);
    uint4* yv = reinterpret_cast<uint4*>(y + (size_t)row * hidden);
    for (int i = threadIdx.x; i < hidden / 8; i += blockDim.x) {
        uint4 g = gv[i];
        uint32_t gw[4] = {g.x, g.y, g.z, g.w}, ow[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float2 fg = unpack2<T>(gw[j]);
            ow[j] = pack2<T>(s_row[i * 8 + j * 2] * inv * fg.x, s_row[i * 8 + j * 2 + 1] * inv * fg.y);
        }
        yv[i] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
}

// Per-head RMSNorm of the q and k heads of a qkv row, in place (Qwen3-style QK-norm before RoPE): the op
// rtp_llm/models_py/bindings/cuda/kernels/fused_qk_rmsnorm.cu:24-78 implements (call site model_desc/qwen3.py:57-79).
// One warp per (row, head); fp32 sum of squares; val * rsqrt(mean + eps) * gamma (+ bias), rounded to T; v heads untouched.
template <typename T>
__global__ void qk_rmsnorm_kernel(T* __restrict__ qkv, const T* __restrict__ q_gamma, const T* __restrict__ k_gamma,
                                  const T* __restrict__ q_bias, const T* __restrict__ k_bias, int rows, int head_num,
                                  int kv_head_num, int head_dim, float eps) {
    pdl_launch_dependents();
    pdl_wait();
    const int warps_per_block = blockDim.x >> 5, lane = threadIdx.x & 31;
    const int unit = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const int per_row = head_num + kv_head_num;
    if (unit >= rows * per_row) return;
    const int r = unit / per_row, h = unit % per_row;
    T* x = qkv + ((size_t)r * (head_num + 2 * kv_head_num) + h) * head_dim;
    const T* gamma = h < head_num ? q_gamma : k_gamma;
    const T* bias = h < head_num ? q_bias : k_bias;
    float ss = 0.f;
    for (int c = lane * 2; c < head_dim; c += 64) {
        const float2 v = unpack2<T>(*reinterpret_cast<const uint32_t*>(x + c));
        ss = fmaf(v.x, v.x, fmaf(v.y, v.y, ss));
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float scale = rsqrtf(ss / (float)head_dim + eps);
    for (int c = lane * 2; c < head_dim; c += 64) {
        const float2 v = unpack2<T>(*reinterpret_cast<const uint32_t*>(x + c));
        const float2 g = unpack2<T>(*reinterpret_cast<const uint32_t*>(gamma + c));
        float y0 = v.x * scale * g.x, y1 = v.y * scale * g.y;
        if (bias) {
            const float2 bb = unpack2<T>(*reinterpret_cast<const uint32_t*>(bias + c));
            y0 += bb.x;
            y1 += bb.y;
        }
        *reinterpret_cast<uint32_t*>(x + c) = pack2<T>(y0, y1);
    }
}

// y[r][c] = silu(gate_up[r][c]) * gate_up[r][inter + c]   (activation_kernels.cu silu_and_mul semantics)
template <typename T>
__global__ void silu_and_mul_kernel(const T* __restrict__ gate_up, T* __restrict__ y, int rows, int inter) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t total = (size_t)rows * inter / 2;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t r = idx / (inter / 2), c2 = idx % (inter / 2);
        const uint32_t g = reinterpret_cast<const uint32_t*>(gate_up + r * 2 * inter)[c2];
        const uint32_t u = reinterpret_cast<const uint32_t*>(gate_up + r * 2 * inter + inter)[c2];
        float2 fg = unpack2<T>(g), fu = unpack2<T>(u);
        float s0 = fg.x / (1.f + __expf(-fg.x)), s1 = fg.y / (1.f + __expf(-fg.y));
        reinterpret_cast<uint32_t*>(y + r * inter)[c2] = pack2<T>(s0 * fu.x, s1 * fu.y);
    }
}

// Decode RoPE (NeoX pairing, RopeStyle::Base, angle = pos * base^(-2i/dim)) + K,V append into the paged cache.
// Contract: SURVEY.md appendix C; call site /root/reference/rtp_llm/ops/fused_rope_kvcache_op.py:202-246.
// grid (B, Hq + 2*Hkv), D/2 threads.
template <typename T>
__global__ void rope_append_kernel(const T* __restrict__ qkv, T* __restrict__ q_out, T* __restrict__ kv_pool,
                                   const int32_t* __restrict__ page_list, const int32_t* __restrict__ seq_lens,
                                   int head_num, int kv_head_num, int head_dim, int max_blocks, int tokens_per_block,
                                   float log2_base) {
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.x, h = blockIdx.y, i = threadIdx.x, half = head_dim / 2;
    const int pos = seq_lens[b];
    const T* src = qkv + ((size_t)b * (head_num + 2 * kv_head_num) + h) * head_dim;
    const size_t page_elems = (size_t)kv_head_num * tokens_per_block * head_dim;
    T* dst;
    bool rotate = true;
    if (h < head_num) {
        dst = q_out + ((size_t)b * head_num + h) * head_dim;
    } else {
        const bool is_v = h >= head_num + kv_head_num;
        const int kvh = is_v ? h - head_num - kv_head_num : h - head_num;
        const int32_t page = page_list[((size_t)b * 2 + (is_v ? 1 : 0)) * max_blocks + pos / tokens_per_block];
        dst = kv_pool + (size_t)page * page_elems + ((size_t)kvh * tokens_per_block + pos % tokens_per_block) * head_dim;
        rotate = !is_v;
    }
    const float x0 = to_f32<T>(src[i]), x1 = to_f32<T>(src[i + half]);
    if (rotate) {
        const float inv_freq = exp2f(-2.0f * (float)i / (float)head_dim * log2_base);
        float sn, cs;
        sincosf((float)pos * inv_freq, &sn, &cs);
        float r0, r1;
        rope_rotate(x0, x1, cs, sn, r0, r1);
        dst[i] = from_f32<T>(r0);
        dst[i + half] = from_f32<T>(r1);
    } else {
        dst[i] = src[i];
        dst[i + half] = src[i + half];
    }
}


// ---- the full decode RoPE contract of FusedRopeKVCacheDecodeOp (rtp_llm/ops/fused_rope_kvcache_op.py:202-246 ->
// decode_fused_rope_kvcache; device code rtp_llm/models_py/bindings/common/kernels/rotary_position_embedding.h):
//   position      sequence_lengths[b], overridden by position_ids[b] when that is > 0              (:1040-1042)
//   coefficient   cos_sin_cache[pos * dim/2 + i] when a cache is given (Base / Yarn: cpp/model_utils/RopeCache.cc:16-85), else
//                 angle = pos / powf(base, 2i/dim) transformed per style (:322-440), (cos, sin) scaled by mscale (Yarn)
//   styles        Base / Mrope (linear scale), DynamicNTK, QwenDynamicNTK (base rescaled when pos+1 > max_pos, :889-903),
//                 Yarn (:365-419), Llama3 (:421-442)
//   pairing       NeoX halves (i, i + dim/2) inside the first `dim` channels of a head; channels >= dim pass through
//   bias          optional qkv bias added (in the tensor type) before the rotation
//   logn          q *= log(pos + 1) / log(max_pos) when pos > max_pos (decoder_masked_multihead_attention_utils.h:2097-2102)
struct RopeCfg {
    int style;      // RopeStyle (cpp/model_utils/RopeConfig.h:7-16): 0 No, 1 Base, 3 DynamicNTK, 4 QwenDynamicNTK, 5 Yarn, 6 Llama3, 7 Mrope
    int dim;        // rotary dim (<= head_dim)
    float base, scale, factor1, factor2;
    int max_pos;
    float extrapolation_factor, mscale;
};

__device__ __forceinline__ float2 rope_coef(const RopeCfg& c, int zid, int pos) {
    float base = c.base;
    const int seq_len = pos + 1;
    if (c.style == 3 && seq_len > c.max_pos)
        base = c.base * powf((c.scale * seq_len / c.max_pos) - (c.scale - 1.f), c.dim / (c.dim - 2.0f));
    if (c.style == 4 && seq_len > c.max_pos) {
        const float ctx = logf((float)seq_len / c.max_pos) / logf(2.0f) + 1.0f;
        const float ntk = fmaxf(powf(2.0f, ceilf(ctx)) - 1.f, 1.0f);
        base = c.base * powf(ntk, (float)c.dim / (c.dim - 2));
    }
    float a = (float)pos / powf(base, zid / (float)c.dim);
    float sc = 1.f;
    if (c.style == 1 || c.style == 7) {
        a = a / c.scale;
    } else if (c.style == 5) {
        const float pi = 3.141592654f;
        const int ibase = (int)c.base;
        const float t1 = 2.f * logf((float)ibase);
        int low = (int)floorf(c.dim * logf((float)c.max_pos / (c.factor2 * 2 * pi)) / t1);
        int high = (int)ceilf(c.dim * logf((float)c.max_pos / (c.factor1 * 2 * pi)) / t1);
        float lo = (float)max(low, 0), hi = (float)min(high, c.dim - 1);
        if (lo == hi) hi += 0.001f;
        const float ramp = fminf(1.f, fmaxf(0.f, (zid / 2 - lo) / (hi - lo)));
        const float mask = (1.f - ramp) * c.extrapolation_factor;
        a = (a / c.scale) * (1.f - mask) + a * mask;
        sc = c.mscale;
    } else if (c.style == 6) {
        const float pi = 3.141592654f;
        const float wavelen = 2 * pi / a;
        const float low_w = c.max_pos / c.factor1, high_w = c.max_pos / c.factor2;
        if (wavelen < high_w) {
        } else if (wavelen > low_w) {
            a = a / c.scale;
        } else {
            const float smooth = (c.max_pos / wavelen - c.factor1) / (c.factor2 - c.factor1);
            a = (1 - smooth) * a / c.scale + smooth * a;
        }
    }
    float sn, cs;
    sincosf(a, &sn, &cs);
    return make_float2(sc * cs, sc * sn);
}

template <typename T>
__global__ void rope_append_ex_kernel(const T* __restrict__ qkv, const T* __restrict__ bias, T* __restrict__ q_out,
                                      T* __restrict__ kv_pool, const int32_t* __restrict__ page_list,
                                      const int32_t* __restrict__ seq_lens, const int32_t* __restrict__ position_ids,
                                      const float2* __restrict__ cos_sin_cache, int cache_positions, RopeCfg cfg, int use_logn,
                                      int head_num, int kv_head_num, int head_dim, int max_blocks, int tokens_per_block) {
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.x, h = blockIdx.y, i = threadIdx.x, half = head_dim / 2, rhalf = cfg.dim / 2;
    const int slot = seq_lens[b];                                  // where K/V of the new token go (tokens already cached)
    const int pos = (position_ids && position_ids[b] > 0) ? position_ids[b] : slot;
    const T* src = qkv + ((size_t)b * (head_num + 2 * kv_head_num) + h) * head_dim;
    const T* bsrc = bias ? bias + (size_t)h * head_dim : nullptr;
    const size_t page_elems = (size_t)kv_head_num * tokens_per_block * head_dim;
    T* dst;
    bool rotate = cfg.style != 0, is_q = h < head_num;
    if (is_q) {
        dst = q_out + ((size_t)b * head_num + h) * head_dim;
    } else {
        const bool is_v = h >= head_num + kv_head_num;
        const int kvh = is_v ? h - head_num - kv_head_num : h - head_num;
        const int32_t page = page_list[((size_t)b * 2 + (is_v ? 1 : 0)) * max_blocks + slot / tokens_per_block];
        dst = kv_pool + (size_t)page * page_elems + ((size_t)kvh * tokens_per_block + slot % tokens_per_block) * head_dim;
        rotate = rotate && !is_v;
    }
    // thread i owns the channel pair (i, i + dim/2) for i < dim/2, and the pass-through channels dim + i, dim + i + (head_dim-dim)/2 ...
    auto load = [&](int c) {
        float v = to_f32<T>(src[c]);
        if (bsrc) v = to_f32<T>(from_f32<T>(v + to_f32<T>(bsrc[c])));
        return v;
    };
    if (i < rhalf) {
        float x0 = load(i), x1 = load(i + rhalf);
        if (rotate) {
            const float2 cf = (cos_sin_cache && pos < cache_positions) ? cos_sin_cache[(size_t)pos * rhalf + i] : rope_coef(cfg, 2 * i, pos);
            float r0, r1;
            rope_rotate(x0, x1, cf.x, cf.y, r0, r1);
            x0 = r0;
            x1 = r1;
        }
        if (is_q && use_logn && pos > cfg.max_pos) {
            const float logn = logf((float)(pos + 1)) / logf((float)cfg.max_pos);
            x0 *= logn;
            x1 *= logn;
        }
        dst[i] = from_f32<T>(x0);
        dst[i + rhalf] = from_f32<T>(x1);
    }
    for (int c = cfg.dim + i; c < head_dim; c += half) {            // channels beyond the rotary dim
        float v = load(c);
        if (is_q && use_logn && pos > cfg.max_pos) v *= logf((float)(pos + 1)) / logf((float)cfg.max_pos);
        dst[c] = from_f32<T>(v);
    }
}

// token embedding gather: out[b][:] = table[ids[b]][:]
template <typename T>
__global__ void embedding_kernel(const int32_t* __restrict__ ids, const T* __restrict__ table, T* __restrict__ out,
                                 int hidden) {
    const uint4* src = reinterpret_cast<const uint4*>(table + (size_t)ids[blockIdx.x] * hidden);
    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)blockIdx.x * hidden);
    for (int i = threadIdx.x; i < hidden / 8; i += blockDim.x) dst[i] = src[i];
}

// greedy sampling: argmax over the vocabulary, lowest index wins ties (torch.argmax semantics used by
// /root/reference/rtp_llm/models_py/bindings/core/CudaSampleOp.cc:330,453). One CTA per row.
template <typename T>
__global__ void argmax_kernel(const T* __restrict__ logits, int vocab, int32_t* __restrict__ out) {
    __shared__ float s_v[32];
    __shared__ int s_i[32];
    const T* row = logits + (size_t)blockIdx.x * vocab;
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    auto consider = [&](float v, int i) {
        if (v > bv || (v == bv && i < bi)) {
            bv = v;
            bi = i;
        }
    };
    if (sizeof(T) == 2 && (vocab % 8) == 0 && ((reinterpret_cast<uintptr_t>(row) & 15) == 0)) {
        // 16-byte loads: 8 logits per thread and iteration (the scalar loop moved 64 B per warp instruction)
        const uint4* rv = reinterpret_cast<const uint4*>(row);
        for (int c = threadIdx.x; c < vocab / 8; c += blockDim.x) {
            const uint4 q = rv[c];
            const T* e = reinterpret_cast<const T*>(&q);
#pragma unroll
            for (int j = 0; j < 8; ++j) consider(to_f32<T>(e[j]), c * 8 + j);
        }
    } else {
        for (int i = threadIdx.x; i < vocab; i += blockDim.x) consider(to_f32<T>(row[i]), i);
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
        bv = threadIdx.x < (blockDim.x >> 5) ? s_v[threadIdx.x] : -INFINITY;
        bi = threadIdx.x < (blockDim.x >> 5) ? s_i[threadIdx.x] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) {
                bv = ov;
                bi = oi;
            }
        }
        if (threadIdx.x == 0) out[blockIdx.x] = bi;
    }
}
}  // namespace b200
