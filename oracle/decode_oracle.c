/*
 * decode_oracle.c -- CPU restatement of rtp-llm's decode hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and there only as the checker / the timed CPU arm.
 * The product path (rtp_llm_b200/) never falls back to it.
 *
 * Parity status
 *   - indexing (block table -> page list, flashinfer plan): pinned against the
 *     reference's own expected-value builders (tests/golden/indexing_*.npz).
 *   - paged decode attention: pinned against the reference's torch oracle
 *     attention_prefill_ref (atten_test_util.py:55-116) on the reference's test
 *     shapes/seeds (tests/golden/attn_*.npz, made by oracle/make_golden.py).
 *   - GPTQ/AWQ/INT8 unpack + dequant formula: pinned against the reference's
 *     loader code device_impl.py:148-300 run in this container (golden fixtures).
 *   - weight-only GEMM *kernel numerics*: "parity unpinned" -- the reference
 *     snapshot holds no weight-only GEMM kernel, test or stored vector
 *     (SURVEY.md section 0 / 8c); the GEMM is defined here as
 *     Y = X . W' with W' from the loader's dequant formula, fp32+ accumulation.
 *
 * Every function cites the reference file:line (paths relative to /root/reference)
 * whose behaviour it restates.  No reference source is copied.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ */
/* 16-bit float helpers                                                */
/* ------------------------------------------------------------------ */
typedef uint16_t h16; /* raw bits of fp16 or bf16 */

static inline float half_to_float(h16 h) {
    _Float16 f;
    memcpy(&f, &h, 2);
    return (float)f;
}
static inline h16 float_to_half(float x) {
    _Float16 f = (_Float16)x; /* round-to-nearest-even */
    h16 h;
    memcpy(&h, &f, 2);
    return h;
}
static inline float bf16_to_float(h16 h) {
    uint32_t u = ((uint32_t)h) << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}
static inline h16 float_to_bf16(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (h16)((u >> 16) | 0x40); /* nan */
    uint32_t lsb = (u >> 16) & 1u;
    u += 0x7fffu + lsb;
    return (h16)(u >> 16);
}
static inline float elem_to_float(h16 h, int is_bf16) { return is_bf16 ? bf16_to_float(h) : half_to_float(h); }
static inline h16 float_to_elem(float x, int is_bf16) { return is_bf16 ? float_to_bf16(x) : float_to_half(x); }
static inline float round_elem(float x, int is_bf16) { return elem_to_float(float_to_elem(x, is_bf16), is_bf16); }

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------ */
/* Indexing                                                            */
/* ------------------------------------------------------------------ */

/* rtp_llm/models_py/bindings/common/kernels/kv_cache_kernels.cu:49-63
 * block table [B,M] -> page list [B,2,M]: K page = 2*id, V page = 2*id+1. */
void oracle_convert_block_table(int32_t* page_list, const int32_t* block_ids, int batch, int max_blocks) {
    for (int b = 0; b < batch; ++b)
        for (int j = 0; j < max_blocks; ++j) {
            int32_t id = block_ids[(size_t)b * max_blocks + j];
            page_list[((size_t)b * 2 + 0) * max_blocks + j] = id * 2;
            page_list[((size_t)b * 2 + 1) * max_blocks + j] = id * 2 + 1;
        }
}

/* rtp_llm/models_py/bindings/cuda/kernels/mha_paged_attn_plan.cu:28-97
 * flashinfer plan metadata.  prefix_lengths == NULL selects decode mode
 * (one new token per sequence, seq_len = sequence_lengths[b] + 1).
 * kv_cache_block_id may be NULL (page_indice untouched). Returns total pages. */
int oracle_paged_attn_plan(const int32_t* input_lengths, const int32_t* sequence_lengths,
                           const int32_t* prefix_lengths, const int32_t* kv_cache_block_id, int batch,
                           int max_blocks, int tokens_per_block, int32_t* last_page_len, int32_t* page_indptr,
                           int32_t* page_indice, int32_t* batch_indice, int32_t* positions) {
    int t_off = 0, p_off = 0;
    page_indptr[0] = 0;
    for (int b = 0; b < batch; ++b) {
        int input_len, seq_len, prefix_len = 0;
        if (prefix_lengths) {
            input_len = input_lengths[b];
            prefix_len = prefix_lengths[b];
            seq_len = input_len + prefix_len;
        } else {
            input_len = 1;
            seq_len = sequence_lengths[b] + 1;
        }
        int pages = (seq_len + tokens_per_block - 1) / tokens_per_block;
        last_page_len[b] = (seq_len - 1) % tokens_per_block + 1;
        if (prefix_lengths) {
            for (int j = 0; j < input_len; ++j) {
                batch_indice[t_off + j] = b;
                positions[t_off + j] = j + prefix_len;
            }
        } else {
            batch_indice[t_off] = b;
            positions[t_off] = sequence_lengths[b];
        }
        if (kv_cache_block_id)
            for (int j = 0; j < pages; ++j) page_indice[p_off + j] = kv_cache_block_id[(size_t)b * max_blocks + j];
        t_off += input_len;
        p_off += pages;
        page_indptr[b + 1] = p_off;
    }
    return p_off;
}

/* ------------------------------------------------------------------ */
/* Paged decode attention                                              */
/* ------------------------------------------------------------------ */

/* Addressing: rtp_llm/models_py/bindings/common/kernels/kv_cache/kv_cache_utils.h:171-206
 *   pool viewed as pages of [Hkv][T][D]; page pointer = pool + page_idx * Hkv*T*D;
 *   in-page offset = head*T*D + (tok & (T-1))*D + c.
 * Page list: [B][1][2][M] with K/V page ids (kv_cache_kernels.cu:49-63, CudaXqa.h:53-54).
 * Lengths: sequence_lengths[b] = tokens already cached; attention covers
 *   0..sequence_lengths[b] inclusive (3rdparty/xqa/mha_sm90.cu:653, trtllm_gen.py:472-486).
 * Math: 3rdparty/xqa/ref.py:75-84,151-158 -- fp32 scores * D^-1/2 * q_scale, max-subtract,
 *   P rounded to the cache element type before the PV product, fp32 accumulation, / rowsum.
 * GQA: q head h -> kv head h / (Hq/Hkv) (atten_test_util.py:77-80).
 * q [B][Hq][D], out [B][Hq*D]  (XQAAttnOp.cc:128). */
void oracle_paged_decode_attn(const h16* q, int is_bf16, h16* out, int head_num, int kv_head_num, int head_dim,
                              int batch, int max_blocks, int tokens_per_block, const h16* kv_pool,
                              const int32_t* page_list, const int32_t* sequence_lengths, float q_scale) {
    const int group = head_num / kv_head_num;
    const float scale = q_scale / sqrtf((float)head_dim);
    const size_t page_elems = (size_t)kv_head_num * tokens_per_block * head_dim;
#pragma omp parallel for collapse(2) schedule(dynamic, 1)
    for (int b = 0; b < batch; ++b) {
        for (int h = 0; h < head_num; ++h) {
            const int kvh = h / group;
            const int len = sequence_lengths[b] + 1;
            float* s = (float*)malloc(sizeof(float) * (size_t)len);
            float* acc = (float*)calloc((size_t)head_dim, sizeof(float));
            float* qf = (float*)malloc(sizeof(float) * (size_t)head_dim);
            const h16* qrow = q + ((size_t)b * head_num + h) * head_dim;
            for (int c = 0; c < head_dim; ++c) qf[c] = elem_to_float(qrow[c], is_bf16);
            const int32_t* kpages = page_list + ((size_t)b * 2 + 0) * max_blocks;
            const int32_t* vpages = page_list + ((size_t)b * 2 + 1) * max_blocks;
            float m = -INFINITY;
            for (int t = 0; t < len; ++t) {
                const h16* krow = kv_pool + (size_t)kpages[t / tokens_per_block] * page_elems +
                                  ((size_t)kvh * tokens_per_block + (t % tokens_per_block)) * head_dim;
                float d = 0.f;
                for (int c = 0; c < head_dim; ++c) d += qf[c] * elem_to_float(krow[c], is_bf16);
                s[t] = d * scale;
                if (s[t] > m) m = s[t];
            }
            float rowsum = 0.f;
            for (int t = 0; t < len; ++t) {
                float p = round_elem(expf(s[t] - m), is_bf16);
                rowsum += p;
                const h16* vrow = kv_pool + (size_t)vpages[t / tokens_per_block] * page_elems +
                                  ((size_t)kvh * tokens_per_block + (t % tokens_per_block)) * head_dim;
                for (int c = 0; c < head_dim; ++c) acc[c] += p * elem_to_float(vrow[c], is_bf16);
            }
            h16* orow = out + ((size_t)b * head_num + h) * head_dim;
            for (int c = 0; c < head_dim; ++c) orow[c] = float_to_elem(acc[c] / rowsum, is_bf16);
            free(s);
            free(acc);
            free(qf);
        }
    }
}

/* ------------------------------------------------------------------ */
/* Weight-only quantisation: unpack + dequant formula                  */
/* ------------------------------------------------------------------ */

/* rtp_llm/device/device_impl.py:148-161 (unpack_int32_into_int16, low nibble first),
 * :163-171 (reverse_awq_order), :204-209 (pack_int8_tensor_to_packed_int4),
 * :242-300 (preprocess_groupwise_weight_params), 4-bit only.
 *
 * Inputs
 *   gptq: qweight int32 [K/8][N]  (nibble i of a word = row 8r+i)
 *   awq : qweight int32 [K][N/8]  (nibble order within a word [0,2,4,6,1,3,5,7])
 *   qzeros int32 [K/g][N/8] (same per-word order as the format), scales fp16 [K/g][N]
 * Outputs (the *un-permuted* tensors the loader hands to preprocess_weights_for_mixed_gemm)
 *   q_packed   uint8 [K][N/2]: byte = (q_s[k][2j+1] & 0xF) << 4 | (q_s[k][2j] & 0xF), q_s = q_u - 8
 *   zeros_x_scales fp16 [K/g][N] = fp16( (8 - z_u - [gptq]) * s )      (device_impl.py:286-292)
 */
static inline int awq_logical_col(int pos_in_word) {
    /* reverse_awq_order: reshape(-1,2,4).transpose(2,1): unpacked position p=(a*4+b) -> logical b*2+a */
    return (pos_in_word % 4) * 2 + (pos_in_word / 4);
}

void oracle_unpack_groupwise_int4(const int32_t* qweight, const int32_t* qzeros, const h16* scales, int K, int N,
                                  int group, int is_gptq, uint8_t* q_packed, h16* zeros_x_scales) {
    const int G = K / group;
    /* weights */
    for (int k = 0; k < K; ++k) {
        for (int n = 0; n < N; ++n) {
            uint32_t nib;
            if (is_gptq) {
                uint32_t w = (uint32_t)qweight[(size_t)(k / 8) * N + n];
                nib = (w >> (4 * (k % 8))) & 0xF;
            } else {
                /* find the unpacked position whose logical column is n%8 */
                int word = n / 8, lc = n % 8, pos = (lc % 2) * 4 + lc / 2;
                uint32_t w = (uint32_t)qweight[(size_t)k * (N / 8) + word];
                nib = (w >> (4 * pos)) & 0xF;
            }
            int qs = (int)nib - 8;
            uint8_t* byte = &q_packed[(size_t)k * (N / 2) + n / 2];
            if (n % 2 == 0)
                *byte = (uint8_t)((*byte & 0xF0) | (qs & 0xF));
            else
                *byte = (uint8_t)((*byte & 0x0F) | ((qs & 0xF) << 4));
        }
    }
    /* zeros * scales; torch computes (int16 - ...) * fp16 -> fp16 product, one rounding */
    for (int g = 0; g < G; ++g) {
        for (int n = 0; n < N; ++n) {
            int word = n / 8, lc = n % 8;
            int pos = is_gptq ? lc : ((lc % 2) * 4 + lc / 2);
            uint32_t w = (uint32_t)qzeros[(size_t)g * (N / 8) + word];
            int z = (int)((w >> (4 * pos)) & 0xF);
            float s = half_to_float(scales[(size_t)g * N + n]);
            float zf = (float)(-z + 8 - (is_gptq ? 1 : 0));
            zeros_x_scales[(size_t)g * N + n] = float_to_half(zf * s);
        }
    }
    (void)awq_logical_col;
}

/* rtp_llm/device/device_impl.py:183-202 (symmetric_quantize_last_axis_of_batched_matrix, int8 branch)
 * weight fp32 [K][N] -> q int8 [K][N], scale fp32 [N]; scale = max(amax,1e-8)/128, q = clamp(rint(w/scale)). */
void oracle_quantize_int8_per_col(const float* w, int K, int N, int8_t* q, float* scale) {
    for (int n = 0; n < N; ++n) {
        float amax = 0.f;
        for (int k = 0; k < K; ++k) {
            float a = fabsf(w[(size_t)k * N + n]);
            if (a > amax) amax = a;
        }
        if (amax < 1e-8f) amax = 1e-8f;
        scale[n] = amax / 128.0f;
    }
    for (int k = 0; k < K; ++k)
        for (int n = 0; n < N; ++n) {
            float r = nearbyintf(w[(size_t)k * N + n] / scale[n]); /* torch.round = half-to-even */
            if (r > 127.f) r = 127.f;
            if (r < -128.f) r = -128.f;
            q[(size_t)k * N + n] = (int8_t)r;
        }
}

/* Dequantised weight W'[k][n] for the three formats:
 *   int4: W' = elem( q_s * s[k/g][n] + zs[k/g][n] )   (fp32 fma, one rounding to the activation type)
 *   int8: W' = elem( q * scale[n] )
 *   f16 : W' = W
 * Y[b][n] = elem( sum_k X[b][k] * W'[k][n] + bias[n] ), accumulation in fp32 pairwise-by-k-blocks of
 * double (we use double to make the oracle the "truth" side of the tolerance). */

static void gemm_accumulate(const float* wrow /*[N]*/, const float* xcol /*[B]*/, int B, int N, double* acc /*[B][N]*/) {
    for (int b = 0; b < B; ++b) {
        const double xv = xcol[b];
        double* a = acc + (size_t)b * N;
        for (int n = 0; n < N; ++n) a[n] += xv * (double)wrow[n];
    }
}

/* fmt: 0 = f16/bf16 weight W[K][N]; 1 = int8 q[K][N] + scale[N]; 2 = int4 packed [K][N/2] + s,zs [K/g][N];
 *      3 = int8 group-wise q_s[K][N] + s,zs [K/g][N] (device_impl.py:256-258,284-291: W' = q_s*s + zeros_x_scales) */
void oracle_dequant_gemm(const h16* x, int is_bf16, int B, int K, int N, int fmt, const void* weight,
                         const h16* scales, const h16* zeros_x_scales, int group, const h16* bias, h16* y) {
    /* column blocks so the double accumulator stays in cache and threads are independent */
    const int NB = 256;
    const int nblocks = (N + NB - 1) / NB;
#pragma omp parallel for schedule(dynamic, 1)
    for (int blk = 0; blk < nblocks; ++blk) {
        const int n0 = blk * NB, n1 = (n0 + NB < N) ? n0 + NB : N, nn = n1 - n0;
        double* acc = (double*)calloc((size_t)B * nn, sizeof(double));
        float* wrow = (float*)malloc(sizeof(float) * (size_t)nn);
        float* xcol = (float*)malloc(sizeof(float) * (size_t)B);
        for (int k = 0; k < K; ++k) {
            if (fmt == 0) {
                const h16* w = (const h16*)weight + (size_t)k * N + n0;
                for (int n = 0; n < nn; ++n) wrow[n] = elem_to_float(w[n], is_bf16);
            } else if (fmt == 1) {
                const int8_t* w = (const int8_t*)weight + (size_t)k * N + n0;
                for (int n = 0; n < nn; ++n)
                    wrow[n] = round_elem((float)w[n] * elem_to_float(scales[n0 + n], is_bf16), is_bf16);
            } else if (fmt == 3) {
                const int8_t* w = (const int8_t*)weight + (size_t)k * N + n0;
                const h16* s = scales + (size_t)(k / group) * N + n0;
                const h16* z = zeros_x_scales + (size_t)(k / group) * N + n0;
                for (int n = 0; n < nn; ++n)
                    wrow[n] = round_elem(fmaf((float)w[n], elem_to_float(s[n], is_bf16), elem_to_float(z[n], is_bf16)), is_bf16);
            } else {
                const uint8_t* w = (const uint8_t*)weight + (size_t)k * (N / 2);
                const h16* s = scales + (size_t)(k / group) * N;
                const h16* z = zeros_x_scales + (size_t)(k / group) * N;
                for (int n = 0; n < nn; ++n) {
                    int col = n0 + n;
                    uint8_t byte = w[col / 2];
                    int nib = (col & 1) ? (byte >> 4) : (byte & 0xF);
                    int qs = (nib & 8) ? nib - 16 : nib;
                    wrow[n] = round_elem(fmaf((float)qs, elem_to_float(s[col], is_bf16), elem_to_float(z[col], is_bf16)),
                                         is_bf16);
                }
            }
            for (int b = 0; b < B; ++b) xcol[b] = elem_to_float(x[(size_t)b * K + k], is_bf16);
            gemm_accumulate(wrow, xcol, B, nn, acc);
        }
        for (int b = 0; b < B; ++b)
            for (int n = 0; n < nn; ++n) {
                double v = acc[(size_t)b * nn + n];
                if (bias) v += elem_to_float(bias[n0 + n], is_bf16);
                y[(size_t)b * N + n0 + n] = float_to_elem((float)v, is_bf16);
            }
        free(acc);
        free(wrow);
        free(xcol);
    }
}

/* Faster fp32 variant of the same arithmetic for the timed CPU baseline (bench.py cpu_baseline /
 * --impl reference): identical dequant formula, fp32 accumulation, B-blocked so the compiler can
 * vectorise over n.  Results agree with oracle_dequant_gemm to fp32 rounding. */
void oracle_dequant_gemm_fast(const h16* x, int is_bf16, int B, int K, int N, int fmt, const void* weight,
                              const h16* scales, const h16* zeros_x_scales, int group, h16* y) {
    const int NB = 128;
    const int nblocks = (N + NB - 1) / NB;
    float* xf = (float*)malloc(sizeof(float) * (size_t)B * K);
    for (size_t i = 0; i < (size_t)B * K; ++i) xf[i] = elem_to_float(x[i], is_bf16);
#pragma omp parallel for schedule(dynamic, 1)
    for (int blk = 0; blk < nblocks; ++blk) {
        const int n0 = blk * NB, n1 = (n0 + NB < N) ? n0 + NB : N, nn = n1 - n0;
        float* acc = (float*)calloc((size_t)B * NB, sizeof(float));
        float wrow[128];
        for (int k = 0; k < K; ++k) {
            if (fmt == 0) {
                const h16* w = (const h16*)weight + (size_t)k * N + n0;
                for (int n = 0; n < nn; ++n) wrow[n] = elem_to_float(w[n], is_bf16);
            } else if (fmt == 1) {
                const int8_t* w = (const int8_t*)weight + (size_t)k * N + n0;
                for (int n = 0; n < nn; ++n) wrow[n] = (float)w[n] * elem_to_float(scales[n0 + n], is_bf16);
            } else if (fmt == 3) {
                const int8_t* w = (const int8_t*)weight + (size_t)k * N + n0;
                const h16* s = scales + (size_t)(k / group) * N + n0;
                const h16* z = zeros_x_scales + (size_t)(k / group) * N + n0;
                for (int n = 0; n < nn; ++n) wrow[n] = (float)w[n] * elem_to_float(s[n], is_bf16) + elem_to_float(z[n], is_bf16);
            } else {
                const uint8_t* w = (const uint8_t*)weight + (size_t)k * (N / 2) + n0 / 2;
                const h16* s = scales + (size_t)(k / group) * N + n0;
                const h16* z = zeros_x_scales + (size_t)(k / group) * N + n0;
                for (int n = 0; n < nn; n += 2) {
                    uint8_t byte = w[n / 2];
                    int lo = byte & 0xF, hi = byte >> 4;
                    lo = (lo & 8) ? lo - 16 : lo;
                    hi = (hi & 8) ? hi - 16 : hi;
                    wrow[n] = (float)lo * elem_to_float(s[n], is_bf16) + elem_to_float(z[n], is_bf16);
                    wrow[n + 1] = (float)hi * elem_to_float(s[n + 1], is_bf16) + elem_to_float(z[n + 1], is_bf16);
                }
            }
            for (int b = 0; b < B; ++b) {
                const float xv = xf[(size_t)b * K + k];
                float* a = acc + (size_t)b * NB;
                for (int n = 0; n < nn; ++n) a[n] += xv * wrow[n];
            }
        }
        for (int b = 0; b < B; ++b)
            for (int n = 0; n < nn; ++n) y[(size_t)b * N + n0 + n] = float_to_elem(acc[(size_t)b * NB + n], is_bf16);
        free(acc);
    }
    free(xf);
}

/* ------------------------------------------------------------------ */
/* Glue ops of the decode step          */
/* ------------------------------------------------------------------ */

/* RMSNorm with optional fused residual add, as bound in rtp_llm/models_py/bindings/cuda/RegisterBaseBindings.hpp:45-60
 * (rmsnorm / fused_add_rmsnorm -> 3rdparty/flashinfer/flashinfer.h:30, flashinfer norm.cuh FusedAddRMSNormKernel):
 * x = float(in) + float(residual) is kept UNROUNDED for the sum of squares and for the output; only the value stored back
 * to `residual` is rounded to the element type.  y = x * rsqrt(mean(x^2) + eps) * gamma. */
void oracle_add_rmsnorm(const h16* x, h16* residual, const h16* gamma, h16* y, int is_bf16, int rows, int hidden,
                        float eps, int has_residual) {
    float* row = (float*)malloc((size_t)hidden * sizeof(float));
    for (int r = 0; r < rows; ++r) {
        double ss = 0.0;
        for (int c = 0; c < hidden; ++c) {
            float v = elem_to_float(x[(size_t)r * hidden + c], is_bf16);
            if (has_residual) {
                v = v + elem_to_float(residual[(size_t)r * hidden + c], is_bf16);
                residual[(size_t)r * hidden + c] = float_to_elem(v, is_bf16);
            }
            row[c] = v;
            ss += (double)v * v;
        }
        float inv = 1.0f / sqrtf((float)(ss / hidden) + eps);
        for (int c = 0; c < hidden; ++c)
            y[(size_t)r * hidden + c] = float_to_elem(row[c] * inv * elem_to_float(gamma[c], is_bf16), is_bf16);
    }
    free(row);
}

/* SiLU(gate) * up on a [rows][2*inter] buffer (gate first), activation_kernels.cu silu_and_mul semantics. */
void oracle_silu_and_mul(const h16* gate_up, h16* y, int is_bf16, int rows, int inter) {
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < inter; ++c) {
            float g = elem_to_float(gate_up[(size_t)r * 2 * inter + c], is_bf16);
            float u = elem_to_float(gate_up[(size_t)r * 2 * inter + inter + c], is_bf16);
            float s = g / (1.0f + expf(-g));
            y[(size_t)r * inter + c] = float_to_elem(s * u, is_bf16);
        }
}

/* Decode-time RoPE (NeoX / rotate-half pairing, RopeStyle::Base) + K,V append into the paged cache.
 * Contract: SURVEY.md appendix C; rtp_llm/ops/fused_rope_kvcache_op.py:202-246 (call),
 * common/kernels/rotary_position_embedding.h:330-345,974-1000 (angle = pos * base^(-2i/dim), element i pairs
 * with i+dim/2), rocm/kernels/fused_rope_kvcache_kernel.cu:1297-1466 (structure: rotate q and k heads, copy v,
 * write at slot sequence_lengths[b]).  qkv [B][(Hq+2Hkv)*D] -> q_out [B][Hq*D]; K,V written in place. */
void oracle_rope_append(const h16* qkv, h16* q_out, h16* kv_pool, const int32_t* page_list,
                        const int32_t* sequence_lengths, int is_bf16, int batch, int head_num, int kv_head_num,
                        int head_dim, int max_blocks, int tokens_per_block, float rope_base) {
    const int half = head_dim / 2;
    const size_t page_elems = (size_t)kv_head_num * tokens_per_block * head_dim;
    for (int b = 0; b < batch; ++b) {
        const int pos = sequence_lengths[b];
        const h16* row = qkv + (size_t)b * (head_num + 2 * kv_head_num) * head_dim;
        const int32_t kpage = page_list[((size_t)b * 2 + 0) * max_blocks + pos / tokens_per_block];
        const int32_t vpage = page_list[((size_t)b * 2 + 1) * max_blocks + pos / tokens_per_block];
        for (int h = 0; h < head_num + kv_head_num; ++h) {
            const h16* src = row + (size_t)h * head_dim;
            h16* dst;
            if (h < head_num)
                dst = q_out + ((size_t)b * head_num + h) * head_dim;
            else
                dst = kv_pool + (size_t)kpage * page_elems +
                      ((size_t)(h - head_num) * tokens_per_block + pos % tokens_per_block) * head_dim;
            for (int i = 0; i < half; ++i) {
                float inv_freq = powf(rope_base, -2.0f * (float)i / (float)head_dim);
                float ang = (float)pos * inv_freq;
                float c = cosf(ang), s = sinf(ang);
                float x0 = elem_to_float(src[i], is_bf16), x1 = elem_to_float(src[i + half], is_bf16);
                dst[i] = float_to_elem(x0 * c - x1 * s, is_bf16);
                dst[i + half] = float_to_elem(x1 * c + x0 * s, is_bf16);
            }
        }
        for (int h = 0; h < kv_head_num; ++h) {
            const h16* src = row + (size_t)(head_num + kv_head_num + h) * head_dim;
            h16* dst = kv_pool + (size_t)vpage * page_elems +
                       ((size_t)h * tokens_per_block + pos % tokens_per_block) * head_dim;
            memcpy(dst, src, sizeof(h16) * (size_t)head_dim);
        }
    }
}

/* Greedy sampling = argmax over the vocabulary, first maximum wins
 * (rtp_llm/models_py/bindings/core/CudaSampleOp.cc:330,453 use torch argmax semantics). */
void oracle_argmax(const float* logits, int rows, int vocab, int32_t* out) {
    for (int r = 0; r < rows; ++r) {
        int best = 0;
        float bv = logits[(size_t)r * vocab];
        for (int c = 1; c < vocab; ++c) {
            float v = logits[(size_t)r * vocab + c];
            if (v > bv) {
                bv = v;
                best = c;
            }
        }
        out[r] = best;
    }
}


/* ------------------------------------------------------------------ */
/* Sampling: restatement of sampleGreedy's CUDA path
 * (rtp_llm/models_py/bindings/core/CudaSampleOp.cc:423-463; processLogits :186-279; flashinferSampleGreedy :287-421;
 *  penalty kernels rtp_llm/models_py/bindings/common/kernels/sampling_penalty_kernels.cu:26-53,129-185).
 * Temperature -> penalties over the distinct history tokens -> softmax (written back over the logits) -> top-k (keep
 * p >= k-th largest) -> top-p inside what is left -> renormalise -> inverse-CDF draw in index order with uniform[r].
 * The reference draws with flashinfer's rejection sampler (Philox); the kept set and its renormalised probabilities are
 * deterministic and pinned by the reference's test vectors (CudaSamplerTest.cc:518-568 top_k = 1; :905-976 penalties +
 * output_all_probs; :986-1057 do_sample + top_k renormalisation), the draw itself is distribution-equivalent. */
static int cmp_desc(const void* a, const void* b) {
    float x = *(const float*)a, y = *(const float*)b;
    return x < y ? 1 : (x > y ? -1 : 0);
}
void oracle_sample(float* logits, int rows, int vocab, const int32_t* history, const int32_t* hist_len, int hist_stride,
                   const float* temperature, const float* repetition, const float* presence, const float* frequency,
                   const int32_t* top_k, const float* top_p, const float* uniform, const uint8_t* process, int32_t* token_out,
                   float* token_prob_out, float* probs_out) {
    int* cnt = (int*)calloc((size_t)vocab, sizeof(int));
    float* sorted = (float*)malloc((size_t)vocab * sizeof(float));
    for (int r = 0; r < rows; ++r) {
        float* lg = logits + (size_t)r * vocab;
        const int proc = !process || process[r];
        if (proc) {
            if (temperature && temperature[r] != 1.0f) {
                const float inv_t = 1.0f / (temperature[r] + 1e-6f);
                for (int i = 0; i < vocab; ++i) lg[i] *= inv_t;
            }
            const float rep = repetition ? repetition[r] : 1.f, pre = presence ? presence[r] : 0.f, fre = frequency ? frequency[r] : 0.f;
            if (history && (rep != 1.f || pre != 0.f || fre != 0.f)) {
                const int32_t* h = history + (size_t)r * hist_stride;
                for (int i = 0; i < hist_len[r]; ++i)
                    if (h[i] >= 0 && h[i] < vocab) cnt[h[i]]++;
                for (int i = 0; i < hist_len[r]; ++i) {
                    const int t = h[i];
                    if (t < 0 || t >= vocab || cnt[t] == 0) continue;
                    float v = lg[t];
                    v = v < 0.f ? v * rep : v / rep;
                    v -= pre;
                    v -= fre * (float)cnt[t];
                    lg[t] = v;
                    cnt[t] = 0;
                }
            }
        }
        float m = -INFINITY;
        for (int i = 0; i < vocab; ++i) m = lg[i] > m ? lg[i] : m;
        double z = 0.0;
        for (int i = 0; i < vocab; ++i) z += exp((double)lg[i] - m);
        for (int i = 0; i < vocab; ++i) lg[i] = (float)(exp((double)lg[i] - m) / z);
        int k = top_k[r];
        if (k <= 0 || k > vocab) k = vocab;
        float tp = top_p[r];
        if (fabsf(tp) < 1e-7f) tp = 1.f;
        float thr = 0.f;
        if (k < vocab || tp < 1.f) {
            memcpy(sorted, lg, (size_t)vocab * sizeof(float));
            qsort(sorted, (size_t)vocab, sizeof(float), cmp_desc);
            if (k < vocab) thr = sorted[k - 1];
            if (tp < 1.f) {
                double zk = 0.0;
                for (int i = 0; i < vocab && sorted[i] >= thr; ++i) zk += sorted[i];
                double acc = 0.0;
                for (int i = 0; i < vocab && sorted[i] >= thr; ++i) {
                    acc += sorted[i];
                    if (acc >= (double)tp * zk) {
                        thr = sorted[i];
                        break;
                    }
                }
            }
        }
        double total = 0.0;
        for (int i = 0; i < vocab; ++i) total += lg[i] >= thr ? lg[i] : 0.f;
        const double target = k == 1 ? 0.0 : (double)uniform[r] * total;
        double acc = 0.0;
        int pick = -1;
        for (int i = 0; i < vocab; ++i) {
            if (lg[i] >= thr) {
                acc += lg[i];
                pick = i;
                if (acc > target) break;
            }
        }
        token_out[r] = pick;
        if (token_prob_out) token_prob_out[r] = pick >= 0 ? (float)(lg[pick] / total) : 0.f;
        if (probs_out)
            for (int i = 0; i < vocab; ++i) probs_out[(size_t)r * vocab + i] = lg[i] >= thr ? (float)(lg[i] / total) : 0.f;
    }
    free(cnt);
    free(sorted);
}


/* ------------------------------------------------------------------ */
/* The full decode RoPE contract (fused_rope_kvcache_op.py:202-246; rotary_position_embedding.h:322-442,889-1062;
 * logn: decoder_masked_multihead_attention_utils.h:2097-2102). Same argument meaning as b200_rope_append_ex. */
typedef struct {
    int style, dim;
    float base, scale, factor1, factor2;
    int max_pos;
    float extrapolation_factor, mscale;
} oracle_rope_config;

static void oracle_rope_coef(const oracle_rope_config* c, int zid, int pos, float* cs, float* sn) {
    float base = c->base;
    const int seq_len = pos + 1;
    if (c->style == 3 && seq_len > c->max_pos)
        base = c->base * powf((c->scale * seq_len / c->max_pos) - (c->scale - 1.f), c->dim / (c->dim - 2.0f));
    if (c->style == 4 && seq_len > c->max_pos) {
        const float ctx = logf((float)seq_len / c->max_pos) / logf(2.0f) + 1.0f;
        float ntk = powf(2.0f, ceilf(ctx)) - 1.f;
        if (ntk < 1.0f) ntk = 1.0f;
        base = c->base * powf(ntk, (float)c->dim / (c->dim - 2));
    }
    float a = (float)pos / powf(base, zid / (float)c->dim);
    float sc = 1.f;
    if (c->style == 1 || c->style == 7) {
        a = a / c->scale;
    } else if (c->style == 5) {
        const float pi = 3.141592654f;
        const float t1 = 2.f * logf((float)(int)c->base);
        int low = (int)floorf(c->dim * logf((float)c->max_pos / (c->factor2 * 2 * pi)) / t1);
        int high = (int)ceilf(c->dim * logf((float)c->max_pos / (c->factor1 * 2 * pi)) / t1);
        float lo = (float)(low > 0 ? low : 0), hi = (float)(high < c->dim - 1 ? high : c->dim - 1);
        if (lo == hi) hi += 0.001f;
        float ramp = (zid / 2 - lo) / (hi - lo);
        ramp = ramp < 0.f ? 0.f : (ramp > 1.f ? 1.f : ramp);
        const float mask = (1.f - ramp) * c->extrapolation_factor;
        a = (a / c->scale) * (1.f - mask) + a * mask;
        sc = c->mscale;
    } else if (c->style == 6) {
        const float pi = 3.141592654f;
        const float wavelen = 2 * pi / a;
        const float low_w = c->max_pos / c->factor1, high_w = c->max_pos / c->factor2;
        if (wavelen < high_w) {
        } else if (wavelen > low_w) {
            a = a / c->scale;
        } else {
            const float smooth = (c->max_pos / wavelen - c->factor1) / (c->factor2 - c->factor1);
            a = (1 - smooth) * a / c->scale + smooth * a;
        }
    }
    *cs = sc * cosf(a);
    *sn = sc * sinf(a);
}

void oracle_rope_append_ex(const h16* qkv, const h16* bias, h16* q_out, h16* kv_pool, const int32_t* page_list,
                           const int32_t* sequence_lengths, const int32_t* position_ids, const float* cos_sin_cache,
                           int cache_positions, const oracle_rope_config* cfg, int use_logn, int is_bf16, int batch,
                           int head_num, int kv_head_num, int head_dim, int max_blocks, int tokens_per_block) {
    const int rhalf = cfg->dim / 2;
    const size_t page_elems = (size_t)kv_head_num * tokens_per_block * head_dim;
    for (int b = 0; b < batch; ++b) {
        const int slot = sequence_lengths[b];
        const int pos = (position_ids && position_ids[b] > 0) ? position_ids[b] : slot;
        const h16* row = qkv + (size_t)b * (head_num + 2 * kv_head_num) * head_dim;
        for (int h = 0; h < head_num + 2 * kv_head_num; ++h) {
            const h16* src = row + (size_t)h * head_dim;
            const int is_q = h < head_num, is_v = h >= head_num + kv_head_num;
            h16* dst;
            if (is_q) {
                dst = q_out + ((size_t)b * head_num + h) * head_dim;
            } else {
                const int kvh = is_v ? h - head_num - kv_head_num : h - head_num;
                const int32_t page = page_list[((size_t)b * 2 + (is_v ? 1 : 0)) * max_blocks + slot / tokens_per_block];
                dst = kv_pool + (size_t)page * page_elems + ((size_t)kvh * tokens_per_block + slot % tokens_per_block) * head_dim;
            }
            const int rotate = cfg->style != 0 && !is_v;
            const float logn = (is_q && use_logn && pos > cfg->max_pos) ? logf((float)(pos + 1)) / logf((float)cfg->max_pos) : 1.f;
            for (int c = 0; c < head_dim; ++c) {
                float v = elem_to_float(src[c], is_bf16);
                if (bias) v = round_elem(v + elem_to_float(bias[(size_t)h * head_dim + c], is_bf16), is_bf16);
                if (c >= cfg->dim) dst[c] = float_to_elem(logn != 1.f ? v * logn : v, is_bf16);
            }
            for (int i = 0; i < rhalf; ++i) {
                float x0 = elem_to_float(src[i], is_bf16), x1 = elem_to_float(src[i + rhalf], is_bf16);
                if (bias) {
                    x0 = round_elem(x0 + elem_to_float(bias[(size_t)h * head_dim + i], is_bf16), is_bf16);
                    x1 = round_elem(x1 + elem_to_float(bias[(size_t)h * head_dim + i + rhalf], is_bf16), is_bf16);
                }
                if (rotate) {
                    float cs, sn;
                    if (cos_sin_cache && pos < cache_positions) {
                        cs = cos_sin_cache[((size_t)pos * rhalf + i) * 2];
                        sn = cos_sin_cache[((size_t)pos * rhalf + i) * 2 + 1];
                    } else {
                        oracle_rope_coef(cfg, 2 * i, pos, &cs, &sn);
                    }
                    const float r0 = cs * x0 - sn * x1, r1 = cs * x1 + sn * x0;
                    x0 = r0;
                    x1 = r1;
                }
                if (logn != 1.f) {
                    x0 *= logn;
                    x1 *= logn;
                }
                dst[i] = float_to_elem(x0, is_bf16);
                dst[i + rhalf] = float_to_elem(x1, is_bf16);
            }
        }
    }
}


/* Per-head RMSNorm of q and k heads, in place: rtp_llm/models_py/bindings/cuda/kernels/fused_qk_rmsnorm.cu:24-78
 * (fp32 sum of squares over the head, val * rsqrt(mean + eps) * gamma (+ bias), rounded to the element type). */
void oracle_qk_rmsnorm(h16* qkv, const h16* q_gamma, const h16* k_gamma, const h16* q_bias, const h16* k_bias, int is_bf16,
                       int rows, int head_num, int kv_head_num, int head_dim, float eps) {
    for (int r = 0; r < rows; ++r)
        for (int h = 0; h < head_num + kv_head_num; ++h) {
            h16* x = qkv + ((size_t)r * (head_num + 2 * kv_head_num) + h) * head_dim;
            const h16* g = h < head_num ? q_gamma : k_gamma;
            const h16* bb = h < head_num ? q_bias : k_bias;
            float ss = 0.f;
            for (int c = 0; c < head_dim; ++c) {
                const float v = elem_to_float(x[c], is_bf16);
                ss += v * v;
            }
            const float scale = 1.0f / sqrtf(ss / (float)head_dim + eps);
            for (int c = 0; c < head_dim; ++c) {
                float y = elem_to_float(x[c], is_bf16) * scale * elem_to_float(g[c], is_bf16);
                if (bb) y += elem_to_float(bb[c], is_bf16);
                x[c] = float_to_elem(y, is_bf16);
            }
        }
}
