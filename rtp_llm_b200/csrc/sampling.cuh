// Token sampling for the decode step: logits post-processing + softmax + top-k / top-p filtering + one draw per row.
//
// Replaces the CUDA path behind sampleGreedy (/root/reference/rtp_llm/models_py/bindings/core/CudaSampleOp.cc:423-463 ->
// processLogits :186-279 and flashinferSampleGreedy :287-421):
//   temperature      logit *= 1 / (T + 1e-6)                       (common/kernels/sampling_penalty_kernels.cu:26-53)
//   repetition       every DISTINCT token of the row's history once: logit = logit < 0 ? logit*r : logit/r; logit -= presence;
//   presence /       logit -= frequency * count                     (sampling_penalty_kernels.cu:129-185)
//   frequency
//   softmax, then the probabilities replace the logits in place      (CudaSampleOp.cc:291-295)
//   top_k == 1       argmax, lowest index on ties                    (:330-331, :453)
//   top_k            keep p >= (k-th largest p), renormalise         (flashinfer top_k_renorm_probs semantics)
//   top_p            of what is left keep the smallest set of largest p whose mass reaches top_p, renormalise
//   draw             inverse CDF over the kept tokens in index order with a caller-supplied uniform u in [0, 1)
// The reference draws through flashinfer's rejection sampler with a Philox stream; the DISTRIBUTION is the same
// (its own accuracy test compares distributions, CudaSamplerTest.cc:143-147), the random stream is not: here the caller
// owns the randomness (one float per row), which makes the op deterministic and replayable in a CUDA graph.
// One CTA per row; thresholds are found by bisection on the float bit pattern (no sort, no vocabulary-sized scratch).
#pragma once
#include "ptx.cuh"

namespace b200 {

constexpr int kSampleThreads = 1024;

struct SampleParams {
    float* logits;              // [rows][vocab] fp32, in: logits, out: probabilities (softmax after penalties)
    float* probs_out;           // optional [rows][vocab]: renormalised probabilities of the kept set (0 elsewhere)
    const int32_t* history;     // optional [rows][hist_stride] token ids generated / prompted so far
    const int32_t* hist_len;    // [rows] valid entries of history
    int32_t* count_ws;          // [rows][vocab] int32, zero on entry, zero on exit (only needed with penalties)
    const float* temperature;   // optional [rows]
    const float* repetition;    // optional [rows] (1 = off)
    const float* presence;      // optional [rows] (0 = off)
    const float* frequency;     // optional [rows] (0 = off)
    const int32_t* top_k;       // [rows]; <= 0: no limit
    const float* top_p;         // [rows]; >= 1 or ~0: no limit
    const float* uniform;       // [rows] in [0, 1)
    const uint8_t* process;     // optional [rows]: 0 = skip temperature / penalties for this row (do_sample == false)
    int32_t* token_out;         // [rows]
    float* token_prob_out;      // optional [rows]: renormalised probability of the drawn token (for cum_log_probs)
    int rows, vocab, hist_stride;
};

__device__ __forceinline__ float block_reduce_sum(float v, float* s_red) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kSampleThreads / 32; ++w) t += s_red[w];      // same order in every thread: identical result
    return t;
}
__device__ __forceinline__ float block_reduce_max(float v, float* s_red) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = -INFINITY;
#pragma unroll
    for (int w = 0; w < kSampleThreads / 32; ++w) t = fmaxf(t, s_red[w]);
    return t;
}

__global__ void __launch_bounds__(kSampleThreads) sample_kernel(const SampleParams p) {
    __shared__ float s_red[kSampleThreads / 32];
    __shared__ float s_scan[kSampleThreads];
    __shared__ int s_pick;
    const int row = blockIdx.x, tid = threadIdx.x, V = p.vocab;
    float* lg = p.logits + (size_t)row * V;
    const bool process = !p.process || p.process[row];

    // ---- temperature + penalties (in place)
    if (process) {
        if (p.temperature) {
            const float inv_t = 1.0f / (p.temperature[row] + 1e-6f);
            if (p.temperature[row] != 1.0f)
                for (int i = tid; i < V; i += kSampleThreads) lg[i] *= inv_t;
        }
        const float rep = p.repetition ? p.repetition[row] : 1.f, pre = p.presence ? p.presence[row] : 0.f,
                    fre = p.frequency ? p.frequency[row] : 0.f;
        if (p.history && p.count_ws && (rep != 1.f || pre != 0.f || fre != 0.f)) {
            __syncthreads();
            int* cnt = p.count_ws + (size_t)row * V;
            const int32_t* h = p.history + (size_t)row * p.hist_stride;
            const int L = p.hist_len[row];
            for (int i = tid; i < L; i += kSampleThreads) {
                const int t = h[i];
                if (t >= 0 && t < V) atomicAdd(&cnt[t], 1);
            }
            __syncthreads();
            // whoever fetches a non-zero count owns that token: penalise it once and leave the workspace clean
            for (int i = tid; i < L; i += kSampleThreads) {
                const int t = h[i];
                if (t < 0 || t >= V) continue;
                const int c = atomicExch(&cnt[t], 0);
                if (c > 0) {
                    float v = lg[t];
                    v = v < 0.f ? v * rep : v / rep;
                    v -= pre;
                    v -= fre * (float)c;
                    lg[t] = v;
                }
            }
        }
        __syncthreads();
    }

    // ---- softmax in place
    float m = -INFINITY;
    for (int i = tid; i < V; i += kSampleThreads) m = fmaxf(m, lg[i]);
    m = block_reduce_max(m, s_red);
    float z = 0.f;
    for (int i = tid; i < V; i += kSampleThreads) {
        const float e = __expf(lg[i] - m);
        lg[i] = e;
        z += e;
    }
    z = block_reduce_sum(z, s_red);
    const float inv_z = 1.f / z;
    float pmax = 0.f;
    for (int i = tid; i < V; i += kSampleThreads) {
        const float q = lg[i] * inv_z;
        lg[i] = q;
        pmax = fmaxf(pmax, q);
    }
    pmax = block_reduce_max(pmax, s_red);
    __syncthreads();

    // ---- thresholds by bisection on the bit pattern (probabilities are non-negative: bit order == value order)
    int k = p.top_k[row];
    if (k <= 0 || k > V) k = V;
    float tp = p.top_p[row];
    if (fabsf(tp) < 1e-7f) tp = 1.f;                       // CudaSampleOp.cc:317
    float thr = 0.f;                                       // keep p >= thr
    if (k < V) {
        uint32_t lo = 0u, hi = __float_as_uint(pmax);      // invariant: count(p >= lo) >= k
        while (lo < hi) {
            const uint32_t mid = lo + ((hi - lo + 1u) >> 1);
            const float fm = __uint_as_float(mid);
            float c = 0.f;
            for (int i = tid; i < V; i += kSampleThreads) c += lg[i] >= fm ? 1.f : 0.f;
            c = block_reduce_sum(c, s_red);
            if (c >= (float)k) lo = mid;
            else hi = mid - 1u;
        }
        thr = __uint_as_float(lo);
    }
    if (tp < 1.f) {
        float zk = 0.f;
        for (int i = tid; i < V; i += kSampleThreads) zk += lg[i] >= thr ? lg[i] : 0.f;
        zk = block_reduce_sum(zk, s_red);
        const float need = tp * zk;
        uint32_t lo = __float_as_uint(thr), hi = __float_as_uint(pmax);   // invariant: mass(p >= lo) >= need
        while (lo < hi) {
            const uint32_t mid = lo + ((hi - lo + 1u) >> 1);
            const float fm = __uint_as_float(mid);
            float s = 0.f;
            for (int i = tid; i < V; i += kSampleThreads) s += lg[i] >= fm ? lg[i] : 0.f;
            s = block_reduce_sum(s, s_red);
            if (s >= need) lo = mid;
            else hi = mid - 1u;
        }
        thr = __uint_as_float(lo);
    }

    // ---- draw: inverse CDF over the kept tokens in index order (contiguous chunk per thread, block scan of chunk sums)
    const int per = (V + kSampleThreads - 1) / kSampleThreads;
    const int i0 = tid * per, i1 = min(V, i0 + per);
    float mine = 0.f;
    for (int i = i0; i < i1; ++i) mine += lg[i] >= thr ? lg[i] : 0.f;
    s_scan[tid] = mine;
    __syncthreads();
    for (int off = 1; off < kSampleThreads; off <<= 1) {   // Hillis-Steele inclusive scan
        const float add = tid >= off ? s_scan[tid - off] : 0.f;
        __syncthreads();
        s_scan[tid] += add;
        __syncthreads();
    }
    const float total = s_scan[kSampleThreads - 1];
    const float target = k == 1 ? 0.f : p.uniform[row] * total;   // top_k == 1: the lowest index among the maxima (argmax semantics)
    if (tid == 0) s_pick = 0;
    __syncthreads();
    // exactly one thread owns the target: the first whose inclusive chunk sum exceeds it
    const float before = s_scan[tid] - mine;
    if (mine > 0.f && before <= target && target < s_scan[tid]) {
        float c = before;
        int pick = -1;
        for (int i = i0; i < i1; ++i) {
            if (lg[i] >= thr) {
                c += lg[i];
                pick = i;                  // the chunk's last kept token is the fallback when rounding undershoots
                if (c > target) break;
            }
        }
        s_pick = pick;
    }
    __syncthreads();
    const int tok = s_pick;
    const float inv_total = 1.f / total;
    if (tid == 0) {
        p.token_out[row] = tok;
        if (p.token_prob_out) p.token_prob_out[row] = tok >= 0 ? lg[tok] * inv_total : 0.f;
    }
    if (p.probs_out) {
        float* po = p.probs_out + (size_t)row * V;
        for (int i = tid; i < V; i += kSampleThreads) po[i] = lg[i] >= thr ? lg[i] * inv_total : 0.f;
    }
}

}  // namespace b200
