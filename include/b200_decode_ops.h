/*
 * b200_decode_ops.h -- C ABI of libb200_decode.so: the B200-native (sm_100a) decode hot path of rtp-llm.
 *
 * This is the drop-in boundary.  Every entry point takes plain device pointers, sizes and a cudaStream_t (passed as
 * void*), returns 0 on success or a negative B200_E* code (message via b200_last_error()), never synchronises the
 * stream, never allocates device memory and is CUDA-graph capturable.  No torch types cross this boundary.
 *
 * Each function names the reference interface it replaces (paths relative to the rtp-llm repo root).
 * INTEGRATION.md shows the reference-side binding for each.
 */
#ifndef B200_DECODE_OPS_H
#define B200_DECODE_OPS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK 0
#define B200_EINVAL (-1)      /* bad argument (shape, alignment, unsupported configuration) */
#define B200_ECUDA (-2)       /* a CUDA runtime / driver call failed */
#define B200_EUNSUPPORTED (-3) /* not an sm_100 device, or feature outside the built scope */

/* weight formats of b200_wo_gemm */
#define B200_FMT_F16 0   /* W stored [N][K] (K contiguous), element type = activation type */
#define B200_FMT_INT8 1  /* per-column symmetric INT8 (device_impl.py:183-222), packed by b200_pack_w8 */
#define B200_FMT_INT4 2  /* GPTQ/AWQ group-128 INT4 (device_impl.py:242-300), packed by b200_pack_w4 */
#define B200_FMT_INT8G 3 /* GPTQ/AWQ group-128 INT8 (device_impl.py:256-258 is_int8 branch: q_s = q_u - 128), packed by b200_pack_w8g */

/* flags of b200_wo_gemm */
#define B200_GEMM_PDL 1  /* launch with programmatic dependent launch (weights prefetched before the upstream grid ends) */
#define B200_GEMM_SILU_MUL 2 /* fused SiLU(gate)*up epilogue: the weight's columns were interleaved per 128-feature tile as
                                [64 gate | 64 matching up] BEFORE packing (rtp_llm_b200.ops.interleave_gate_up); y is [B][N/2] */

/* Last error message of the calling thread ("" if none). */
const char* b200_last_error(void);

/* 0 if device `device` is sm_100 (B200) and the library's kernels can run there, else B200_EUNSUPPORTED. */
int b200_device_check(int device);

/* Library-wide switch: launch the kernels of the decode chain (norms, rope, attention, GEMMs) with programmatic dependent
 * launch so each kernel's launch latency / prologue / weight prefetch overlaps its predecessor's tail. Every such kernel
 * executes griddepcontrol.wait before touching dependent data, so results are unchanged. Default off. */
int b200_set_pdl(int enable);

/* Number of kernels this library has launched since load (all threads); used by bench.py for "gpu_launches". */
uint64_t b200_launch_count(void);

/* Launch shapes the library will choose (host-only, no CUDA work): how many sequence splits / 64-token tiles per CTA the
 * attention takes for `units` = batch*kv_heads and a given max_seq_len, and the split-K of a GEMM. For capacity planning
 * (workspaces) and for tests of the heuristics. */
int b200_plan_attn_split(int units, int max_seq_len, int* nsplit, int* tiles_per_split);
int b200_plan_gemm_split(int K, int N, int* nsplit, int* k_blocks_per_split);

/* ------------------------------------------------------------------------------------------------ indexing */

/* Replaces invokeConvertOffsetToBlockArrayData
 *   rtp_llm/models_py/bindings/common/kernels/kv_cache_kernels.cu:66-80 (kernel :49-63).
 * block_ids [batch][max_blocks] int32 -> page_list [batch][1][2][max_blocks] int32, K page = 2*id, V page = 2*id+1. */
int b200_convert_block_table(int32_t* page_list, const int32_t* block_ids, int batch, int max_blocks, void* stream);

/* Replaces invokeMhaPagedAttnPlan  rtp_llm/models_py/bindings/cuda/kernels/mha_paged_attn_plan.cu:100-180 (kernel :16-97).
 * prefix_lengths == NULL selects decode mode (one token per sequence, seq_len = sequence_lengths[b]+1).
 * block_ids may be NULL (page_indice untouched). batch <= 1024. */
int b200_paged_attn_plan(const int32_t* input_lengths, const int32_t* sequence_lengths, const int32_t* prefix_lengths,
                         const int32_t* block_ids, int batch, int max_blocks, int tokens_per_block,
                         int32_t* paged_kv_last_page_len, int32_t* decode_page_indptr, int32_t* page_indice,
                         int32_t* batch_indice, int32_t* positions, void* stream);

/* ------------------------------------------------------------------------------------------------ attention */

/* Bytes of scratch b200_paged_decode_attn needs for this problem size. The buffer must be zero-filled ONCE after
 * allocation (cudaMemset); the kernels leave it zeroed where it matters (semaphores self-reset). */
size_t b200_paged_decode_attn_workspace_bytes(size_t batch, size_t head_num, size_t kv_head_num, size_t max_seq_len);

/* Replaces runXqa  rtp_llm/models_py/bindings/cuda/ops/CudaXqa.h:65-84 (called from XQAAttnOp::forward,
 * rtp_llm/models_py/bindings/cuda/XQAAttnOp.cc:119-156) and the flashinfer trtllm-gen call at
 * rtp_llm/models_py/modules/factory/attention/cuda_impl/trtllm_gen.py:531-545.
 *   q            [batch][head_num][head_dim]                        fp16 / bf16
 *   out          [batch][head_num*head_dim]                         same type
 *   kv_pool      layer cache [pages][2][kv_head_num][page_size][head_dim] (OpDefs.h:201-202), addressed as a pool of
 *                2*pages pages of [kv_head_num][page_size][head_dim]
 *   page_list    [batch][1][2][max_blocks_per_seq] int32 from b200_convert_block_table
 *   sequence_lengths [batch] tokens ALREADY in the cache (device-accessible; pinned host memory works, as in the
 *                reference); attention covers positions 0..sequence_lengths[b] inclusive
 *   max_seq_len  host upper bound of sequence_lengths[b]+1 (sizes the sequence split; as XQAAttnOp.cc:149)
 *   softmax scale = q_scale * head_dim^-1/2.
 * Supported: head_dim 64 / 128 / 256, head_num/kv_head_num in 1..16, page_size in {16,32,64,128}. */
int b200_paged_decode_attn(const void* q, int is_bf16, void* out, size_t head_num, size_t kv_head_num, size_t head_dim,
                           size_t batch, size_t max_blocks_per_seq, size_t max_seq_len, size_t page_size,
                           const void* kv_pool, const int32_t* page_list, const uint32_t* sequence_lengths,
                           float q_scale, void* workspace, size_t workspace_bytes, void* stream);

/* q_len > 1 query tokens per sequence (verification of speculative tokens: runXqa's max_q_len, CudaXqa.h:59-84; trtllm-gen's
 * q_len_per_req, trtllm_gen.py:544): q [batch][q_len][head_num][head_dim], out [batch][q_len][head_num*head_dim]; the K/V of all
 * q_len new tokens are already in the cache; sequence_lengths[b] = tokens cached BEFORE them, query j attends to positions
 * 0 .. sequence_lengths[b] + j. (head_num / kv_head_num) * q_len <= 16; max_seq_len bounds sequence_lengths[b] + q_len; size the
 * workspace with b200_paged_decode_attn_workspace_bytes(batch * q_len, ...). q_len == 1 is b200_paged_decode_attn. */
int b200_paged_decode_attn_multi(const void* q, int is_bf16, void* out, size_t head_num, size_t kv_head_num, size_t head_dim,
                                 size_t batch, size_t q_len, size_t max_blocks_per_seq, size_t max_seq_len, size_t page_size,
                                 const void* kv_pool, const int32_t* page_list, const uint32_t* sequence_lengths, float q_scale,
                                 void* workspace, size_t workspace_bytes, void* stream);

/* The same attention with RoPE + KV append FUSED IN (what XQA does with USE_INPUT_KV + ROPE_STYLE, 3rdparty/xqa/mha.h:82-86):
 * replaces FusedRopeKVCacheDecodeOp.forward + XQAAttnOp.forward (rtp_llm/ops/fused_rope_kvcache_op.py:202-246 then
 * bindings/cuda/XQAAttnOp.cc:119-156) by one launch. `qkv` is the qkv GEMM output [batch][(head_num + 2 kv_head_num) * head_dim],
 * un-rotated; q is rotated on load, the CTA that owns the tile of position sequence_lengths[b] rotates k and appends k, v to
 * `kv_pool` before streaming it. RopeStyle::Base, NeoX pairing, same arithmetic as b200_rope_append: cache contents and attention
 * output are bit-identical to the two-call sequence. */
int b200_paged_decode_attn_rope(const void* qkv, int is_bf16, void* out, size_t head_num, size_t kv_head_num, size_t head_dim,
                                size_t batch, size_t max_blocks_per_seq, size_t max_seq_len, size_t page_size, void* kv_pool,
                                const int32_t* page_list, const uint32_t* sequence_lengths, float q_scale, float rope_base,
                                void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------ weight-only GEMM */

/* Load-time weight re-layout.  Replace CudaImpl.preprocess_weights_for_mixed_gemm (rtp_llm/device/device_impl.py:392-479):
 * instead of the FT sm80 interleave, weights become per-(128 features x 128 k) blobs that one TMA bulk copy stages.
 * Inputs are the loader's UN-permuted tensors (device pointers):
 *   b200_pack_w4: q_packed uint8 [K][N/2] (low nibble = even column, two's complement q_s = q_u - 8, device_impl.py:204-209),
 *                 scales, zeros_x_scales [K/128][N] 16-bit in the ACTIVATION type (fp16 or bf16)
 *   b200_pack_w8: q int8 [K][N] (device_impl.py:183-202); the per-column scale stays a separate [N] tensor
 *   b200_pack_w8g: q int8 [K][N] (8-bit group-wise checkpoints, q_s = q_u - 128, device_impl.py:256-274) + scales,
 *                 zeros_x_scales [K/128][N] as for b200_pack_w4; dequant W' = q_s * s + zeros_x_scales (device_impl.py:284-291)
 * K % 128 == 0; N % 2 == 0 (int4). Output size: b200_wo_gemm_packed_bytes(fmt, K, N). */
size_t b200_wo_gemm_packed_bytes(int fmt, int K, int N);
int b200_pack_w4(const uint8_t* q_packed, const void* scales, const void* zeros_x_scales, int K, int N, int group,
                 void* blob, void* stream);
int b200_pack_w8(const int8_t* q, int K, int N, void* blob, void* stream);
int b200_pack_w8g(const int8_t* q, const void* scales, const void* zeros_x_scales, int K, int N, int group, void* blob,
                  void* stream);

/* Scratch for split-K / stream-K partials + semaphores (+ the grid-barrier words of a stand-alone call); zero-fill once
 * after allocation. Always required for INT8 / INT4 weights (at least 16 KiB). */
size_t b200_wo_gemm_workspace_bytes(int max_batch, int N, int K);

/* Y[B][N] = X[B][K] . W' (+ bias).  The compute behind LinearBase.forward
 * (rtp_llm/models_py/modules/factory/linear/linear_base.py:81; call sites modules/hybrid/causal_attention.py:83,90,
 * dense_mlp.py:99,103; lm_head cpp/models/PyWrappedModel.cc:1041-1046) for weight-only INT4/INT8 and FP16 weights.
 *   w         B200_FMT_INT4/INT8/INT8G: blob from b200_pack_w4/8/8g; B200_FMT_F16: W [N][K] (K contiguous)
 *   col_scale INT8 only: [N] per-column scale in the activation type
 *   bias      optional [N]
 * B <= 128, K % 128 == 0, x / y / w 16-byte aligned. */
int b200_wo_gemm(int fmt, int is_bf16, const void* x, int B, int K, int N, const void* w, const void* col_scale,
                 const void* bias, void* y, void* workspace, size_t workspace_bytes, int flags, void* stream);

/* ------------------------------------------------------------------------------------------------ glue ops */

/* fused_add_rmsnorm / rmsnorm (rtp_llm/models_py/bindings/cuda/RegisterBaseBindings.hpp:45-160).
 * residual may be NULL; otherwise residual += x (in place) and y = rmsnorm(residual) * gamma. hidden % 8 == 0. */
int b200_add_rmsnorm(const void* x, void* residual, const void* gamma, void* y, int is_bf16, int rows, int hidden,
                     float eps, void* stream);
/* Per-head RMSNorm of the q and k heads of qkv [rows][(head_num + 2 kv_head_num) * head_dim], in place (QK-norm before RoPE):
 * invokeFusedQkRmsNorm, rtp_llm/models_py/bindings/cuda/kernels/fused_qk_rmsnorm.cu:80-150 (call site model_desc/qwen3.py:57-79).
 * gamma / bias are [head_dim]; biases optional (both or none); head_dim % 64 == 0. */
int b200_qk_rmsnorm(void* qkv, const void* q_gamma, const void* k_gamma, const void* q_bias, const void* k_bias, int is_bf16,
                    int rows, int head_num, int kv_head_num, int head_dim, float eps, void* stream);
/* silu_and_mul: y[r][c] = silu(gate_up[r][c]) * gate_up[r][inter + c]. */
int b200_silu_and_mul(const void* gate_up, void* y, int is_bf16, int rows, int inter, void* stream);
/* FusedRopeKVCacheDecodeOp.forward (rtp_llm/ops/fused_rope_kvcache_op.py:202-246), RopeStyle::Base, NeoX pairing:
 * rotates q and k of qkv [B][(Hq+2Hkv)*D] at position sequence_lengths[b], writes q to q_out [B][Hq*D] and K,V of the
 * new token into the paged cache. */
int b200_rope_append(const void* qkv, void* q_out, void* kv_pool, const int32_t* page_list,
                     const int32_t* sequence_lengths, int is_bf16, int batch, int head_num, int kv_head_num,
                     int head_dim, int max_blocks_per_seq, int page_size, float rope_base, void* stream);
/* The full decode contract of FusedRopeKVCacheDecodeOp.forward (rtp_llm/ops/fused_rope_kvcache_op.py:202-246; device code
 * rtp_llm/models_py/bindings/common/kernels/rotary_position_embedding.h:322-442,889-1062). b200_rope_config mirrors the
 * fields of RopeConfig (rtp_llm/cpp/model_utils/RopeConfig.h:21-42) the decode path reads.
 *   position_ids     optional [batch]; an entry > 0 replaces sequence_lengths[b] as the rotation position (:1040-1042);
 *                    K/V are still appended at slot sequence_lengths[b]
 *   cos_sin_cache    optional float2 [cache_positions][dim/2] = (cos, sin), the layout getRopeCache produces with
 *                    interleave = true (cpp/model_utils/RopeCache.cc:16-85; used for Base and Yarn); without it the
 *                    coefficients are computed in the kernel with the reference's formulas
 *   styles           0 No, 1 Base (linear scale), 3 DynamicNTK, 4 QwenDynamicNTK, 5 Yarn, 6 Llama3, 7 Mrope (as Base);
 *                    factor1 / factor2 = beta_slow / beta_fast (Yarn) or low / high frequency factor (Llama3); max_pos =
 *                    original_max_position_embeddings
 *   qkv_bias         optional [(head_num + 2 kv_head_num) * head_dim], added before the rotation
 *   use_logn_attn    q *= log(pos + 1) / log(max_pos) for pos > max_pos
 * NeoX pairing (i, i + dim/2) within the first `dim` channels of a head; the rest pass through. */
typedef struct b200_rope_config {
    int style;
    int dim;
    float base, scale, factor1, factor2;
    int max_pos;
    float extrapolation_factor, mscale;
} b200_rope_config;
int b200_rope_append_ex(const void* qkv, const void* qkv_bias, void* q_out, void* kv_pool, const int32_t* page_list,
                        const int32_t* sequence_lengths, const int32_t* position_ids, const float* cos_sin_cache,
                        int cache_positions, const b200_rope_config* cfg, int use_logn_attn, int is_bf16, int batch, int head_num,
                        int kv_head_num, int head_dim, int max_blocks_per_seq, int page_size, void* stream);
/* embedding gather out[b] = table[ids[b]] (rtp_ops.embedding). */
int b200_embedding(const int32_t* ids, const void* table, void* out, int is_bf16, int rows, int hidden, void* stream);
/* greedy sampling: out[r] = argmax(logits[r]) (lowest index on ties; CudaSampleOp.cc:330,453). dtype: 0 fp16, 1 bf16, 2 fp32 */
int b200_argmax(const void* logits, int dtype, int rows, int vocab, int32_t* out, void* stream);

/* Token sampling: the CUDA path behind sampleGreedy (rtp_llm/models_py/bindings/core/CudaSampleOp.cc:423-463; processLogits
 * :186-279, flashinferSampleGreedy :287-421; penalty kernels common/kernels/sampling_penalty_kernels.cu:26-53,129-185).
 * Per row, in this order: temperature (logit *= 1/(T+1e-6)), repetition / presence / frequency penalties over the DISTINCT
 * tokens of `history` (hist_len[r] valid ids of [rows][hist_stride]), softmax (the probabilities REPLACE the logits, as in the
 * reference), top-k (keep p >= the k-th largest; top_k <= 0 = no limit; top_k == 1 = argmax, lowest index on ties), top-p
 * within what is left (smallest set of largest p reaching top_p; |top_p| < 1e-7 or >= 1 = no limit), renormalise, and one
 * inverse-CDF draw in index order with the caller's uniform[r] in [0, 1).  Same distribution as the reference's flashinfer
 * rejection sampler, different random stream: the caller owns the randomness, so the op is deterministic and graph-replayable.
 * process[r] == 0 (do_sample false) skips temperature / penalties for that row. count_ws: [rows][vocab] int32, zero on entry
 * and on exit (only read with penalties). Optional outputs: token_prob_out[r] (renormalised probability of the drawn token,
 * for cum_log_probs += log p), probs_out [rows][vocab] (output_all_probs after renormalisation). logits fp32, contiguous. */
int b200_sample(float* logits, int rows, int vocab, const int32_t* history, const int32_t* hist_len, int hist_stride,
                int32_t* count_ws, const float* temperature, const float* repetition, const float* presence,
                const float* frequency, const int32_t* top_k, const float* top_p, const float* uniform, const uint8_t* process,
                int32_t* token_out, float* token_prob_out, float* probs_out, void* stream);

/* ------------------------------------------------------------------------------------------------ TP all-reduce over NVLink peer memory */

/* One-shot SUM all-reduce of a small [rows][hidden] fp16/bf16 tensor, the exchange after each row-parallel GEMM
 * (replaces all_reduce(t, Group.TP), rtp_llm/models_py/distributed/collective_torch.py:694-722; the reference's fast path
 * is torch symmetric memory, distributed/symm_mem.py:126-185). One process per GPU. Set-up, once per process:
 *   b200_peer_alloc(b200_peer_ar_region_bytes(max_msg), &mine, handle)  -> exchange the 64-byte handles between ranks
 *   (any host channel) -> b200_peer_open(handle_of_rank_r, &regions[r]) for r != rank; regions[rank] = mine.
 * Per call: every rank must issue the same sequence of peer_* calls on its stream. `call_parity` is ignored (kept for ABI
 * compatibility): the two data slots alternate by a device-side call counter, so CUDA graphs holding any number of calls
 * replay correctly. Deterministic: all ranks sum in rank order and get identical bits. A rank whose peers never arrive
 * traps after a few seconds of polling instead of hanging. bytes % 16 == 0, bytes <= max_message_bytes, world <= 8. */
size_t b200_peer_ar_region_bytes(size_t max_message_bytes);
int b200_peer_alloc(size_t bytes, void** ptr, void* ipc_handle_out);
int b200_peer_open(const void* ipc_handle, void** ptr);
int b200_peer_allreduce(const void* in, void* out, size_t bytes, int is_bf16, void* const* regions, size_t max_message_bytes,
                        int call_parity, int rank, int world, void* stream);

/* The same exchange fused with what follows it in the decoder layer: y = rmsnorm(allreduce(in) + residual) * gamma and
 * residual += allreduce(in), in ONE kernel (replaces all_reduce + fused_add_rmsnorm: collective_torch.py:694-722 then
 * RegisterBaseBindings.hpp:54; call sites hybrid/causal_attention.py:91-92, dense_mlp.py:104-105). Numerics are those of
 * the unfused sequence (the reduced value is rounded to the tensor type first). hidden % (8*world) == 0, hidden <= 8192. */
int b200_peer_allreduce_norm(const void* in, void* residual, const void* gamma, void* y, int is_bf16, int rows, int hidden,
                             float eps, void* const* regions, size_t max_message_bytes, int rank, int world, void* stream);

/* GEMM + reduce-scatter in ONE kernel, for the row-parallel GEMMs (o, w2) under tensor parallelism (reference sequence:
 * LinearBase.forward, then all_reduce(t, Group.TP): hybrid/causal_attention.py:90-92, dense_mlp.py:103-105). Same arguments
 * and arithmetic as b200_wo_gemm; the epilogue pushes every output element another rank owns (rank r owns the r-th 1/world
 * of the columns of every row) straight into that rank's reduce-scatter slot over NVLink while the tile is still in
 * registers; the columns this rank owns are stored to y. MUST be followed on the same stream by b200_peer_gather_norm over
 * the same y / regions with no other collective of this communicator in between (both derive the exchange epoch from the
 * communicator's device-side call counter). Bits are those of b200_wo_gemm + b200_peer_allreduce_norm.
 * N % 128 == 0, N % (8*world) == 0, N <= 8192, B*N*2 <= max_message_bytes, no SILU_MUL. */
int b200_wo_gemm_rs(int fmt, int is_bf16, const void* x, int B, int K, int N, const void* w, const void* col_scale,
                    const void* bias, void* y, void* workspace, size_t workspace_bytes, int flags, void* const* regions,
                    size_t max_message_bytes, int rank, int world, void* stream);
/* Second half of the fused exchange: reduce the own slice in rank order, all-gather, residual add, RMSNorm (the kernel of
 * b200_peer_allreduce_norm without its scatter phase, which the GEMM epilogue already did). */
int b200_peer_gather_norm(const void* in, void* residual, const void* gamma, void* y, int is_bf16, int rows, int hidden,
                          float eps, void* const* regions, size_t max_message_bytes, int rank, int world, void* stream);

/* Vocab-parallel greedy sampling: out[r] = argmax over the CONCATENATED vocabulary (rank r holds columns
 * [r*vocab_local, (r+1)*vocab_local), columns >= vocab_total are padding), lowest index on ties, identical on every rank.
 * Replaces the logits all-gather + argmax of cpp/models/PyWrappedModel.cc:915-936,1001-1052 for top_k == 1. */
int b200_peer_argmax(const void* logits, int is_bf16, int rows, int vocab_local, int vocab_total, int32_t* out,
                     void* const* regions, size_t max_message_bytes, int rank, int world, void* stream);

/* ------------------------------------------------------------------------------------------------ decode programs */

/* A decode program is a recorded sequence of the op calls above, replayed with far fewer launches -- the B200 counterpart
 * of the reference's CUDA-graph capture of the decode step (rtp_llm/cpp/cuda_graph/cuda_graph_runner.cc: capture once per
 * batch size, replay every step).  Between b200_program_begin and b200_program_end every op call of the CALLING THREAD is
 * appended to the program instead of being launched (arguments are checked and launch shapes planned at record time;
 * pointers and sizes are frozen, exactly as under stream capture).  Consecutive weight-only GEMMs, norms, rope+append,
 * embedding and block-table ops are fused into ONE persistent kernel (2 CTAs per SM, grid-wide barriers between ops,
 * stream-K GEMMs, weights prefetched across op boundaries -- csrc/decode_program.cuh); attention, FP16 GEMMs, argmax and
 * all-reduce calls are replayed as their own launches in order.  b200_program_launch never allocates or synchronises and
 * may itself be captured into a CUDA graph.  b200_program_end allocates the device-side op table (once).
 * Results are identical to issuing the same calls one by one (same kernels' arithmetic; tests compare bit for bit). */
typedef struct b200_program b200_program;
int b200_program_create(b200_program** out);
int b200_program_begin(b200_program* p);
int b200_program_end(b200_program* p);
int b200_program_launch(b200_program* p, void* stream);
int b200_program_num_ops(const b200_program* p);      /* ops fused into persistent-kernel segments */
int b200_program_num_launches(const b200_program* p); /* kernel launches per b200_program_launch */
int b200_program_set_trace(b200_program* p, void* device_buffer); /* developer: per-op, per-CTA globaltimer stamps */
int b200_program_destroy(b200_program* p);

#ifdef __cplusplus
}
#endif
#endif /* B200_DECODE_OPS_H */
