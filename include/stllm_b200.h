/* stllm_b200 - C ABI of the B200-native ST-LLM video-token hot path.
 *
 * The reference (TencentARC/ST-LLM) has no FFI/operator layer of its own: its hot path is Python calling
 * torch/cuBLAS/cuDNN ops.  This header is the drop-in boundary *under* the reference's Python
 * model API: every entry point replaces the library call(s) the reference makes at the cited file:line.
 * The Python classes in stllm_b200/ (same names/attributes/state-dict keys as stllm/models/*.py) bind these
 * with ctypes; INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - All pointers are BORROWED device pointers (the caller - torch - owns memory; nothing is allocated here).
 *   - fp16 = IEEE binary16 ("half"); matrices are row-major; ld* are row strides in ELEMENTS.
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it; no internal sync.
 *   - Return value: 0 on success, negative on error (never throws); stllm_last_error() describes the failure.
 *   - Thread-safe for distinct streams.  There is NO CPU fallback: without a CUDA device every call fails.
 */
#ifndef STLLM_B200_H
#define STLLM_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define STLLM_OK 0
#define STLLM_ERR_BAD_ARG (-1)
#define STLLM_ERR_CUDA (-2)
#define STLLM_ERR_UNSUPPORTED (-3)
#define STLLM_ERR_DRIVER (-4)

/* GEMM epilogues */
#define STLLM_EPI_NONE 0   /* C = A W^T (+bias) (+residual) */
#define STLLM_EPI_GELU 1   /* C = gelu_erf(A W^T + bias) (+residual)                 eva_vit.py:55-57, Qformer.py:359 */
#define STLLM_EPI_SWIGLU 2 /* W rows interleaved (gate_j, up_j): C[:,j] = silu(g_j) * u_j   modeling_llama_mem.py:143 */
#define STLLM_EPI_RELU 3   /* C = relu(A W^T + bias)                                 st_llm.py:242,475 */

int stllm_abi_version(void);
const char* stllm_last_error(void);
int stllm_device_sm_count(void);

/* C[M,N] = epi(A[M,K] * W[N,K]^T + bias[N]) (+ residual[M,N]); fp16 operands, fp32 accumulate (tcgen05/TMEM).
 * Replaces F.linear / nn.Linear / nn.Conv2d-as-GEMM at eva_vit.py:124,146,55-59,196; Qformer.py:186-208,285,
 * 357,371; st_llm.py:368 (llama_proj), :38 (mvm head), :108,122 (lm_head); modeling_llama_mem.py:160-163,
 * 139-143 (q/k/v/o, gate/up/down).  K, lda, ldw multiples of 8; A, W 16-byte aligned.
 * For STLLM_EPI_SWIGLU the output has N/2 columns.  out_f32: 0 -> fp16 C, 1 -> fp32 C.
 * tile_n: 0 = auto (host cost model); 64 / 128 / 256 = one-CTA tiles of that many columns; 512 = CTA-pair
 * (cta_group::2) 256x256 tiles; 1024 + bn = CTA-pair tiles of bn columns (bn % 16 == 0, bn <= 256); 2048 + bn =
 * CTA-pair with SWAPPED operands (C^T = W A^T: 256 features x bn tokens per tile, transposed epilogue); +1 on any
 * code forbids the stream-K treatment of the last wave.
 * PDL note: the first weight tiles are requested before griddepcontrol.wait -- W must not be produced by the kernel
 * launched immediately before on the same stream with programmatic serialization (weights are static on this path);
 * STLLM_GEMM_WPREFETCH=0 disables this. */
int stllm_gemm_f16(const void* A, long long lda, const void* W, long long ldw, void* C, long long ldc, int M, int N,
                   int K, const void* bias, const void* residual, long long ldr, int epilogue, int out_f32,
                   int tile_n, void* stream);

/* stllm_gemm_f16 with the LayerNorm / RMSNorm in front of the Linear FOLDED into the GEMM, so that the norms of
 * eva_vit.py:157,163 (Block.norm1/norm2) and modeling_llama_mem.py:61-78 (input / post-attention RMSNorm) need no
 * kernel of their own.  A holds the UN-normalised rows x; W holds W' = gamma (.) W (packed once by the host);
 *   norm_mode 1 (RMSNorm):   C = epi(rstd_m * (x W'^T) + bias)            rstd_m = rsqrt(sum x^2 / K + eps)
 *   norm_mode 2 (LayerNorm): C = epi(rstd_m * (x W'^T - mu_m s_n) + b'_n) col_sb[n] = (s_n = sum_k W'_nk,
 *                            b'_n = bias_n + sum_k beta_k W_nk); `bias` must be NULL
 * The statistics come from the GEMM that PRODUCED x: rs_out_mode 1 / 2 makes its epilogue store sum(c^2) / (sum c,
 * sum c^2) of the values it wrote (after bias, activation and residual), per row and per (column tile, warp half) slot
 * -- plain stores to distinct slots, summed in slot order by the consumer: deterministic, no atomics, no zeroing.  A
 * statistics row is [slots][mode] floats with 2 slots per column tile of the producer; a row stride (rs_*_stride, in
 * floats, multiple of 8) of >= 2 * ceil(columns / 128) * mode covers every tile width the planner picks for a producer
 * (>= 128 columns); floats of unused slots must be zero.  Both features need tokens on the accumulator lanes (no
 * swapped operands): the planner restricts itself accordingly. */
int stllm_gemm_norm_f16(const void* A, long long lda, const void* W, long long ldw, void* C, long long ldc, int M,
                        int N, int K, const void* bias, const void* residual, long long ldr, int epilogue,
                        int out_f32, int tile_n, int norm_mode, const float* rs_in, int rs_in_stride,
                        const float* col_sb, float norm_eps, float* rs_out, int rs_out_mode, int rs_out_stride,
                        void* stream);

/* Stream-K scheduling of the ragged last wave of stllm_gemm_f16 (fp32 partial tiles through the workspace): -1 = follow
 * the STLLM_GEMM_SK environment variable (default: off -- it loses under the power cap), 0 = off,
 * 1 = selective rule, 2 = every shape (used by the parity tests to exercise the fix-up path). */
int stllm_set_gemm_streamk(int mode);

/* The configuration the cost model picks for a shape (host only, no launch): out3 = {engine: 0 one-CTA, 1 CTA pair,
 * 2 CTA pair with swapped operands; tile columns; estimated SM cycles}. */
int stllm_gemm_plan(int M, int N, int K, int epilogue, long long* out3);

/* Stream-K scheduling of the GEMM needs a caller-owned, ZEROED device workspace of at least
 * stllm_gemm_workspace_bytes() bytes (256-byte aligned), registered once per process/device.  Passing NULL
 * unregisters it (data-parallel tiles only).  STLLM_GEMM_SK=1|2 enables the stream-K schedule (off by default). */
long long stllm_gemm_workspace_bytes(void);
int stllm_set_gemm_workspace(void* ptr, long long bytes);

/* Debug aid: copies and resets 8 cycle counters of the GEMM kernels ([0] producer waits, [1] MMA waits for TMA data,
 * [2] MMA waits for a free accumulator, [3] epilogue waits, [4] CTA-0 kernel cycles, [5] launches).  Synchronises. */
int stllm_gemm_debug_counters(unsigned long long* out8);

/* Debug aid: per-CTA phase timestamps (%globaltimer ns) of the last 4 CTA-pair GEMM launches: out[4][160][12] =
 * {entry, set-up done, producer past griddepcontrol.wait, first stage landed, last MMA committed, last epilogue done,
 * exit, tiles, MMA-phase SM cycles, cycles waiting for operands, cycles waiting for an accumulator, -}.  on != 0 enables tracing for the following launches; next_slot = slot the next launch writes. */
int stllm_gemm_debug_trace(int on, unsigned long long* out, int* next_slot);

/* out[b,s,h,:] = softmax(scale * q k^T + mask) v   (flash-style, fp32 softmax).  Strides: {batch, seq, head} in
 * elements, multiples of 8.  q_prescale != 0: q <- fp16(q * q_prescale) first (eva_vit.py:128).
 * key_mask: optional uint8 [B,Skv], 0 = masked -> additive mask_value (Qformer.py:801 -10000; HF Llama fp16 min).
 * causal: key j visible to query i iff j <= i + (Skv - Sq).  D in {64, 88, 128}.
 * Replaces eva_vit.py:128-145, Qformer.py:210-268, modeling_llama_mem.py:205-244. */
int stllm_attention_f16(const void* q, const void* k, const void* v, void* out, int B, int H, int Sq, int Skv, int D,
                        const long long* q_strides, const long long* k_strides, const long long* v_strides,
                        const long long* o_strides, float scale, float q_prescale, int causal,
                        const void* key_mask, float mask_value, void* stream);

/* A/B switch (default 1): route ViT / Llama prefill attention (D in {88,128}, Sq >= 32, heads contiguous) to the
 * tcgen05 kernel; 0 = always the mma.sync kernel. */
int stllm_set_attention_tc(int on);

/* Variants of the ViT-shape tensor-core attention (A/B and parity; profiles/r02_attention.md): wide = 128-key tiles with
 * the leftover key / trailing query rows on CUDA cores in one kernel (measured slower, default off);
 * tail: what happens to the query rows after the last full 128-row tile (the ViT's 257th row): 0 = a third tensor-core
 * tile, 1 = CUDA-core tail CTAs inside the same grid, 2 = a second small launch (default: measured fastest).  -1 = follow
 * the STLLM_ATTN_WIDE (default 0) / STLLM_ATTN_TAIL (default 2) environment variables. */
int stllm_set_attention_variant(int wide, int tail);
/* kernels launched by the most recent stllm_attention_f16 call (1, or 2 when the trailing query rows of the ViT shape
   went to their own small launch); bookkeeping for launch counts, not a synchronisation point */
int stllm_attention_last_launches(void);

/* y = LayerNorm(x) with fp32 statistics; gamma/beta fp32.  eva_vit.py:157,163; blip2.py:103-109; Qformer.py:65,
 * 282,366; st_llm.py:39. */
int stllm_layernorm_f16(const void* x, long long ldx, void* y, long long ldy, const float* gamma, const float* beta,
                        int rows, int cols, float eps, void* stream);

/* y = w * fp16(x * rsqrt(mean(x^2) + eps)).  modeling_llama_mem.py:61-78. */
int stllm_rmsnorm_f16(const void* x, long long ldx, void* y, long long ldy, const void* weight, int rows, int cols,
                      float eps, void* stream);

/* frames [n,C,H,W] (fp16, or fp32 when img_is_f32) -> patch matrix [n*(H/P)*(W/P), kpad] fp16, column =
 * c*P*P + py*P + px (the flattening of Conv2d weight [D,C,P,P]); columns >= C*P*P are zero.  eva_vit.py:196,203. */
int stllm_im2col_patch(const void* img, int img_is_f32, void* out, int n_img, int C, int H, int W, int P, int kpad,
                       void* stream);

/* x[n,0,:] = cls + pos[0]; x[n,1+p,:] = patches[n*NP+p,:] + pos[1+p].  eva_vit.py:328-332. */
int stllm_vit_assemble_f16(const void* patches, long long ldp, const void* cls, const void* pos, void* x, int n_img,
                           int n_patches, int dim, void* stream);

/* in-place rotate-half RoPE on nheads consecutive heads per row; cos/sin fp32 [max_pos, head_dim/2].
 * modeling_llama_mem.py:81-127. */
int stllm_rope_f16(void* x, long long ldx, const int* pos, const float* cos_t, const float* sin_t, int rows,
                   int nheads, int head_dim, int max_pos, void* stream);

/* dst[i,:] = src[idx[i],:] (+ addend[i % add_period,:] when addend != NULL), int64 idx.
 * st_llm.py:491 (img_embeds[~mask]), nn.Embedding lookups, Qformer.py:95-99 (word + position embeddings).
 * err_flag (optional, device int): set to 1 on an out-of-range index. */
int stllm_gather_rows_f16(const void* src, long long lds, long long src_rows, const long long* idx, void* dst,
                          long long ldd, int n_rows, int cols, const void* addend, int add_period, int* err_flag,
                          void* stream);

/* dst[i,:] = idx[i] >= 0 ? src0[idx[i],:] : src1[-1 - idx[i],:]  (two-source row gather, int64 idx).
 * Replaces the per-sample torch.cat / slice-assign chains of st_llm.py:379-432,519-530 (prompt_wrap,
 * concat_emb_input_output, BOS prepend): src0 = token-embedding table, src1 = flattened video tokens, idx = the row plan
 * built on the host.  err_flag (optional, device int): set to 1 on an out-of-range index. */
int stllm_gather_rows2_f16(const void* src0, long long lds0, long long rows0, const void* src1, long long lds1,
                           long long rows1, const long long* idx, void* dst, long long ldd, int n_rows, int cols,
                           int* err_flag, void* stream);

/* y[b,:] = mean_t x[b,t,:] (fp32 accumulate).  st_llm.py:466,470. */
int stllm_mean_frames_f16(const void* x, void* y, int B, int T, long long inner, void* stream);

/* out[b,r,:] = x[b,idx[r],:] + g[b,:].  st_llm.py:472-476 (global-local merge). */
int stllm_residual_merge_f16(const void* x, const void* g, const long long* idx, void* out, int B, int T, int R,
                             long long inner, void* stream);

/* Frame preprocessing on the GPU, bit-exact with the reference's CPU pipeline GroupScale(224, BICUBIC) ->
 * GroupCenterCrop(224) -> Stack -> ToTorchFormatTensor -> GroupNormalize (conversation.py:192-200,274-278;
 * video_transforms.py:66-71,94-123,367-406), i.e. Pillow's 8-bit two-pass bicubic resampler (22-bit fixed-point
 * weights, rounding and clamping to uint8 after each pass) followed by x/255 and (x - mean)/std in fp32.
 * frames [n,H,W,3] uint8 -> out [n,3,S,S] fp16.  xbounds/xcoef [new_w][2]/[new_w][kx], ybounds/ycoef
 * [new_h][2]/[new_h][ky]: int32 DEVICE tables of the resampling windows (first input index, tap count / weights),
 * computed by the host in double precision as Pillow's precompute_coeffs does.  crop_left/crop_top: offset of the
 * centre crop in the resized image; [y_first, y_last): input rows the cropped rows read.  tmp: scratch of
 * n*(y_last-y_first)*S*3 bytes.  mean3/std3: HOST pointers. */
int stllm_preprocess_frames_u8(const void* frames, int n_img, int H, int W, int S, const int* xbounds,
                               const int* xcoef, int kx, int crop_left, const int* ybounds, const int* ycoef, int ky,
                               int crop_top, int y_first, int y_last, const float* mean3, const float* std3, void* tmp,
                               long long tmp_bytes, void* out, void* stream);

/* BT-Adapter encoder (eva_btadapter.py; vit_model "eva_btadapter_g").  ViT stream x: rows (b*T + t)*(P+1) + l, l = 0
 * the cls token.  Branch z: rows [0,B) one cls token per clip, then B + (b*P + l)*T + t ('b (p t) m').
 *   stllm_bt_vit_to_branch_f16: z_out = transpose(x) + z_prev, the cls row = fp16(mean_t x_cls) + z_prev_cls
 *     (forward_branch :184-193; also the residual of BTAdapter_Spatial :273-276 with x = the spatial attention output).
 *     z_prev == NULL: init_input (:203-226): patch rows + pos[1+l] + time_emb[t], cls = (mean_t x_cls + bt_cls + pos[0])/2.
 *     z_out may alias z_prev.
 *   stllm_bt_branch_to_vit_avg_f16: out = (x + transpose^-1(z)) / 2 (forward_features :178-181); out may alias x. */
int stllm_bt_vit_to_branch_f16(const void* x, const void* z_prev, void* z_out, const void* pos, const void* time_emb,
                               const void* bt_cls, int B, int T, int P, int dim, void* stream);
int stllm_bt_branch_to_vit_avg_f16(const void* x, const void* z, void* out, int B, int T, int P, int dim,
                                   void* stream);

/* out_sum[0] = mean_rows(2 - 2 cos(a_row, b_row)).  st_llm.py:89-91. */
int stllm_cosine_loss_f16(const void* a, long long lda, const void* b, long long ldb, int rows, int cols,
                          float* out_sum, void* stream);

/* out[0] = sum of CE over rows with label != ignore_index, out[1] = number of such rows.  st_llm.py:124-138. */
int stllm_cross_entropy_f16(const void* logits, long long ldl, const long long* labels, int rows, int V,
                            long long ignore_index, float* out_sum_count, void* stream);

/* ---- Single-token decode step (HF generate() as called at conversation.py:231-243,
 * mvbench_conversation.py:114-125).  R <= 8 rows (beams) per step: the step is bound by streaming the weights from
 * HBM once per token, so these are byte-moving kernels (16-byte streaming loads, fp32 accumulate), not 128-row MMA
 * tiles. ---- */

/* y[R,N] = epi(norm(x)[R,K] * W[N,K]^T) (+ residual[R,N_out]); R <= 8, K % 32 == 0 (else STLLM_ERR_UNSUPPORTED).
 * norm_mode 0: none.  1: RMSNorm folded into W (W = gamma (.) W as packed for stllm_gemm_norm_f16): y = rstd_r * (x W^T),
 * rstd_r = rsqrt(mean(x_r^2) + eps) computed in the prologue.  2: explicit RMSNorm, x' = norm_weight (.) fp16(x rstd)
 * (modeling_llama_mem.py:61-78), then x' W^T.  epilogue: STLLM_EPI_NONE or STLLM_EPI_SWIGLU (N/2 outputs).
 * y may alias residual (each element is read before it is written); y must not alias x.
 * Replaces the q/k/v/o, gate/up/down and lm_head nn.Linear calls of a cached decode step
 * (modeling_llama_mem.py:139-143,160-163; st_llm.py:108,122). */
int stllm_decode_linear_f16(const void* x, long long ldx, const void* W, long long ldw, void* y, long long ldy, int R,
                            int N, int K, int norm_mode, const void* norm_weight, float norm_eps,
                            const void* residual, long long ldr, int epilogue, void* stream);

/* One new token per beam.  qkv [R, 3*H*D] (row stride ldq): rotate-half RoPE at position prefix_len + *n_gen on q (in
 * place) and on k; the roped k and v are stored in the suffix caches k_suffix / v_suffix [rows][l_max][H*D] (strides in
 * elements) at (row r, token *n_gen), and src[r*src_ld + *n_gen] = r records where beam r's new token lives.
 * n_gen is a DEVICE scalar (tokens generated before this step) so that a captured CUDA graph can be replayed.
 * modeling_llama_mem.py:81-127 (RoPE), 205-217 (cache append). */
int stllm_decode_rope_append_f16(void* qkv, long long ldq, const float* cos_t, const float* sin_t, int max_pos,
                                 int prefix_len, const int* n_gen, void* k_suffix, void* v_suffix,
                                 long long s_beam_stride, long long s_tok_stride, int l_max, int* src, int src_ld, int R,
                                 int H, int D, void* stream);

/* out[r, h*D + :] = softmax(scale q_r K^T) V over the keys [prompt (prefix_len, ONE copy shared by all beams; optional
 * uint8 key mask, 0 = padded) || generated tokens 0 .. *n_gen of beam r], the generated ones read through the ancestry
 * table: token j of beam r is row src[r*src_ld + j] of the suffix caches (a beam re-ordering permutes rows of `src`;
 * HF _reorder_cache copies the whole cache).  D = 128.  The keys are split n_split ways (grid = H x n_split) and merged
 * by a second kernel; workspace: stllm_decode_attention_workspace_bytes(H, n_split) bytes, 16-byte aligned.
 * modeling_llama_mem.py:205-244 with q_len = 1. */
long long stllm_decode_attention_workspace_bytes(int H, int n_split);
int stllm_decode_attention_f16(const void* q, long long ldq, const void* k_prefix, const void* v_prefix,
                               long long p_stride, const void* prefix_mask, int prefix_len, const void* k_suffix,
                               const void* v_suffix, long long s_beam_stride, long long s_tok_stride, int l_max,
                               const int* src, int src_ld, const int* n_gen, void* out, long long ldo, void* workspace,
                               long long workspace_bytes, int R, int H, int D, int n_split, float scale, void* stream);

/* stllm_decode_rope_append_f16 + stllm_decode_attention_f16 in ONE launch per layer: every CTA (head, key split) ropes
 * the query rows of its head itself; the CTA of the last split ropes k of the new token, stores k / v into the suffix
 * caches (and src[r, *n_gen] = r) and scores that key from shared memory; the last CTA of a head to finish merges the
 * splits (atomic ticket, reset for the next launch).  qkv is NOT modified.  The workspace holds the split partials and
 * the tickets: stllm_decode_attention_fused_workspace_bytes(H, n_split) bytes, ZEROED once by the caller.
 * prefetch_ptr / prefetch_bytes, prefetch_ptr2 / prefetch_bytes2 (optional, 16-byte aligned): static memory the NEXT
 * kernels will stream (the o_proj weights, the head of the gate|up weights); pulled into L2 with
 * cp.async.bulk.prefetch.L2 while this kernel's short latency chain leaves HBM idle. */
long long stllm_decode_attention_fused_workspace_bytes(int H, int n_split);
int stllm_decode_attention_fused_f16(void* qkv, long long ldq, const float* cos_t, const float* sin_t, int max_pos,
                                     const void* k_prefix, const void* v_prefix, long long p_stride,
                                     const void* prefix_mask, int prefix_len, void* k_suffix, void* v_suffix,
                                     long long s_beam_stride, long long s_tok_stride, int l_max, int* src, int src_ld,
                                     const int* n_gen, void* out, long long ldo, void* workspace,
                                     long long workspace_bytes, int R, int H, int D, int n_split, float scale,
                                     const void* prefetch_ptr, long long prefetch_bytes, const void* prefetch_ptr2,
                                     long long prefetch_bytes2, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* STLLM_B200_H */
