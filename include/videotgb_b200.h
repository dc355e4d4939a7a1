/*
 * videotgb_b200 -- C ABI of the B200-native VideoTGB hot path (sm_100a only).
 *
 * The reference (bigai-nlco/VideoTGB) is pure Python: it has no FFI / plugin interface, its "operator API"
 * for this path is the set of torch calls its modules make.  Each entry point below therefore cites the
 * reference statement(s) it replaces (paths relative to the reference root).  The Python host layer in
 * videotgb_b200/ mirrors the reference *module* API on top of these.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the caller owns all memory
 *   - no hidden allocation, no hidden synchronisation: work is enqueued on `stream`
 *   - returns TGB_OK (0) or a negative error; tgb_last_error() has the text (thread-local)
 *   - bf16 tensors are passed as `const void*` / `void*` (raw __nv_bfloat16 storage)
 *   - there is no CPU fallback: on a non-sm_100 device every compute call returns TGB_EARCH
 */
#ifndef VIDEOTGB_B200_H_
#define VIDEOTGB_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* tgb_stream_t; /* == cudaStream_t */

#if defined(__GNUC__)
#define TGB_API __attribute__((visibility("default")))
#else
#define TGB_API
#endif

enum {
  TGB_OK = 0,
  TGB_EINVAL = -1, /* bad shape / alignment / null pointer */
  TGB_EARCH = -2,  /* device is not sm_100 (B200) */
  TGB_ECUDA = -3   /* CUDA runtime / driver error, see tgb_last_error() */
};

TGB_API const char* tgb_last_error(void);
TGB_API int tgb_version(void);
/* TGB_OK iff the current CUDA device is compute capability 10.x. */
TGB_API int tgb_device_check(void);
TGB_API int tgb_sm_count(void);

/* ---------------------------------------------------------------- A0  frame normalisation -------------
 * src/gadgets/functional_video.py:76-90 (to_tensor: float()/255.0, permute) and :93-109 (normalize: sub_(mean)
 * .div_(std)), composed at src/data/videoinstruct_datamodule.py:125-132, then the caller's permute(1,0,2,3).
 * u8 [T,H,W,3] -> f32 [T,3,H,W], bit-exact with the reference op order ((u/255) - mean) / std in fp32.     */
TGB_API int tgb_normalize_frames_u8(const uint8_t* frames_thwc, float* out_tchw, int T, int H, int W,
                            const float* mean3_host, const float* std3_host, tgb_stream_t stream);

/* ---------------------------------------------------------------- N3  resize of decoded frames -------
 * ResizeVideo -> ToUint8 -> ToTHWC (src/gadgets/functional_video.py:33-40, src/gadgets/transforms.py:33-57,213-238;
 * composed at eval/utils/builder_utils.py:117-124): bilinear, align_corners=False, no antialias, to S x S regardless of
 * the aspect ratio, then truncation to uint8.  u8 [T,H0,W0,3] (decoded RGB) -> u8 [T,S,S,3], bit-exact with
 * F.interpolate's CPU arithmetic followed by .to(torch.uint8).  S % 4 == 0.                                 */
TGB_API int tgb_resize_bilinear_u8(const uint8_t* frames_thwc, uint8_t* out_thwc, int T, int H0, int W0, int S,
                           tgb_stream_t stream);

/* ---------------------------------------------------------------- A7  Gumbel span draw ----------------
 * src/models/LSTP_SF_blip2_module.py:424-428: F.gumbel_softmax(cat(start,end), tau, dim=1).argmax(dim=1)
 * with the Gumbel noise made an explicit input.  logits, gumbel: f32 [R, L] (R = 2B rows).  index: i64 [R].
 * Computes softmax((logits+gumbel)/tau) in fp32 and returns the FIRST index of its maximum.               */
TGB_API int tgb_gumbel_argmax(const float* logits, const float* gumbel, float tau, int R, int L, int64_t* index,
                      tgb_stream_t stream);
/* One round of the draw (LSTP_SF_blip2_module.py:424-430, the body of `for i in range(top_k)`) in ONE launch, without the
 * cat / log / neg / split statements around it: start_logits, end_logits f32 [B, L]; noise f32 [2B, L] (rows [0,B) pair
 * with the start logits, rows [B,2B) with the end logits, as in cat(dim=0)).  noise_is_exponential = 1: `noise` holds the
 * Exp(1) draws of torch.Tensor.exponential_ and the Gumbel noise is formed as -log(e) (torch/nn/functional.py
 * gumbel_softmax); 0: `noise` is the Gumbel noise itself.  cand_start, cand_end: i64 [B] (row i of the [top_k, B] lists). */
TGB_API int tgb_span_draw(const float* start_logits, const float* end_logits, const float* noise, int noise_is_exponential,
                  float tau, int B, int L, int64_t* cand_start, int64_t* cand_end, tgb_stream_t stream);

/* ---------------------------------------------------------------- A8  spans -> frame indices ----------
 * src/models/LSTP_SF_blip2_module.py:433-459 (eval), :284-312 (train, inclusive end),
 * src/models/LSTP_module.py:233-261 and eval/utils/model.py:119-147 (ratio mapping).
 * cand_start / cand_end: i64 [top_k, B] argmax draws.  of_lengths: i32 [B] true flow lengths (>= 2: the reference divides
 * by length - 1 and raises below that; a device-side value below 2 is treated as 2).
 * mapping: 0 = int(s*(N-1)/(L-1)), 1 = int(s/L*N).  inclusive: range(start,end+1) instead of range(start,end).
 * host_ints: the draws are Python ints in the reference (float64 arithmetic) instead of 0-dim tensors (float32).
 * indices: i64 [B, nframe], sorted.  Returns TGB_EINVAL when num_frames > 4096 (bitmap limit).             */
enum { TGB_MAP_SCALED = 0, TGB_MAP_RATIO = 1 };
TGB_API int tgb_spans_to_indices(const int64_t* cand_start, const int64_t* cand_end, int top_k, const int32_t* of_lengths,
                         int B, int num_frames, int nframe, int mapping, int inclusive, int host_ints,
                         int64_t* indices, tgb_stream_t stream);

/* ---------------------------------------------------------------- N2  self-refinement pseudo labels --
 * src/models/LSTP_SF_blip2_module.py:195-222: per-frame answer scores f32 [B, num_frames] -> largest-rectangle
 * monotone stack over the min-shifted zero-padded scores -> (start, end) frame span -> rescale to flow positions
 * int(t / (num_frames-1) * (of_lengths[b]-1)) in float64 (the reference computes on Python floats).
 * start_targets / end_targets: i64 [B].  Bit-exact.  2 <= num_frames <= 3000.                              */
TGB_API int tgb_pseudo_span(const float* scores, const int32_t* of_lengths, int B, int num_frames,
                            int64_t* start_targets, int64_t* end_targets, tgb_stream_t stream);
/* src/models/LSTP_SF_blip2_module.py:238-253 ("mrc_loss"): logits f32 [B, L, 2] as the mrc head emits them;
 * CrossEntropyLoss(ignore_index=L) on start and end rows with targets clamped to [0, L]; loss f32 [1] =
 * (CE_start + CE_end) / 2.  fp32.                                                                          */
TGB_API int tgb_span_ce_loss(const float* logits, const int64_t* start_targets, const int64_t* end_targets, int B, int L,
                             float* loss, tgb_stream_t stream);

/* ---------------------------------------------------------------- A9  frame gather --------------------
 * src/models/LSTP_SF_blip2_module.py:417-419,460-464: sampled[j] = index_select(pixel_values[j], 0, idx_j).
 * src: [B, N, frame_bytes], indices: i64 [B, k] -> dst: [B*k, frame_bytes].  Pure byte copy (bit-exact);
 * frame_bytes must be a multiple of 16.                                                                 */
TGB_API int tgb_gather_frames(const void* src, const int64_t* indices, void* dst, int B, int N, int k, size_t frame_bytes,
                      tgb_stream_t stream);

/* ---------------------------------------------------------------- A10 patchify ------------------------
 * The im2col half of Blip2VisionEmbeddings' Conv2d(3,1408,k14,s14) (src/models/components/xblip2.py:108-112).
 * pixels f32 [F,3,S,S] -> bf16 [F*(S/P)^2, ld] with column (c*P + ky)*P + kx, zero padded up to ld.
 * The u8 variant fuses A0 (normalise) in front: frames u8 [F,S,S,3].                                     */
TGB_API int tgb_patchify_f32(const float* pixels_tchw, void* out_bf16, int F, int S, int P, int ld, tgb_stream_t stream);
TGB_API int tgb_patchify_u8(const uint8_t* frames_thwc, void* out_bf16, int F, int S, int P, int ld,
                    const float* mean3_host, const float* std3_host, tgb_stream_t stream);
/* hidden[f, 0, :] = class_embedding + position_embedding[0]   (xblip2.py:114-116)   hidden: f32 [F, tokens, D] */
TGB_API int tgb_fill_cls_rows(float* hidden, const float* cls, const float* pos0, int F, int tokens, int D,
                      tgb_stream_t stream);

/* ---------------------------------------------------------------- GEMM (tcgen05 + TMA) ---------------
 * out[M,N] = epilogue(A[M,K] * W[N,K]^T): every nn.Linear / the patch-embed conv on the path
 * (xblip2.py:172 qkv, :196 projection, :211-213 fc1/fc2, :632-641 Q-Former q/k/v, :707 dense, :773, :787,
 *  :1552 language_projection; xropebert.py:222-230, :546, :560, :575).
 * A, W: bf16 row-major with leading dims lda, ldw (elements, multiples of 8; base pointers 16-B aligned).
 * bias: f32 [N] or NULL.
 *   TGB_EPI_BF16        out bf16 = acc + bias
 *   TGB_EPI_BF16_GELU   out bf16 = gelu_erf(acc + bias)                          (ACT2FN["gelu"])
 *   TGB_EPI_F32         out f32  = acc + bias
 *   TGB_EPI_F32_RESID   out f32  = acc + bias + resid[m,n]   (resid f32, leading dim ldr; may alias out)
 *   TGB_EPI_F32_ROWBIAS out f32 row (m/period)*(period+skip)+skip+m%period = acc + rowbias[m%period, n]
 *                       (patch-embed: bias+position_embedding folded into rowbias, CLS row skipped)
 *   TGB_EPI_BF16_RELU   out bf16 = relu(acc + bias)                               (RAFT convolutions)
 * N % 8 == 0.  bn: 0 = choose, else force the N tile (256/128/64).                                      */
enum { TGB_EPI_BF16 = 0, TGB_EPI_BF16_GELU = 1, TGB_EPI_F32 = 2, TGB_EPI_F32_RESID = 3, TGB_EPI_F32_ROWBIAS = 4,
       TGB_EPI_BF16_RELU = 5, TGB_EPI_BF16_LN = 6, TGB_EPI_BF16_GELU_LN = 7, TGB_EPI_F32_RESID_STATS = 8 };
TGB_API int tgb_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K, int epilogue,
                  const float* bias, const float* resid, int64_t ldr, const float* rowbias, int period, int skip,
                  void* out, int64_t ldo, int bn, tgb_stream_t stream);

/* LayerNorm folded into the GEMMs either side of it (the pre-LN blocks of Blip2EncoderLayer, xblip2.py:236-258:
 * layer_norm1 -> qkv, layer_norm2 -> fc1).  With W' = W * gamma (per input column), colsum[n] = sum_k W'[n,k] and
 * bias'[n] = sum_k W[n,k] beta[k] + b[n]:   LN(x) W^T + b = rstd_m * (x W'^T - mean_m * colsum) + bias'.
 *   TGB_EPI_BF16_LN / TGB_EPI_BF16_GELU_LN   consumer: A = bf16 copy of the un-normalised residual rows, W = W',
 *                       bias = bias', rowstats f32 [M,2] = (mean, rstd) per row, colsum f32 [N]
 *   TGB_EPI_F32_RESID_STATS                  producer: out f32 = acc + bias + resid (as TGB_EPI_F32_RESID), plus
 *                       out_bf16 [M, ldo2] = bf16(out) and partial_stats f32 [slots, M, 2] (slot-major) = (sum, sum of
 *                       squares) of each row over the columns of one slot; slots = tgb_gemm_ln_stats_slots(M, N) (two per
 *                       128-column block of N, whatever the tile width: a row's statistics do not depend on M); every slot
 *                       is written exactly once per launch (no atomics: results are run-to-run deterministic)
 * The consumer multiplies the UN-normalised rows on the tensor cores, so its error grows with |row mean| / row std
 * (measured: 1x the separate-LayerNorm error at 0.25, 2x at 4, 4.6e-2 at 32); tgb_ln_stats_finalize can track the largest
 * ratio it sees in max_ratio (f32 scalar, zero it first; NULL = off) so that the host can fall back to tgb_layernorm_f32.
 * tgb_ln_stats_finalize reduces the slots to rowstats (float64 combine, biased variance, rstd = 1/sqrt(var + eps));
 * tgb_rowstats_cast_f32 produces rowstats and the bf16 copy for rows that did not come out of a GEMM (first layer).  */
TGB_API int tgb_gemm_bf16_ln(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K, int epilogue,
                     const float* bias, const float* resid, int64_t ldr, const float* rowstats, const float* colsum,
                     void* out, int64_t ldo, void* out_bf16, int64_t ldo2, float* partial_stats, tgb_stream_t stream);
TGB_API int tgb_gemm_ln_stats_slots(int M, int N);
TGB_API int tgb_ln_stats_finalize(const float* partial_stats, int M, int slots, int N, float eps, float* rowstats,
                          float* max_ratio, tgb_stream_t stream);
TGB_API int tgb_rowstats_cast_f32(const float* x, int64_t ldx, int M, int N, float eps, void* out_bf16, int64_t ldo,
                          float* rowstats, tgb_stream_t stream);

/* ---------------------------------------------------------------- LayerNorm ---------------------------
 * nn.LayerNorm over the last dim (xblip2.py:222-224,511; :703,:782; xropebert.py:91).  x: f32 [M,N] (ldx).
 * Writes bf16 and/or f32 outputs (either may be NULL).  N % 4 == 0, N <= 8192.                            */
TGB_API int tgb_layernorm_f32(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps, int M, int N,
                      void* out_bf16, int64_t ld_bf16, float* out_f32, int64_t ld_f32, tgb_stream_t stream);

/* ---------------------------------------------------------------- attention ---------------------------
 * softmax(Q K^T * scale + mask) V for one (batch, head) per CTA; fp32 softmax, bf16 tensor-core operands.
 * (xblip2.py:174-189 ViT, :647-686 Q-Former).  q/k/v: bf16, row (b*S + i), head h at columns [h*D, (h+1)*D).
 * mask: f32 additive [B, Sk] or NULL.  out: bf16 [B*Sq, ldo].  D in {64, 88}; long key sequences stream K/V through shared memory.
 * The Q-Former cross-attention shape (Sq = 32, D = 64, H % 4 == 0, 64 < Sk <= 257: xblip2.py:647-686 with the 257 image
 * tokens as keys) and its self-attention shape (Sq = Sk = 32, D = 64, H % 4 == 0) run on tcgen05 (four heads per 128-row
 * MMA tile, csrc/attention_qf_tc.cu); same contract.                                                                          */
TGB_API int tgb_attention_bf16(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                       void* out, int64_t ldo, const float* mask, int B, int H, int Sq, int Sk, int D, float scale,
                       tgb_stream_t stream);
/* Self-attention over the concatenation of two separately stored row groups, without the concatenation: the
 * InstructBLIP Q-Former attends over cat(query rows, instruction rows) (xinstructblip.py:1168-1176, :836-851), whose two
 * halves go through different feed-forward blocks afterwards (:858-876).  qkv1: bf16 [B*n1, ld_qkv] fused q|k|v of the
 * first n1 sequence rows of every batch element, qkv2: [B*n2, ld_qkv] of the remaining n2; out1 [B*n1, ldo], out2
 * [B*n2, ldo]; mask f32 additive [B, n1+n2] or NULL.  Same arithmetic as tgb_attention_bf16 on the concatenated rows.  */
TGB_API int tgb_self_attention_2seg_bf16(const void* qkv1, const void* qkv2, int64_t ld_qkv, void* out1, void* out2,
                                 int64_t ldo, const float* mask, int B, int H, int n1, int n2, int D, float scale,
                                 tgb_stream_t stream);

/* ViT self-attention on tcgen05 (xblip2.py:156-192), specialised for the EVA-ViT-g shape: 257 tokens, 16 heads x 88.
 * qkv: bf16 [frames*257, ld_qkv] fused projection, q|k|v of head h at columns {0,1408,2816}+88h (the tensor is read
 * through a 3-D TMA map, ld_qkv >= 4224).  out: bf16 [frames*257, ld_out], head h at columns 88h.                    */
TGB_API int tgb_vit_attention_bf16(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, int frames, float scale,
                                   tgb_stream_t stream);

/* ---------------------------------------------------------------- A1/A2/A4 temporal sampler -----------
 * TemporalOFEmbedding (src/models/components/xropebert.py:103-129): Conv2d(2,768,k16,s16) -> flatten(2) ->
 * Linear(196,1) has no non-linearity in between, so it reassociates exactly to
 *   z[f,(c,ky,kx)] = sum_p fc_w[p] * flow[f,c,py*16+ky,px*16+kx]     (tgb_of_patch_reduce, HBM-bound, 1 pass)
 *   y = z * W_conv^T + (b_conv * sum(fc_w) + fc_b)                      (tgb_gemm_bf16, [T,512]x[512,768])
 * flow f32 [frames, C, S, S]; fc_weight f32 [(S/16)^2]; z bf16 [frames, ldz], ldz >= C*256.               */
TGB_API int tgb_of_patch_reduce(const float* flow, const float* fc_weight, void* z_bf16, int frames, int C, int S, int P,
                                int ldz, tgb_stream_t stream);
/* A3 masks.  out[b,i] = (1.0 - mask[b,i]) * value in fp32: get_extended_attention_mask (xropebert.py:986-1012, value
 * -10000.0) and invert_attention_mask (value = finfo.min; the Q-Former twins xblip2.py:1033-1060).  last_index[b] =
 * sum(mask[b,:]) - 1, the `of_mask.sum(dim=1) - 1` of TemporalOFEmbedding (xropebert.py:121).  mask i64 [B,S] of 0/1;
 * out f32 [B,S] or NULL; last_index i32 [B] or NULL.                                                              */
TGB_API int tgb_additive_mask(const int64_t* mask, int B, int S, float value, float* out, int32_t* last_index,
                              tgb_stream_t stream);
/* xropebert.py:118-127: cat(bos, y, zero row); row ends[b] := eos; += frame_pos_embed[t]; LayerNorm(eps).
 * y f32 [B, L, N]; ends i32 [B] (= of_mask.sum(1)-1); out [B, L+2, N] as f32 and/or bf16.                   */
TGB_API int tgb_sampler_assemble(const float* y, const float* bos, const float* eos, const float* pos, const int32_t* ends,
                                 const float* gamma, const float* beta, float eps, int B, int L, int N, float* out_f32,
                                 void* out_bf16, tgb_stream_t stream);
/* LayerNorm(table[ids[r]] + pos[r % period] + addvec): RopeBertEmbeddings (xropebert.py:190-208: word + token_type[0],
 * no position) and InstructBlipQFormerEmbeddings (xinstructblip.py:1018-1046: word + position).  pos/addvec may be NULL.
 * An id outside [0, vocab) -- where nn.Embedding raises / trips a device-side assert -- yields a NaN output row.           */
TGB_API int tgb_embed_layernorm(const int64_t* ids, const float* table, int64_t vocab, const float* pos, int period,
                                const float* addvec, const float* gamma, const float* beta, float eps, int64_t rows, int N,
                                float* out_f32, int64_t ld_f32, void* out_bf16, int64_t ld_bf16, tgb_stream_t stream);
/* apply_rotary_position_embeddings / apply_rope (xropebert.py:334-377), in place on bf16 [rows, ldx] head vectors
 * (H heads of HD at columns [h*HD,(h+1)*HD)); position = row % S; table f32 [>=S, HD] = [sin(HD/2) | cos(HD/2)].     */
TGB_API int tgb_rope_bf16(void* x, int64_t ldx, const float* table, int64_t rows, int S, int H, int HD, tgb_stream_t stream);

/* mrc_head = Linear(768, 2) in fp32 (xropebert.py:953,1164): out f32 [M,N] = x f32 [M,K] * w[N,K]^T + bias, N <= 8.   */
TGB_API int tgb_linear_small_f32(const float* x, int64_t ldx, const float* w, const float* bias, int64_t M, int N, int K,
                                 float* out, int64_t ldo, tgb_stream_t stream);

/* ---------------------------------------------------------------- small elementwise ops ---------------*/
/* out bf16 [M,N] (ldo) = x f32 [M,N] (ldx) */
TGB_API int tgb_cast_f32_bf16(const float* x, int64_t ldx, void* out, int64_t ldo, int M, int N, tgb_stream_t stream);
/* out f32 [G, R, N] -> mean over groups of `width` consecutive leading rows: in [G*width, R*N] -> [G, R*N]
 * (eval/utils/model.py:186-190 mean pooling of a video's frames) */
TGB_API int tgb_group_mean_f32(const float* x, float* out, int G, int width, int64_t inner, tgb_stream_t stream);

/* ---------------------------------------------------------------- N4  LLM hand-off --------------------
 * src/models/LSTP_SF_blip2_module.py:482-495: inputs_embeds = cat([language_model_inputs, embed(question)], 1),
 * attention_mask = cat([ones, question_attention_mask], 1).  visual f32 [B,Nv,H] (the projected query tokens),
 * table f32 [vocab,H] (the unchanged LLM's input embedding), ids / qmask i64 [B,Tq] ->
 * embeds f32 [B,Nv+Tq,H], mask i64 [B,Nv+Tq].  Byte copies (bit-exact).  H % 4 == 0.                        */
TGB_API int tgb_llm_inputs_f32(const float* visual, const float* table, const int64_t* ids, const int64_t* qmask,
                               int64_t vocab, int B, int Nv, int Tq, int H, float* embeds, int64_t* mask,
                               tgb_stream_t stream);
/* :503-504  outputs[outputs == from] = to, in place on i64 [n] (generated id 0 -> 2 for LLaMA tokenizers). */
TGB_API int tgb_replace_ids(int64_t* ids, int64_t n, int64_t from, int64_t to, tgb_stream_t stream);

/* ---------------------------------------------------------------- N1  RAFT optical-flow extractor -----
 * src/models/components/xraft.py:102-158 and raft_utils/{extractor,corr,update}.py.  Activations are NHWC pixel rows
 * [N*H*W, C]; every nn.Conv2d is tgb_im2col_nhwc_bf16 + tgb_gemm_bf16 (weights reordered to [Cout, (ky,kx,c)]).          */
/* out bf16 [N*Ho*Wo, ldo]: column (ky*kw + kx)*C + c = x[(n, oy*sh-ph+ky, ox*sw-pw+kx), c] (0 outside), rest zero.
 * x: bf16 pixel rows of pitch ldx >= C.  ldo % 8 == 0.                                                        */
TGB_API int tgb_im2col_nhwc_bf16(const void* x, int64_t ldx, int N, int H, int W, int C, int kh, int kw, int sh, int sw,
                                 int ph, int pw, void* out, int64_t ldo, tgb_stream_t stream);
/* nn.InstanceNorm2d(affine=False) (+ ReLU) over H*W per (n, c) (extractor.py:27-31,125-126): x f32 [N, HW, C].
 * workspace: f32 [N * 17 * C * 2] scratch for the split statistics (the library never allocates).  C % 4 == 0.      */
TGB_API int tgb_instance_norm_nhwc(const float* x, int N, int HW, int C, float eps, int relu, float* workspace,
                                   float* out_f32, void* out_bf16, tgb_stream_t stream);
/* out = act(scale * x (+ y)) on [M,N] with row pitches; act 0 identity / 1 relu / 2 sigmoid / 3 tanh.               */
TGB_API int tgb_act_f32(const float* x, int64_t ldx, const float* y, int64_t ldy, int64_t M, int N, int mode, float scale,
                        float* out_f32, int64_t ld_f32, void* out_bf16, int64_t ld_bf16, tgb_stream_t stream);
/* SepConvGRU gates (update.py:42-59); zr f32 [M, 2C] = [z_pre | r_pre].  rh = sigmoid(r_pre) * h (bf16, pitch ld);
 * h <- (1 - sigmoid(z_pre)) h + sigmoid(z_pre) tanh(q_pre) in place (f32) plus a bf16 copy (pitch ld).              */
TGB_API int tgb_gru_rh(const float* zr, const float* h, int64_t M, int C, void* rh_bf16, int64_t ld, tgb_stream_t stream);
TGB_API int tgb_gru_update(const float* zr, const float* q, float* h, int64_t M, int C, void* h_bf16, int64_t ld,
                           tgb_stream_t stream);
/* CorrBlock (corr.py:12-60): 2x2 average pooling of [R, h, w] and the 4-level 9x9 bilinear lookup around grid + flow
 * -> bf16 [N*h*w, ldo] (324 columns, zero padded), x taking the first window index as in the reference.            */
TGB_API int tgb_corr_pool(const float* in, float* out, int64_t R, int h, int w, tgb_stream_t stream);
TGB_API int tgb_corr_lookup(const float* l0, const float* l1, const float* l2, const float* l3, const float* flow, int N,
                            int h, int w, void* out_bf16, int64_t ldo, tgb_stream_t stream);
/* Convex 8x upsampling (xraft.py:88-99): mask f32 [N*h*w, ldm>=576], flow f32 [N*h*w, 2] -> f32 [N,2,8h,8w].        */
TGB_API int tgb_raft_upsample(const float* mask, int64_t ldm, const float* flow, int N, int h, int w, float* out,
                              tgb_stream_t stream);
/* 2*(img/255) - 1 (xraft.py:106-107): f32 [N,C,H*W] -> bf16 pixel rows [N*H*W, ldo].                               */
TGB_API int tgb_raft_prep_image(const float* img_nchw, int N, int C, int HW, void* out_bf16, int64_t ldo, tgb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VIDEOTGB_B200_H_ */
