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

/* ---------------------------------------------------------------- A7  Gumbel span draw ----------------
 * src/models/LSTP_SF_blip2_module.py:424-428: F.gumbel_softmax(cat(start,end), tau, dim=1).argmax(dim=1)
 * with the Gumbel noise made an explicit input.  logits, gumbel: f32 [R, L] (R = 2B rows).  index: i64 [R].
 * Computes softmax((logits+gumbel)/tau) in fp32 and returns the FIRST index of its maximum.               */
TGB_API int tgb_gumbel_argmax(const float* logits, const float* gumbel, float tau, int R, int L, int64_t* index,
                      tgb_stream_t stream);

/* ---------------------------------------------------------------- A8  spans -> frame indices ----------
 * src/models/LSTP_SF_blip2_module.py:433-459 (eval), :284-312 (train, inclusive end),
 * src/models/LSTP_module.py:233-261 and eval/utils/model.py:119-147 (ratio mapping).
 * cand_start / cand_end: i64 [top_k, B] argmax draws.  of_lengths: i32 [B] true flow lengths (>= 2).
 * mapping: 0 = int(s*(N-1)/(L-1)), 1 = int(s/L*N).  inclusive: range(start,end+1) instead of range(start,end).
 * host_ints: the draws are Python ints in the reference (float64 arithmetic) instead of 0-dim tensors (float32).
 * indices: i64 [B, nframe], sorted.  Returns TGB_EINVAL when num_frames > 4096 (bitmap limit).             */
enum { TGB_MAP_SCALED = 0, TGB_MAP_RATIO = 1 };
TGB_API int tgb_spans_to_indices(const int64_t* cand_start, const int64_t* cand_end, int top_k, const int32_t* of_lengths,
                         int B, int num_frames, int nframe, int mapping, int inclusive, int host_ints,
                         int64_t* indices, tgb_stream_t stream);

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
 * N % 8 == 0.  bn: 0 = choose, else force the N tile (256/176/128/64).                                      */
enum { TGB_EPI_BF16 = 0, TGB_EPI_BF16_GELU = 1, TGB_EPI_F32 = 2, TGB_EPI_F32_RESID = 3, TGB_EPI_F32_ROWBIAS = 4 };
TGB_API int tgb_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K, int epilogue,
                  const float* bias, const float* resid, int64_t ldr, const float* rowbias, int period, int skip,
                  void* out, int64_t ldo, int bn, tgb_stream_t stream);

/* ---------------------------------------------------------------- LayerNorm ---------------------------
 * nn.LayerNorm over the last dim (xblip2.py:222-224,511; :703,:782; xropebert.py:91).  x: f32 [M,N] (ldx).
 * Writes bf16 and/or f32 outputs (either may be NULL).  N % 4 == 0, N <= 8192.                            */
TGB_API int tgb_layernorm_f32(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps, int M, int N,
                      void* out_bf16, int64_t ld_bf16, float* out_f32, int64_t ld_f32, tgb_stream_t stream);

/* ---------------------------------------------------------------- attention ---------------------------
 * softmax(Q K^T * scale + mask) V for one (batch, head) per CTA; fp32 softmax, bf16 tensor-core operands.
 * (xblip2.py:174-189 ViT, :647-686 Q-Former).  q/k/v: bf16, row (b*S + i), head h at columns [h*D, (h+1)*D).
 * mask: f32 additive [B, Sk] or NULL.  out: bf16 [B*Sq, ldo].  D in {64, 88}.  Sk <= 1024.               */
TGB_API int tgb_attention_bf16(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                       void* out, int64_t ldo, const float* mask, int B, int H, int Sq, int Sk, int D, float scale,
                       tgb_stream_t stream);

/* ---------------------------------------------------------------- small elementwise ops ---------------*/
/* out bf16 [M,N] (ldo) = x f32 [M,N] (ldx) */
TGB_API int tgb_cast_f32_bf16(const float* x, int64_t ldx, void* out, int64_t ldo, int M, int N, tgb_stream_t stream);
/* out f32 [G, R, N] -> mean over groups of `width` consecutive leading rows: in [G*width, R*N] -> [G, R*N]
 * (eval/utils/model.py:186-190 mean pooling of a video's frames) */
TGB_API int tgb_group_mean_f32(const float* x, float* out, int G, int width, int64_t inner, tgb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VIDEOTGB_B200_H_ */
