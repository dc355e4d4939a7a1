/*
 * Test / A-B knobs of the kernels.  NOT part of the drop-in contract (include/videotgb_b200.h): they select between
 * kernel variants that compute the same result, are process-global (atomics), and exist so that the parity tests can
 * exercise every variant.  The library exports them; tests and tools/ bind them by name.
 */
#ifndef TGB_DEBUG_H_
#define TGB_DEBUG_H_

#include "../../include/videotgb_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* 0 = choose (word-load kernel when the input is 4-byte aligned; default), 1 = generic byte-load kernel. */
TGB_API int tgb_resize_set_variant(int variant);
/* 0 = choose (default), 1 = never use the single-fma pair kernel for u8 frames. */
TGB_API int tgb_patchify_set_variant(int variant);
/* 0 = choose per problem (default), 1 = single-CTA 128xBN tiles, 2 = CTA-pair 256xBN tiles. */
TGB_API int tgb_gemm_set_cta_group(int cta_group);
/* persistent tile schedule: 1 = balanced tile schedule (default); 0 = plain round-robin; 2 = whole A row blocks per
 * cluster when M fills whole rounds of them, else 1 (measured slower, kept for A/B). */
TGB_API int tgb_gemm_set_schedule(int mode);
/* epilogue warps per CTA, 0 = choose (8; the fp32-residual epilogues always run 8). */
TGB_API int tgb_gemm_set_epilogue_warps(int warps);
/* column width of the fp32-residual epilogue's pieces: 0 = choose (32 when K <= 2048 and N tiles are 256 wide, else 16),
 * 16 / 32 = force (32 needs 256-wide N tiles; other tile widths keep 16). */
TGB_API int tgb_gemm_set_resid_piece(int width);
/* 0 = ping-pong kernel (default), 1 = one-shot kernel (whole score row in tensor memory), 2 = three-block kernel,
 * 3 = CTA-pair kernel (cta_group::2), 4 = split-row ping-pong kernel (two threads per query row), 5 = ping-pong kernel
 * with the software-pipelined exp pass.  Measurements: profiles/r2_attention_experiments.txt. */
TGB_API int tgb_vit_attention_set_variant(int variant);
/* tgb_attention_bf16: 1 = the Q-Former shapes (32 queries, d = 64, H % 4 == 0; 64 < Sk <= 257 or Sk == 32) run the
 * tcgen05 kernels (default), 0 = they run the generic mma.sync kernel like every other shape (cross-check). */
TGB_API int tgb_attention_set_tc(int on);
/* split-row kernel: 1 = prefetch the next item's operands into the L2 (measured 30 % slower), 0 = off (default). */
TGB_API int tgb_vit_attention_set_prefetch(int on);

#ifdef __cplusplus
}
#endif
#endif
