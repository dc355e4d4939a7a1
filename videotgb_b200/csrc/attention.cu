// Fused attention  out = softmax(Q K^T * scale + mask) V  for the short sequences of this path
// (ViT: S=257, d=88, 16 heads; Q-Former: 32(+Tt) queries x 32..257 keys, d=64, 12 heads).
//
// One CTA per (batch, head): K and V of that head are staged ONCE in shared memory with cp.async (16-byte, padded
// pitches chosen so every ldmatrix phase is bank-conflict free), each warp owns 16-query-row tiles and runs a
// flash-style online softmax over 64-key blocks: S = Q K^T with mma.sync.m16n8k16 (bf16 in, fp32 accumulate),
// fp32 softmax statistics, P re-used from the accumulator registers as the A operand of P V.  The score matrix
// never reaches HBM (the reference materialises [F,16,257,257], xblip2.py:174-189).
//
// TODO(round 2): move QK^T / PV to tcgen05 with S in TMEM; at 2.8 % of the path's FLOPs this kernel is sized to
// be HBM/latency-balanced first.
#include "common.cuh"

namespace tgb {

struct AttnParams {
  const __nv_bfloat16* q; const __nv_bfloat16* k; const __nv_bfloat16* v;
  __nv_bfloat16* out;
  const float* mask;          // additive [B, Sk] or null
  long long ldq, ldk, ldv, ldo;
  int B, H, Sq, Sk;
  float scale_log2;           // scale * log2(e)
  // optional second row segment (self-attention over the concatenation of two separately stored row groups, e.g. the
  // Q-Former's 32 query rows and Tt instruction rows): sequence row r < n1 is row b*n1 + r of q/k/v/out, row r >= n1 is
  // row b*(S - n1) + (r - n1) of q2/k2/v2/out2 (same leading dimensions).  q2 == nullptr: one segment.
  const __nv_bfloat16* q2; const __nv_bfloat16* k2; const __nv_bfloat16* v2;
  __nv_bfloat16* out2;
  int n1;
  // key slices per 16-row query tile (single-pass kernel): with few query rows (the Q-Former's 32 queries against 257
  // image tokens) one warp per query tile leaves a 64-thread CTA that first waits for 66 KB of K/V and then walks all the
  // keys alone; nslices > 1 gives every query tile `nslices` warps, each owning every nslices-th 64-key block, and merges
  // their (max, sum, O) partials through shared memory.  blockDim = query tiles * nslices warps.
  int nslices;
};

// address of sequence row r of batch b in a (possibly two-segment) operand with S rows per batch
template <typename T>
__device__ __forceinline__ T* attn_row(T* a, T* a2, long long ld, int b, int r, int S, int n1) {
  if (a2 == nullptr) return a + ((long long)b * S + r) * ld;
  return r < n1 ? a + ((long long)b * n1 + r) * ld : a2 + ((long long)b * (S - n1) + (r - n1)) * ld;
}

// additive mask value in log2 units, kept FINITE: the reference's (1 - mask) * finfo.min masks (xblip2.py:1106-1113) absorb
// the score, so a row whose keys are ALL masked sees equal scores and softmax gives the uniform distribution over the
// keys; finfo.min * log2(e) would overflow to -inf here and turn that row into exp2(-inf - (-inf)) = NaN / zeros instead.
__device__ __forceinline__ float mask_log2(float m) { return fmaxf(m * 1.4426950408889634f, -3.4028234663852886e38f); }

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int D>
struct AttnCfg {
  static constexpr int DK = (D + 15) / 16 * 16;        // head dim padded to the MMA K step (88 -> 96)
  static constexpr int KP = DK + 8;                    // K pitch (elements): 208 B / 144 B rows -> conflict-free ldmatrix
  static constexpr int VP = (D * 2) % 128 == 0 ? D + 8 : D;   // V pitch: 176 B (d=88) / 144 B (d=64)
  static constexpr int ND = D / 8;                     // PV n-tiles
  static constexpr int KS = DK / 16;                   // QK^T k-steps
  static_assert(D % 8 == 0, "head dim must be a multiple of 8");
};

// NT = number of 8-key n-tiles in this key block (8 for a full 64-key block, 2/4/6 for the tail).
template <int D, int NT>
__device__ __forceinline__ void attn_block(const uint32_t (&qf)[AttnCfg<D>::KS][4], float (&o)[AttnCfg<D>::ND][4],
                                           float (&mrow)[2], float (&lrow)[2], uint32_t sK, uint32_t sV, int key0,
                                           int Sk, const float* smask, float scale_log2, int lane, int kbase = 0) {
  using C = AttnCfg<D>;
  float s[NT][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) { s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f; }
  // ---- S = Q K^T
  const int lrow8 = lane & 7, lmat = lane >> 3;
#pragma unroll
  for (int ks = 0; ks < C::KS; ++ks) {
#pragma unroll
    for (int nt = 0; nt < NT; nt += 2) {
      // matrices: 0 = (keys nt, d lo) 1 = (keys nt, d hi) 2 = (keys nt+1, d lo) 3 = (keys nt+1, d hi)
      const int key = key0 + (nt + (lmat >> 1)) * 8 + lrow8;
      const int dcol = ks * 16 + (lmat & 1) * 8;
      uint32_t b0, b1, b2, b3;
      ldsm_x4(sK + (uint32_t)(key * C::KP + dcol) * 2, b0, b1, b2, b3);
      mma_bf16_16816(s[nt], qf[ks], b0, b1);
      mma_bf16_16816(s[nt + 1], qf[ks], b2, b3);
    }
  }
  // ---- scale, mask, block max
  const int kq = (lane & 3) * 2;
  float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int key = key0 + nt * 8 + kq + e;
      float add = 0.f;
      if (kbase + key >= Sk) add = -INFINITY;
      else if (smask) add = smask[key];
      s[nt][e] = fmaf(s[nt][e], scale_log2, add);
      s[nt][2 + e] = fmaf(s[nt][2 + e], scale_log2, add);
      bm0 = fmaxf(bm0, s[nt][e]);
      bm1 = fmaxf(bm1, s[nt][2 + e]);
    }
  }
  bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
  bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
  bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
  bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
  const float mn0 = fmaxf(mrow[0], bm0), mn1 = fmaxf(mrow[1], bm1);
  const float base0 = (mn0 == -INFINITY) ? 0.f : mn0, base1 = (mn1 == -INFINITY) ? 0.f : mn1;
  const float al0 = ex2_approx(mrow[0] - base0), al1 = ex2_approx(mrow[1] - base1);
  mrow[0] = mn0; mrow[1] = mn1;
  float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    s[nt][0] = ex2_approx(s[nt][0] - base0); s[nt][1] = ex2_approx(s[nt][1] - base0);
    s[nt][2] = ex2_approx(s[nt][2] - base1); s[nt][3] = ex2_approx(s[nt][3] - base1);
    ps0 += s[nt][0] + s[nt][1];
    ps1 += s[nt][2] + s[nt][3];
  }
  lrow[0] = lrow[0] * al0 + ps0;
  lrow[1] = lrow[1] * al1 + ps1;
#pragma unroll
  for (int nd = 0; nd < C::ND; ++nd) { o[nd][0] *= al0; o[nd][1] *= al0; o[nd][2] *= al1; o[nd][3] *= al1; }
  // ---- O += P V   (P from the S accumulators, V through ldmatrix.trans)
#pragma unroll
  for (int kk = 0; kk < NT / 2; ++kk) {
    uint32_t pa[4];
    pa[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
    pa[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
    pa[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
    pa[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
    const int key = key0 + kk * 16 + (lmat & 1) * 8 + lrow8;
#pragma unroll
    for (int nd = 0; nd + 1 < C::ND; nd += 2) {
      uint32_t b0, b1, b2, b3;   // (keys lo, d nd) (keys hi, d nd) (keys lo, d nd+1) (keys hi, d nd+1)
      ldsm_x4_t(sV + (uint32_t)(key * C::VP + (nd + (lmat >> 1)) * 8) * 2, b0, b1, b2, b3);
      mma_bf16_16816(o[nd], pa, b0, b1);
      mma_bf16_16816(o[nd + 1], pa, b2, b3);
    }
    if (C::ND & 1) {
      uint32_t b0, b1;
      const int key2 = key0 + kk * 16 + ((lane >> 3) & 1) * 8 + lrow8;
      ldsm_x2_t(sV + (uint32_t)(key2 * C::VP + (C::ND - 1) * 8) * 2, b0, b1);
      mma_bf16_16816(o[C::ND - 1], pa, b0, b1);
    }
  }
}

// (d = 64 -- the Q-Former and sampler shapes -- is compiled for two resident CTAs per SM: one CTA's K / V staging then
//  overlaps the other's math; ncu showed a single resident CTA alternating between waiting for 66 KB and computing.
//  Q-Former cross attention, 1024 frames x 12 heads of 32 x 257: 611 -> 370 us per layer.  A persistent CTA per SM with
//  two K / V buffers and the next item's cp.async copies in flight during the math measured 393 us: the shape is bound by
//  the legacy mma.sync rate -- 27 GFLOP per layer at the ~155 TFLOP/s this path reaches is 178 us -- not by its loads.)
template <int D>
__global__ void __launch_bounds__(D == 64 ? 256 : 288, D == 64 ? 2 : 1) attention_kernel(const AttnParams p) {
  using C = AttnCfg<D>;
  extern __shared__ __align__(16) uint8_t smem[];
  const int SkPad = (p.Sk + 15) & ~15;
  __nv_bfloat16* sKp = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* sVp = sKp + (size_t)SkPad * C::KP;
  float* smask = reinterpret_cast<float*>(sVp + (size_t)SkPad * C::VP);
  const int b = blockIdx.x / p.H, h = blockIdx.x - b * p.H;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;

  // ---- stage K, V (and the mask row, pre-multiplied by log2 e) in shared memory
  constexpr int CH = D / 8;                    // 16-byte chunks per row
  constexpr int KCH = C::KP / 8;               // chunks per padded K row (zero beyond D)
  for (int i = tid; i < SkPad * KCH; i += blockDim.x) {
    const int r = i / KCH, c = i - r * KCH;
    __nv_bfloat16* dst = sKp + (size_t)r * C::KP + c * 8;
    if (r < p.Sk && c < CH) cp_async16(dst, attn_row(p.k, p.k2, p.ldk, b, r, p.Sk, p.n1) + h * D + c * 8);
    else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
  }
  constexpr int VCH = C::VP / 8;
  for (int i = tid; i < SkPad * VCH; i += blockDim.x) {
    const int r = i / VCH, c = i - r * VCH;
    __nv_bfloat16* dst = sVp + (size_t)r * C::VP + c * 8;
    if (r < p.Sk && c < CH) cp_async16(dst, attn_row(p.v, p.v2, p.ldv, b, r, p.Sk, p.n1) + h * D + c * 8);
    else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
  }
  if (p.mask)
    for (int i = tid; i < SkPad; i += blockDim.x)
      smask[i] = i < p.Sk ? mask_log2(p.mask[(long long)b * p.Sk + i]) : 0.f;
  cp_async_wait_all();
  __syncthreads();

  const uint32_t sK = smem_u32(sKp), sV = smem_u32(sVp);
  const float* sm = p.mask ? smask : nullptr;
  const int mtiles = (p.Sq + 15) >> 4;
  const int g = lane >> 2, t2 = (lane & 3) * 2;
  const int nsl = p.nslices;
  for (int mt = (nsl > 1 ? warp % mtiles : warp); mt < mtiles; mt += nwarps) {
    const int sl = nsl > 1 ? warp / mtiles : 0;
    // ---- Q fragments straight from global (row-major A layout of m16n8k16)
    uint32_t qf[C::KS][4];
    const int r0 = mt * 16 + g, r1 = r0 + 8;
    // (rows >= Sq are never dereferenced: clamp them so that the two-segment address stays inside its segment)
    const __nv_bfloat16* q0 = attn_row(p.q, p.q2, p.ldq, b, r0 < p.Sq ? r0 : 0, p.Sq, p.n1) + h * D;
    const __nv_bfloat16* q1 = attn_row(p.q, p.q2, p.ldq, b, r1 < p.Sq ? r1 : 0, p.Sq, p.n1) + h * D;
#pragma unroll
    for (int ks = 0; ks < C::KS; ++ks) {
      const int c0 = ks * 16 + t2, c1 = c0 + 8;
      qf[ks][0] = (r0 < p.Sq && c0 < D) ? __ldg(reinterpret_cast<const uint32_t*>(q0 + c0)) : 0u;
      qf[ks][1] = (r1 < p.Sq && c0 < D) ? __ldg(reinterpret_cast<const uint32_t*>(q1 + c0)) : 0u;
      qf[ks][2] = (r0 < p.Sq && c1 < D) ? __ldg(reinterpret_cast<const uint32_t*>(q0 + c1)) : 0u;
      qf[ks][3] = (r1 < p.Sq && c1 < D) ? __ldg(reinterpret_cast<const uint32_t*>(q1 + c1)) : 0u;
    }
    float o[C::ND][4];
#pragma unroll
    for (int nd = 0; nd < C::ND; ++nd) { o[nd][0] = o[nd][1] = o[nd][2] = o[nd][3] = 0.f; }
    float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};
    const int nfull = SkPad >> 6;               // whole 64-key blocks; this warp owns blocks sl, sl + nsl, ...
    int key0 = sl * 64;
    for (; key0 + 64 <= SkPad; key0 += 64 * nsl)
      attn_block<D, 8>(qf, o, mrow, lrow, sK, sV, key0, p.Sk, sm, p.scale_log2, lane);
    const int rem = SkPad - nfull * 64;         // 0, 16, 32 or 48 keys left: they belong to slice nfull % nsl
    if (nfull % nsl == sl) {
      key0 = nfull * 64;
      if (rem == 16) attn_block<D, 2>(qf, o, mrow, lrow, sK, sV, key0, p.Sk, sm, p.scale_log2, lane);
      else if (rem == 32) attn_block<D, 4>(qf, o, mrow, lrow, sK, sV, key0, p.Sk, sm, p.scale_log2, lane);
      else if (rem == 48) attn_block<D, 6>(qf, o, mrow, lrow, sK, sV, key0, p.Sk, sm, p.scale_log2, lane);
    }
    if (nsl > 1) {
      // ---- merge the slices of this query tile (all warps take this path together: blockDim = mtiles * nsl warps).
      // Partials go through the K / V staging area, which every warp has finished reading after the first barrier.
      constexpr int XS = C::ND * 4 + 4;           // floats per thread: O fragment, row maxima, row sums
      __syncthreads();
      float* xw = reinterpret_cast<float*>(smem) + ((size_t)warp * 32 + lane) * XS;
      if (sl > 0) {
#pragma unroll
        for (int nd = 0; nd < C::ND; ++nd) {
          xw[nd * 4 + 0] = o[nd][0]; xw[nd * 4 + 1] = o[nd][1]; xw[nd * 4 + 2] = o[nd][2]; xw[nd * 4 + 3] = o[nd][3];
        }
        xw[C::ND * 4 + 0] = mrow[0]; xw[C::ND * 4 + 1] = mrow[1]; xw[C::ND * 4 + 2] = lrow[0]; xw[C::ND * 4 + 3] = lrow[1];
      }
      __syncthreads();
      if (sl > 0) continue;                       // (single pass: mt + nwarps >= mtiles, the loop ends here for every warp)
      for (int s2 = 1; s2 < nsl; ++s2) {
        const float* xo = reinterpret_cast<const float*>(smem) + ((size_t)(mt + s2 * mtiles) * 32 + lane) * XS;
        const float m0 = xo[C::ND * 4 + 0], m1 = xo[C::ND * 4 + 1];
        if (m0 == -INFINITY && m1 == -INFINITY) continue;       // that slice owned no key block
        const float n0 = fmaxf(mrow[0], m0), n1 = fmaxf(mrow[1], m1);
        const float a0 = ex2_approx(mrow[0] - n0), a1 = ex2_approx(mrow[1] - n1);
        const float b0 = ex2_approx(m0 - n0), b1 = ex2_approx(m1 - n1);
        mrow[0] = n0; mrow[1] = n1;
        lrow[0] = lrow[0] * a0 + xo[C::ND * 4 + 2] * b0;
        lrow[1] = lrow[1] * a1 + xo[C::ND * 4 + 3] * b1;
#pragma unroll
        for (int nd = 0; nd < C::ND; ++nd) {
          o[nd][0] = o[nd][0] * a0 + xo[nd * 4 + 0] * b0; o[nd][1] = o[nd][1] * a0 + xo[nd * 4 + 1] * b0;
          o[nd][2] = o[nd][2] * a1 + xo[nd * 4 + 2] * b1; o[nd][3] = o[nd][3] * a1 + xo[nd * 4 + 3] * b1;
        }
      }
    }
    // ---- normalise and store
    float l0 = lrow[0], l1 = lrow[1];
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = l0 > 0.f ? 1.f / l0 : 0.f, inv1 = l1 > 0.f ? 1.f / l1 : 0.f;
    __nv_bfloat16* o0 = attn_row(p.out, p.out2, p.ldo, b, r0 < p.Sq ? r0 : 0, p.Sq, p.n1) + h * D;
    __nv_bfloat16* o1 = attn_row(p.out, p.out2, p.ldo, b, r1 < p.Sq ? r1 : 0, p.Sq, p.n1) + h * D;
#pragma unroll
    for (int nd = 0; nd < C::ND; ++nd) {
      if (r0 < p.Sq) *reinterpret_cast<uint32_t*>(o0 + nd * 8 + t2) = pack_bf16(o[nd][0] * inv0, o[nd][1] * inv0);
      if (r1 < p.Sq) *reinterpret_cast<uint32_t*>(o1 + nd * 8 + t2) = pack_bf16(o[nd][2] * inv1, o[nd][3] * inv1);
    }
  }
}


// Long-key variant (sampler, up to ~1k candidate frames): one 16-row query tile per warp, K/V streamed through
// shared memory in chunks of KC keys, online-softmax state carried in registers across chunks.
template <int D>
__global__ void __launch_bounds__(256) attention_chunked_kernel(const AttnParams p, int KC) {
  using C = AttnCfg<D>;
  extern __shared__ __align__(16) uint8_t smem[];
  const int SkPad = (p.Sk + 15) & ~15;
  __nv_bfloat16* sKp = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* sVp = sKp + (size_t)KC * C::KP;
  float* smask = reinterpret_cast<float*>(sVp + (size_t)KC * C::VP);
  const int b = blockIdx.x / p.H, h = blockIdx.x - b * p.H;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int mt = blockIdx.y * nwarps + warp;
  const int g = lane >> 2, t2 = (lane & 3) * 2;
  const int r0 = mt * 16 + g, r1 = r0 + 8;
  uint32_t qf[C::KS][4];
  {
    const __nv_bfloat16* q0 = p.q + ((long long)b * p.Sq + r0) * p.ldq + h * D;
    const __nv_bfloat16* q1 = p.q + ((long long)b * p.Sq + r1) * p.ldq + h * D;
#pragma unroll
    for (int ks = 0; ks < C::KS; ++ks) {
      const int c0 = ks * 16 + t2, c1 = c0 + 8;
      qf[ks][0] = (r0 < p.Sq && c0 < D) ? __ldg(reinterpret_cast<const uint32_t*>(q0 + c0)) : 0u;
      qf[ks][1] = (r1 < p.Sq && c0 < D) ? __ldg(reinterpret_cast<const uint32_t*>(q1 + c0)) : 0u;
      qf[ks][2] = (r0 < p.Sq && c1 < D) ? __ldg(reinterpret_cast<const uint32_t*>(q0 + c1)) : 0u;
      qf[ks][3] = (r1 < p.Sq && c1 < D) ? __ldg(reinterpret_cast<const uint32_t*>(q1 + c1)) : 0u;
    }
  }
  float o[C::ND][4];
#pragma unroll
  for (int nd = 0; nd < C::ND; ++nd) { o[nd][0] = o[nd][1] = o[nd][2] = o[nd][3] = 0.f; }
  float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};
  const uint32_t sK = smem_u32(sKp), sV = smem_u32(sVp);
  const float* sm = p.mask ? smask : nullptr;
  constexpr int CH = D / 8, KCH = C::KP / 8, VCH = C::VP / 8;
  for (int c0 = 0; c0 < SkPad; c0 += KC) {
    const int len = (SkPad - c0) < KC ? (SkPad - c0) : KC;
    __syncthreads();
    for (int i = tid; i < len * KCH; i += blockDim.x) {
      const int r = i / KCH, c = i - r * KCH;
      __nv_bfloat16* dst = sKp + (size_t)r * C::KP + c * 8;
      if (c0 + r < p.Sk && c < CH) cp_async16(dst, p.k + ((long long)b * p.Sk + c0 + r) * p.ldk + h * D + c * 8);
      else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
    for (int i = tid; i < len * VCH; i += blockDim.x) {
      const int r = i / VCH, c = i - r * VCH;
      __nv_bfloat16* dst = sVp + (size_t)r * C::VP + c * 8;
      if (c0 + r < p.Sk && c < CH) cp_async16(dst, p.v + ((long long)b * p.Sk + c0 + r) * p.ldv + h * D + c * 8);
      else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
    if (p.mask)
      for (int i = tid; i < len; i += blockDim.x)
        smask[i] = (c0 + i) < p.Sk ? mask_log2(p.mask[(long long)b * p.Sk + c0 + i]) : 0.f;
    cp_async_wait_all();
    __syncthreads();
    int key0 = 0;
    for (; key0 + 64 <= len; key0 += 64)
      attn_block<D, 8>(qf, o, mrow, lrow, sK, sV, key0, p.Sk, sm, p.scale_log2, lane, c0);
    const int rem = len - key0;
    if (rem == 16) attn_block<D, 2>(qf, o, mrow, lrow, sK, sV, key0, p.Sk, sm, p.scale_log2, lane, c0);
    else if (rem == 32) attn_block<D, 4>(qf, o, mrow, lrow, sK, sV, key0, p.Sk, sm, p.scale_log2, lane, c0);
    else if (rem == 48) attn_block<D, 6>(qf, o, mrow, lrow, sK, sV, key0, p.Sk, sm, p.scale_log2, lane, c0);
  }
  float l0 = lrow[0], l1 = lrow[1];
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float inv0 = l0 > 0.f ? 1.f / l0 : 0.f, inv1 = l1 > 0.f ? 1.f / l1 : 0.f;
  __nv_bfloat16* o0 = p.out + ((long long)b * p.Sq + r0) * p.ldo + h * D;
  __nv_bfloat16* o1 = p.out + ((long long)b * p.Sq + r1) * p.ldo + h * D;
#pragma unroll
  for (int nd = 0; nd < C::ND; ++nd) {
    if (r0 < p.Sq) *reinterpret_cast<uint32_t*>(o0 + nd * 8 + t2) = pack_bf16(o[nd][0] * inv0, o[nd][1] * inv0);
    if (r1 < p.Sq) *reinterpret_cast<uint32_t*>(o1 + nd * 8 + t2) = pack_bf16(o[nd][2] * inv1, o[nd][3] * inv1);
  }
}

template <int D>
static int launch_attention_chunked(const AttnParams& p, cudaStream_t stream) {
  using C = AttnCfg<D>;
  const int KC = 256;
  const size_t smem = (size_t)KC * (C::KP + C::VP) * 2 + (size_t)KC * 4;
  auto kern = attention_chunked_kernel<D>;
  TGB_ONCE_PER_DEVICE(TGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
  const int mtiles = (p.Sq + 15) / 16;
  const int warps = mtiles < 8 ? mtiles : 8;
  dim3 grid(p.B * p.H, (mtiles + warps - 1) / warps);
  kern<<<grid, warps * 32, smem, stream>>>(p, KC);
  TGB_CUDA(cudaGetLastError());
  return TGB_OK;
}

template <int D>
static int launch_attention(const AttnParams& p, cudaStream_t stream) {
  using C = AttnCfg<D>;
  const int SkPad = (p.Sk + 15) & ~15;
  const size_t smem = (size_t)SkPad * (C::KP + C::VP) * 2 + (size_t)SkPad * 4;
  if (smem > 112 * 1024) return launch_attention_chunked<D>(p, stream);     // long key sequences: stream K/V
  auto kern = attention_kernel<D>;
  TGB_ONCE_PER_DEVICE({
    TGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    // ask for the largest shared-memory carve-out: the driver's default sized it for ONE CTA of this kernel per SM
    TGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  });
  const int mtiles = (p.Sq + 15) / 16;
  const int max_warps = D == 64 ? 8 : 9;          // (9 warps x 2 rounds cover the ViT's 17 row tiles)
  int warps = mtiles < max_warps ? mtiles : max_warps;
  AttnParams q = p;
  q.nslices = 1;
  size_t smem_launch = smem;
  const int nblocks = (SkPad + 63) / 64;
  if (mtiles <= 4 && nblocks >= 2) {              // few query rows, several key blocks: split the keys between warps
    int nsl = 8 / mtiles;
    if (nsl > nblocks) nsl = nblocks;
    if (nsl > 1) {
      q.nslices = nsl;
      warps = mtiles * nsl;
      const size_t merge = (size_t)warps * 32 * (C::ND * 4 + 4) * 4;
      if (merge > smem_launch) smem_launch = merge;
    }
  }
  kern<<<p.B * p.H, warps * 32, smem_launch, stream>>>(q);
  TGB_CUDA(cudaGetLastError());
  return TGB_OK;
}

}  // namespace tgb

using namespace tgb;

extern "C" int tgb_attention_bf16(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                  void* out, int64_t ldo, const float* mask, int B, int H, int Sq, int Sk, int D,
                                  float scale, tgb_stream_t stream) {
  int st = check_device();
  if (st != TGB_OK) return st;
  TGB_CHECK_ARG(q && k && v && out, "attention: null pointer");
  TGB_CHECK_ARG(B >= 0 && H > 0 && Sq > 0 && Sk > 0 && (long long)B * H <= 2147483647, "attention: bad shape B=%d H=%d Sq=%d Sk=%d", B, H, Sq, Sk);
  TGB_CHECK_ARG(D == 64 || D == 88, "attention: head dim %d not supported (64, 88)", D);
  TGB_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 2 == 0, "attention: leading dims must be multiples of 8");
  TGB_CHECK_ARG(((uintptr_t)q % 16) == 0 && ((uintptr_t)k % 16) == 0 && ((uintptr_t)v % 16) == 0 && ((uintptr_t)out % 4) == 0,
                "attention: q/k/v must be 16-byte aligned");
  if (B == 0) return TGB_OK;
  // the Q-Former cross-attention shape (32 queries x <= 257 image tokens) runs on tcgen05 (attention_qf_tc.cu)
  if (qf_cross_attention_tc_eligible(q, ldq, k, ldk, v, ldv, out, ldo, B, H, Sq, Sk, D, scale))
    return qf_cross_attention_tc(q, ldq, k, ldk, v, ldv, out, ldo, mask, B, H, Sk, scale, (cudaStream_t)stream);
  if (qf_self_attention_tc_eligible(q, ldq, k, ldk, v, ldv, out, ldo, B, H, Sq, Sk, D, scale))
    return qf_self_attention_tc(q, ldq, k, ldk, v, ldv, out, ldo, mask, B, H, scale, (cudaStream_t)stream);
  AttnParams p;
  p.q = (const __nv_bfloat16*)q; p.k = (const __nv_bfloat16*)k; p.v = (const __nv_bfloat16*)v;
  p.out = (__nv_bfloat16*)out; p.mask = mask;
  p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.ldo = ldo;
  p.B = B; p.H = H; p.Sq = Sq; p.Sk = Sk;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.q2 = p.k2 = p.v2 = nullptr; p.out2 = nullptr; p.n1 = 0;
  if (D == 88) return launch_attention<88>(p, (cudaStream_t)stream);
  return launch_attention<64>(p, (cudaStream_t)stream);
}

extern "C" int tgb_self_attention_2seg_bf16(const void* qkv1, const void* qkv2, int64_t ld_qkv, void* out1, void* out2,
                                            int64_t ldo, const float* mask, int B, int H, int n1, int n2, int D, float scale,
                                            tgb_stream_t stream) {
  int st = check_device();
  if (st != TGB_OK) return st;
  TGB_CHECK_ARG(qkv1 && qkv2 && out1 && out2, "self_attention_2seg: null pointer");
  TGB_CHECK_ARG(B >= 0 && H > 0 && n1 > 0 && n2 > 0 && (long long)B * H <= 2147483647, "self_attention_2seg: bad shape B=%d H=%d n1=%d n2=%d", B, H, n1, n2);
  TGB_CHECK_ARG(D == 64 || D == 88, "self_attention_2seg: head dim %d not supported (64, 88)", D);
  TGB_CHECK_ARG(ld_qkv % 8 == 0 && ld_qkv >= 3 * (int64_t)H * D && ldo % 2 == 0 && ldo >= (int64_t)H * D,
                "self_attention_2seg: leading dims must cover 3*H*D / H*D columns (multiples of 8 / 2)");
  TGB_CHECK_ARG(((uintptr_t)qkv1 % 16) == 0 && ((uintptr_t)qkv2 % 16) == 0 && ((uintptr_t)out1 % 4) == 0 && ((uintptr_t)out2 % 4) == 0,
                "self_attention_2seg: buffers must be 16-byte aligned");
  using C64 = AttnCfg<64>;
  const int S = n1 + n2, SkPad = (S + 15) & ~15;
  TGB_CHECK_ARG((size_t)SkPad * (C64::KP + C64::VP + 16) * 2 + (size_t)SkPad * 4 <= 112 * 1024,
                "self_attention_2seg: %d rows do not fit the single-pass kernel", S);
  if (B == 0) return TGB_OK;
  const __nv_bfloat16* a = (const __nv_bfloat16*)qkv1;
  const __nv_bfloat16* b2 = (const __nv_bfloat16*)qkv2;
  const int hd = H * D;
  AttnParams p;
  p.q = a; p.k = a + hd; p.v = a + 2 * hd;
  p.q2 = b2; p.k2 = b2 + hd; p.v2 = b2 + 2 * hd;
  p.out = (__nv_bfloat16*)out1; p.out2 = (__nv_bfloat16*)out2; p.mask = mask;
  p.ldq = p.ldk = p.ldv = ld_qkv; p.ldo = ldo;
  p.B = B; p.H = H; p.Sq = S; p.Sk = S; p.n1 = n1;
  p.scale_log2 = scale * 1.4426950408889634f;
  if (D == 88) return launch_attention<88>(p, (cudaStream_t)stream);
  return launch_attention<64>(p, (cudaStream_t)stream);
}
