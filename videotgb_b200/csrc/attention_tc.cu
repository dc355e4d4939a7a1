// ViT self-attention on the 5th-gen tensor cores (S = 257 tokens, 16 heads x 88, fused QKV layout [F*257, 3*16*88]).
//
// One persistent CTA per SM walks (frame, head) items.  Per item the 256 patch-token queries form two 128-row tiles
// (the CLS query row is done by a SIMT warp, see below); all 257 keys are processed in one shot, so there is no online
// rescaling:
//   warp 0      TMA producer: 3-D tensor map over [rows][q|k|v x head][88] -> Q tiles, K and V in shared memory
//               (SWIZZLE_128B, 64-wide d blocks; columns 88..127 are zero-filled by TMA bounds checking)
//   warp 1      MMA issuer (one thread):
//                 S[128 x 272] = Q K^T     tcgen05.mma kind::f16, A/B K-major from shared memory (N=256 + N=16)
//                 O[128 x  96] = P V       A = P in TENSOR MEMORY (bf16, written by the softmax warps), B = V as an
//                                          MN-major shared-memory operand (no transpose of V anywhere)
//   warp 2      TMEM allocator: S fp32 [0,272) | P bf16 [272,408) | O fp32 [416,512)
//   warp 3      CLS query row (row 0 of every image) in fp32 on CUDA cores, reading the swizzled K/V tiles in smem
//   warps 4-11  softmax + epilogue: two threads per query row (136 keys each) read S from TMEM ONCE into registers,
//               exchange the row max through shared memory, exponentiate (fp32, exp2), write P back to TMEM as bf16,
//               later read O, scale by 1/rowsum and store bf16.
// The score matrix never leaves the SM (the reference materialises [F,16,257,257], xblip2.py:174-189).
#include "common.cuh"
#include "tc_helpers.cuh"

#include <cuda.h>
#include <cudaTypedefs.h>

namespace tgb {

constexpr int VA_S = 257;
constexpr int VA_D = 88;
constexpr int VA_H = 16;
constexpr int VA_BOX_ROWS = 136;                       // TMA box rows (272 keys = 2 boxes; Q tiles use the first 128)
constexpr int VA_BOX_BYTES = VA_BOX_ROWS * 128;        // 17408, a multiple of 1024
constexpr int VA_KBLK_BYTES = 2 * VA_BOX_BYTES;        // one 64-wide d block of K or V: 272 rows x 128 B
constexpr int VA_THREADS = 384;
constexpr uint32_t VA_COL_S = 0, VA_COL_P = 272, VA_COL_O = 416;

constexpr int VA_SMEM_Q = 0;                           // [2 tiles][2 dblk][136 x 128 B]
constexpr int VA_SMEM_K = 4 * VA_BOX_BYTES;            // [2 dblk][272 x 128 B]
constexpr int VA_SMEM_V = VA_SMEM_K + 2 * VA_KBLK_BYTES;
constexpr int VA_SMEM_BAR = VA_SMEM_V + 2 * VA_KBLK_BYTES;      // 208896
constexpr int VA_SMEM_XMAX = VA_SMEM_BAR + 256;        // float [2][2][128]  (double buffered by tile parity)
constexpr int VA_SMEM_XSUM = VA_SMEM_XMAX + 2048;      // float [2][2][128]
constexpr int VA_SMEM_CLS = VA_SMEM_XSUM + 2048;       // float q0[96] + p[288]
constexpr int VA_SMEM_TOTAL = VA_SMEM_CLS + (96 + 288) * 4 + 1024;   // + alignment slack

struct VitAttnParams {
  const __nv_bfloat16* qkv;
  __nv_bfloat16* out;
  long long ld_qkv, ld_out;
  int frames;
  int num_items;            // frames * 16
  float scale_log2;
  int l2_prefetch;          // split-row kernel: L2 prefetch of the next item's operands (A/B knob, default off: slower)
  int dbg;                  // timing experiments only (results are wrong): 1 = operands are loaded for the first item only
};


__global__ void __launch_bounds__(VA_THREADS, 1)
vit_attention_tc_kernel(const __grid_constant__ CUtensorMap tmap, const VitAttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + VA_SMEM_Q;
  uint8_t* sK = smem + VA_SMEM_K;
  uint8_t* sV = smem + VA_SMEM_V;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + VA_SMEM_BAR);
  uint64_t* q_full = bars + 0;     // [2]
  uint64_t* q_free = bars + 2;     // [2]
  uint64_t* k_full = bars + 4;
  uint64_t* k_free = bars + 5;
  uint64_t* v_full = bars + 6;
  uint64_t* v_free = bars + 7;
  uint64_t* s_full = bars + 8;
  uint64_t* s_free = bars + 9;
  uint64_t* p_full = bars + 10;
  uint64_t* o_full = bars + 11;
  uint64_t* o_free = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
  float* xmax = reinterpret_cast<float*>(smem + VA_SMEM_XMAX);
  float* xsum = reinterpret_cast<float*>(smem + VA_SMEM_XSUM);
  float* cls_q = reinterpret_cast<float*>(smem + VA_SMEM_CLS);
  float* cls_p = cls_q + 96;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmap);
  if (warp == 1 && lane == 0) {
    mbar_init(&q_full[0], 1); mbar_init(&q_full[1], 1);
    mbar_init(&q_free[0], 1); mbar_init(&q_free[1], 1);
    mbar_init(k_full, 1); mbar_init(k_free, 2);      // k_free / v_free: tcgen05.commit + the CLS warp
    mbar_init(v_full, 1); mbar_init(v_free, 2);
    mbar_init(s_full, 1); mbar_init(s_free, 256);
    mbar_init(p_full, 256);
    mbar_init(o_full, 1); mbar_init(o_free, 256);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================================================ TMA producer
    if (lane == 0) {
      int i = 0;
      for (int w = blockIdx.x; w < p.num_items; w += gridDim.x, ++i) {
        const int b = w / VA_H, h = w - b * VA_H;
        const int row0 = b * VA_S;
        const uint32_t ph = i & 1;
        mbar_wait(k_free, ph ^ 1);
        mbar_arrive_expect_tx(k_full, 2 * VA_KBLK_BYTES);
#pragma unroll
        for (int db = 0; db < 2; ++db)
#pragma unroll
          for (int hf = 0; hf < 2; ++hf)
            tma_load_3d(sK + db * VA_KBLK_BYTES + hf * VA_BOX_BYTES, &tmap, k_full, db * 64, VA_H + h, row0 + hf * VA_BOX_ROWS);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&q_free[t], ph ^ 1);
          mbar_arrive_expect_tx(&q_full[t], 2 * VA_BOX_BYTES);
#pragma unroll
          for (int db = 0; db < 2; ++db)
            tma_load_3d(sQ + (t * 2 + db) * VA_BOX_BYTES, &tmap, &q_full[t], db * 64, h, row0 + 1 + t * 128);
        }
        mbar_wait(v_free, ph ^ 1);
        mbar_arrive_expect_tx(v_full, 2 * VA_KBLK_BYTES);
#pragma unroll
        for (int db = 0; db < 2; ++db)
#pragma unroll
          for (int hf = 0; hf < 2; ++hf)
            tma_load_3d(sV + db * VA_KBLK_BYTES + hf * VA_BOX_BYTES, &tmap, v_full, db * 64, 2 * VA_H + h, row0 + hf * VA_BOX_ROWS);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s256 = umma_idesc_bf16(128, 256);
      constexpr uint32_t idesc_s16 = umma_idesc_bf16(128, 16);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 96) | (1u << 16);     // B (= V) is MN-major
      const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), v_addr = smem_u32(sV);
      auto issue_pv = [&](int gp) {
        const int ip = gp >> 1, tp = gp & 1;
        if (tp == 0) mbar_wait(v_full, ip & 1);
        mbar_wait(p_full, gp & 1);
        mbar_wait(o_free, (gp & 1) ^ 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 17; ++kk) {
          const uint64_t db = umma_desc_mn_sw128(v_addr + kk * 2048, VA_KBLK_BYTES, 1024);
          umma_bf16_ts(tmem_base + VA_COL_O, tmem_base + VA_COL_P + kk * 8, db, idesc_pv, kk > 0 ? 1u : 0u);
        }
        umma_commit(o_full);
        if (tp == 1) umma_commit(v_free);
      };
      int g = 0, i = 0;
      for (int w = blockIdx.x; w < p.num_items; w += gridDim.x, ++i) {
        const uint32_t ph = i & 1;
        mbar_wait(k_full, ph);
#pragma unroll 1
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&q_full[t], ph);
          mbar_wait(s_free, (g & 1) ^ 1);
          tc_fence_after();
          uint32_t accum = 0;
#pragma unroll
          for (int db = 0; db < 2; ++db) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              if (db == 1 && ks >= 2) break;                       // d = 96..127 is all zero
              const uint64_t da = umma_desc_sw128(q_addr + (t * 2 + db) * VA_BOX_BYTES + ks * 32);
              const uint64_t dk = umma_desc_sw128(k_addr + db * VA_KBLK_BYTES + ks * 32);
              const uint64_t dk2 = umma_desc_sw128(k_addr + db * VA_KBLK_BYTES + 256 * 128 + ks * 32);
              umma_bf16_ss(tmem_base + VA_COL_S, da, dk, idesc_s256, accum);
              umma_bf16_ss(tmem_base + VA_COL_S + 256, da, dk2, idesc_s16, accum);
              accum = 1;
            }
          }
          umma_commit(s_full);
          umma_commit(&q_free[t]);
          if (t == 1) umma_commit(k_free);
          if (g > 0) issue_pv(g - 1);
          ++g;
        }
      }
      if (g > 0) issue_pv(g - 1);
    }
    __syncwarp();
  } else if (warp == 3) {
    // ================================================================ CLS query row (fp32, CUDA cores)
    // Reads K and V from the SAME swizzled shared-memory tiles the tensor cores use (element (key r, d c) of a
    // 64-wide d block lives at (r/8)*1024 + (r%8)*128 + (((c/8) ^ (r%8)) * 16) + (c%8)*2), so the CLS row costs no
    // extra global traffic; it co-signs k_free / v_free.
    int i = 0;
    for (int w = blockIdx.x; w < p.num_items; w += gridDim.x, ++i) {
      const int b = w / VA_H, h = w - b * VA_H;
      const long long row0 = (long long)b * VA_S;
      const __nv_bfloat16* qrow = p.qkv + row0 * p.ld_qkv + h * VA_D;
      for (int c = lane; c < VA_D / 2; c += 32) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(qrow + 2 * c));
        cls_q[2 * c] = f.x;
        cls_q[2 * c + 1] = f.y;
      }
      __syncwarp();
      mbar_wait(k_full, i & 1);
      float sc[9];
#pragma unroll
      for (int it = 0; it < 9; ++it) sc[it] = 0.f;
#pragma unroll 1
      for (int c = 0; c < VA_D / 8; ++c) {                      // 11 chunks of 8 d
        const float4 qa = *reinterpret_cast<const float4*>(cls_q + c * 8);
        const float4 qb = *reinterpret_cast<const float4*>(cls_q + c * 8 + 4);
        const uint8_t* kblk = sK + (c >> 3) * VA_KBLK_BYTES;
#pragma unroll
        for (int it = 0; it < 9; ++it) {
          const int r = it * 32 + lane;                          // rows up to 287 < 272? guard below
          if (r < VA_S) {
            const uint4 kk = *reinterpret_cast<const uint4*>(kblk + (r >> 3) * 1024 + (r & 7) * 128 + (((c & 7) ^ (r & 7)) << 4));
            const float2 k0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&kk.x));
            const float2 k1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&kk.y));
            const float2 k2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&kk.z));
            const float2 k3 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&kk.w));
            float d = sc[it];
            d = fmaf(k0.x, qa.x, d); d = fmaf(k0.y, qa.y, d); d = fmaf(k1.x, qa.z, d); d = fmaf(k1.y, qa.w, d);
            d = fmaf(k2.x, qb.x, d); d = fmaf(k2.y, qb.y, d); d = fmaf(k3.x, qb.z, d); d = fmaf(k3.y, qb.w, d);
            sc[it] = d;
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(k_free);                        // done reading K (second arrival: tcgen05.commit)
      float m = -INFINITY;
#pragma unroll
      for (int it = 0; it < 9; ++it) {
        sc[it] = (it * 32 + lane < VA_S) ? sc[it] * p.scale_log2 : -INFINITY;
        m = fmaxf(m, sc[it]);
      }
      m = warp_max(m);
      float l = 0.f;
#pragma unroll
      for (int it = 0; it < 9; ++it) {
        const int j = it * 32 + lane;
        const float pj = (j < VA_S) ? ex2_approx(sc[it] - m) : 0.f;
        cls_p[j] = pj;
        l += pj;
      }
      l = warp_sum(l);
      __syncwarp();
      mbar_wait(v_full, i & 1);
      // O[d] = sum_j p_j V[j][d]: lane -> (d chunk of 8 = lane % 11, key half = lane / 11), lanes 22..31 idle
      float acc[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = 0.f;
      if (lane < 22) {
        const int dc = lane % 11, kh = lane / 11;
        const uint8_t* vblk = sV + (dc >> 3) * VA_KBLK_BYTES;
        const int j0 = kh * 129, j1 = kh == 0 ? 129 : VA_S;
#pragma unroll 4
        for (int j = j0; j < j1; ++j) {
          const uint4 vv = *reinterpret_cast<const uint4*>(vblk + (j >> 3) * 1024 + (j & 7) * 128 + (((dc & 7) ^ (j & 7)) << 4));
          const float pj = cls_p[j];
          const float2 v0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&vv.x));
          const float2 v1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&vv.y));
          const float2 v2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&vv.z));
          const float2 v3 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&vv.w));
          acc[0] = fmaf(pj, v0.x, acc[0]); acc[1] = fmaf(pj, v0.y, acc[1]);
          acc[2] = fmaf(pj, v1.x, acc[2]); acc[3] = fmaf(pj, v1.y, acc[3]);
          acc[4] = fmaf(pj, v2.x, acc[4]); acc[5] = fmaf(pj, v2.y, acc[5]);
          acc[6] = fmaf(pj, v3.x, acc[6]); acc[7] = fmaf(pj, v3.y, acc[7]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(v_free);                        // done reading V
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += __shfl_down_sync(0xffffffffu, acc[e], 11);
      if (lane < 11) {
        const float inv = 1.f / l;
        uint4 o;
        o.x = pack_bf16(acc[0] * inv, acc[1] * inv);
        o.y = pack_bf16(acc[2] * inv, acc[3] * inv);
        o.z = pack_bf16(acc[4] * inv, acc[5] * inv);
        o.w = pack_bf16(acc[6] * inv, acc[7] * inv);
        *reinterpret_cast<uint4*>(p.out + row0 * p.ld_out + h * VA_D + lane * 8) = o;
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ================================================================ softmax + epilogue (2 threads per query row)
    const int quarter = warp & 3;
    const int half = (warp - 4) >> 2;            // keys [0,136) or [136,272); O columns [0,48) or [48,96)
    const int row = quarter * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const int nvalid = half == 0 ? 136 : (VA_S - 136);     // valid keys in this half (121 in the second)
    int g = 0;
    int pb = 0, ph_ = 0, pt = 0;                 // (frame, head, tile) of the previous tile, for its deferred epilogue

    auto epilogue = [&](int gp, int eb, int eh, int et) {
      const float l = xsum[((gp & 1) * 2 + 0) * 128 + row] + xsum[((gp & 1) * 2 + 1) * 128 + row];
      const float inv = 1.f / l;
      mbar_wait(o_full, gp & 1);
      tc_fence_after();
      // 48 (or 40) fp32 columns -> 24 packed bf16 registers, 16 columns at a time to keep register pressure low
      uint32_t ob[24];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        uint32_t o[16];
        tmem_ld16<0>(lane_base + VA_COL_O + half * 48 + c * 16, o);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 8; ++e)
          ob[c * 8 + e] = pack_bf16(__uint_as_float(o[2 * e]) * inv, __uint_as_float(o[2 * e + 1]) * inv);
      }
      tc_fence_before();
      mbar_arrive(o_free);
      __nv_bfloat16* dst = p.out + ((long long)eb * VA_S + 1 + et * 128 + row) * p.ld_out + eh * VA_D + half * 48;
      const int nchunks = half == 0 ? 6 : 5;     // 48 or 40 valid d columns (d 88..95 is padding)
#pragma unroll
      for (int c = 0; c < 6; ++c)
        if (c < nchunks) reinterpret_cast<uint4*>(dst)[c] = make_uint4(ob[4 * c], ob[4 * c + 1], ob[4 * c + 2], ob[4 * c + 3]);
    };

    for (int w = blockIdx.x; w < p.num_items; w += gridDim.x) {
      const int b = w / VA_H, h = w - b * VA_H;
#pragma unroll 1
      for (int t = 0; t < 2; ++t) {
        mbar_wait(s_full, g & 1);
        tc_fence_after();
        uint32_t sv[136];
        const uint32_t s_addr = lane_base + VA_COL_S + half * 136;
        tmem_ld32<0>(s_addr, sv);
        tmem_ld32<32>(s_addr + 32, sv);
        tmem_ld32<64>(s_addr + 64, sv);
        tmem_ld32<96>(s_addr + 96, sv);
        tmem_ld8<128>(s_addr + 128, sv);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(s_free);                      // S is in registers: the next tile's QK^T may overwrite it
        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;      // 4 independent chains
#pragma unroll
        for (int j = 0; j < 136; j += 4) {
          if (j + 0 < nvalid) m0 = fmaxf(m0, __uint_as_float(sv[j + 0]));
          if (j + 1 < nvalid) m1 = fmaxf(m1, __uint_as_float(sv[j + 1]));
          if (j + 2 < nvalid) m2 = fmaxf(m2, __uint_as_float(sv[j + 2]));
          if (j + 3 < nvalid) m3 = fmaxf(m3, __uint_as_float(sv[j + 3]));
        }
        float m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        float* xm = xmax + (g & 1) * 256;
        xm[half * 128 + row] = m;
        named_bar_sync(1 + quarter, 64);
        m = fmaxf(xm[row], xm[128 + row]);
        const float negmc = -m * p.scale_log2;
        float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
        const float sl2 = p.scale_log2;
        // exponentiate in place: sv[j] (j < 68) is overwritten with the packed bf16 pair (P[2j], P[2j+1]) as soon as
        // its two source values are dead.  The previous tile's epilogue sits between the two halves so that its
        // temporaries fit next to the (by then half-packed) score registers.
#pragma unroll
        for (int j = 0; j < 34; j += 2) {
          const float p0 = (2 * j + 0 < nvalid) ? ex2_approx(fmaf(__uint_as_float(sv[2 * j + 0]), sl2, negmc)) : 0.f;
          const float p1 = (2 * j + 1 < nvalid) ? ex2_approx(fmaf(__uint_as_float(sv[2 * j + 1]), sl2, negmc)) : 0.f;
          const float p2 = (2 * j + 2 < nvalid) ? ex2_approx(fmaf(__uint_as_float(sv[2 * j + 2]), sl2, negmc)) : 0.f;
          const float p3 = (2 * j + 3 < nvalid) ? ex2_approx(fmaf(__uint_as_float(sv[2 * j + 3]), sl2, negmc)) : 0.f;
          l0 += p0; l1 += p1; l2 += p2; l3 += p3;
          sv[j] = pack_bf16(p0, p1);
          sv[j + 1] = pack_bf16(p2, p3);
        }
        if (g > 0) epilogue(g - 1, pb, ph_, pt);  // also guarantees P V of the previous tile is done with the P region
#pragma unroll
        for (int j = 34; j < 68; j += 2) {
          const float p0 = (2 * j + 0 < nvalid) ? ex2_approx(fmaf(__uint_as_float(sv[2 * j + 0]), sl2, negmc)) : 0.f;
          const float p1 = (2 * j + 1 < nvalid) ? ex2_approx(fmaf(__uint_as_float(sv[2 * j + 1]), sl2, negmc)) : 0.f;
          const float p2 = (2 * j + 2 < nvalid) ? ex2_approx(fmaf(__uint_as_float(sv[2 * j + 2]), sl2, negmc)) : 0.f;
          const float p3 = (2 * j + 3 < nvalid) ? ex2_approx(fmaf(__uint_as_float(sv[2 * j + 3]), sl2, negmc)) : 0.f;
          l0 += p0; l1 += p1; l2 += p2; l3 += p3;
          sv[j] = pack_bf16(p0, p1);
          sv[j + 1] = pack_bf16(p2, p3);
        }
        const float l = (l0 + l1) + (l2 + l3);
        const uint32_t p_addr = lane_base + VA_COL_P + half * 68;
        tmem_st32<0>(p_addr, sv);
        tmem_st32<32>(p_addr + 32, sv);
        tmem_st4<64>(p_addr + 64, sv);
        xsum[((g & 1) * 2 + half) * 128 + row] = l;
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(p_full);
        pb = b; ph_ = h; pt = t;
        ++g;
      }
    }
    if (g > 0) {
      named_bar_sync(1 + quarter, 64);            // the partner's row sum of the last tile is visible
      epilogue(g - 1, pb, ph_, pt);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

#ifdef TGB_ATTN_TRACE
// Debug timeline (tools/attn_trace.py builds a private copy of the library with -DTGB_ATTN_TRACE): CTA 0 logs
// (tag, clock64) per role; role 0 TMA, 1 MMA, 2/3 CLS warps, 4/5 lane 0 of the first warp of softmax warpgroup 0/1.
constexpr int TR_ROLES = 6, TR_CAP = 1024;
__device__ long long g_attn_trace[TR_ROLES][TR_CAP][2];
__device__ int g_attn_trace_n[TR_ROLES];
#define TRACE_DECL int tr_n = 0;
#define TRACE(role, tag)                                                        \
  do {                                                                          \
    if (blockIdx.x == 0 && tr_n < TR_CAP) {                                     \
      g_attn_trace[role][tr_n][0] = (tag);                                      \
      g_attn_trace[role][tr_n][1] = clock64();                                  \
      g_attn_trace_n[role] = ++tr_n;                                            \
    }                                                                           \
  } while (0)
#else
#define TRACE_DECL
#define TRACE(role, tag) do { } while (0)
#endif

// =====================================================================================================================
// Ping-pong variant (default).  The two 128-row query tiles of an item are two independent chains, each with its own
// S/P and O accumulators in tensor memory and its own softmax warpgroup (one thread per query row, no cross-thread
// exchange), so the TMEM reads + exponentials of one tile overlap the tensor-core work of the other:
//   TMEM   S_t fp32 [t*144, t*144+144)  (P_t bf16 aliases its first 72 columns)   O_t fp32 [288 + t*96, +96)
//   keys   block 0 = keys [0,144) (one N=144 MMA per k-step), block 1 = keys [144,272) (N=128, 113 valid)
//   chain  S(t,0) -> softmax -> P V(t,0) -> S(t,1) -> online softmax (running max m; O_t and l are rescaled only if the
//          new block raises the max by more than 2^8 -- then p <= 256 is still exact enough in bf16/fp32) -> P V(t,1)
//          -> epilogue O_t / l.  The tensor pipe executes in issue order, so overwriting the P/S alias needs no barrier.
//   warp 0 TMEM alloc + TMA producer, warp 1 MMA issuer, warps 2/3 CLS query row (alternate items), warps 4-7 / 8-11
//   softmax warpgroups of tile 0 / tile 1.
// =====================================================================================================================
constexpr uint32_t PP_COL_S = 0, PP_S_STRIDE = 144, PP_COL_O = 288, PP_O_STRIDE = 96;
constexpr int PP_KEYS0 = 144;                  // keys in block 0; block 1 holds keys [144, 272)
constexpr int PP_VALID1 = VA_S - PP_KEYS0;     // 113 valid keys in block 1
constexpr float PP_RESCALE_THRESHOLD = 8.0f;   // log2 units
// shared memory: Q [tile][dblk][128 x 128 B] | K [dblk][272 x 128 B] | V [dblk][272 x 128 B] | barriers.  K and V are
// loaded (and released) per key block -- rows [0,144) through a 144-row TMA box, rows [144,272) through a 128-row box --
// so the next item's operands stream in while the current item is still being computed.
constexpr int PP_Q_BYTES = 128 * 128;                          // one (tile, d block)
constexpr int PP_SMEM_Q = 0;
constexpr int PP_SMEM_K = 4 * PP_Q_BYTES;
constexpr int PP_SMEM_V = PP_SMEM_K + 2 * VA_KBLK_BYTES;
constexpr int PP_SMEM_BAR = PP_SMEM_V + 2 * VA_KBLK_BYTES;     // 204800
constexpr int PP_SMEM_CLS = PP_SMEM_BAR + 256;                 // float [2 warps][272] CLS-row scores / probabilities
constexpr int PP_SMEM_STAGE = PP_SMEM_CLS + 2 * 272 * 4 + 128; // bf16 [128][88] output tile staged for the TMA store (128-B aligned)
constexpr int PP_STAGE_BYTES = 128 * VA_D * 2;                 // 22528
constexpr int PP_SMEM_TOTAL = PP_SMEM_STAGE + PP_STAGE_BYTES + 1024;   // + alignment slack
constexpr int PP_BLK0_BYTES = PP_KEYS0 * 128;                  // one d block of key block 0
constexpr int PP_BLK1_BYTES = (272 - PP_KEYS0) * 128;
// Split-row form (HALVES = 2, the default): every query row is shared by TWO threads of two different warpgroups (same
// TMEM lane quarter), half A owning the low score columns of a key block and half B the high ones, so each softmax phase
// of the serial S -> softmax -> P V chain takes about half as long:
//   key block 0 (144 columns): A [0,80)  -> P columns [0,40)     B [80,144) -> P columns [80,112)
//   key block 1 (128 columns): A [0,64)  -> P columns [0,32)     B [64,128) -> P columns [64,96)   (keys 113.. masked)
// (every half writes its bf16 P behind its own read position, so the two halves never touch each other's columns; the
// MMA issuer picks the P column of each 16-key step accordingly).  Row maxima are exchanged through shared memory
// (float [tile][half][128], one named barrier per tile), the two partial row sums through two spare TMEM columns per tile
// (ordered by the P-published -> P V committed mbarrier chain), the epilogue is split by d columns (A [0,40), B [40,88)).
constexpr int PP_SPLIT0 = 80, PP_SPLIT1 = 64;
constexpr uint32_t PP_COL_L = 480;                             // fp32 partial row sums: column 480 + 2*tile + half
constexpr int PP2_THREADS = 640;                               // 4 service warps + 4 softmax warpgroups
constexpr int PP2_SMEM_XMAX = PP_SMEM_BAR + 192;               // float [2 tiles][2 halves][128]
constexpr int PP2_SMEM_CLS = PP2_SMEM_XMAX + 2048;             // float [2][144] scores + [2][100] exchange = 1952 B
constexpr int PP2_SMEM_STAGE = (PP2_SMEM_CLS + 1952 + 127) / 128 * 128;
// the dynamic shared memory base is 1024-byte aligned in practice; the kernel traps if it is not even 128-byte aligned
constexpr int PP2_SMEM_TOTAL = PP2_SMEM_STAGE + PP_STAGE_BYTES + 896;
static_assert(PP2_SMEM_TOTAL <= 227 * 1024, "shared memory budget (split-row ping-pong kernel)");

// Row max of the first NVALID of NCOLS fp32 score columns at `s_addr`: every tcgen05.ld is in flight at once (the only
// moment a thread holds a whole score row), 3-input max tree.
template <int NCOLS, int NVALID>
__device__ __forceinline__ float pp_row_max(uint32_t s_addr) {
  static_assert(NCOLS == 128 || NCOLS == 144, "two key-block widths");
  uint32_t v[NCOLS];
  tmem_ld32<0>(s_addr, v);
  tmem_ld32<32>(s_addr + 32, v);
  tmem_ld32<64>(s_addr + 64, v);
  tmem_ld32<96>(s_addr + 96, v);
  if (NCOLS == 144) tmem_ld16<128>(s_addr + 128, v);
  tmem_ld_wait();
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
  for (int j = 0; j < NVALID; j += 4) {
    m0 = fmaxf(m0, __uint_as_float(v[j]));
    if (j + 1 < NVALID) m1 = fmaxf(m1, __uint_as_float(v[j + 1]));
    if (j + 2 < NVALID) m2 = fmaxf(m2, __uint_as_float(v[j + 2]));
    if (j + 3 < NVALID) m3 = fmaxf(m3, __uint_as_float(v[j + 3]));
  }
  return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}

// Row max of the first NVALID of NCOLS (64 or 80) fp32 score columns at `s_addr` (split-row form: at most 64 score
// registers live, the thread budget is 96).
template <int NCOLS, int NVALID>
__device__ __forceinline__ float pp_row_max_part(uint32_t s_addr) {
  static_assert(NCOLS == 64 || NCOLS == 80, "column counts of the row halves");
  uint32_t v[32], w[32];
  tmem_ld32<0>(s_addr, v);
  tmem_ld32<0>(s_addr + 32, w);
  tmem_ld_wait();
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    if (j < NVALID) m0 = fmaxf(m0, __uint_as_float(v[j]));
    if (j + 1 < NVALID) m1 = fmaxf(m1, __uint_as_float(v[j + 1]));
    if (j + 2 < NVALID) m2 = fmaxf(m2, __uint_as_float(v[j + 2]));
    if (j + 3 < NVALID) m3 = fmaxf(m3, __uint_as_float(v[j + 3]));
  }
  if (NCOLS == 80) tmem_ld16<0>(s_addr + 64, v);
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    if (32 + j < NVALID) m0 = fmaxf(m0, __uint_as_float(w[j]));
    if (32 + j + 1 < NVALID) m1 = fmaxf(m1, __uint_as_float(w[j + 1]));
    if (32 + j + 2 < NVALID) m2 = fmaxf(m2, __uint_as_float(w[j + 2]));
    if (32 + j + 3 < NVALID) m3 = fmaxf(m3, __uint_as_float(w[j + 3]));
  }
  if (NCOLS == 80) {
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      if (64 + j < NVALID) m0 = fmaxf(m0, __uint_as_float(v[j]));
      if (64 + j + 1 < NVALID) m1 = fmaxf(m1, __uint_as_float(v[j + 1]));
      if (64 + j + 2 < NVALID) m2 = fmaxf(m2, __uint_as_float(v[j + 2]));
      if (64 + j + 3 < NVALID) m3 = fmaxf(m3, __uint_as_float(v[j + 3]));
    }
  }
  return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t& a, uint32_t& b) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "r"(taddr));
}

// 16 score columns -> 8 packed bf16 probability registers; staged (fma, ex2, add/pack) so the MUFU pipe is fed back to
// back.  NV = valid columns of the chunk (the rest produce p = 0).
template <int NV>
__device__ __forceinline__ void pp_exp16(uint32_t (&v)[16], float sl2, float negm, float2& la, float2& lb) {
  // packed fp32 (FFMA2 / FADD2): half the FP32-pipe instructions of the scalar form, same roundings and summation order
  // (la = (l0, l1), lb = (l2, l3) of the four interleaved partial sums)
  float2 x[8];
  const float2 s2 = make_float2(sl2, sl2), m2 = make_float2(negm, negm);
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = __ffma2_rn(make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), s2, m2);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    x[j].x = (2 * j < NV) ? ex2_approx(x[j].x) : 0.f;
    x[j].y = (2 * j + 1 < NV) ? ex2_approx(x[j].y) : 0.f;
  }
#pragma unroll
  for (int j = 0; j < 8; j += 2) {
    la = __fadd2_rn(la, x[j]);
    lb = __fadd2_rn(lb, x[j + 1]);
    v[j] = pack_bf16(x[j].x, x[j].y);
    v[j + 1] = pack_bf16(x[j + 1].x, x[j + 1].y);
  }
}

// Pass 2 over one key block: NFULL all-valid 16-column chunks in a ROLLED loop (the unrolled form was instruction-fetch
// bound: every warp streamed ~10 KB of straight-line code per block) plus one peeled chunk with TAIL valid columns
// (TAIL = 16: a plain extra chunk; 0 < TAIL < 16: masked; the P columns of a fully masked remainder are zeroed).
// P chunk c (8 bf16x2 columns) is stored behind the read position; the next chunk's tcgen05.ld is issued before the
// current chunk is processed.  Returns the row sum.
template <int NFULL, int TAIL>
__device__ __forceinline__ float pp_exp_block(uint32_t s_addr, float sl2, float negm) {
  float2 la = make_float2(0.f, 0.f), lb = make_float2(0.f, 0.f);
  uint32_t cur[16], nxt[16];
  tmem_ld16<0>(s_addr, cur);
#pragma unroll 1
  for (int c = 0; c < NFULL; ++c) {
    tmem_ld_wait();
    if (TAIL > 0 || c + 1 < NFULL) tmem_ld16<0>(s_addr + (c + 1) * 16, nxt);   // (never reads past the block)
    pp_exp16<16>(cur, sl2, negm, la, lb);
    tmem_st8<0>(s_addr + c * 8, cur);
#pragma unroll
    for (int j = 0; j < 16; ++j) cur[j] = nxt[j];
  }
  if (TAIL > 0) {
    tmem_ld_wait();
    pp_exp16<TAIL>(cur, sl2, negm, la, lb);
    tmem_st8<0>(s_addr + NFULL * 8, cur);
  }
  return (la.x + la.y) + (lb.x + lb.y);
}

// ---- software-pipelined form of pass 2.  In pp_exp_block a chunk's 16 MUFU.EX2 (8 issue cycles each on the SMSP: 128
// clk) are followed by its adds / packs / TMEM store and the next chunk's scaling, all of which depend on -- or feed --
// them, and the warp is alone on its SMSP during its softmax phase: measured 217-240 clk per chunk, MUFU idle 45 % of
// the phase.  Here the three stages of three different chunks sit in ONE basic block, mutually independent:
//   pack + store of chunk c-1   |   ex2 of chunk c   |   wait + scale of chunk c+1, load of chunk c+2
// so the scheduler can keep the MUFU pipe fed.  Chunks alternate between two register sets (rolled loop over pairs).
__device__ __forceinline__ void pp_scale16(const uint32_t (&raw)[16], float2 (&x)[8], float sl2, float negm) {
  const float2 s2 = make_float2(sl2, sl2), m2 = make_float2(negm, negm);
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = __ffma2_rn(make_float2(__uint_as_float(raw[2 * j]), __uint_as_float(raw[2 * j + 1])), s2, m2);
}
#ifndef TGB_EXP_EXPERIMENT
#define TGB_EXP_EXPERIMENT 0
#endif
template <int NV>
__device__ __forceinline__ void pp_ex2_16(float2 (&x)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
#if TGB_EXP_EXPERIMENT == 1      // timing experiment: no MUFU
    x[j].x = (2 * j < NV) ? x[j].x * 0.001f : 0.f;
    x[j].y = (2 * j + 1 < NV) ? x[j].y * 0.001f : 0.f;
#else
    x[j].x = (2 * j < NV) ? ex2_approx(x[j].x) : 0.f;
    x[j].y = (2 * j + 1 < NV) ? ex2_approx(x[j].y) : 0.f;
#endif
  }
}
__device__ __forceinline__ uint32_t pp_pack2(float lo, float hi) {
#if TGB_EXP_EXPERIMENT == 2      // timing experiment: truncating pack on the ALU pipe (PRMT) instead of F2FP
  return __byte_perm(__float_as_uint(lo), __float_as_uint(hi), 0x7632);
#else
  return pack_bf16(lo, hi);
#endif
}
__device__ __forceinline__ void pp_pack_store16(const float2 (&e)[8], uint32_t p_addr, float2& la, float2& lb) {
  uint32_t pk[8];
#pragma unroll
  for (int j = 0; j < 8; j += 2) {                 // same summation order as pp_exp16
    la = __fadd2_rn(la, e[j]);
    lb = __fadd2_rn(lb, e[j + 1]);
    pk[j] = pp_pack2(e[j].x, e[j].y);
    pk[j + 1] = pp_pack2(e[j + 1].x, e[j + 1].y);
  }
  tmem_st8<0>(p_addr, pk);
}
// one pipeline step: chunk c is scaled in `cur`; `prev` holds the exponentials of chunk c-1 (HAVE_PREV) and receives the
// scaled chunk c+1 (have_next); `raw` is the in-flight load of chunk c+1, re-issued for chunk c+2 (have_next2).
template <int NV, bool HAVE_PREV>
__device__ __forceinline__ void pp_sp_step(uint32_t s_addr, int c, float2 (&cur)[8], float2 (&prev)[8], uint32_t (&raw)[16],
                                           bool have_next, bool have_next2, float sl2, float negm, float2& la, float2& lb) {
  if (HAVE_PREV) pp_pack_store16(prev, s_addr + (c - 1) * 8, la, lb);
  pp_ex2_16<NV>(cur);
  if (have_next) {
    tmem_ld_wait_dep(raw);
    pp_scale16(raw, prev, sl2, negm);
    if (have_next2) tmem_ld16<0>(s_addr + (c + 2) * 16, raw);
  }
}
// NCHUNK 16-column chunks, the last one with TAILV (1..16) valid columns.  Returns the row sum.
template <int NCHUNK, int TAILV>
__device__ __forceinline__ float pp_exp_block_sp(uint32_t s_addr, float sl2, float negm) {
  static_assert(NCHUNK >= 3 && TAILV >= 1 && TAILV <= 16, "pipeline shape");
  float2 la = make_float2(0.f, 0.f), lb = make_float2(0.f, 0.f);
  uint32_t raw[16];
  float2 xa[8], xb[8];
  tmem_ld16<0>(s_addr, raw);
  tmem_ld_wait_dep(raw);
  pp_scale16(raw, xa, sl2, negm);
  tmem_ld16<0>(s_addr + 16, raw);
  pp_sp_step<16, false>(s_addr, 0, xa, xb, raw, true, true, sl2, negm, la, lb);        // chunk 0 (xa); chunk 1 -> xb
  constexpr int MID = NCHUNK - 2;                                                      // chunks 1 .. NCHUNK-2, all full
  int c = 1;
#pragma unroll 1
  for (int k = 0; k < MID / 2; ++k, c += 2) {
    pp_sp_step<16, true>(s_addr, c, xb, xa, raw, true, true, sl2, negm, la, lb);
    pp_sp_step<16, true>(s_addr, c + 1, xa, xb, raw, true, c + 3 < NCHUNK, sl2, negm, la, lb);
  }
  if (MID & 1) {
    pp_sp_step<16, true>(s_addr, NCHUNK - 2, xb, xa, raw, true, false, sl2, negm, la, lb);
    pp_sp_step<TAILV, true>(s_addr, NCHUNK - 1, xa, xb, raw, false, false, sl2, negm, la, lb);
    pp_pack_store16(xa, s_addr + (NCHUNK - 1) * 8, la, lb);
  } else {
    pp_sp_step<TAILV, true>(s_addr, NCHUNK - 1, xb, xa, raw, false, false, sl2, negm, la, lb);
    pp_pack_store16(xb, s_addr + (NCHUNK - 1) * 8, la, lb);
  }
  return (la.x + la.y) + (lb.x + lb.y);
}

// One CLS query row (row 0 of the image) on the warp-level tensor-core path (mma.sync.m16n8k16), one warp per item,
// reading the SAME swizzled K/V tiles TMA staged for tcgen05 (element (key r, d c) of a 64-wide d block lives at
// (r/8)*1024 + (r%8)*128 + (((c/8) ^ (r%8)) * 16) + (c%8)*2) -- no extra global traffic.  Transposed formulation so that
// the single query occupies the N side (8 columns, one valid) instead of the 16-row M side:
//   S^T[keys x 1] = K[keys x d] q^T      A = K through ldmatrix,        B = q (lanes 0..3 hold it, the rest are zero)
//   O^T[d x 1]    = V^T[d x keys] p      A = V through ldmatrix.trans,  B = p (moved from the lanes that own column 0)
// Column 0 of an accumulator tile lives in lanes 0,4,..,28 (c0: row lane/4, c2: row lane/4 + 8).
__device__ __forceinline__ void va_ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void va_ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void va_mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// byte offset of the 16-byte chunk holding d [8*chunk, 8*chunk+8) of key `r` (chunk in 0..11: 0..7 in d block 0)
__device__ __forceinline__ uint32_t va_tile_off(int r, int chunk) {
  return (uint32_t)((chunk >> 3) * VA_KBLK_BYTES + (r >> 3) * 1024 + (r & 7) * 128 + ((((chunk & 7) ^ (r & 7))) << 4));
}

// mma without `volatile`: a pure function of its operands, so independent accumulator chains may be interleaved
__device__ __forceinline__ void va_mma_16816_nv(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// The CLS query row of one item is split between two warps BY KEY BLOCK (warp BLK handles keys of block BLK and only
// touches that block's K / V buffers, so each block is released to the TMA producer as early as possible); the two
// partial results (own max m, row sum l, unnormalised O) are merged by the block-0 warp through `xch`.
// sc_s: warp-private float[144] (scores, then bf16 probabilities over the same storage); xch: float[2][100].
// Loops are rolled: the routine runs once per item per warp and straight-line code would be fetched cold every time.
template <int BLK>
__device__ __forceinline__ void cls_row_block(const VitAttnParams& p, int w, int i, uint32_t sK, uint32_t sV,
                                              uint64_t* k_full, uint64_t* k_free, uint64_t* v_full, uint64_t* v_free,
                                              float* sc_s, float* xch, int lane) {
  constexpr int NT = BLK == 0 ? PP_KEYS0 / 16 : (272 - PP_KEYS0) / 16;     // 16-key tiles: 9 / 8
  constexpr int KEY0 = BLK * PP_KEYS0;
  constexpr int NVALID = BLK == 0 ? PP_KEYS0 : PP_VALID1;                  // 144 / 113
  const uint32_t ph = i & 1;
  const int b = w / VA_H, h = w - b * VA_H;
  const long long row0 = (long long)b * VA_S;
  const int g = lane >> 2, t = lane & 3;
  // B fragments of q: b0 = d (16ks + 2t, +1), b1 = d (16ks + 8 + 2t, +1), column n = g -> only g == 0 is the real query
  uint32_t qb[6][2];
  {
    const uint32_t* q32 = reinterpret_cast<const uint32_t*>(p.qkv + row0 * p.ld_qkv + h * VA_D);
#pragma unroll
    for (int ks = 0; ks < 6; ++ks) {
      qb[ks][0] = (g == 0) ? __ldg(q32 + ks * 8 + t) : 0u;
      qb[ks][1] = (g == 0 && ks < 5) ? __ldg(q32 + ks * 8 + 4 + t) : 0u;      // d 88..95 does not exist
    }
  }
  const int lrow = lane & 7, lmat = lane >> 3;
  mbar_wait(&k_full[BLK], ph);
  // ---- S^T = K q^T: two key tiles per iteration, two accumulators per tile (even / odd k-steps) = 4 independent chains
  auto qk_pair = [&](int mt, bool second) {
    uint32_t a0[6][4], a1[6][4];
    const int key = KEY0 + mt * 16 + (lmat & 1) * 8 + lrow;     // matrices: (keys lo, d lo) (keys hi, d lo) (keys lo, d hi) (keys hi, d hi)
#pragma unroll
    for (int ks = 0; ks < 6; ++ks) va_ldsm_x4(sK + va_tile_off(key, 2 * ks + (lmat >> 1)), a0[ks]);
    if (second) {
#pragma unroll
      for (int ks = 0; ks < 6; ++ks) va_ldsm_x4(sK + va_tile_off(key + 16, 2 * ks + (lmat >> 1)), a1[ks]);
    }
    float c0e[4] = {0.f, 0.f, 0.f, 0.f}, c0o[4] = {0.f, 0.f, 0.f, 0.f}, c1e[4] = {0.f, 0.f, 0.f, 0.f}, c1o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ks = 0; ks < 6; ks += 2) {
      va_mma_16816_nv(c0e, a0[ks], qb[ks][0], qb[ks][1]);
      va_mma_16816_nv(c0o, a0[ks + 1], qb[ks + 1][0], qb[ks + 1][1]);
      if (second) {
        va_mma_16816_nv(c1e, a1[ks], qb[ks][0], qb[ks][1]);
        va_mma_16816_nv(c1o, a1[ks + 1], qb[ks + 1][0], qb[ks + 1][1]);
      }
    }
    if (t == 0) {
      sc_s[mt * 16 + g] = c0e[0] + c0o[0];
      sc_s[mt * 16 + 8 + g] = c0e[2] + c0o[2];
      if (second) {
        sc_s[mt * 16 + 16 + g] = c1e[0] + c1o[0];
        sc_s[mt * 16 + 24 + g] = c1e[2] + c1o[2];
      }
    }
  };
#pragma unroll 1
  for (int mt = 0; mt + 1 < NT; mt += 2) qk_pair(mt, true);
  if (NT & 1) qk_pair(NT - 1, false);
  __syncwarp();
  if (lane == 0) mbar_arrive(&k_free[BLK]);                      // done reading this K block (second arrival: tcgen05.commit)
  // ---- local softmax over this block's valid keys, 5 per lane
  float x[5];
  float m = -INFINITY;
#pragma unroll
  for (int it = 0; it < 5; ++it) {
    const int j = it * 32 + lane;
    x[it] = (j < NVALID) ? sc_s[j] * p.scale_log2 : -INFINITY;
    m = fmaxf(m, x[it]);
  }
  m = warp_max(m);
  float l = 0.f;
#pragma unroll
  for (int it = 0; it < 5; ++it) {
    x[it] = ex2_approx(x[it] - m);                               // exp2(-inf) = 0 for masked keys
    l += x[it];
  }
  l = warp_sum(l);
  __syncwarp();                                                  // every lane has read its scores before p overwrites them
  __nv_bfloat16* p_s = reinterpret_cast<__nv_bfloat16*>(sc_s);
#pragma unroll
  for (int it = 0; it < 5; ++it)
    if (it * 32 + lane < NT * 16) p_s[it * 32 + lane] = __float2bfloat16(x[it]);
  __syncwarp();
  mbar_wait(&v_full[BLK], ph);
  // ---- O^T = V^T p over this block's keys
  float o[6][4];
#pragma unroll
  for (int mt = 0; mt < 6; ++mt) { o[mt][0] = o[mt][1] = o[mt][2] = o[mt][3] = 0.f; }
  const uint32_t* p32 = reinterpret_cast<const uint32_t*>(p_s);
#pragma unroll 1
  for (int kk = 0; kk < NT; ++kk) {
    // B fragment of p for keys [16kk, 16kk+16): b0 = keys (2t, 2t+1), b1 = keys (8+2t, 9+2t), column n = g
    const uint32_t b0 = (g == 0) ? p32[kk * 8 + t] : 0u;
    const uint32_t b1 = (g == 0) ? p32[kk * 8 + 4 + t] : 0u;
    const int key = KEY0 + kk * 16 + (lmat >> 1) * 8 + lrow;     // matrices: (keys lo, d lo) (keys lo, d hi) (keys hi, d lo) (keys hi, d hi)
    uint32_t a[6][4];
#pragma unroll
    for (int mt = 0; mt < 6; ++mt) va_ldsm_x4_t(sV + va_tile_off(key, 2 * mt + (lmat & 1)), a[mt]);
#pragma unroll
    for (int mt = 0; mt < 6; ++mt) va_mma_16816_nv(o[mt], a[mt], b0, b1);
  }
  __syncwarp();
  if (lane == 0) mbar_arrive(&v_free[BLK]);                      // done reading this V block
  // ---- merge: block-1 warp publishes (O, m, l); block-0 warp combines and stores
  float* xb = xch + (i & 1) * 100;
  if (BLK == 1) {
    if (t == 0) {
#pragma unroll
      for (int mt = 0; mt < 6; ++mt) { xb[mt * 16 + g] = o[mt][0]; xb[mt * 16 + 8 + g] = o[mt][2]; }
    }
    if (lane == 0) { xb[96] = m; xb[97] = l; }
    __syncwarp();
    named_bar_sync(3, 64);
  } else {
    named_bar_sync(3, 64);
    const float m1 = xb[96], l1 = xb[97];
    const float mm = fmaxf(m, m1);
    const float f0 = ex2_approx(m - mm), f1 = ex2_approx(m1 - mm);
    const float inv = 1.f / (l * f0 + l1 * f1);
    if (t == 0) {
      __nv_bfloat16* dst = p.out + row0 * p.ld_out + h * VA_D;
#pragma unroll
      for (int mt = 0; mt < 6; ++mt) {
        dst[mt * 16 + g] = __float2bfloat16((o[mt][0] * f0 + xb[mt * 16 + g] * f1) * inv);
        if (mt < 5) dst[mt * 16 + 8 + g] = __float2bfloat16((o[mt][2] * f0 + xb[mt * 16 + 8 + g] * f1) * inv);
      }
    }
  }
  __syncwarp();
}

template <int HALVES, bool SP>
__global__ void __launch_bounds__(HALVES == 1 ? VA_THREADS : PP2_THREADS, 1)
vit_attention_pp_kernel(const __grid_constant__ CUtensorMap tmap144, const __grid_constant__ CUtensorMap tmap128,
                        const __grid_constant__ CUtensorMap tmap_out, const VitAttnParams p) {
  static_assert(HALVES == 1 || HALVES == 2, "threads per query row");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  if (HALVES == 2 && (smem_u32(smem_raw) & 127u) != 0) __trap();      // PP2_SMEM_TOTAL carries 896 bytes of alignment slack
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + PP_SMEM_Q;
  uint8_t* sK = smem + PP_SMEM_K;
  uint8_t* sV = smem + PP_SMEM_V;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + PP_SMEM_BAR);
  uint64_t* q_full = bars + 0;     // [2] per query tile
  uint64_t* q_free = bars + 2;     // [2]
  uint64_t* k_full = bars + 4;     // [2] per key block
  uint64_t* k_free = bars + 6;     // [2]
  uint64_t* v_full = bars + 8;     // [2]
  uint64_t* v_free = bars + 10;    // [2]
  uint64_t* s_full = bars + 12;    // [2] per query tile, one completion per (item, key block)
  uint64_t* p_full = bars + 14;    // [2]
  uint64_t* o_full = bars + 16;    // [2] one completion per item
  uint64_t* o_free = bars + 18;    // [2]
  uint64_t* stage_free = bars + 20; // the output staging tile has been read by its TMA store (one completion per use)
  uint64_t* stage_ready = bars + 21; // ... has been written by a softmax warpgroup (128 arrivals per use)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);
  float* cls_scores = reinterpret_cast<float*>(smem + (HALVES == 1 ? PP_SMEM_CLS : PP2_SMEM_CLS));
  uint8_t* stage = smem + (HALVES == 1 ? PP_SMEM_STAGE : PP2_SMEM_STAGE);
  float* xmax = reinterpret_cast<float*>(smem + PP2_SMEM_XMAX);   // (HALVES == 2 only) [tile][half][128]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 1 && lane == 0) {
    tma_prefetch_desc(&tmap144);
    tma_prefetch_desc(&tmap128);
    tma_prefetch_desc(&tmap_out);
    mbar_init(stage_free, 1);
    mbar_init(stage_ready, 128 * HALVES);
    for (int t = 0; t < 2; ++t) {
      mbar_init(&q_full[t], 1); mbar_init(&q_free[t], 1);
      mbar_init(&k_full[t], 1); mbar_init(&k_free[t], 2);      // tcgen05.commit + the CLS warp of the item
      mbar_init(&v_full[t], 1); mbar_init(&v_free[t], 2);
      mbar_init(&s_full[t], 1); mbar_init(&p_full[t], 128 * HALVES);
      mbar_init(&o_full[t], 1); mbar_init(&o_free[t], 128 * HALVES);
    }
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================================================ TMA producer
    // Loads are issued in the order the previous item releases their buffers: K0, Q0, V0, K1, Q1, V1.
    if (lane == 0) {
      TRACE_DECL
      // Output tiles: a softmax warpgroup stages its bf16 [128][88] tile in shared memory (use n = 2*item + tile of the
      // single staging buffer); this thread turns it into one asynchronous bulk tensor store and hands the buffer back.
      // The stores sit at points of the load schedule where this thread has slack.
      auto store_tile = [&](int ws, int is, int t) {
        const int use = 2 * is + t;
        mbar_wait(stage_ready, use & 1);
        const int bs = ws / VA_H, hs = ws - bs * VA_H;
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                     ::"l"(reinterpret_cast<uint64_t>(&tmap_out)), "r"(smem_u32(stage)), "r"(0), "r"(hs),
                       "r"(bs * VA_S + 1 + t * 128) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        mbar_arrive(stage_free);
      };
      int i = 0;
      for (int w = blockIdx.x; w < p.num_items; w += gridDim.x, ++i) {
        const int b = w / VA_H, h = w - b * VA_H;
        const int row0 = b * VA_S;
        const uint32_t ph = i & 1;
        const bool skip_loads = (p.dbg & 1) && i > 1;
        auto load_kv = [&](uint8_t* dst, uint64_t* full, int blk, int col_h) {
          if (skip_loads) { mbar_arrive(full); return; }
          mbar_arrive_expect_tx(full, 2 * (blk == 0 ? PP_BLK0_BYTES : PP_BLK1_BYTES));
#pragma unroll
          for (int db = 0; db < 2; ++db)
            tma_load_3d(dst + db * VA_KBLK_BYTES + blk * PP_BLK0_BYTES, blk == 0 ? &tmap144 : &tmap128, full, db * 64, col_h,
                        row0 + blk * PP_KEYS0);
        };
        auto load_q = [&](int t) {
          mbar_wait(&q_free[t], ph ^ 1);
          TRACE(0, i * 100 + 2 + t);                             // Q_t load issued
          if (skip_loads) { mbar_arrive(&q_full[t]); return; }
          mbar_arrive_expect_tx(&q_full[t], 2 * PP_Q_BYTES);
#pragma unroll
          for (int db = 0; db < 2; ++db)
            tma_load_3d(sQ + (t * 2 + db) * PP_Q_BYTES, &tmap128, &q_full[t], db * 64, h, row0 + 1 + t * 128);
        };
        mbar_wait(&k_free[0], ph ^ 1);
        TRACE(0, i * 100 + 1);                                   // K block 0 load issued
        load_kv(sK, &k_full[0], 0, VA_H + h);
        load_q(0);
        if (HALVES == 2 && p.l2_prefetch && w + (int)gridDim.x < p.num_items) {
          // The loads of an item are issued as the previous item frees their buffers, i.e. at most one softmax phase
          // before they are needed, and a load from DRAM takes 3-6k clk (176-byte pieces at an 8 448-byte stride): pull
          // the NEXT item's operands into the L2 now, so that its loads are L2 hits.
          const int wn = w + (int)gridDim.x;
          const int bn = wn / VA_H, hn = wn - bn * VA_H, rown = bn * VA_S;
#pragma unroll
          for (int db = 0; db < 2; ++db) {
            tma_prefetch_3d(&tmap144, db * 64, VA_H + hn, rown);
            tma_prefetch_3d(&tmap128, db * 64, hn, rown + 1);
            tma_prefetch_3d(&tmap128, db * 64, hn, rown + 129);
            tma_prefetch_3d(&tmap144, db * 64, 2 * VA_H + hn, rown);
            tma_prefetch_3d(&tmap128, db * 64, VA_H + hn, rown + PP_KEYS0);
            tma_prefetch_3d(&tmap128, db * 64, 2 * VA_H + hn, rown + PP_KEYS0);
          }
        }
        mbar_wait(&v_free[0], ph ^ 1);
        TRACE(0, i * 100 + 4);                                   // V block 0 load issued
        load_kv(sV, &v_full[0], 0, 2 * VA_H + h);
        mbar_wait(&k_free[1], ph ^ 1);
        TRACE(0, i * 100 + 5);                                   // K block 1 load issued
        load_kv(sK, &k_full[1], 1, VA_H + h);
        load_q(1);
        if (i > 0) store_tile(w - (int)gridDim.x, i - 1, 0);     // previous item's tile 0 (staged around now)
        mbar_wait(&v_free[1], ph ^ 1);
        TRACE(0, i * 100 + 6);                                   // V block 1 load issued
        load_kv(sV, &v_full[1], 1, 2 * VA_H + h);
        if (i > 0) store_tile(w - (int)gridDim.x, i - 1, 1);
      }
      if (i > 0) {                                               // the last item's tiles
        const int wl = blockIdx.x + (i - 1) * (int)gridDim.x;
        store_tile(wl, i - 1, 0);
        store_tile(wl, i - 1, 1);
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // every store has fully completed before the CTA may exit
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    if (lane == 0) {
      TRACE_DECL
      constexpr uint32_t idesc_s144 = umma_idesc_bf16(128, 144);
      constexpr uint32_t idesc_s128 = umma_idesc_bf16(128, 128);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 96) | (1u << 16);     // B (= V) is MN-major
      const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK), v_addr = smem_u32(sV);
      auto issue_s = [&](int t, int blk) {                      // S_t = Q_t K_blk^T, then publish it
        const uint32_t d = tmem_base + PP_COL_S + t * PP_S_STRIDE;
        uint32_t accum = 0;
#pragma unroll
        for (int db = 0; db < 2; ++db) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            if (db == 1 && ks >= 2) break;                       // d = 96..127 is all zero
            const uint64_t da = umma_desc_sw128(q_addr + (t * 2 + db) * PP_Q_BYTES + ks * 32);
            const uint64_t dk = umma_desc_sw128(k_addr + db * VA_KBLK_BYTES + blk * PP_BLK0_BYTES + ks * 32);
            umma_bf16_ss(d, da, dk, blk == 0 ? idesc_s144 : idesc_s128, accum);
            accum = 1;
          }
        }
        umma_commit(&s_full[t]);
      };
      auto issue_pv = [&](int t, int blk) {                     // O_t (+)= P_t V_blk
        const uint32_t d = tmem_base + PP_COL_O + t * PP_O_STRIDE;
        const uint32_t a = tmem_base + PP_COL_S + t * PP_S_STRIDE;
        const int kk0 = blk == 0 ? 0 : PP_KEYS0 / 16, kk1 = blk == 0 ? PP_KEYS0 / 16 : 17;
        for (int kk = kk0; kk < kk1; ++kk) {
          const uint64_t db = umma_desc_mn_sw128(v_addr + kk * 2048, VA_KBLK_BYTES, 1024);
          uint32_t pcol = (kk - kk0) * 8;                        // bf16 P of 16 keys = 8 columns
          if (HALVES == 2) {                                     // half B's P sits behind ITS score columns
            const int split = (blk == 0 ? PP_SPLIT0 : PP_SPLIT1) / 16;
            if (kk - kk0 >= split) pcol = (blk == 0 ? PP_SPLIT0 : PP_SPLIT1) + (kk - kk0 - split) * 8;
          }
          umma_bf16_ts(d, a + pcol, db, idesc_pv, (blk > 0 || kk > kk0) ? 1u : 0u);
        }
      };
      int i = 0;
      if (blockIdx.x < p.num_items) {
        mbar_wait(&k_full[0], 0);
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&q_full[t], 0);
          tc_fence_after();
          issue_s(t, 0);
        }
        umma_commit(&k_free[0]);
      }
      for (int w = blockIdx.x; w < p.num_items; w += gridDim.x, ++i) {
        const uint32_t ph = i & 1;
        const bool has_next = w + (int)gridDim.x < p.num_items;
#pragma unroll 1
        for (int t = 0; t < 2; ++t) {                            // key block 0 done by the softmax warps -> P V, next S
          mbar_wait(&p_full[t], 0);
          TRACE(1, i * 100 + 10 + t);                            // P(t,0) seen
          if (t == 0) { mbar_wait(&v_full[0], ph); mbar_wait(&k_full[1], ph); }
          mbar_wait(&o_free[t], ph ^ 1);                         // the previous item's epilogue has read O_t
          TRACE(1, i * 100 + 12 + t);                            // V / K1 / O_t available -> issue
          tc_fence_after();
          issue_pv(t, 0);
          issue_s(t, 1);
          umma_commit(&q_free[t]);
          if (t == 1) { umma_commit(&v_free[0]); umma_commit(&k_free[1]); }
        }
#pragma unroll 1
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&p_full[t], 1);
          TRACE(1, i * 100 + 20 + t);                            // P(t,1) seen
          if (t == 0) mbar_wait(&v_full[1], ph);
          tc_fence_after();
          issue_pv(t, 1);
          umma_commit(&o_full[t]);
          if (t == 1) umma_commit(&v_free[1]);
          if (has_next) {
            if (t == 0) mbar_wait(&k_full[0], ph ^ 1);
            mbar_wait(&q_full[t], ph ^ 1);
            TRACE(1, i * 100 + 22 + t);                          // next item's K0 / Q_t landed -> issue S(t,0)
            tc_fence_after();
            issue_s(t, 0);
            if (t == 1) umma_commit(&k_free[0]);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp < 4) {
    // ================================================================ CLS query row: warp 2 key block 0, warp 3 key block 1
    const uint32_t sK32 = smem_u32(sK), sV32 = smem_u32(sV);
    float* sc_s = cls_scores + (warp - 2) * 144;
    float* xch = cls_scores + 2 * 144;
    TRACE_DECL
    int i = 0;
    for (int w = blockIdx.x; w < p.num_items; w += gridDim.x, ++i) {
      if (lane == 0) TRACE(warp, i * 100 + 30);
      if (warp == 2) cls_row_block<0>(p, w, i, sK32, sV32, k_full, k_free, v_full, v_free, sc_s, xch, lane);
      else cls_row_block<1>(p, w, i, sK32, sV32, k_full, k_free, v_full, v_free, sc_s, xch, lane);
      if (lane == 0) TRACE(warp, i * 100 + 31);
    }
  } else {
    // ================================================================ softmax warpgroups: one thread per query row
    // Streaming two-pass softmax over 32-column chunks of S_t: pass 1 finds the row max, pass 2 re-reads S (tensor
    // memory reads are cheap: ~68 B/clk per warp), exponentiates and writes the packed bf16 P chunk behind the read
    // position (P aliases the first half of S).  Each tcgen05.ld is issued one chunk ahead of its use; at most two
    // chunks are live, so nothing spills.
    const int half = (HALVES == 2 && warp >= 12) ? 1 : 0;   // which columns of the row this thread owns (split-row form)
    const int t = ((warp - 4) & 7) >> 2;         // query tile of this warpgroup (warps 4-7 / 12-15: tile 0, 8-11 / 16-19: tile 1)
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t s_addr = lane_base + PP_COL_S + t * PP_S_STRIDE;
    const uint32_t o_addr = lane_base + PP_COL_O + t * PP_O_STRIDE;
    const uint32_t l_addr = lane_base + PP_COL_L + 2 * t;
    float* xm_mine = xmax + t * 256 + half * 128 + row;
    const float* xm_other = xmax + t * 256 + (half ^ 1) * 128 + row;
    const float sl2 = p.scale_log2;
    TRACE_DECL
#ifdef TGB_ATTN_TRACE
    const bool tr_on = (quarter == 0 && lane == 0 && half == 0);
#define STRACE(tag) do { if (tr_on) TRACE(4 + t, tag); } while (0)
#else
#define STRACE(tag) do { } while (0)
#endif
    // row max over both halves of the row: own part -> shared memory -> the tile's named barrier -> the other part.
    // One slot per (tile, half, row) is enough: between two uses of a slot lies an mbarrier chain through the MMA
    // thread (the reader's P-published arrival -> the next S of this tile -> the writer's wait on it).
    auto merged_max = [&](float part) -> float {
      *xm_mine = part;
      named_bar_sync(4 + t, 256);
      return fmaxf(part, *xm_other);
    };
    int i = 0;
    for (int w = blockIdx.x; w < p.num_items; w += gridDim.x, ++i) {
      float mx, l;
      {
        // ------------------------------------------------------------ key block 0: keys [0,144), all valid
        mbar_wait(&s_full[t], 0);
        STRACE(i * 100 + 40);                                    // S(t,0) seen
        tc_fence_after();
        if (HALVES == 1) mx = pp_row_max<PP_KEYS0, PP_KEYS0>(s_addr) * sl2;
        else if (half == 0) mx = merged_max(pp_row_max_part<PP_SPLIT0, PP_SPLIT0>(s_addr)) * sl2;
        else mx = merged_max(pp_row_max_part<PP_KEYS0 - PP_SPLIT0, PP_KEYS0 - PP_SPLIT0>(s_addr + PP_SPLIT0)) * sl2;
        STRACE(i * 100 + 41);                                    // row max known
        if constexpr (SP) {
          if (HALVES == 1) l = pp_exp_block_sp<PP_KEYS0 / 16, 16>(s_addr, sl2, -mx);
          else if (half == 0) l = pp_exp_block_sp<PP_SPLIT0 / 16, 16>(s_addr, sl2, -mx);
          else l = pp_exp_block_sp<(PP_KEYS0 - PP_SPLIT0) / 16, 16>(s_addr + PP_SPLIT0, sl2, -mx);
        } else {
          if (HALVES == 1) l = pp_exp_block<PP_KEYS0 / 16 - 1, 16>(s_addr, sl2, -mx);
          else if (half == 0) l = pp_exp_block<PP_SPLIT0 / 16 - 1, 16>(s_addr, sl2, -mx);
          else l = pp_exp_block<(PP_KEYS0 - PP_SPLIT0) / 16 - 1, 16>(s_addr + PP_SPLIT0, sl2, -mx);
        }
        STRACE(i * 100 + 42);                                    // exponentials done
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_full[t]);
        STRACE(i * 100 + 43);                                    // P(t,0) published
      }
      {
        // ------------------------------------------------------------ key block 1: keys [144,272), 113 valid
        mbar_wait(&s_full[t], 1);                                 // also implies P V of block 0 has completed
        STRACE(i * 100 + 50);                                    // S(t,1) seen
        tc_fence_after();
        float m1x;
        if (HALVES == 1) m1x = pp_row_max<128, PP_VALID1>(s_addr) * sl2;
        else if (half == 0) m1x = merged_max(pp_row_max_part<PP_SPLIT1, PP_SPLIT1>(s_addr)) * sl2;
        else m1x = merged_max(pp_row_max_part<128 - PP_SPLIT1, PP_VALID1 - PP_SPLIT1>(s_addr + PP_SPLIT1)) * sl2;
        STRACE(i * 100 + 51);
        float f = 1.f;
        const bool need = m1x > mx + PP_RESCALE_THRESHOLD;       // (same decision in both halves of a row: same operands)
        if (need) {
          f = ex2_approx(mx - m1x);
          mx = m1x;
          l *= f;
        }
        if (__any_sync(0xffffffffu, need)) {                     // rare: rescale this warp's 32 rows of O_t in place
          const int c0 = (HALVES == 2 && half == 1) ? 48 : 0, c1 = (HALVES == 2 && half == 0) ? 48 : 96;
#pragma unroll 1
          for (int c = c0; c < c1; c += 8) {
            uint32_t o[8];
            tmem_ld8<0>(o_addr + c, o);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * f);
            tmem_st8<0>(o_addr + c, o);
          }
          tmem_st_wait();
        }
        static_assert(PP_VALID1 % 16 != 0 && (PP_VALID1 - PP_SPLIT1) % 16 != 0, "the last chunk of key block 1 is partial");
        if constexpr (SP) {
          if (HALVES == 1) l += pp_exp_block_sp<PP_VALID1 / 16 + 1, PP_VALID1 % 16>(s_addr, sl2, -mx);
          else if (half == 0) l += pp_exp_block_sp<PP_SPLIT1 / 16, 16>(s_addr, sl2, -mx);
          else l += pp_exp_block_sp<(PP_VALID1 - PP_SPLIT1) / 16 + 1, (PP_VALID1 - PP_SPLIT1) % 16>(s_addr + PP_SPLIT1, sl2, -mx);
        } else {
          if (HALVES == 1) l += pp_exp_block<PP_VALID1 / 16, PP_VALID1 % 16>(s_addr, sl2, -mx);
          else if (half == 0) l += pp_exp_block<PP_SPLIT1 / 16 - 1, 16>(s_addr, sl2, -mx);
          else l += pp_exp_block<(PP_VALID1 - PP_SPLIT1) / 16, (PP_VALID1 - PP_SPLIT1) % 16>(s_addr + PP_SPLIT1, sl2, -mx);
        }
        if (HALVES == 2) tmem_st1(l_addr + half, __float_as_uint(l));   // this half's partial row sum
        STRACE(i * 100 + 52);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_full[t]);
        STRACE(i * 100 + 53);                                    // P(t,1) published
      }
      {
        // ------------------------------------------------------------ epilogue: O_t / l -> bf16 -> staging tile -> TMA store
        // Per-row 16-byte global stores touch 32 cache lines per instruction and kept the LSU (and with it the tensor
        // memory traffic of the other warpgroup) busy for ~3k cycles; instead each thread writes its row into a dense
        // [128][88] shared tile (conflict-free: 176-byte rows) and the TMA producer thread issues an asynchronous bulk
        // tensor store (waiting for the store to drain inside this warpgroup would sit on the critical chain).
        mbar_wait(&o_full[t], i & 1);
        STRACE(i * 100 + 60);                                    // O_t complete
        tc_fence_after();
        constexpr int NO = HALVES == 1 ? 96 : 48;                // O columns held by one thread
        uint32_t o[NO];
        uint32_t la = 0, lb = 0;
        int c_first, c_count;                                    // 16-byte d chunks of the row this thread writes
        if constexpr (HALVES == 1) {
          tmem_ld32<0>(o_addr, o);
          tmem_ld32<32>(o_addr + 32, o);
          tmem_ld32<64>(o_addr + 64, o);
          c_first = 0; c_count = 11;
        } else {
          tmem_ld2(l_addr, la, lb);
          if (half == 0) {                                       // d [0,40)
            tmem_ld32<0>(o_addr, o);
            tmem_ld8<32>(o_addr + 32, o);
            c_first = 0; c_count = 5;
          } else {                                               // d [40,88)
            tmem_ld32<0>(o_addr + 40, o);
            tmem_ld16<32>(o_addr + 72, o);
            c_first = 5; c_count = 6;
          }
        }
        tmem_ld_wait();
        STRACE(i * 100 + 61);                                    // O_t in registers
        tc_fence_before();
        mbar_arrive(&o_free[t]);
        if (HALVES == 2) l = __uint_as_float(la) + __uint_as_float(lb);   // (A's part + B's part in both halves)
        const float inv = 1.f / l;
        const int use = 2 * i + t;                               // uses of the staging tile alternate between the tiles
        mbar_wait(stage_free, (use & 1) ^ 1);                    // use - 1 has been read out
        uint4* srow = reinterpret_cast<uint4*>(stage + row * (VA_D * 2)) + c_first;
#pragma unroll
        for (int c = 0; c < (HALVES == 1 ? 11 : 6); ++c) {       // 88 valid d columns = 11 x 16 bytes
          if (c < c_count) {
            const float2 i2 = make_float2(inv, inv);            // packed fp32 multiplies (FMUL2)
            const float2 a = __fmul2_rn(make_float2(__uint_as_float(o[8 * c + 0]), __uint_as_float(o[8 * c + 1])), i2);
            const float2 b = __fmul2_rn(make_float2(__uint_as_float(o[8 * c + 2]), __uint_as_float(o[8 * c + 3])), i2);
            const float2 d = __fmul2_rn(make_float2(__uint_as_float(o[8 * c + 4]), __uint_as_float(o[8 * c + 5])), i2);
            const float2 e = __fmul2_rn(make_float2(__uint_as_float(o[8 * c + 6]), __uint_as_float(o[8 * c + 7])), i2);
            uint4 v;
            v.x = pack_bf16(a.x, a.y);
            v.y = pack_bf16(b.x, b.y);
            v.z = pack_bf16(d.x, d.y);
            v.w = pack_bf16(e.x, e.y);
            srow[c] = v;
          }
        }
        fence_proxy_async();                                     // generic-proxy writes -> visible to the TMA (async proxy)
        mbar_arrive(stage_ready);                                // the TMA producer thread issues the bulk store
        STRACE(i * 100 + 62);                                    // stores issued
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// =====================================================================================================================
// Three-block variant (cross-check / experiment).  Same roles as the ping-pong kernel, re-cut so that no warpgroup ever waits for a
// tensor-core round trip it does not depend on:
//   keys   three blocks of 96 / 96 / 80 rows (65 valid in the last); a "unit" is (query tile t, key block kb), units of
//          an item run in the order r = 2*kb + t:  (0,0) (1,0) (0,1) (1,1) (0,2) (1,2)
//   TMEM   three S/P buffers of 96 columns used as a RING over the units (unit u -> buffer u % 3; the bf16 P aliases
//          the first 48 columns) | O_t fp32 [288 + 96 t, +96).  S of unit u+3 is issued right behind P V of unit u --
//          the tensor pipe executes in order, so the write of S(u+3) cannot pass the read of P(u) -- which puts the S
//          of a warpgroup's NEXT unit in flight one whole unit ahead: when a softmax warpgroup publishes P it finds its
//          next score block already waiting (the ping-pong kernel idled ~1.9k of 9.4k clk per item there).
//   warps  0 TMEM alloc + TMA producer | 1 MMA issuer | 2 / 3 CLS query row (key blocks 0 + 2 / block 1) |
//          4-7 / 8-11 softmax warpgroups of tile 0 / 1 (one thread per query row) |
//          12-15 epilogue warpgroup: O_t / l -> bf16 -> staging tile -> bulk tensor store, for both tiles -- the
//          softmax warpgroups go straight on to the next item instead of waiting for their last P V.
//   online softmax over three blocks with the conditional accumulator rescale of the ping-pong kernel; the (rare)
//   rescale first waits for the tile's previous P V on its own barrier (pv_done), because S(u) no longer implies it.
// =====================================================================================================================
constexpr int P3_THREADS = 16 * 32;
constexpr int P3_KB = 96;                          // keys per block (blocks 0 and 1); block 2 = keys [192, 272)
constexpr int P3_KB2 = 272 - 2 * P3_KB;            // 80
constexpr int P3_VALID2 = VA_S - 2 * P3_KB;        // 65 valid keys in block 2
constexpr uint32_t P3_COL_S = 0, P3_S_STRIDE = 96, P3_COL_O = 288, P3_O_STRIDE = 96;
constexpr int P3_SMEM_Q = 0;
constexpr int P3_SMEM_K = 4 * PP_Q_BYTES;
constexpr int P3_SMEM_V = P3_SMEM_K + 2 * VA_KBLK_BYTES;
constexpr int P3_SMEM_BAR = P3_SMEM_V + 2 * VA_KBLK_BYTES;       // 204800: 30 mbarriers + the TMEM slot
constexpr int P3_SMEM_CLS = P3_SMEM_BAR + 256;                   // float [2 warps][96] scores | float [100] exchange
constexpr int P3_SMEM_L = P3_SMEM_CLS + (2 * 96 + 100) * 4;  // float [2 tiles][128] row sums for the epilogue warpgroup
constexpr int P3_SMEM_STAGE = (P3_SMEM_L + 2 * 128 * 4 + 127) / 128 * 128;
constexpr int P3_SMEM_TOTAL = P3_SMEM_STAGE + PP_STAGE_BYTES + 1024;     // + alignment slack
static_assert(P3_SMEM_TOTAL <= 227 * 1024, "shared memory budget");
constexpr int P3_BLK_BYTES = P3_KB * 128;                         // one d block of key block 0 / 1 (12288 = 12 swizzle atoms)
constexpr int P3_BLK2_BYTES = P3_KB2 * 128;


template <int NCOLS, int NVALID>
__device__ __forceinline__ float p3_row_max(uint32_t s_addr) {
  static_assert(NCOLS == 96 || NCOLS == 80, "two key-block widths");
  uint32_t v[NCOLS];
  tmem_ld32<0>(s_addr, v);
  tmem_ld32<32>(s_addr + 32, v);
  if (NCOLS == 96) tmem_ld32<64>(s_addr + 64, v);
  else tmem_ld16<64>(s_addr + 64, v);
  tmem_ld_wait();
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
  for (int j = 0; j < NVALID; j += 4) {
    m0 = fmaxf(m0, __uint_as_float(v[j]));
    if (j + 1 < NVALID) m1 = fmaxf(m1, __uint_as_float(v[j + 1]));
    if (j + 2 < NVALID) m2 = fmaxf(m2, __uint_as_float(v[j + 2]));
    if (j + 3 < NVALID) m3 = fmaxf(m3, __uint_as_float(v[j + 3]));
  }
  return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}

// CLS query row over key block BLK of three (see cls_row_block for the formulation): unnormalised O^T in `o`, the block's
// own max `m` (log2 units) and row sum `l`.  qb = the B fragments of q.
template <int BLK>
__device__ __forceinline__ void cls3_part(const VitAttnParams& p, int i, uint32_t sK, uint32_t sV, uint64_t* k_full,
                                          uint64_t* k_free, uint64_t* v_full, uint64_t* v_free, float* sc_s, int lane,
                                          const uint32_t (&qb)[6][2], float (&o)[6][4], float& m, float& l) {
  constexpr int NT = BLK < 2 ? P3_KB / 16 : P3_KB2 / 16;                   // 16-key tiles: 6 / 6 / 5
  constexpr int KEY0 = BLK * P3_KB;
  constexpr int NVALID = BLK < 2 ? P3_KB : P3_VALID2;                      // 96 / 96 / 65
  const uint32_t ph = i & 1;
  const int g = lane >> 2, t = lane & 3;
  const int lrow = lane & 7, lmat = lane >> 3;
  mbar_wait(&k_full[BLK], ph);
  auto qk_pair = [&](int mt, bool second) {
    uint32_t a0[6][4], a1[6][4];
    const int key = KEY0 + mt * 16 + (lmat & 1) * 8 + lrow;
#pragma unroll
    for (int ks = 0; ks < 6; ++ks) va_ldsm_x4(sK + va_tile_off(key, 2 * ks + (lmat >> 1)), a0[ks]);
    if (second) {
#pragma unroll
      for (int ks = 0; ks < 6; ++ks) va_ldsm_x4(sK + va_tile_off(key + 16, 2 * ks + (lmat >> 1)), a1[ks]);
    }
    float c0e[4] = {0.f, 0.f, 0.f, 0.f}, c0o[4] = {0.f, 0.f, 0.f, 0.f}, c1e[4] = {0.f, 0.f, 0.f, 0.f}, c1o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ks = 0; ks < 6; ks += 2) {
      va_mma_16816_nv(c0e, a0[ks], qb[ks][0], qb[ks][1]);
      va_mma_16816_nv(c0o, a0[ks + 1], qb[ks + 1][0], qb[ks + 1][1]);
      if (second) {
        va_mma_16816_nv(c1e, a1[ks], qb[ks][0], qb[ks][1]);
        va_mma_16816_nv(c1o, a1[ks + 1], qb[ks + 1][0], qb[ks + 1][1]);
      }
    }
    if (t == 0) {
      sc_s[mt * 16 + g] = c0e[0] + c0o[0];
      sc_s[mt * 16 + 8 + g] = c0e[2] + c0o[2];
      if (second) {
        sc_s[mt * 16 + 16 + g] = c1e[0] + c1o[0];
        sc_s[mt * 16 + 24 + g] = c1e[2] + c1o[2];
      }
    }
  };
#pragma unroll 1
  for (int mt = 0; mt + 1 < NT; mt += 2) qk_pair(mt, true);
  if (NT & 1) qk_pair(NT - 1, false);
  __syncwarp();
  if (lane == 0) mbar_arrive(&k_free[BLK]);                      // done reading this K block (second arrival: tcgen05.commit)
  float x[3];
  m = -INFINITY;
#pragma unroll
  for (int it = 0; it < 3; ++it) {
    const int j = it * 32 + lane;
    x[it] = (j < NVALID) ? sc_s[j] * p.scale_log2 : -INFINITY;
    m = fmaxf(m, x[it]);
  }
  m = warp_max(m);
  l = 0.f;
#pragma unroll
  for (int it = 0; it < 3; ++it) {
    x[it] = ex2_approx(x[it] - m);                               // exp2(-inf) = 0 for masked keys
    l += x[it];
  }
  l = warp_sum(l);
  __syncwarp();                                                  // every lane has read its scores before p overwrites them
  __nv_bfloat16* p_s = reinterpret_cast<__nv_bfloat16*>(sc_s);
#pragma unroll
  for (int it = 0; it < 3; ++it)
    if (it * 32 + lane < NT * 16) p_s[it * 32 + lane] = __float2bfloat16(x[it]);
  __syncwarp();
  mbar_wait(&v_full[BLK], ph);
#pragma unroll
  for (int mt = 0; mt < 6; ++mt) { o[mt][0] = o[mt][1] = o[mt][2] = o[mt][3] = 0.f; }
  const uint32_t* p32 = reinterpret_cast<const uint32_t*>(p_s);
#pragma unroll 1
  for (int kk = 0; kk < NT; ++kk) {
    const uint32_t b0 = (g == 0) ? p32[kk * 8 + t] : 0u;
    const uint32_t b1 = (g == 0) ? p32[kk * 8 + 4 + t] : 0u;
    const int key = KEY0 + kk * 16 + (lmat >> 1) * 8 + lrow;
    uint32_t a[6][4];
#pragma unroll
    for (int mt = 0; mt < 6; ++mt) va_ldsm_x4_t(sV + va_tile_off(key, 2 * mt + (lmat & 1)), a[mt]);
#pragma unroll
    for (int mt = 0; mt < 6; ++mt) va_mma_16816_nv(o[mt], a[mt], b0, b1);
  }
  __syncwarp();                                                  // (also: every lane is done with sc_s before its next use)
  if (lane == 0) mbar_arrive(&v_free[BLK]);                      // done reading this V block
}

// The CLS row of one item on two warps: HALF 0 takes key blocks 0 and 2 (merged in registers) and writes the result,
// HALF 1 takes block 1 and hands (O, m, l) over through `xch` (float[100]; two barriers per item guard its reuse).
template <int HALF>
__device__ __forceinline__ void cls3_row(const VitAttnParams& p, int w, int i, uint32_t sK, uint32_t sV, uint64_t* k_full,
                                         uint64_t* k_free, uint64_t* v_full, uint64_t* v_free, float* sc_s, float* xch,
                                         int lane) {
  const int b = w / VA_H, h = w - b * VA_H;
  const long long row0 = (long long)b * VA_S;
  const int g = lane >> 2, t = lane & 3;
  // B fragments of q: b0 = d (16ks + 2t, +1), b1 = d (16ks + 8 + 2t, +1), column n = g -> only g == 0 is the real query
  uint32_t qb[6][2];
  {
    const uint32_t* q32 = reinterpret_cast<const uint32_t*>(p.qkv + row0 * p.ld_qkv + h * VA_D);
#pragma unroll
    for (int ks = 0; ks < 6; ++ks) {
      qb[ks][0] = (g == 0) ? __ldg(q32 + ks * 8 + t) : 0u;
      qb[ks][1] = (g == 0 && ks < 5) ? __ldg(q32 + ks * 8 + 4 + t) : 0u;      // d 88..95 does not exist
    }
  }
  float o[6][4], m, l;
  if (HALF == 1) {
    cls3_part<1>(p, i, sK, sV, k_full, k_free, v_full, v_free, sc_s, lane, qb, o, m, l);
    if (t == 0) {
#pragma unroll
      for (int mt = 0; mt < 6; ++mt) { xch[mt * 16 + g] = o[mt][0]; xch[mt * 16 + 8 + g] = o[mt][2]; }
    }
    if (lane == 0) { xch[96] = m; xch[97] = l; }
    __syncwarp();
    named_bar_sync(3, 64);                                       // published
    named_bar_sync(3, 64);                                       // consumed
  } else {
    cls3_part<0>(p, i, sK, sV, k_full, k_free, v_full, v_free, sc_s, lane, qb, o, m, l);
    float o2[6][4], m2, l2;
    cls3_part<2>(p, i, sK, sV, k_full, k_free, v_full, v_free, sc_s, lane, qb, o2, m2, l2);
    named_bar_sync(3, 64);
    const float m1 = xch[96], l1 = xch[97];
    const float mm = fmaxf(m, fmaxf(m1, m2));
    const float f0 = ex2_approx(m - mm), f1 = ex2_approx(m1 - mm), f2 = ex2_approx(m2 - mm);
    const float inv = 1.f / (l * f0 + l1 * f1 + l2 * f2);
    if (t == 0) {
      __nv_bfloat16* dst = p.out + row0 * p.ld_out + h * VA_D;
#pragma unroll
      for (int mt = 0; mt < 6; ++mt) {
        dst[mt * 16 + g] = __float2bfloat16((o[mt][0] * f0 + xch[mt * 16 + g] * f1 + o2[mt][0] * f2) * inv);
        if (mt < 5)
          dst[mt * 16 + 8 + g] = __float2bfloat16((o[mt][2] * f0 + xch[mt * 16 + 8 + g] * f1 + o2[mt][2] * f2) * inv);
      }
    }
    __syncwarp();
    named_bar_sync(3, 64);
  }
}

// ---- MMA issuer of the three-block kernel.  ONE thread issues ~70 tcgen05.mma + ~22 tcgen05.commit per item, and a single
// thread retires a dependent instruction only every ~4-6 clk: with descriptors rebuilt per instruction (64-bit shifts /
// ors, run-time unit indices) each MMA cost ~150-200 clk of issue time -- more than its 48 clk on the tensor pipe -- and
// the issuer, not the tensor cores or the softmax warps, set the period of both this kernel's first version (13.8k clk per
// item) and of the ping-pong kernel (9.4k).  Here every unit index is a template parameter and the shared-memory
// descriptors are (low word + compile-time constant, constant high word) pairs: ~3 instructions per MMA.
struct P3Issue {
  uint32_t tmem_base;
  uint32_t q_lo, k_lo, v_lo;        // low descriptor words of the Q / K / V buffers' first bytes
  uint32_t hi_k, hi_mn;             // high descriptor words: K-major SWIZZLE_128B (SBO 1024) / MN-major (SBO 1024), LBO in the low word
  uint64_t *q_full, *q_free, *k_full, *k_free, *v_full, *v_free, *s_full, *p_full, *pv_done, *o_full, *o_free;
};
__device__ __forceinline__ uint64_t p3_desc(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}
// S of unit R (= 2 kb + t) of item parity `ph` into ring buffer R % 3; waits for the operands it is the first to use.
template <int R>
__device__ __forceinline__ void p3_issue_s(const P3Issue& c, uint32_t ph) {
  constexpr int t = R & 1, kb = R >> 1;
  constexpr uint32_t idesc = umma_idesc_bf16(128, kb < 2 ? P3_KB : P3_KB2);
  if (kb == 0) mbar_wait(&c.q_full[t], ph);
  if (t == 0) mbar_wait(&c.k_full[kb], ph);
  tc_fence_after();
  const uint32_t d = c.tmem_base + P3_COL_S + (R % 3) * P3_S_STRIDE;
#pragma unroll
  for (int ks6 = 0; ks6 < 6; ++ks6) {                            // d block 0: four k-steps, d block 1: two (d 96..127 is zero)
    const int db = ks6 >> 2, ks = ks6 & 3;
    const uint64_t da = p3_desc(c.q_lo + (((t * 2 + db) * PP_Q_BYTES + ks * 32) >> 4), c.hi_k);
    const uint64_t dk = p3_desc(c.k_lo + ((db * VA_KBLK_BYTES + kb * P3_BLK_BYTES + ks * 32) >> 4), c.hi_k);
    umma_bf16_ss(d, da, dk, idesc, ks6 > 0 ? 1u : 0u);
  }
  umma_commit(&c.s_full[R % 3]);
  if (t == 1) umma_commit(&c.k_free[kb]);
  if (kb == 2) umma_commit(&c.q_free[t]);
}
// O_t (+)= P V of unit R; waits for P, the V block (first use) and, for the item's first block, the epilogue of O_t.
template <int R>
__device__ __forceinline__ void p3_issue_pv(const P3Issue& c, uint32_t ph) {
  constexpr int t = R & 1, kb = R >> 1;
  constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 96) | (1u << 16);       // B (= V) is MN-major
  constexpr int nsteps = kb < 2 ? P3_KB / 16 : P3_KB2 / 16;
  mbar_wait(&c.p_full[R % 3], R >= 3 ? 1u : 0u);
  if (t == 0) mbar_wait(&c.v_full[kb], ph);
  if (kb == 0) mbar_wait(&c.o_free[t], ph ^ 1);                  // the previous item's epilogue has read O_t
  tc_fence_after();
  const uint32_t d = c.tmem_base + P3_COL_O + t * P3_O_STRIDE;
  const uint32_t a = c.tmem_base + P3_COL_S + (R % 3) * P3_S_STRIDE;
#pragma unroll
  for (int j = 0; j < nsteps; ++j)
    umma_bf16_ts(d, a + j * 8, p3_desc(c.v_lo + (((kb * (P3_KB / 16) + j) * 2048) >> 4), c.hi_mn), idesc_pv,
                 (kb > 0 || j > 0) ? 1u : 0u);
  umma_commit(&c.pv_done[t]);
  if (kb == 2) umma_commit(&c.o_full[t]);
  if (t == 1) umma_commit(&c.v_free[kb]);
}
// unit R of item i: P V(R), then S of the unit three places on (same item for R < 3, the next item otherwise)
template <int R>
__device__ __forceinline__ void p3_unit(const P3Issue& c, uint32_t ph, bool has_next) {
  p3_issue_pv<R>(c, ph);
  if constexpr (R < 3) {
    p3_issue_s<R + 3>(c, ph);
  } else {
    if (has_next) p3_issue_s<R - 3>(c, ph ^ 1);
  }
}

__global__ void __launch_bounds__(P3_THREADS, 1)
vit_attention_p3_kernel(const __grid_constant__ CUtensorMap tmap96, const __grid_constant__ CUtensorMap tmap80,
                        const __grid_constant__ CUtensorMap tmap128, const __grid_constant__ CUtensorMap tmap_out,
                        const VitAttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + P3_SMEM_Q;
  uint8_t* sK = smem + P3_SMEM_K;
  uint8_t* sV = smem + P3_SMEM_V;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P3_SMEM_BAR);
  uint64_t* q_full = bars + 0;      // [2] per query tile, one completion per item
  uint64_t* q_free = bars + 2;      // [2]
  uint64_t* k_full = bars + 4;      // [3] per key block
  uint64_t* k_free = bars + 7;      // [3] tcgen05.commit + the block's CLS warp
  uint64_t* v_full = bars + 10;     // [3]
  uint64_t* v_free = bars + 13;     // [3]
  uint64_t* s_full = bars + 16;     // [3] per S/P ring buffer, two completions per item
  uint64_t* p_full = bars + 19;     // [3] 128 arrivals
  uint64_t* pv_done = bars + 22;    // [2] per tile, one completion per unit (three per item)
  uint64_t* o_full = bars + 24;     // [2] per tile, one completion per item
  uint64_t* o_free = bars + 26;     // [2] 128 arrivals (epilogue warpgroup)
  uint64_t* l_full = bars + 28;     // [2] 128 arrivals: the tile's row sums are in shared memory
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 30);
  float* cls_scores = reinterpret_cast<float*>(smem + P3_SMEM_CLS);
  float* l_s = reinterpret_cast<float*>(smem + P3_SMEM_L);
  uint8_t* stage = smem + P3_SMEM_STAGE;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_my = p.num_items > (int)blockIdx.x ? (p.num_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == 1 && lane == 0) {
    tma_prefetch_desc(&tmap96);
    tma_prefetch_desc(&tmap80);
    tma_prefetch_desc(&tmap128);
    tma_prefetch_desc(&tmap_out);
    for (int t = 0; t < 2; ++t) {
      mbar_init(&q_full[t], 1); mbar_init(&q_free[t], 1);
      mbar_init(&pv_done[t], 1); mbar_init(&o_full[t], 1);
      mbar_init(&o_free[t], 128); mbar_init(&l_full[t], 128);
    }
    for (int k = 0; k < 3; ++k) {
      mbar_init(&k_full[k], 1); mbar_init(&k_free[k], 2);
      mbar_init(&v_full[k], 1); mbar_init(&v_free[k], 2);
      mbar_init(&s_full[k], 1); mbar_init(&p_full[k], 128);
    }
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================================================ TMA producer
    // Per item, loads are issued in the order the previous item releases their buffers:
    //   K0 K1 V0 Q0 K2 Q1 V1 V2   (K_kb after S(1,kb), V_kb after P V(1,kb), Q_t after S(t,2); S(u+3) rides behind P V(u))
    if (lane == 0) {
      TRACE_DECL
      for (int i = 0; i < n_my; ++i) {
        const int w = (int)blockIdx.x + i * (int)gridDim.x;
        const int b = w / VA_H, h = w - b * VA_H;
        const int row0 = b * VA_S;
        const uint32_t ph = i & 1;
        auto load_kv = [&](uint8_t* dst, uint64_t* full, uint64_t* free_, int kb, int col_h) {
          mbar_wait(&free_[kb], ph ^ 1);
          mbar_arrive_expect_tx(&full[kb], 2 * (kb < 2 ? P3_BLK_BYTES : P3_BLK2_BYTES));
#pragma unroll
          for (int db = 0; db < 2; ++db)
            tma_load_3d(dst + db * VA_KBLK_BYTES + kb * P3_BLK_BYTES, kb < 2 ? &tmap96 : &tmap80, &full[kb], db * 64, col_h,
                        row0 + kb * P3_KB);
        };
        auto load_q = [&](int t) {
          mbar_wait(&q_free[t], ph ^ 1);
          mbar_arrive_expect_tx(&q_full[t], 2 * PP_Q_BYTES);
#pragma unroll
          for (int db = 0; db < 2; ++db)
            tma_load_3d(sQ + (t * 2 + db) * PP_Q_BYTES, &tmap128, &q_full[t], db * 64, h, row0 + 1 + t * 128);
        };
        TRACE(0, i * 100 + 1);
        load_kv(sK, k_full, k_free, 0, VA_H + h);
        if (i == 0) load_q(0);                                   // the very first item: Q0 next to K0 so that S(0,0) can start
        load_kv(sK, k_full, k_free, 1, VA_H + h);
        load_kv(sV, v_full, v_free, 0, 2 * VA_H + h);
        if (i != 0) load_q(0);
        TRACE(0, i * 100 + 2);
        load_kv(sK, k_full, k_free, 2, VA_H + h);
        load_q(1);
        load_kv(sV, v_full, v_free, 1, 2 * VA_H + h);
        load_kv(sV, v_full, v_free, 2, 2 * VA_H + h);
        TRACE(0, i * 100 + 3);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    if (lane == 0) {
      TRACE_DECL
      P3Issue c;
      c.tmem_base = tmem_base;
      c.q_lo = ((smem_u32(sQ) & 0x3FFFFu) >> 4) | (1u << 16);                      // K-major: LBO field unused (1)
      c.k_lo = ((smem_u32(sK) & 0x3FFFFu) >> 4) | (1u << 16);
      c.v_lo = ((smem_u32(sV) & 0x3FFFFu) >> 4) | ((uint32_t)(VA_KBLK_BYTES >> 4) << 16);   // MN-major: LBO = next 64-wide d block
      c.hi_k = (1024u >> 4) | (1u << 14) | (2u << 29);                              // SBO 1024 | version | SWIZZLE_128B
      c.hi_mn = (1024u >> 4) | (1u << 14) | (2u << 29);
      c.q_full = q_full; c.q_free = q_free; c.k_full = k_full; c.k_free = k_free; c.v_full = v_full; c.v_free = v_free;
      c.s_full = s_full; c.p_full = p_full; c.pv_done = pv_done; c.o_full = o_full; c.o_free = o_free;
      if (n_my > 0) {
        p3_issue_s<0>(c, 0);
        p3_issue_s<1>(c, 0);
        p3_issue_s<2>(c, 0);
      }
#pragma unroll 1
      for (int i = 0; i < n_my; ++i) {
        const uint32_t ph = i & 1;
        const bool has_next = i + 1 < n_my;
        p3_unit<0>(c, ph, has_next);
        TRACE(1, i * 100 + 20);
        p3_unit<1>(c, ph, has_next);
        TRACE(1, i * 100 + 21);
        p3_unit<2>(c, ph, has_next);
        TRACE(1, i * 100 + 22);
        p3_unit<3>(c, ph, has_next);
        TRACE(1, i * 100 + 23);
        p3_unit<4>(c, ph, has_next);
        TRACE(1, i * 100 + 24);
        p3_unit<5>(c, ph, has_next);
        TRACE(1, i * 100 + 25);
      }
    }
    __syncwarp();
  } else if (warp < 4) {
    // ================================================================ CLS query row: warp 2 key blocks 0 and 2, warp 3 block 1
    const uint32_t sK32 = smem_u32(sK), sV32 = smem_u32(sV);
    float* sc_s = cls_scores + (warp - 2) * 96;
    float* xch = cls_scores + 2 * 96;
    for (int i = 0; i < n_my; ++i) {
      const int w = (int)blockIdx.x + i * (int)gridDim.x;
      if (warp == 2) cls3_row<0>(p, w, i, sK32, sV32, k_full, k_free, v_full, v_free, sc_s, xch, lane);
      else cls3_row<1>(p, w, i, sK32, sV32, k_full, k_free, v_full, v_free, sc_s, xch, lane);
    }
  } else if (warp < 12) {
    // ================================================================ softmax warpgroups: one thread per query row
    const int t = (warp - 4) >> 2;               // query tile of this warpgroup
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t o_addr = lane_base + P3_COL_O + t * P3_O_STRIDE;
    const float sl2 = p.scale_log2;
    TRACE_DECL
#ifdef TGB_ATTN_TRACE
    const bool tr_on = (quarter == 0 && lane == 0);
#define STRACE3(tag) do { if (tr_on) TRACE(4 + t, tag); } while (0)
#else
#define STRACE3(tag) do { } while (0)
#endif
    for (int i = 0; i < n_my; ++i) {
      float mx = 0.f, l = 0.f;
#pragma unroll
      for (int kb = 0; kb < 3; ++kb) {
        const int r = 2 * kb + t;
        const int buf = r % 3;
        const uint32_t s_addr = lane_base + P3_COL_S + buf * P3_S_STRIDE;
        mbar_wait(&s_full[buf], r >= 3 ? 1u : 0u);
        STRACE3(i * 100 + 40 + kb * 10);                         // S(t,kb) seen
        tc_fence_after();
        if (kb == 0) {
          mx = p3_row_max<P3_KB, P3_KB>(s_addr) * sl2;
          l = pp_exp_block<P3_KB / 16 - 1, 16>(s_addr, sl2, -mx);
        } else {
          const float m1x = (kb == 1 ? p3_row_max<P3_KB, P3_KB>(s_addr) : p3_row_max<P3_KB2, P3_VALID2>(s_addr)) * sl2;
          float f = 1.f;
          const bool need = m1x > mx + PP_RESCALE_THRESHOLD;
          if (need) {
            f = ex2_approx(mx - m1x);
            mx = m1x;
            l *= f;
          }
          if (__any_sync(0xffffffffu, need)) {                   // rare: rescale this warp's 32 rows of O_t in place
            mbar_wait(&pv_done[t], (uint32_t)(i + kb - 1) & 1u); // P V of the tile's previous block has completed
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < 96; c += 8) {
              uint32_t o[8];
              tmem_ld8<0>(o_addr + c, o);
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * f);
              tmem_st8<0>(o_addr + c, o);
            }
            tmem_st_wait();
          }
          if (kb == 1) l += pp_exp_block<P3_KB / 16 - 1, 16>(s_addr, sl2, -mx);
          else l += pp_exp_block<P3_VALID2 / 16, P3_VALID2 % 16>(s_addr, sl2, -mx);
        }
        STRACE3(i * 100 + 42 + kb * 10);                         // exponentials done
        if (kb == 2) l_s[t * 128 + row] = l;
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_full[buf]);
        if (kb == 2) mbar_arrive(&l_full[t]);
        STRACE3(i * 100 + 43 + kb * 10);                         // P(t,kb) published
      }
    }
  } else {
    // ================================================================ epilogue warpgroup (warps 12-15), both tiles
    // O_t / l -> bf16 -> dense [128][88] staging tile (176-byte rows: conflict-free) -> one bulk tensor store per tile.
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const bool leader = (warp == 12 && lane == 0);
    TRACE_DECL
    for (int i = 0; i < n_my; ++i) {
      const int w = (int)blockIdx.x + i * (int)gridDim.x;
      const int b = w / VA_H, h = w - b * VA_H;
#pragma unroll 1
      for (int t = 0; t < 2; ++t) {
        mbar_wait(&l_full[t], i & 1);
        const float inv = 1.f / l_s[t * 128 + row];
        mbar_wait(&o_full[t], i & 1);
        if (leader) TRACE(2, i * 100 + 60 + t);                  // O_t complete
        tc_fence_after();
        uint32_t o[96];
        const uint32_t o_addr = lane_base + P3_COL_O + t * P3_O_STRIDE;
        tmem_ld32<0>(o_addr, o);
        tmem_ld32<32>(o_addr + 32, o);
        tmem_ld32<64>(o_addr + 64, o);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&o_free[t]);
        if (leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous store has read the tile
        named_bar_sync(4, 128);
        uint4* srow = reinterpret_cast<uint4*>(stage + row * (VA_D * 2));
#pragma unroll
        for (int c = 0; c < 11; ++c) {                           // 88 valid d columns = 11 x 16 bytes
          const float2 i2 = make_float2(inv, inv);
          const float2 a = __fmul2_rn(make_float2(__uint_as_float(o[8 * c + 0]), __uint_as_float(o[8 * c + 1])), i2);
          const float2 bq = __fmul2_rn(make_float2(__uint_as_float(o[8 * c + 2]), __uint_as_float(o[8 * c + 3])), i2);
          const float2 d = __fmul2_rn(make_float2(__uint_as_float(o[8 * c + 4]), __uint_as_float(o[8 * c + 5])), i2);
          const float2 e = __fmul2_rn(make_float2(__uint_as_float(o[8 * c + 6]), __uint_as_float(o[8 * c + 7])), i2);
          uint4 v;
          v.x = pack_bf16(a.x, a.y);
          v.y = pack_bf16(bq.x, bq.y);
          v.z = pack_bf16(d.x, d.y);
          v.w = pack_bf16(e.x, e.y);
          srow[c] = v;
        }
        fence_proxy_async();                                     // generic-proxy writes -> visible to the TMA (async proxy)
        named_bar_sync(4, 128);
        if (leader) {
          asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                       ::"l"(reinterpret_cast<uint64_t>(&tmap_out)), "r"(smem_u32(stage)), "r"(0), "r"(h),
                         "r"(b * VA_S + 1 + t * 128) : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          TRACE(2, i * 100 + 62 + t);                            // store issued
        }
      }
    }
    if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // every store has landed before the CTA may exit
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// =====================================================================================================================
// CTA-pair variant (default): one (frame, head) item per 2-CTA cluster, `tcgen05.mma.cta_group::2` with M = 256.
//
// Why: a tcgen05.mma with M = 128 costs 60-110 clk on the tensor pipe whatever its N (profiles/r2_mma_chain_microbench.txt:
// the A read of 4 KB from shared / tensor memory, not the math, sets the time), and one item needs 58-70 of them on one
// SM: 6-7k clk per item, above the item's HBM time.  With the pair, ONE instruction covers both query tiles (CTA r owns
// tile r), each CTA stages only half of K (B rows = keys, split across the pair) and a 64-column half of V (B rows = d),
// and an item needs 29 instructions: 12 for S (6 k-steps x key blocks of 144 + 128) and 17 for P V.
//   smem   per CTA and item 100 KB: Q tile [2 d blocks][128 x 128 B] | K halves [2 d blocks][72 + 64 rows x 128 B] |
//          V d-half [272 keys x 128 B] -- DOUBLE BUFFERED over items, so the loads run a whole item ahead
//   TMEM   per CTA: S0 fp32 [0,144) | S1 [144,272) | O fp32 [272,400) | row sums [400,404).  P (bf16) aliases the first
//          half of the columns each softmax thread reads (two threads per query row: A keys [0,80) / [0,64) of the block,
//          B the rest; P of thread B's keys sits at the start of B's own columns, each P V k-step has its own address)
//   sync   full[stage] local (TMA bytes) + peer_full[stage] on the leader (forwarded by the peer's idle MMA warp);
//          tcgen05.commit multicast -> s_full / pv_done / o_full / empty in both CTAs; the softmax and epilogue warps of
//          both CTAs arrive (one lane per warp) on the LEADER's p_full / o_free.
//   warps  0 TMA producer (+ TMEM alloc) | 1 MMA issuer (leader) / full-barrier forwarder (peer) | 2, 3 CLS query row of
//          key block 0 / 1 | 4-7, 8-11 softmax threads A / B | 12-15 epilogue (O / l -> bf16 -> staging tile -> bulk store)
//   CLS    the CLS query row needs every key's score but K is split across the pair: each CTA's CLS warps compute the
//          scores of their local keys (mma.sync, transposed formulation as in the ping-pong kernel), write them into BOTH
//          CTAs' score buffers (st.shared::cluster), and each CTA then runs softmax + P V for its own d-half.
// =====================================================================================================================
constexpr int C2_THREADS = 512;
constexpr int C2_KEYS0 = 144, C2_KEYS1 = 128;                      // key blocks (272 padded keys, 257 valid)
constexpr int C2_HALF0 = C2_KEYS0 / 2, C2_HALF1 = C2_KEYS1 / 2;    // 72 / 64 keys of a block held by one CTA
constexpr int C2_SPLIT0 = 80, C2_SPLIT1 = 64;                      // softmax thread A takes the first SPLIT columns of the block
constexpr int C2_VALID1 = VA_S - C2_KEYS0;                         // 113 valid keys in block 1
constexpr uint32_t C2_COL_S0 = 0, C2_COL_S1 = 144, C2_COL_O = 272, C2_COL_L = 400;
constexpr int C2_Q_BYTES = 2 * PP_Q_BYTES;                         // 32768
constexpr int C2_KD_BYTES = (C2_HALF0 + C2_HALF1) * 128;           // 17408: one d block of this CTA's K rows
constexpr int C2_K_BYTES = 2 * C2_KD_BYTES;
constexpr int C2_V_BYTES = 272 * 128;                              // 34816
constexpr int C2_STAGE_BYTES = C2_Q_BYTES + C2_K_BYTES + C2_V_BYTES;   // 102400
constexpr int C2_SMEM_BAR = 2 * C2_STAGE_BYTES;                    // 204800
constexpr int C2_SMEM_CLS = C2_SMEM_BAR + 256;                     // CLS row: float [2 item parities][2 blocks][144] scores |
                                                                   // bf16 [2 blocks][144] probabilities | float [2 blocks][2] (max, sum)
constexpr int C2_SMEM_XMAX = C2_SMEM_CLS + 2 * 2 * 144 * 4 + 2 * 144 * 2 + 32;   // float [2 halves][128] row maxima of the thread pairs
constexpr int C2_SMEM_STAGE = (C2_SMEM_XMAX + 2 * 128 * 4 + 127) / 128 * 128;
constexpr int C2_SMEM_LAYOUT = C2_SMEM_STAGE + PP_STAGE_BYTES;
// The layout is laid down from the first 1024-byte boundary of the dynamic window.  The window starts on such a boundary on
// sm_100 (no static shared memory in this kernel), so the slack below is normally unused; the kernel checks and traps if a
// misaligned base would push the layout past the allocation.
constexpr int C2_SMEM_TOTAL = C2_SMEM_LAYOUT + 1024 <= 227 * 1024 ? C2_SMEM_LAYOUT + 1024 : 227 * 1024;
static_assert(C2_SMEM_LAYOUT <= C2_SMEM_TOTAL, "shared memory budget");

__device__ __forceinline__ void umma_bf16_ts_cg2(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
template <int OFF, int NREG>
__device__ __forceinline__ void tmem_st1(uint32_t taddr, const uint32_t (&v)[NREG]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v[OFF]) : "memory");
}
template <int OFF, int NREG>
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t (&v)[NREG]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(v[OFF]), "=r"(v[OFF + 1]) : "r"(taddr));
}
__device__ __forceinline__ void st_shared_cluster_f32(uint32_t cluster_addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(cluster_addr), "f"(v) : "memory");
}

// ---- code-size diet.  The first version of this kernel instantiated the score-row routines per (key block, thread half)
// and inlined every barrier wait's time-out path: 76 KB of SASS whose roles run concurrently on one SM, and ncu showed
// 7.8 "no instruction" stall cycles per issued instruction (the ping-pong kernel: 0.7).  The routines below take their
// column counts at run time (one copy for every caller) and keep the rare paths out of line.
__device__ __noinline__ void c2_wait_slow(uint64_t* bar, uint32_t parity) { mbar_wait_impl<true>(bar, parity); }
__device__ __forceinline__ void c2_wait(uint64_t* bar, uint32_t parity) {
  if (!mbar_try_wait(bar, parity)) c2_wait_slow(bar, parity);
}
// row max of the first NVALID of NCOLS (80 / 64) fp32 score columns at s_addr: every tcgen05.ld in flight at once (a rolled
// 16-column loop with one chunk of lookahead took 1.1k clk per block instead of 0.45k: it sits on the softmax chain)
template <int NCOLS, int NVALID>
__device__ __forceinline__ float c2_row_max(uint32_t s_addr) {
  static_assert(NCOLS == 80 || NCOLS == 64, "column split widths");
  uint32_t v[NCOLS];
  tmem_ld32<0>(s_addr, v);
  tmem_ld32<32>(s_addr + 32, v);
  if (NCOLS == 80) tmem_ld16<64>(s_addr + 64, v);
  tmem_ld_wait();
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
  for (int j = 0; j < NVALID; j += 4) {
    m0 = fmaxf(m0, __uint_as_float(v[j]));
    if (j + 1 < NVALID) m1 = fmaxf(m1, __uint_as_float(v[j + 1]));
    if (j + 2 < NVALID) m2 = fmaxf(m2, __uint_as_float(v[j + 2]));
    if (j + 3 < NVALID) m3 = fmaxf(m3, __uint_as_float(v[j + 3]));
  }
  return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}
// (rolled alternative, kept for reference / tiny code: max over the first `nvalid` columns, 16-column chunks)
__device__ __forceinline__ float c2_row_max_rt(uint32_t s_addr, int nvalid) {
  const int nfull = nvalid >> 4, ntail = nvalid & 15;
  const int nchunk = nfull + (ntail ? 1 : 0);
  uint32_t cur[16], nxt[16];
  tmem_ld16<0>(s_addr, cur);
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll 1
  for (int c = 0; c < nfull; ++c) {
    tmem_ld_wait();
    if (c + 1 < nchunk) tmem_ld16<0>(s_addr + (c + 1) * 16, nxt);
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      m0 = fmaxf(m0, __uint_as_float(cur[j]));
      m1 = fmaxf(m1, __uint_as_float(cur[j + 1]));
      m2 = fmaxf(m2, __uint_as_float(cur[j + 2]));
      m3 = fmaxf(m3, __uint_as_float(cur[j + 3]));
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) cur[j] = nxt[j];
  }
  if (ntail) {
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) m0 = fmaxf(m0, j < ntail ? __uint_as_float(cur[j]) : -INFINITY);
  }
  return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}
// exp pass over `nfull` all-valid 16-column chunks plus (ntail > 0) one chunk with ntail valid columns (the rest give p = 0);
// P chunk c (8 bf16x2 columns) is stored behind the read position.  Returns the row sum.
__device__ __forceinline__ float c2_exp_rt(uint32_t s_addr, int nfull, int ntail, float sl2, float negm) {
  float2 la = make_float2(0.f, 0.f), lb = make_float2(0.f, 0.f);
  uint32_t cur[16], nxt[16];
  tmem_ld16<0>(s_addr, cur);
#pragma unroll 1
  for (int c = 0; c < nfull; ++c) {
    tmem_ld_wait();
    if (ntail > 0 || c + 1 < nfull) tmem_ld16<0>(s_addr + (c + 1) * 16, nxt);
    pp_exp16<16>(cur, sl2, negm, la, lb);
    tmem_st8<0>(s_addr + c * 8, cur);
#pragma unroll
    for (int j = 0; j < 16; ++j) cur[j] = nxt[j];
  }
  if (ntail > 0) {
    tmem_ld_wait();
    float l = 0.f;
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      const float x0 = j < ntail ? ex2_approx(fmaf(__uint_as_float(cur[j]), sl2, negm)) : 0.f;
      const float x1 = j + 1 < ntail ? ex2_approx(fmaf(__uint_as_float(cur[j + 1]), sl2, negm)) : 0.f;
      l += x0 + x1;
      cur[j >> 1] = pack_bf16(x0, x1);
    }
    la.x += l;
    tmem_st8<0>(s_addr + nfull * 8, cur);
  }
  return (la.x + la.y) + (lb.x + lb.y);
}

// byte offset of the 16-byte chunk holding d [8*chunk, 8*chunk+8) of local K row r (d block stride C2_KD_BYTES) / V row r
__device__ __forceinline__ uint32_t c2_k_off(int r, int chunk) {
  return (uint32_t)((chunk >> 3) * C2_KD_BYTES + (r >> 3) * 1024 + (r & 7) * 128 + ((((chunk & 7) ^ (r & 7))) << 4));
}
__device__ __forceinline__ uint32_t c2_v_off(int r, int chunk) {      // chunk 0..7 of this CTA's 64-column d half
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + (((chunk ^ (r & 7))) << 4));
}

// MMA issuer state of the pair kernel (see P3Issue): descriptor low words of stage 0, constant high words
struct C2Issue {
  uint32_t tmem_base, q_lo, k_lo, v_lo, hi;
  uint64_t* s_full;
};
// S of key block kb of the item in `stage` into its score buffer (one out-of-line copy for its three call sites)
__device__ __noinline__ void c2_issue_s(const C2Issue& c, int stage, int kb) {
  const uint32_t idesc = kb == 0 ? umma_idesc_bf16(256, C2_KEYS0) : umma_idesc_bf16(256, C2_KEYS1);
  const uint32_t d = c.tmem_base + (kb == 0 ? C2_COL_S0 : C2_COL_S1);
  const uint32_t qo = c.q_lo + (uint32_t)stage * (C2_STAGE_BYTES >> 4);
  const uint32_t ko = c.k_lo + (uint32_t)stage * (C2_STAGE_BYTES >> 4) + (uint32_t)kb * ((C2_HALF0 * 128) >> 4);
#pragma unroll
  for (int ks6 = 0; ks6 < 6; ++ks6) {                            // d block 0: four k-steps, d block 1: two (d 96..127 is zero)
    const int db = ks6 >> 2, ks = ks6 & 3;
    const uint64_t da = p3_desc(qo + ((db * PP_Q_BYTES + ks * 32) >> 4), c.hi);
    const uint64_t dk = p3_desc(ko + ((db * C2_KD_BYTES + ks * 32) >> 4), c.hi);
    umma_bf16_ss_cg<2>(d, da, dk, idesc, ks6 > 0 ? 1u : 0u);
  }
  umma_commit_cg<2>(&c.s_full[kb]);
}
// O (+)= P V over key block KB.  P of softmax thread A's keys starts at the block's first column, P of thread B's keys at
// B's first column (SPLIT); V rows are the item's keys in order.
template <int KB>
__device__ __forceinline__ void c2_issue_pv(const C2Issue& c, int stage) {
  constexpr uint32_t idesc_pv = umma_idesc_bf16(256, 128) | (1u << 16);      // B (= V) is MN-major
  constexpr int NSTEP = (KB == 0 ? C2_KEYS0 : C2_KEYS1) / 16;                // 9 / 8
  constexpr int ASTEP = (KB == 0 ? C2_SPLIT0 : C2_SPLIT1) / 16;              // k-steps over thread A's keys: 5 / 4
  constexpr int SPLIT = KB == 0 ? C2_SPLIT0 : C2_SPLIT1;
  constexpr int KK0 = KB == 0 ? 0 : C2_KEYS0 / 16;
  const uint32_t d = c.tmem_base + C2_COL_O;
  const uint32_t a = c.tmem_base + (KB == 0 ? C2_COL_S0 : C2_COL_S1);
  const uint32_t so = (uint32_t)stage * (C2_STAGE_BYTES >> 4);
#pragma unroll
  for (int j = 0; j < NSTEP; ++j) {
    const uint32_t pa = j < ASTEP ? a + j * 8 : a + SPLIT + (j - ASTEP) * 8;
    umma_bf16_ts_cg2(d, pa, p3_desc(c.v_lo + so + (((KK0 + j) * 2048) >> 4), c.hi), idesc_pv, (KB > 0 || j > 0) ? 1u : 0u);
  }
}

// CLS query row, key block KB: scores of this CTA's keys -> both CTAs' score buffers; then (every key's score present)
// softmax of the block and O^T = V^T p for this CTA's 64-column d half.  Returns the block's (max, sum, O).
__device__ __forceinline__ void st_async_f32(uint32_t cluster_addr, float v, uint32_t cluster_mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
               ::"r"(cluster_addr), "r"(__float_as_uint(v)), "r"(cluster_mbar) : "memory");
}
// The scores travel to the peer with st.async: every 4-byte store completes on the PEER's cls_ready barrier (transaction
// bytes), so no cluster-scope release fence is needed (an mbarrier.arrive.release.cluster here cost ~2k clk per item).
// CLS row scores on the FP32 pipe.  (A first version used mma.sync in the transposed formulation of the ping-pong kernel:
// warp-level MMAs queue behind the tcgen05.mma stream on the tensor pipe -- 30 of them took 1.9k clk, the 17 of the P V
// 4.8k -- which made the CLS row the bottleneck of the whole kernel.  The FP32 pipe is idle next to the MUFU-bound softmax.)
// Lane = (key = lane & 7 of the round's eight keys, quarter qd = lane >> 3 of the d chunks: 16-byte chunks qd, qd+4, qd+8);
// a phase of eight lanes reads eight different K rows at the same chunk index: conflict-free under the 128-byte swizzle.
__device__ __forceinline__ void c2_cls_block(const VitAttnParams& p, int i, int kb, uint32_t rank, uint32_t sK, float* sc_local,
                                             uint32_t sc_peer, uint64_t* cls_ready, uint32_t cls_ready_peer, uint64_t* kq_empty,
                                             uint64_t* clsp_free, uint64_t* clsp_ready, __nv_bfloat16* p_s, float* ml,
                                             int lane, const float (&qf)[24], int& tr_n) {
  const int half = kb == 0 ? C2_HALF0 : C2_HALF1;              // local keys of the block: 72 / 64
  const int row0 = kb == 0 ? 0 : C2_HALF0;                     // first local K row of the block
  const int nkey = kb == 0 ? C2_KEYS0 : C2_KEYS1;
  const int nvalid = kb == 0 ? C2_KEYS0 : C2_VALID1;
  const int kl = lane & 7, qd = lane >> 3;
  // per-lane byte offsets of its three chunks within an 8-row group (row kl): the round only adds 1024 bytes
  uint32_t coff[3];
#pragma unroll
  for (int c3 = 0; c3 < 3; ++c3) {
    const int chunk = qd + 4 * c3;                             // 0..10 hold d 0..87; chunk 11 (qd == 3, c3 == 2) does not exist
    coff[c3] = sK + (uint32_t)((chunk >> 3) * C2_KD_BYTES + (row0 >> 3) * 1024 + kl * 128 + (((chunk & 7) ^ kl) << 4));
  }
  const bool has3 = qd < 3;
  // all rounds first, the stores afterwards: a st.async between two rounds is a scheduling barrier (volatile + memory
  // clobber) that exposed every round's load and shuffle latency -- 3.7k clk for 520 instructions
  float scs[9];
#pragma unroll
  for (int r = 0; r < 9; ++r) {
    if (r == 8 && kb != 0) { scs[8] = 0.f; break; }            // 9 rounds of eight keys (72) / 8 rounds (64)
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int c3 = 0; c3 < 3; ++c3) {
      if (c3 == 2 && !has3) break;
      uint32_t w[4];
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3])
                   : "r"(coff[c3] + r * 1024));
      a0 = fmaf(__uint_as_float(w[0] << 16), qf[c3 * 8 + 0], a0);
      a1 = fmaf(__uint_as_float(w[0] & 0xffff0000u), qf[c3 * 8 + 1], a1);
      a2 = fmaf(__uint_as_float(w[1] << 16), qf[c3 * 8 + 2], a2);
      a3 = fmaf(__uint_as_float(w[1] & 0xffff0000u), qf[c3 * 8 + 3], a3);
      a0 = fmaf(__uint_as_float(w[2] << 16), qf[c3 * 8 + 4], a0);
      a1 = fmaf(__uint_as_float(w[2] & 0xffff0000u), qf[c3 * 8 + 5], a1);
      a2 = fmaf(__uint_as_float(w[3] << 16), qf[c3 * 8 + 6], a2);
      a3 = fmaf(__uint_as_float(w[3] & 0xffff0000u), qf[c3 * 8 + 7], a3);
    }
    scs[r] = (a0 + a1) + (a2 + a3);
  }
#pragma unroll
  for (int r = 0; r < 9; ++r) {
    scs[r] += __shfl_xor_sync(0xffffffffu, scs[r], 8);
    scs[r] += __shfl_xor_sync(0xffffffffu, scs[r], 16);
  }
  // every lane now holds the scores of keys 8 r + kl; lane quarter qd stores rounds qd, qd + 4, qd + 8
#pragma unroll
  for (int r = 0; r < 9; ++r) {
    if ((r & 3) == qd && r * 8 < half) {
      const int j = r * 8 + kl;
      sc_local[rank * half + j] = scs[r];
      st_async_f32(sc_peer + (rank * half + j) * 4, scs[r], cls_ready_peer);
    }
  }
  __syncwarp();
  if (lane == 0) {
    mbar_arrive(kq_empty);                                       // done reading this stage's K rows
    mbar_arrive_expect_tx(cls_ready, half * 4);                  // my half is written; the peer's scores arrive as bytes
  }
  if (lane == 0 && kb == 0) TRACE(3, i * 100 + 31);              // scores of the local keys sent
  c2_wait(cls_ready, i & 1);                                     // both halves of the block's scores are here
  if (lane == 0 && kb == 0) TRACE(3, i * 100 + 32);              // peer's scores here
  // ---- softmax over the block's valid keys, 5 per lane
  float x[5];
  float m = -INFINITY;
#pragma unroll
  for (int it = 0; it < 5; ++it) {
    const int j = it * 32 + lane;
    x[it] = (j < nvalid) ? sc_local[j] * p.scale_log2 : -INFINITY;
    m = fmaxf(m, x[it]);
  }
  m = warp_max(m);
  float l = 0.f;
#pragma unroll
  for (int it = 0; it < 5; ++it) {
    x[it] = ex2_approx(x[it] - m);                               // exp2(-inf) = 0 for masked keys
    l += x[it];
  }
  l = warp_sum(l);
  // ---- hand the block's probabilities (bf16) and (max, sum) to the epilogue warps, which run O = p V
  c2_wait(clsp_free, (i & 1) ^ 1);                               // they are done with the previous item's
#pragma unroll
  for (int it = 0; it < 5; ++it)
    if (it * 32 + lane < nkey) p_s[it * 32 + lane] = __float2bfloat16(x[it]);
  if (lane == 0) { ml[0] = m; ml[1] = l; }
  __syncwarp();
  if (lane == 0) mbar_arrive(clsp_ready);
  if (lane == 0 && kb == 0) TRACE(3, i * 100 + 33);              // probabilities published
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(C2_THREADS, 1)
vit_attention_c2_kernel(const __grid_constant__ CUtensorMap tmap72, const __grid_constant__ CUtensorMap tmap64,
                        const __grid_constant__ CUtensorMap tmap128, const __grid_constant__ CUtensorMap tmap136,
                        const __grid_constant__ CUtensorMap tmap_out, const VitAttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  if (threadIdx.x == 0 && (smem - smem_raw) + C2_SMEM_LAYOUT > C2_SMEM_TOTAL) {
    printf("tgb: vit_attention_c2: dynamic shared memory base %p is not aligned as expected\n", (void*)smem_raw);
    __trap();
  }
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C2_SMEM_BAR);
  // Q + K and V of an item have separate barriers: Q / K are released as soon as the item's S MMAs have completed (they are
  // issued during the PREVIOUS item), so the producer runs Q / K two items and V one item ahead.
  uint64_t* kq_full = bars + 0;     // [2 stages] this CTA's TMA bytes of an item's Q tile and K rows
  uint64_t* v_full = bars + 2;      // [2] ... and of its V half
  uint64_t* peer_kq = bars + 4;     // [2] leader only: the peer CTA's kq_full has completed (forwarded)
  uint64_t* peer_v = bars + 6;      // [2] leader only: ... v_full
  uint64_t* kq_empty = bars + 8;    // [2] tcgen05.commit (multicast) + this CTA's two CLS warps
  uint64_t* v_empty = bars + 10;    // [2] same
  uint64_t* s_full = bars + 12;     // [2 key blocks] multicast commit, one completion per item
  uint64_t* p_full = bars + 14;     // [2] leader only: 16 arrivals (8 softmax warps x 2 CTAs)
  uint64_t* pv_done = bars + 16;    // P V of key block 0 has completed (multicast; the rare accumulator rescale waits for it)
  uint64_t* o_full = bars + 17;     // multicast commit, one completion per item
  uint64_t* o_free = bars + 18;     // leader only: 8 arrivals (4 epilogue warps x 2 CTAs)
  uint64_t* cls_ready = bars + 19;  // [2 key blocks] 1 arrival (this CTA's CLS warp) + the peer's scores as transaction bytes
  uint64_t* clsp_ready = bars + 21; // the CLS row's probabilities of both key blocks are in shared memory (2 CLS warps)
  uint64_t* clsp_free = bars + 22;  // ... have been consumed (4 epilogue warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 23);
  float* cls_sc = reinterpret_cast<float*>(smem + C2_SMEM_CLS);             // [2 item parities][2 key blocks][144] scores
  __nv_bfloat16* cls_p = reinterpret_cast<__nv_bfloat16*>(cls_sc + 2 * 2 * 144);   // [2 key blocks][144] probabilities
  float* cls_ml = reinterpret_cast<float*>(cls_p + 2 * 144);                // [2 key blocks][2] block max (log2 units), block sum
  float* xmax = reinterpret_cast<float*>(smem + C2_SMEM_XMAX);              // [2][128]
  uint8_t* stage_out = smem + C2_SMEM_STAGE;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int n_my = p.num_items > cluster_id ? (p.num_items - cluster_id + num_clusters - 1) / num_clusters : 0;

  if (warp == 1 && lane == 0) {
    tma_prefetch_desc(&tmap72);
    tma_prefetch_desc(&tmap64);
    tma_prefetch_desc(&tmap128);
    tma_prefetch_desc(&tmap136);
    tma_prefetch_desc(&tmap_out);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&kq_full[s], 1); mbar_init(&v_full[s], 1); mbar_init(&peer_kq[s], 1); mbar_init(&peer_v[s], 1);
      mbar_init(&kq_empty[s], 3); mbar_init(&v_empty[s], 5);
      mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 16); mbar_init(&cls_ready[s], 1);
    }
    mbar_init(pv_done, 1); mbar_init(o_full, 1); mbar_init(o_free, 8);
    mbar_init(clsp_ready, 2); mbar_init(clsp_free, 4);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc_cg<2>(tmem_slot, 512);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================================================ TMA producer: this CTA's share of every item
    // order K/Q(0) K/Q(1) V(0) K/Q(2) V(1) K/Q(3) ...: each load is issued as soon as its buffer can be free
    if (lane == 0) {
      TRACE_DECL
      auto load_kq = [&](int i) {
        const int w = cluster_id + i * num_clusters;
        const int b = w / VA_H, h = w - b * VA_H;
        const int row0 = b * VA_S;
        const int s = i & 1;
        uint8_t* sQ = smem + s * C2_STAGE_BYTES;
        uint8_t* sK = sQ + C2_Q_BYTES;
        c2_wait(&kq_empty[s], ((i >> 1) & 1) ^ 1);
        TRACE(0, i * 100 + 1);                                   // Q / K buffers free -> loads issued
        mbar_arrive_expect_tx(&kq_full[s], C2_Q_BYTES + C2_K_BYTES);
#pragma unroll
        for (int db = 0; db < 2; ++db) {                         // K rows of block 0, Q tile, K rows of block 1
          tma_load_3d(sK + db * C2_KD_BYTES, &tmap72, &kq_full[s], db * 64, VA_H + h, row0 + (int)rank * C2_HALF0);
          tma_load_3d(sQ + db * PP_Q_BYTES, &tmap128, &kq_full[s], db * 64, h, row0 + 1 + (int)rank * 128);
          tma_load_3d(sK + db * C2_KD_BYTES + C2_HALF0 * 128, &tmap64, &kq_full[s], db * 64, VA_H + h,
                      row0 + C2_KEYS0 + (int)rank * C2_HALF1);
        }
      };
      auto load_v = [&](int i) {
        const int w = cluster_id + i * num_clusters;
        const int b = w / VA_H, h = w - b * VA_H;
        const int row0 = b * VA_S;
        const int s = i & 1;
        uint8_t* sV = smem + s * C2_STAGE_BYTES + C2_Q_BYTES + C2_K_BYTES;
        c2_wait(&v_empty[s], ((i >> 1) & 1) ^ 1);
        TRACE(0, i * 100 + 2);                                   // V buffer free -> loads issued
        mbar_arrive_expect_tx(&v_full[s], C2_V_BYTES);
#pragma unroll
        for (int hb = 0; hb < 2; ++hb)                           // every key, this CTA's 64-column d half (88.. zero-filled)
          tma_load_3d(sV + hb * VA_BOX_BYTES, &tmap136, &v_full[s], (int)rank * 64, 2 * VA_H + h, row0 + hb * VA_BOX_ROWS);
      };
      if (n_my > 0) load_kq(0);
      if (n_my > 1) load_kq(1);
      for (int i = 0; i < n_my; ++i) {
        load_v(i);
        if (i + 2 < n_my) load_kq(i + 2);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && !leader) {
      // ============================================================== peer: forward "my loads have landed" to the leader
      // (same order as the producer issues them; the bytes are TMA (async-proxy) writes, complete once the barrier fires)
      const uint32_t pkq0 = mapa_u32(smem_u32(&peer_kq[0]), 0), pv0 = mapa_u32(smem_u32(&peer_v[0]), 0);
      auto fwd_kq = [&](int i) { c2_wait(&kq_full[i & 1], (i >> 1) & 1); mbar_arrive_cluster_relaxed(pkq0 + (i & 1) * 8); };
      auto fwd_v = [&](int i) { c2_wait(&v_full[i & 1], (i >> 1) & 1); mbar_arrive_cluster_relaxed(pv0 + (i & 1) * 8); };
      if (n_my > 0) fwd_kq(0);
      if (n_my > 1) fwd_kq(1);
      for (int i = 0; i < n_my; ++i) {
        fwd_v(i);
        if (i + 2 < n_my) fwd_kq(i + 2);
      }
    } else if (lane == 0) {
      // ============================================================== MMA issuer (leader CTA, for the pair)
      TRACE_DECL
      C2Issue c;
      c.tmem_base = tmem_base;
      c.q_lo = ((smem_u32(smem) & 0x3FFFFu) >> 4) | (1u << 16);
      c.k_lo = ((smem_u32(smem + C2_Q_BYTES) & 0x3FFFFu) >> 4) | (1u << 16);
      c.v_lo = ((smem_u32(smem + C2_Q_BYTES + C2_K_BYTES) & 0x3FFFFu) >> 4) | (1u << 16);
      c.hi = (1024u >> 4) | (1u << 14) | (2u << 29);               // SBO 1024 | version | SWIZZLE_128B (K-major and MN-major alike)
      c.s_full = s_full;
      auto kq_landed = [&](int i) {                              // non-blocking
        return mbar_try_wait(&kq_full[i & 1], (i >> 1) & 1) && mbar_try_wait(&peer_kq[i & 1], (i >> 1) & 1);
      };
      auto wait_kq = [&](int i) {
        c2_wait(&kq_full[i & 1], (i >> 1) & 1);
        c2_wait(&peer_kq[i & 1], (i >> 1) & 1);
        tc_fence_after();
      };
      if (n_my > 0) {
        wait_kq(0);
        c2_issue_s(c, 0, 0);
        c2_issue_s(c, 0, 1);
        umma_commit_cg<2>(&kq_empty[0]);
      }
#pragma unroll 1
      for (int i = 0; i < n_my; ++i) {
        const uint32_t ph = i & 1;
        const int s = i & 1;
        const bool has_next = i + 1 < n_my;
        c2_wait(&p_full[0], ph);
        TRACE(1, i * 100 + 10);                                  // P(0) seen
        c2_wait(o_free, ph ^ 1);                               // both CTAs' epilogues have read the previous item's O
        c2_wait(&v_full[s], (i >> 1) & 1);
        c2_wait(&peer_v[s], (i >> 1) & 1);
        tc_fence_after();
        TRACE(1, i * 100 + 11);                                  // O free, V landed
        c2_issue_pv<0>(c, s);
        umma_commit_cg<2>(pv_done);
        TRACE(1, i * 100 + 12);                                  // P V(0) issued
        // S(0) of the next item goes in here when its operands are already there; otherwise P V(1) must not wait for them
        bool s0_done = false;
        if (has_next && kq_landed(i + 1)) {
          tc_fence_after();
          c2_issue_s(c, s ^ 1, 0);
          s0_done = true;
          TRACE(1, i * 100 + 14);                                // next S(0) issued (early)
        }
        c2_wait(&p_full[1], ph);
        TRACE(1, i * 100 + 15);                                  // P(1) seen
        tc_fence_after();
        c2_issue_pv<1>(c, s);
        umma_commit_cg<2>(o_full);
        umma_commit_cg<2>(&v_empty[s]);                          // every MMA reading this stage's V has completed
        TRACE(1, i * 100 + 16);                                  // P V(1) issued
        if (has_next) {
          if (!s0_done) {
            wait_kq(i + 1);
            TRACE(1, i * 100 + 13);                              // next item's Q / K landed (late)
            c2_issue_s(c, s ^ 1, 0);
          }
          c2_issue_s(c, s ^ 1, 1);
          umma_commit_cg<2>(&kq_empty[s ^ 1]);                   // ... Q / K of the next item
          TRACE(1, i * 100 + 17);                                // next S(1) issued
        }
      }
    }
    __syncwarp();
  } else if (warp < 4) {
    // ================================================================ CLS query row, scores + softmax: warp 2 key block 0, warp 3 block 1
    const int kb = warp - 2;
    const uint32_t peer = rank ^ 1u;
    int tr_n = 0;
    // the CLS query: this lane's d chunks (16 bytes each: qd, qd + 4, qd + 8), fetched one item ahead as packed bf16
    // (global-memory latency off the per-item chain) and widened to fp32 at the top of the item
    const int qd = lane >> 3;
    uint4 qn[3];
    auto load_q = [&](int i) {
      const int w = cluster_id + i * num_clusters;
      const int b = w / VA_H, h = w - b * VA_H;
      const uint4* q4 = reinterpret_cast<const uint4*>(p.qkv + (long long)b * VA_S * p.ld_qkv + h * VA_D);
#pragma unroll
      for (int c3 = 0; c3 < 3; ++c3) qn[c3] = (qd + 4 * c3 < 11) ? __ldg(q4 + qd + 4 * c3) : make_uint4(0u, 0u, 0u, 0u);
    };
    if (n_my > 0) load_q(0);
    for (int i = 0; i < n_my; ++i) {
      const int s = i & 1;
      const uint32_t sK32 = smem_u32(smem + s * C2_STAGE_BYTES + C2_Q_BYTES);
      float qf[24];
#pragma unroll
      for (int c3 = 0; c3 < 3; ++c3) {
        const uint32_t w4[4] = {qn[c3].x, qn[c3].y, qn[c3].z, qn[c3].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          qf[c3 * 8 + 2 * e] = __uint_as_float(w4[e] << 16);
          qf[c3 * 8 + 2 * e + 1] = __uint_as_float(w4[e] & 0xffff0000u);
        }
      }
      if (i + 1 < n_my) load_q(i + 1);
      float* sc_local = cls_sc + ((i & 1) * 2 + kb) * 144;
      const uint32_t sc_peer = mapa_u32(smem_u32(sc_local), peer);
      const uint32_t ready_peer = mapa_u32(smem_u32(&cls_ready[kb]), peer);
      c2_wait(&kq_full[s], (i >> 1) & 1);
      if (lane == 0 && kb == 0) TRACE(3, i * 100 + 30);          // K landed
      c2_cls_block(p, i, kb, rank, sK32, sc_local, sc_peer, &cls_ready[kb], ready_peer, &kq_empty[s], clsp_free, clsp_ready,
                   cls_p + kb * 144, cls_ml + kb * 2, lane, qf, tr_n);
    }
  } else if (warp < 12) {
    // ================================================================ softmax: two threads per query row (A: warps 4-7, B: 8-11)
    const int hf = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t o_addr = lane_base + C2_COL_O + hf * 64;      // this thread's half of the accumulator columns (rescale)
    const float sl2 = p.scale_log2;
    const uint32_t p_full0 = mapa_u32(smem_u32(&p_full[0]), 0), p_full1 = mapa_u32(smem_u32(&p_full[1]), 0);
    TRACE_DECL
#ifdef TGB_ATTN_TRACE
    const bool tr_on = (quarter == 0 && lane == 0);
#define CTRACE(tag) do { if (tr_on) TRACE(4 + hf, tag); } while (0)
#else
#define CTRACE(tag) do { } while (0)
#endif
    for (int i = 0; i < n_my; ++i) {
      const uint32_t ph = i & 1;
      float mx, l;
      {
        // ------------------------------------------------------------ key block 0: keys [0,144), all valid
        const uint32_t s_addr = lane_base + C2_COL_S0 + hf * C2_SPLIT0;
        c2_wait(&s_full[0], ph);
        CTRACE(i * 100 + 40);                                    // S(0) seen
        tc_fence_after();
        const float mine = hf == 0 ? c2_row_max<80, 80>(s_addr) : c2_row_max<64, 64>(s_addr);
        xmax[hf * 128 + row] = mine;
        named_bar_sync(4 + quarter, 64);
        mx = fmaxf(mine, xmax[(hf ^ 1) * 128 + row]) * sl2;
        named_bar_sync(4 + quarter, 64);                         // both have read before the next block's maxima are written
        CTRACE(i * 100 + 41);                                    // row max known
        l = c2_exp_rt(s_addr, hf == 0 ? C2_SPLIT0 / 16 : (C2_KEYS0 - C2_SPLIT0) / 16, 0, sl2, -mx);
        CTRACE(i * 100 + 42);                                    // exponentials done
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_relaxed(p_full0);   // P is in tensor memory (tcgen05.wait::st above): no generic-memory release needed
        CTRACE(i * 100 + 43);                                    // P(0) published
      }
      {
        // ------------------------------------------------------------ key block 1: keys [144,272), 113 valid
        const uint32_t s_addr = lane_base + C2_COL_S1 + hf * C2_SPLIT1;
        c2_wait(&s_full[1], ph);
        CTRACE(i * 100 + 50);                                    // S(1) seen
        tc_fence_after();
        const int nv1 = hf == 0 ? C2_SPLIT1 : C2_VALID1 - C2_SPLIT1;      // valid keys among this thread's 64 columns: 64 / 49
        const float mine = hf == 0 ? c2_row_max<64, 64>(s_addr) : c2_row_max<64, C2_VALID1 - C2_SPLIT1>(s_addr);
        xmax[hf * 128 + row] = mine;
        named_bar_sync(4 + quarter, 64);
        const float m1x = fmaxf(mine, xmax[(hf ^ 1) * 128 + row]) * sl2;
        named_bar_sync(4 + quarter, 64);
        float f = 1.f;
        const bool need = m1x > mx + PP_RESCALE_THRESHOLD;       // the same decision in both threads of the row
        if (need) {
          f = ex2_approx(mx - m1x);
          mx = m1x;
          l *= f;
        }
        if (__any_sync(0xffffffffu, need)) {                     // rare: rescale this warp's rows of its O columns in place
          c2_wait(pv_done, ph);                                 // P V of key block 0 has completed
          tc_fence_after();
#pragma unroll 1
          for (int cc = 0; cc < (hf == 0 ? 64 : 32); cc += 8) {  // d 0..63 / 64..95 (88.. are zero)
            uint32_t ov[8];
            tmem_ld8<0>(o_addr + cc, ov);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 8; ++e) ov[e] = __float_as_uint(__uint_as_float(ov[e]) * f);
            tmem_st8<0>(o_addr + cc, ov);
          }
          tmem_st_wait();
        }
        l += c2_exp_rt(s_addr, nv1 >> 4, nv1 & 15, sl2, -mx);
        uint32_t lv[1] = {__float_as_uint(l)};                   // this thread's part of the row sum, for the epilogue
        tmem_st1<0>(lane_base + C2_COL_L + (i & 1) * 2 + hf, lv);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_relaxed(p_full1);
        CTRACE(i * 100 + 53);                                    // P(1) published
      }
    }
  } else {
    // ================================================================ epilogue warpgroup (warps 12-15)
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const bool store_thread = (warp == 12 && lane == 0);
    const uint32_t o_free0 = mapa_u32(smem_u32(o_free), 0);
    TRACE_DECL
    for (int i = 0; i < n_my; ++i) {
      const int w = cluster_id + i * num_clusters;
      const int b = w / VA_H, h = w - b * VA_H;
      c2_wait(o_full, i & 1);
      if (store_thread) TRACE(2, i * 100 + 60);                  // O complete
      tc_fence_after();
      uint32_t o[96], lv[2];
      tmem_ld32<0>(lane_base + C2_COL_O, o);
      tmem_ld32<32>(lane_base + C2_COL_O + 32, o);
      tmem_ld32<64>(lane_base + C2_COL_O + 64, o);
      tmem_ld2<0>(lane_base + C2_COL_L + (i & 1) * 2, lv);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(o_free0);   // the accumulator has been READ (tcgen05.wait::ld above)
      const float inv = 1.f / (__uint_as_float(lv[0]) + __uint_as_float(lv[1]));
      if (store_thread) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous store has read the tile
      named_bar_sync(2, 128);
      uint4* srow = reinterpret_cast<uint4*>(stage_out + row * (VA_D * 2));
#pragma unroll
      for (int cc = 0; cc < 11; ++cc) {                          // 88 valid d columns = 11 x 16 bytes
        const float2 i2 = make_float2(inv, inv);
        const float2 a = __fmul2_rn(make_float2(__uint_as_float(o[8 * cc + 0]), __uint_as_float(o[8 * cc + 1])), i2);
        const float2 bq = __fmul2_rn(make_float2(__uint_as_float(o[8 * cc + 2]), __uint_as_float(o[8 * cc + 3])), i2);
        const float2 d = __fmul2_rn(make_float2(__uint_as_float(o[8 * cc + 4]), __uint_as_float(o[8 * cc + 5])), i2);
        const float2 e = __fmul2_rn(make_float2(__uint_as_float(o[8 * cc + 6]), __uint_as_float(o[8 * cc + 7])), i2);
        uint4 v;
        v.x = pack_bf16(a.x, a.y);
        v.y = pack_bf16(bq.x, bq.y);
        v.z = pack_bf16(d.x, d.y);
        v.w = pack_bf16(e.x, e.y);
        srow[cc] = v;
      }
      fence_proxy_async();                                       // generic-proxy writes -> visible to the TMA (async proxy)
      named_bar_sync(2, 128);
      if (store_thread) {
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                     ::"l"(reinterpret_cast<uint64_t>(&tmap_out)), "r"(smem_u32(stage_out)), "r"(0), "r"(h),
                       "r"(b * VA_S + 1 + (int)rank * 128) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        TRACE(2, i * 100 + 62);                                  // store issued
      }
      // ---- CLS query row, O = p V for this CTA's d half on the FP32 pipe: warp q takes d [16q, 16q+16); lane = (key group
      // kg = lane >> 3, d pair dp = lane & 7); key = 4 j + kg, so a phase of eight lanes reads 32 contiguous bytes of one V row
      // and the four rows of an instruction sit at four different swizzle phases (conflict-free).  Keys 0..143 (j < 36) are
      // block 0, the rest block 1: one accumulator pair per block (each block's p is relative to its own max).
      {
        const int s = i & 1;
        const uint32_t sV32 = smem_u32(smem + s * C2_STAGE_BYTES + C2_Q_BYTES + C2_K_BYTES);
        const int kg = lane >> 3, dp = lane & 7;
        c2_wait(clsp_ready, i & 1);
        const int chunk = 2 * quarter + (dp >> 2);              // 16-byte chunk of the row holding this lane's d pair
        // key = 8 jj + kg (row kg of 8-row group jj) and 8 jj + 4 + kg (row kg + 4): two fixed byte offsets per lane, the
        // group only adds 1024 bytes; the probabilities sit at element == key (block 1's key 144 is element 144 of cls_p)
        const uint32_t voff0 = sV32 + (uint32_t)(kg * 128 + ((chunk ^ kg) << 4) + (dp & 3) * 4);
        const uint32_t voff1 = sV32 + (uint32_t)((kg + 4) * 128 + ((chunk ^ (kg + 4)) << 4) + (dp & 3) * 4);
        const uint32_t poff = smem_u32(cls_p) + kg * 2;
        float2 acc[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
        float2 accb[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
          const int g0 = blk == 0 ? 0 : 18, g1 = blk == 0 ? 18 : 34;      // 8-row groups: keys [0,144) / [144,272)
#pragma unroll 6
          for (int jj = g0; jj < g1; ++jj) {
            uint32_t v0, v1;
            uint16_t p0, p1;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v0) : "r"(voff0 + jj * 1024));
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v1) : "r"(voff1 + jj * 1024));
            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(p0) : "r"(poff + jj * 16));
            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(p1) : "r"(poff + jj * 16 + 8));
            const float pk0 = __uint_as_float(static_cast<uint32_t>(p0) << 16);
            const float pk1 = __uint_as_float(static_cast<uint32_t>(p1) << 16);
            acc[blk].x = fmaf(pk0, __uint_as_float(v0 << 16), acc[blk].x);
            acc[blk].y = fmaf(pk0, __uint_as_float(v0 & 0xffff0000u), acc[blk].y);
            accb[blk].x = fmaf(pk1, __uint_as_float(v1 << 16), accb[blk].x);
            accb[blk].y = fmaf(pk1, __uint_as_float(v1 & 0xffff0000u), accb[blk].y);
          }
          acc[blk].x += accb[blk].x;
          acc[blk].y += accb[blk].y;
        }
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
          acc[blk].x += __shfl_xor_sync(0xffffffffu, acc[blk].x, 8);
          acc[blk].y += __shfl_xor_sync(0xffffffffu, acc[blk].y, 8);
          acc[blk].x += __shfl_xor_sync(0xffffffffu, acc[blk].x, 16);
          acc[blk].y += __shfl_xor_sync(0xffffffffu, acc[blk].y, 16);
        }
        const float m0 = cls_ml[0], l0 = cls_ml[1], m1 = cls_ml[2], l1 = cls_ml[3];
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(clsp_free);                                // p and (m, l) of this item have been read
          mbar_arrive(&v_empty[s]);                              // ... and this stage's V
        }
        const float mm = fmaxf(m0, m1);
        const float f0 = ex2_approx(m0 - mm), f1 = ex2_approx(m1 - mm);
        const float inv = 1.f / (l0 * f0 + l1 * f1);
        if (kg == 0) {
          const int d0 = (int)rank * 64 + quarter * 16 + 2 * dp;
          if (d0 < VA_D) {
            __nv_bfloat16* dst = p.out + (long long)b * VA_S * p.ld_out + h * VA_D + d0;
            *reinterpret_cast<uint32_t*>(dst) = pack_bf16((acc[0].x * f0 + acc[1].x * f1) * inv, (acc[0].y * f0 + acc[1].y * f1) * inv);
          }
        }
        if (store_thread) TRACE(2, i * 100 + 63);                // CLS row done
      }
    }
    if (store_thread) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // every store has landed before the CTA may exit
  }

  tc_fence_before();
  cluster_sync_all();                                            // no CTA leaves while its pair may still touch its smem / barriers
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc_cg<2>(tmem_base, 512);
  }
}

static PFN_cuTensorMapEncodeTiled_v12000 va_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  return fn;
}

}  // namespace tgb

using namespace tgb;

#ifdef TGB_ATTN_TRACE
extern "C" __attribute__((visibility("default"))) int tgb_debug_attn_trace(long long* out_host, int* n_host) {
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(out_host, g_attn_trace, sizeof(long long) * TR_ROLES * TR_CAP * 2) != cudaSuccess) return -1;
  if (cudaMemcpyFromSymbol(n_host, g_attn_trace_n, sizeof(int) * TR_ROLES) != cudaSuccess) return -1;
  return TR_CAP;
}
#endif

static std::atomic<int> g_vit_attention_variant{0};   // 0 = ping-pong kernel (default), 1 = one-shot, 2 = three-block (cross-checks)

static std::atomic<int> g_vit_attention_dbg{0};
extern "C" __attribute__((visibility("default"))) int tgb_debug_vit_attention_dbg(int bits) {
  g_vit_attention_dbg = bits;
  return TGB_OK;
}
static std::atomic<int> g_vit_attention_prefetch{0};   // measured 30 % slower with it (profiles/r2_attention_experiments.txt)
extern "C" int tgb_vit_attention_set_prefetch(int on) {
  g_vit_attention_prefetch = on ? 1 : 0;
  return TGB_OK;
}

extern "C" int tgb_vit_attention_set_variant(int variant) {
  TGB_CHECK_ARG(variant >= 0 && variant <= 5, "vit_attention_set_variant: 0 (ping-pong), 1 (one-shot), 2 (three-block), 3 (CTA pair), 4 (split-row ping-pong) or 5 (ping-pong, pipelined softmax)");
  g_vit_attention_variant = variant;
  return TGB_OK;
}

extern "C" int tgb_vit_attention_bf16(const void* qkv, int64_t ld_qkv, void* out, int64_t ld_out, int frames,
                                      float scale, tgb_stream_t stream) {
  int st = check_device();
  if (st != TGB_OK) return st;
  TGB_CHECK_ARG(qkv && out && frames >= 0, "vit_attention: null pointer");
  TGB_CHECK_ARG(ld_qkv % 8 == 0 && ld_qkv >= 3 * VA_H * VA_D && ld_out % 8 == 0 && ld_out >= VA_H * VA_D,
                "vit_attention: leading dims must be multiples of 8 and cover 3*16*88 / 16*88 columns");
  TGB_CHECK_ARG(((uintptr_t)qkv % 16) == 0 && ((uintptr_t)out % 16) == 0, "vit_attention: buffers must be 16-byte aligned");
  if (frames == 0) return TGB_OK;
  auto fn = va_encode_fn();
  if (!fn) {
    set_last_error("cuTensorMapEncodeTiled not available from the driver");
    return TGB_ECUDA;
  }
  CUtensorMap map, map144, map128, map96, map80, map72, map64;
  cuuint64_t dims[3] = {VA_D, 3 * VA_H, (cuuint64_t)frames * VA_S};
  cuuint64_t strides[2] = {VA_D * 2, (cuuint64_t)ld_qkv * 2};
  cuuint32_t estr[3] = {1, 1, 1};
  const int variant = g_vit_attention_variant.load();
  CUtensorMap* maps[7] = {&map128, &map96, &map80, &map144, &map, &map72, &map64};
  const cuuint32_t box_rows[7] = {128, P3_KB, P3_KB2, PP_KEYS0, VA_BOX_ROWS, C2_HALF0, C2_HALF1};
  // tensor maps each variant reads: ping-pong 128 144 | one-shot 136 | three-block 128 96 80 | pair 128 136 72 64
  static const unsigned char needs[6][7] = {{1, 0, 0, 1, 0, 0, 0}, {0, 0, 0, 0, 1, 0, 0}, {1, 1, 1, 0, 0, 0, 0}, {1, 0, 0, 0, 1, 1, 1},
                                            {1, 0, 0, 1, 0, 0, 0}, {1, 0, 0, 1, 0, 0, 0}};
  for (int i = 0; i < 7; ++i) {
    if (!needs[variant][i]) continue;
    cuuint32_t box[3] = {64, 1, box_rows[i]};
    CUresult r = fn(maps[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(qkv), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_last_error("vit_attention: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
      return TGB_ECUDA;
    }
  }
  CUtensorMap map_out;                                            // [rows][16 heads][88] view of `out`, dense 128 x 88 box
  {
    cuuint64_t odims[3] = {VA_D, VA_H, (cuuint64_t)frames * VA_S};
    cuuint64_t ostrides[2] = {VA_D * 2, (cuuint64_t)ld_out * 2};
    cuuint32_t obox[3] = {VA_D, 1, 128};
    CUresult r = fn(&map_out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, out, odims, ostrides, obox, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_last_error("vit_attention: cuTensorMapEncodeTiled (output) failed (CUresult %d)", (int)r);
      return TGB_ECUDA;
    }
  }
  TGB_ONCE_PER_DEVICE({
    TGB_CUDA(cudaFuncSetAttribute(vit_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, VA_SMEM_TOTAL));
    TGB_CUDA(cudaFuncSetAttribute(vit_attention_pp_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PP_SMEM_TOTAL));
    TGB_CUDA(cudaFuncSetAttribute(vit_attention_pp_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PP_SMEM_TOTAL));
    TGB_CUDA(cudaFuncSetAttribute(vit_attention_pp_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PP2_SMEM_TOTAL));
    TGB_CUDA(cudaFuncSetAttribute(vit_attention_p3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P3_SMEM_TOTAL));
    TGB_CUDA(cudaFuncSetAttribute(vit_attention_c2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C2_SMEM_TOTAL));
  });
  VitAttnParams p;
  p.qkv = (const __nv_bfloat16*)qkv;
  p.out = (__nv_bfloat16*)out;
  p.ld_qkv = ld_qkv;
  p.ld_out = ld_out;
  p.frames = frames;
  p.num_items = frames * VA_H;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.l2_prefetch = g_vit_attention_prefetch.load();
  p.dbg = g_vit_attention_dbg.load();
  const int grid = p.num_items < sm_count() ? p.num_items : sm_count();
  if (variant == 3) {
    const int pairs = sm_count() / 2;
    const int clusters = p.num_items < pairs ? p.num_items : pairs;
    vit_attention_c2_kernel<<<2 * clusters, C2_THREADS, C2_SMEM_TOTAL, (cudaStream_t)stream>>>(map72, map64, map128, map, map_out, p);
  } else if (variant == 1)
    vit_attention_tc_kernel<<<grid, VA_THREADS, VA_SMEM_TOTAL, (cudaStream_t)stream>>>(map, p);
  else if (variant == 2)
    vit_attention_p3_kernel<<<grid, P3_THREADS, P3_SMEM_TOTAL, (cudaStream_t)stream>>>(map96, map80, map128, map_out, p);
  else if (variant == 4)
    vit_attention_pp_kernel<2, true><<<grid, PP2_THREADS, PP2_SMEM_TOTAL, (cudaStream_t)stream>>>(map144, map128, map_out, p);
  else if (variant == 5)
    vit_attention_pp_kernel<1, true><<<grid, VA_THREADS, PP_SMEM_TOTAL, (cudaStream_t)stream>>>(map144, map128, map_out, p);
  else
    vit_attention_pp_kernel<1, false><<<grid, VA_THREADS, PP_SMEM_TOTAL, (cudaStream_t)stream>>>(map144, map128, map_out, p);
  TGB_CUDA(cudaGetLastError());
  return TGB_OK;
}
