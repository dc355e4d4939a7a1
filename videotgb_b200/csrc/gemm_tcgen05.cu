// bf16 GEMM with fused epilogues on the 5th-gen tensor cores:  out = epi(A[M,K] * W[N,K]^T)
//
// Persistent, warp-specialised kernel; one CTA per SM, CTAs paired into 2-CTA clusters (cta_group::2) so that one
// tcgen05.mma computes a 256 x BN tile across two SMs and each SM stages only half of the W tile:
//   warp EW     TMA producer   cp.async.bulk.tensor (SWIZZLE_128B): A [128 x 64] + W [BN/CG x 64] bf16 per stage into a
//                              STAGES-deep shared-memory ring; both CTAs' bytes complete on the LEADER's `full` mbarrier
//   warp EW+1   MMA issuer     (leader CTA only) one thread issues tcgen05.mma.cta_group::CG.kind::f16
//                              (M=128*CG, N=BN, K=16) x4 per stage into TMEM; tcgen05.commit (multicast to both CTAs)
//                              frees the stage / publishes the accumulator
//   warp EW+2   TMEM allocator 512 columns = two BN-column fp32 accumulators (double buffered)
//   warps 0..EW-1 epilogue     tcgen05.ld the finished accumulator (lane == output row; four warps per lane quarter
//                              split the 16-column chunks), transpose each chunk through shared memory so the global
//                              side is coalesced, fuse bias / GELU / residual / row-bias, store bf16 or fp32;
//                              overlaps the next tile's main loop.  Residual rows are prefetched one chunk ahead.
//
// Both operands are K-major (row-major activations [M,K] and nn.Linear weights [N,K]), so the same descriptor
// recipe serves A and B.  Tails: TMA zero-fills out-of-bounds rows/columns, the epilogue masks stores.
// CG=1 (single-CTA, M=128 tiles) is kept for small problems and as a cross-check of the paired path.
#include "common.cuh"

#include <mutex>

#include <cuda.h>
#include <cudaTypedefs.h>

namespace tgb {

constexpr int BM = 128;          // rows per CTA
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr uint32_t TMEM_COLS = 512;
constexpr int STG_PITCH = 20;                        // floats per staged row: 16 + 4 pad -> conflict-free 16-B phases

struct GemmParams {
  int M, N, K;
  int epi;
  int num_m_blocks, num_n_blocks, num_k_blocks;     // m blocks of 128*CG rows
  int period, skip;
  const float* bias;
  const float* resid;
  const float* rowbias;
  void* out;
  long long ldo, ldr;
  // LayerNorm fold (see tgb_gemm_bf16_ln): consumer side reads rowstats [M,2] = (mean, rstd) and colsum [N]; the
  // producer side (residual epilogue) also writes a bf16 copy of the new residual rows and per-slot (sum, sum of squares)
  const float* rowstats;
  const float* colsum;
  __nv_bfloat16* out2;
  long long ldo2;
  float2* pstats;
  int nslots;
  int sched_shift;                                  // persistent tile schedule, see tile_at() in the kernel
};

template <int EPI> struct EpiTraits {
  static constexpr bool LN = EPI == TGB_EPI_BF16_LN || EPI == TGB_EPI_BF16_GELU_LN;      // consumes row statistics
  static constexpr bool STATS = EPI == TGB_EPI_F32_RESID_STATS;                          // produces them
  static constexpr bool RESID = EPI == TGB_EPI_F32_RESID || STATS;
  static constexpr bool GELU = EPI == TGB_EPI_BF16_GELU || EPI == TGB_EPI_BF16_GELU_LN;
  static constexpr bool OUT_BF16 = EPI == TGB_EPI_BF16 || EPI == TGB_EPI_BF16_RELU || GELU || LN;
  static constexpr int FAM = OUT_BF16 ? 1 : (RESID ? 2 : 0);
};

// EW = number of epilogue warps (8 or 16: 2 or 4 per TMEM lane quarter).  16 hides the global-memory latency of the
// residual read-modify-write; 8 leaves room for one more pipeline stage.
// Epilogue families (per-warp staging in shared memory):
//   FAM_T    transposing fp32 epilogue (F32, F32_ROWBIAS): padded [32 x 16] fp32 tile per warp
//   FAM_BF16 lane == row, 64-column pieces staged as a SWIZZLE_128B [32 x 64] bf16 box, one bulk tensor store each
//   FAM_RES  lane == row, 16-column pieces: the fp32 residual piece arrives by bulk tensor LOAD into a SWIZZLE_64B
//            [32 x 16] fp32 box (ring of three: loads run two pieces ahead), is updated in place and leaves by bulk
//            tensor store, plus a bf16 copy box for the statistics-producing variant
constexpr int FAM_BF16 = 1, FAM_RES = 2;           // (0 = FAM_T)
// Piece width PW (template): 16 columns = 64-byte row pieces in SWIZZLE_64B boxes, ring of three (the default), or
// 32 columns = whole 128-byte lines per row in SWIZZLE_128B boxes, ring of two (bf16 copy: 64-byte rows, SWIZZLE_64B).
// The wide form halves the number of bulk operations and doubles the contiguous bytes per row in DRAM; it costs one
// operand stage (4 instead of 5).  Used for short main loops (K <= 2048: the attention projection), whose epilogue
// traffic -- not the tensor pipe -- sets the kernel's time.
template <int PW> struct ResCfg {
  static_assert(PW == 16 || PW == 32, "residual piece width");
  static constexpr int NBOX = PW == 16 ? 3 : 2;
  static constexpr int BOX_BYTES = 32 * PW * 4;
  static constexpr int COPY_BYTES = 32 * PW * 2;
  static constexpr int WARP_BYTES = NBOX * BOX_BYTES + 2 * COPY_BYTES;
};
// residual boxes per warp: loads run RES_NBOX-1 pieces ahead.  Measured (tools/gemm_ab.py, profiles/r1_gemm_ab_nbox.txt):
// rings of 2 / 3 / 4 boxes and an extra L2 prefetch of the tile's residual by the TMA producer all land within 2 % of
// each other -- the residual GEMMs are bound by the power their DRAM traffic takes from the tensor cores, not by latency.
constexpr int RES_MAX_NBOX = 3;

template <int BN, int CG, int EW, int FAM, int PW = 16>
struct GemmCfg {
  static constexpr int THREADS = 128 + EW * 32;
  static constexpr int STG_BYTES = FAM == FAM_BF16 ? EW * 4096 : (FAM == FAM_RES ? EW * ResCfg<PW>::WARP_BYTES : EW * 32 * STG_PITCH * 4);
  static constexpr int B_ROWS = BN / CG;                 // W rows staged per CTA
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = B_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_BYTES = 512;
  static constexpr int BUDGET = 227 * 1024 - 1024 - STG_BYTES - BAR_BYTES;
  static constexpr int STAGES = BUDGET / STAGE_BYTES < 8 ? BUDGET / STAGE_BYTES : 8;
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + STG_BYTES + BAR_BYTES;
  static_assert(BN == 256 || BN == 128 || BN == 64, "N tiles are whole 64-column epilogue pieces");
  static_assert(B_ROWS % 8 == 0, "W half tile must be whole swizzle atoms");
  static_assert(STAGE_BYTES % 1024 == 0, "stages must stay 1024-B aligned for SWIZZLE_128B");
  static_assert(2 * BN <= (int)TMEM_COLS, "two accumulators must fit TMEM");
};

// Epilogue math on 4 consecutive columns of one output row (transposed mapping: 4 lanes cover a 16-column row chunk,
// so a warp instruction touches 8 rows x 64 contiguous bytes instead of 32 scattered rows).
template <int EPI>
__device__ __forceinline__ float4 epilogue_math4(float4 acc, const float4& bias4, const float4& res, const float4& cs4,
                                                 float nmean, float rstd) {
  using T = EpiTraits<EPI>;
  // packed fp32 pairs (FFMA2 / FADD2): half the FP32-pipe instructions of the scalar form
  float2 lo = make_float2(acc.x, acc.y), hi = make_float2(acc.z, acc.w);
  if (T::LN) {       // LN(x) W^T + b = rstd * (x W'^T - mean * colsum(W')) + (W beta + b)
    const float2 m2 = make_float2(nmean, nmean), r2 = make_float2(rstd, rstd);
    lo = __ffma2_rn(__ffma2_rn(m2, make_float2(cs4.x, cs4.y), lo), r2, make_float2(bias4.x, bias4.y));
    hi = __ffma2_rn(__ffma2_rn(m2, make_float2(cs4.z, cs4.w), hi), r2, make_float2(bias4.z, bias4.w));
  } else {
    lo = __fadd2_rn(lo, make_float2(bias4.x, bias4.y));
    hi = __fadd2_rn(hi, make_float2(bias4.z, bias4.w));
  }
  if (T::GELU) {
    lo = gelu_erf2(lo);
    hi = gelu_erf2(hi);
  }
  if (T::RESID) {
    lo = __fadd2_rn(lo, make_float2(res.x, res.y));
    hi = __fadd2_rn(hi, make_float2(res.z, res.w));
  }
  float4 v = make_float4(lo.x, lo.y, hi.x, hi.y);
  if (EPI == TGB_EPI_BF16_RELU) {
    v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
  }
  return v;
}

template <int EPI>
__device__ __forceinline__ float4 epilogue_store4(const GemmParams& p, float4 acc, const float4& bias4, const float4& res,
                                                  const float4& cs4, float nmean, float rstd, long long out_row, int col) {
  using T = EpiTraits<EPI>;
  const float4 v = epilogue_math4<EPI>(acc, bias4, res, cs4, nmean, rstd);
  if (T::OUT_BF16) {
    uint2 o;
    o.x = pack_bf16(v.x, v.y);
    o.y = pack_bf16(v.z, v.w);
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + out_row * p.ldo + col) = o;
  } else {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + out_row * p.ldo + col) = v;
  }
  if (T::STATS) {
    uint2 o;
    o.x = pack_bf16(v.x, v.y);
    o.y = pack_bf16(v.z, v.w);
    *reinterpret_cast<uint2*>(p.out2 + out_row * p.ldo2 + col) = o;
  }
  return v;
}

// bf16 epilogue, one 32-column half of a lane's 64-column row piece: bias / LayerNorm fold / activation in packed fp32,
// rounded to 16 bf16x2 registers.  FULL = the whole 64-column piece lies inside N, so the per-column vectors are read at
// immediate offsets from one base pointer; otherwise every 4-column group clamps its index (columns >= N are clipped by
// the tensor store, so their bias may be anything).
template <int EPI, bool FULL>
__device__ __forceinline__ void bf16_half(const GemmParams& p, const uint32_t (&v)[32], uint32_t (&pk)[16], int c0,
                                          float nmean, float rstd) {
  using T = EpiTraits<EPI>;
  const float4* bp = reinterpret_cast<const float4*>(p.bias + (FULL ? c0 : 0));
  const float4* cp = reinterpret_cast<const float4*>(p.colsum + (FULL ? c0 : 0));
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int i = FULL ? q : (min(c0 + q * 4, p.N - 4) >> 2);
    const float4 bias4 = __ldg(bp + i);
    float4 cs4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (T::LN) cs4 = __ldg(cp + i);
    const float4 a = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                                 __uint_as_float(v[4 * q + 3]));
    const float4 r = epilogue_math4<EPI>(a, bias4, bias4, cs4, nmean, rstd);
    pk[2 * q] = pack_bf16(r.x, r.y);
    pk[2 * q + 1] = pack_bf16(r.z, r.w);
  }
}

template <int BN, int EPI, int CG, int EW, int PW = 16>
__global__ void __launch_bounds__(128 + EW * 32, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                         const __grid_constant__ CUtensorMap tma_c, const __grid_constant__ CUtensorMap tma_r,
                         const __grid_constant__ CUtensorMap tma_c2, const GemmParams p) {
  using Cfg = GemmCfg<BN, CG, EW, EpiTraits<EPI>::FAM, PW>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int STG_BYTES = Cfg::STG_BYTES;
  constexpr int RES_NBOX = ResCfg<PW>::NBOX, RES_BOX_BYTES = ResCfg<PW>::BOX_BYTES, RES_COPY_BYTES = ResCfg<PW>::COPY_BYTES;
  constexpr int RES_WARP_BYTES = ResCfg<PW>::WARP_BYTES;
  static_assert(EpiTraits<EPI>::FAM == FAM_RES || PW == 16, "the piece width only exists in the fp32-residual epilogues");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* staging = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + STG_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint64_t* res_full = full_bar + 24;               // FAM_RES: RES_NBOX residual-box barriers per epilogue warp

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // Warp roles.  The SM's issue arbiter favours HIGHER warp ids, so the latency-critical single-thread TMA and MMA
  // issuers sit above the ALU-heavy epilogue warps (ids 0..EW-1) instead of below them.
  constexpr int W_TMA = EW, W_MMA = EW + 1, W_ALLOC = EW + 2;
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x / CG;
  const int num_clusters = gridDim.x / CG;
  const int num_tiles = p.num_m_blocks * p.num_n_blocks;
  // Persistent schedule: in round i the clusters share the tiles [i*C, (i+1)*C) (C = clusters), cluster c taking position
  // (c + i*shift) mod C.  A plain round-robin (shift = 0) gives cluster c the N blocks (c + i*C) mod NB: with NB = 6 and
  // C = 74 (the N = 1408 GEMMs) half of the clusters never see the cheap half-width last N tile and the others see it on
  // a third of their tiles -- the fast half runs ahead by several tiles and the kernel waits for the slow half.  The host
  // searches the shift that minimises the busiest cluster's work (balanced_shift).
  // Row-block mode (sched_shift < 0; chosen by the host when M is large enough to fill whole rounds of row blocks): a
  // cluster walks ALL N tiles of one A row block, then the row block `clusters` further down.  Every A row block is then
  // read by exactly one cluster -- the N tiles after the first hit the L2 -- so no cluster depends on staying in step with
  // another (the tile-granular schedule let the six clusters sharing a row block of the K = 6144 GEMM drift apart until
  // the block had left the L2: 1.6-2x the algorithmic A traffic).  Cluster c starts its walk at N tile c mod NB: at any
  // moment the clusters are spread over all W tiles (all 74 on the same W tile at once measured 6-8 % slower).
  const int sched_shift = p.sched_shift;
  const int n_blocks = p.num_n_blocks;
  const int n_start = cluster_id % n_blocks;
  auto tile_at = [&](int round) {
    if (sched_shift == -2) {
      // Group mode (A/B knob): the grid is a whole number of groups of NB clusters; a group owns one A row block per round
      // and its members rotate through the N tiles, so the clusters that share a row block never change and stay within
      // one half-width tile of each other.
      const int grp = cluster_id / n_blocks, pos = cluster_id - grp * n_blocks;
      return round * num_clusters + grp * n_blocks + (pos + round) % n_blocks;
    }
    if (sched_shift < 0) {
      const int pass = round / n_blocks;
      int n = round - pass * n_blocks + n_start;
      if (n >= n_blocks) n -= n_blocks;
      return (cluster_id + pass * num_clusters) * n_blocks + n;
    }
    return round * num_clusters + (cluster_id + round * sched_shift) % num_clusters;
  };

  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
  }
  if (warp == W_MMA && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);                 // leader's producer arrival (+ the transaction bytes of both CTAs)
      mbar_init(&empty_bar[s], 1);                // one tcgen05.commit arrival (multicast to both CTAs)
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], CG * EW);  // lane 0 of every epilogue warp of every CTA of the pair
    }
    if (EpiTraits<EPI>::FAM == FAM_RES)
      for (int s = 0; s < RES_MAX_NBOX * EW; ++s) mbar_init(&res_full[s], 1);
    fence_mbar_init();
  }
  if (warp == W_ALLOC) tmem_alloc_cg<CG>(tmem_slot, TMEM_COLS);
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == W_TMA) {
    // ------------------------------------------------------------- TMA producer (every CTA loads its halves)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int round = 0, tile; (tile = tile_at(round)) < num_tiles; ++round) {
        const int m_blk = tile / p.num_n_blocks;
        const int n_blk = tile - m_blk * p.num_n_blocks;
        const int a_row = (m_blk * CG + (int)rank) * BM;
        // a last N tile that is at most half full is computed with a half-width MMA: each CTA then supplies
        // cur_bn/CG rows of W (the TMA box still has B_ROWS rows; the extra rows are never read)
        const int cur_bn = (BN == 256 && p.N - n_blk * BN <= BN / 2) ? BN / 2 : BN;
        const int b_row = n_blk * BN + (int)rank * (cur_bn / CG);
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          if (CG == 1) {
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
            tma_load_2d(sa, &tma_a, &full_bar[stage], kb * BK, a_row);
            tma_load_2d(sa + Cfg::A_BYTES, &tma_b, &full_bar[stage], kb * BK, b_row);
          } else {
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], CG * Cfg::STAGE_BYTES);
            const uint32_t bar0 = mapa_u32(smem_u32(&full_bar[stage]), 0);
            tma_load_2d_cg2(sa, &tma_a, bar0, kb * BK, a_row);
            tma_load_2d_cg2(sa + Cfg::A_BYTES, &tma_b, bar0, kb * BK, b_row);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == W_MMA) {
    // ------------------------------------------------------------- MMA issuer (one thread of the leader CTA)
    if (lane == 0 && leader) {
      constexpr uint32_t idesc_full = umma_idesc_bf16(BM * CG, BN);
      constexpr uint32_t idesc_half = umma_idesc_bf16(BM * CG, BN / 2);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int round = 0, tile; (tile = tile_at(round)) < num_tiles; ++round) {
        const int n_blk_mma = tile % p.num_n_blocks;
        const uint32_t idesc = (BN == 256 && p.N - n_blk_mma * BN <= BN / 2) ? idesc_half : idesc_full;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t da = umma_desc_sw128(a_addr + k * UMMA_K * 2);
            const uint64_t db = umma_desc_sw128(b_addr + k * UMMA_K * 2);
            umma_bf16_ss_cg<CG>(d_tmem, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit_cg<CG>(&empty_bar[stage]);     // stage reusable (in both CTAs) once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_cg<CG>(&tmem_full[acc]);         // accumulator complete (both CTAs' epilogues)
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp < EW && EpiTraits<EPI>::OUT_BF16) {
    // ------------------------------------------------------------- bf16 epilogue (EW warps), lane == output row
    // A warp owns 32 rows (its TMEM lane quarter) x 64-column pieces sub, sub+NSUB, ...  Each lane reads its row piece
    // with two 32-column tcgen05.ld, applies bias / LayerNorm fold / GELU in packed fp32, rounds to bf16 and writes the
    // 128-byte row into the warp's SWIZZLE_128B staging box (conflict-free 16-byte units); one elected lane turns the
    // box into a bulk tensor store, which also clips the M / N tails.  ~2 instructions per element outside the
    // activation math (the transposing epilogue below needs ~6: per-row addresses, predicates, LDS + STG).
    using T = EpiTraits<EPI>;
    const int quarter = warp & 3;
    const int sub = warp >> 2;
    constexpr int NSUB = EW / 4;
    constexpr int CW = 64;
    constexpr int NPIECE = BN / CW;
    uint8_t* box = reinterpret_cast<uint8_t*>(staging) + warp * 4096;
    const uint32_t my_row = smem_u32(box) + lane * 128;
    const uint32_t sw = lane & 7;
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t empty0 = (CG == 2) ? mapa_u32(smem_u32(&tmem_empty[0]), 0) : 0u;
    const uint32_t empty1 = (CG == 2) ? mapa_u32(smem_u32(&tmem_empty[1]), 0) : 0u;
    if (lane == 0) tma_prefetch_desc(&tma_c);
    for (int round = 0, tile; (tile = tile_at(round)) < num_tiles; ++round) {
      const int m_blk = tile / p.num_n_blocks;
      const int n_blk = tile - m_blk * p.num_n_blocks;
      const int row_base = (m_blk * CG + (int)rank) * BM + quarter * 32;
      const int col0 = n_blk * BN;
      float nmean = 0.f, rstd = 0.f;
      if (T::LN && row_base + lane < p.M) {
        const float2 st = __ldg(reinterpret_cast<const float2*>(p.rowstats) + (row_base + lane));
        nmean = -st.x;
        rstd = st.y;
      }
      mbar_wait_parked(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * BN);
#pragma unroll 1
      for (int j = sub; j < NPIECE; j += NSUB) {
        const int col = col0 + j * CW;
        if (col >= p.N) break;                       // warp-uniform
        const bool full = col + CW <= p.N;           // warp-uniform
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t v[32];
          tmem_ld_32x32b_x32_nomem(t_row + j * CW + h * 32, v);
          if (h == 0 && lane == 0) bulk_wait_read0();     // the previous store has read the box (overlaps the TMEM load)
          tmem_ld_wait_dep32(v);
          if (h == 0) __syncwarp();
          if (h == 1 && (j + NSUB >= NPIECE || col + NSUB * CW >= p.N)) {
            // that was this warp's last read of the accumulator: hand it back to the MMA issuer before the math
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (CG == 1) mbar_arrive_relaxed(&tmem_empty[acc]);
              else mbar_arrive_cluster_relaxed(acc == 0 ? empty0 : empty1);
            }
          }
          uint32_t pk[16];
          if (full) bf16_half<EPI, true>(p, v, pk, col + h * 32, nmean, rstd);
          else bf16_half<EPI, false>(p, v, pk, col + h * 32, nmean, rstd);
#pragma unroll
          for (int u = 0; u < 4; ++u)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(my_row + (((h * 4 + u) ^ sw) << 4)),
                         "r"(pk[4 * u]), "r"(pk[4 * u + 1]), "r"(pk[4 * u + 2]), "r"(pk[4 * u + 3]) : "memory");
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0 && row_base < p.M) {
          tma_store_2d(&tma_c, box, col, row_base);
          bulk_commit();
        }
      }
      if (sub * CW + col0 >= p.N || sub >= NPIECE) {   // no piece in this tile (narrow N): still release the accumulator
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 1) mbar_arrive_relaxed(&tmem_empty[acc]);
          else mbar_arrive_cluster_relaxed(acc == 0 ? empty0 : empty1);
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) bulk_wait0();                     // every store has landed before the CTA may exit
    __syncwarp();
  } else if (warp < EW && EpiTraits<EPI>::FAM == FAM_RES) {
    // ------------------------------------------------------------- fp32 residual epilogue (EW warps), lane == output row
    // out = acc + bias + resid.  A warp owns 32 rows x the 16-column pieces sub, sub+NSUB, ... of its tiles.  The residual
    // piece is fetched by a bulk tensor load one piece AHEAD (across tile boundaries: it does not depend on the MMA),
    // each lane adds its accumulator row to it in place in shared memory (SWIZZLE_64B: conflict-free 16-byte units) and
    // an elected lane sends the box back with a bulk tensor store -- no per-row global addresses, predicates or LSU
    // wavefronts per row; TMA zero-fills / clips the M and N tails.  The statistics variant also stages the bf16 copy
    // and accumulates the row's (sum, sum of squares) in registers: the lane owns the row, so no shuffles.
    using T = EpiTraits<EPI>;
    const int quarter = warp & 3;
    const int sub = warp >> 2;
    constexpr int NSUB = EW / 4;
    constexpr int NCHUNK = BN / PW;
    uint8_t* wbase = reinterpret_cast<uint8_t*>(staging) + warp * RES_WARP_BYTES;
    const uint32_t box0 = smem_u32(wbase);
    const uint32_t copy0 = box0 + RES_NBOX * RES_BOX_BYTES;
    // 16-byte unit u of the lane's row sits at unit (u ^ sw): SWIZZLE_64B (64-byte rows) xors address bits [4,6) with
    // [7,9), SWIZZLE_128B (128-byte rows) bits [4,7) with [7,10)
    const uint32_t sw = PW == 16 ? ((lane >> 1) & 3) : (lane & 7);
    const uint32_t sw_copy = (lane >> 1) & 3;          // PW == 32: the bf16 copy box has 64-byte rows (SWIZZLE_64B)
    uint64_t* my_full = res_full + RES_NBOX * warp;
    const bool has_bias = p.bias != nullptr;
    const uint32_t empty0 = (CG == 2) ? mapa_u32(smem_u32(&tmem_empty[0]), 0) : 0u;
    const uint32_t empty1 = (CG == 2) ? mapa_u32(smem_u32(&tmem_empty[1]), 0) : 0u;
    if (lane == 0) {
      tma_prefetch_desc(&tma_c);
      tma_prefetch_desc(&tma_r);
      if (T::STATS) tma_prefetch_desc(&tma_c2);
    }
    // piece iterator: (tile, j) with j = sub, sub+NSUB, ... while the piece starts inside N
    auto piece_rows = [&](int tile) { return ((tile / p.num_n_blocks) * CG + (int)rank) * BM + quarter * 32; };
    auto piece_col = [&](int tile, int j) { return (tile % p.num_n_blocks) * BN + j * PW; };
    auto next_piece = [&](int& round, int& tile, int& j) {       // advance; returns false past the end
      j += NSUB;
      if (j >= NCHUNK || piece_col(tile, j) >= p.N) {
        tile = tile_at(++round);
        j = sub;
      }
      return tile < num_tiles;
    };
    auto issue_load = [&](int tile, int j, int b) {  // lane 0
      mbar_arrive_expect_tx(&my_full[b], RES_BOX_BYTES);
      tma_load_2d(wbase + b * RES_BOX_BYTES, &tma_r, &my_full[b], piece_col(tile, j), piece_rows(tile));
    };
    int round = 0, tile = tile_at(0), j = sub;
    // (a piece that starts at or beyond N -- narrow problems, the half-empty last N tile -- is still walked: its load is
    //  zero-filled and its store clipped, which keeps one accumulator hand-shake per tile for every warp)
    int acc = 0;
    uint32_t acc_phase = 0, box_phase = 0;           // box_phase: one parity bit per residual box
    int n = 0, b = 0;                                // piece counter, its box (n % RES_NBOX)
    {
      bool live = tile < num_tiles;
      // load cursor: runs RES_NBOX-1 pieces ahead of the compute cursor
      int lround = round, ltile = tile, lj = j, lb = 0;
      bool lmore = live;
      for (int k = 0; k < RES_NBOX - 1 && lmore; ++k) {
        if (lane == 0) issue_load(ltile, lj, lb);
        lmore = next_piece(lround, ltile, lj);
        lb = lb + 1 == RES_NBOX ? 0 : lb + 1;
      }
      float2 ssum = make_float2(0.f, 0.f), ssq = make_float2(0.f, 0.f);
      while (live) {
        int nround = round, ntile = tile, nj = j;
        const bool more = next_piece(nround, ntile, nj);
        if (lane == 0) bulk_wait_read0();           // the store of piece n-1 has read its box (= the one loaded next)
        if (lmore) {
          if (lane == 0) issue_load(ltile, lj, lb);
          lmore = next_piece(lround, ltile, lj);
          lb = lb + 1 == RES_NBOX ? 0 : lb + 1;
        }
        const int row_base = piece_rows(tile);
        const int col = piece_col(tile, j);
        if (j == sub) {                              // first piece of a tile: the accumulator must be complete
          if (T::STATS && sub > 0 && (p.N & 127)) {
            // the partial last 128-column block of N may hold fewer pieces than there are warps per lane quarter: the slot
            // this warp never reaches is defined as empty (a later real flush by the same thread would overwrite it)
            const int tail0 = p.N & ~127;
            const int tile_col0 = col - j * PW;
            if (tail0 >= tile_col0 && tail0 < tile_col0 + BN && p.N - tail0 <= sub * PW && row_base + lane < p.M)
              p.pstats[(long long)((tail0 >> 7) * NSUB + sub) * p.M + row_base + lane] = make_float2(0.f, 0.f);
          }
          mbar_wait_parked(&tmem_full[acc], acc_phase);
          tc_fence_after();
        }
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * BN);
        uint32_t v[PW];
        if constexpr (PW == 16) tmem_ld_32x32b_x16_nomem(t_row + j * 16, v);
        else tmem_ld_32x32b_x32_nomem(t_row + j * 32, v);
        mbar_wait(&my_full[b], (box_phase >> b) & 1u);
        box_phase ^= 1u << b;
        if constexpr (PW == 16) tmem_ld_wait_dep(v);
        else tmem_ld_wait_dep32(v);
        const bool tile_done = !more || ntile != tile;
        if (tile_done) {     // that was this warp's last read of the accumulator: hand it back before the math and the stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CG == 1) mbar_arrive_relaxed(&tmem_empty[acc]);
            else mbar_arrive_cluster_relaxed(acc == 0 ? empty0 : empty1);
          }
        }
        const uint32_t my_box = box0 + b * RES_BOX_BYTES + lane * (PW * 4);
        uint32_t pk[PW / 2];
#pragma unroll
        for (int u = 0; u < PW / 4; ++u) {
          const uint32_t addr = my_box + ((u ^ sw) << 4);
          float4 r;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
          const int c = min(col + u * 4, p.N - 4);      // columns >= N are clipped by the store
          float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (has_bias) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + c));
          const float4 a = make_float4(__uint_as_float(v[4 * u]), __uint_as_float(v[4 * u + 1]),
                                       __uint_as_float(v[4 * u + 2]), __uint_as_float(v[4 * u + 3]));
          const float4 o = epilogue_math4<EPI>(a, bias4, r, bias4, 0.f, 0.f);
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
          if (T::STATS) {
            pk[2 * u] = pack_bf16(o.x, o.y);
            pk[2 * u + 1] = pack_bf16(o.z, o.w);
            if (col + u * 4 < p.N) {                  // warp-uniform: only real columns enter the statistics
              const float2 lo = make_float2(o.x, o.y), hi = make_float2(o.z, o.w);
              ssum = __fadd2_rn(ssum, __fadd2_rn(lo, hi));
              ssq = __ffma2_rn(lo, lo, __ffma2_rn(hi, hi, ssq));
            }
          }
        }
        if (T::STATS) {
          const uint32_t my_copy = copy0 + (n & 1) * RES_COPY_BYTES + lane * (PW * 2);
#pragma unroll
          for (int w = 0; w < PW / 8; ++w) {            // 16-byte units of the bf16 row (32-byte rows are not swizzled)
            const uint32_t addr = my_copy + ((PW == 16 ? (uint32_t)w : ((uint32_t)w ^ sw_copy)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * w]), "r"(pk[4 * w + 1]), "r"(pk[4 * w + 2]), "r"(pk[4 * w + 3]) : "memory");
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0 && row_base < p.M) {
          tma_store_2d(&tma_c, wbase + b * RES_BOX_BYTES, col, row_base);
          if (T::STATS) tma_store_2d(&tma_c2, wbase + RES_NBOX * RES_BOX_BYTES + (n & 1) * RES_COPY_BYTES, col, row_base);
        }
        if (lane == 0) bulk_commit();
        if (T::STATS) {
          // One (sum, sum of squares) slot per (128-column block of N, warp of the lane quarter), written when this warp's
          // walk leaves the block.  The partition does not depend on the tile width, so 256- and 128-wide tiles -- large
          // and small passes -- produce bit-identical statistics.
          if (tile_done || (piece_col(ntile, nj) >> 7) != (col >> 7)) {
            if (row_base + lane < p.M && col < p.N)
              p.pstats[(long long)((col >> 7) * NSUB + sub) * p.M + row_base + lane] =      // slot-major: 32 rows = 256 B
                  make_float2(ssum.x + ssum.y, ssq.x + ssq.y);
            ssum = make_float2(0.f, 0.f);
            ssq = make_float2(0.f, 0.f);
          }
        }
        if (tile_done) {
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        round = nround;
        tile = ntile;
        j = nj;
        live = more;
        ++n;
        b = b + 1 == RES_NBOX ? 0 : b + 1;
      }
    }
    if (lane == 0) bulk_wait0();                     // every store has landed before the CTA may exit
    __syncwarp();
  } else if (warp < EW) {
    // ------------------------------------------------------------- fp32 epilogue (EW warps), transposing
    // TMEM hands every thread one ROW (lane == row) of a 16-column chunk.  The chunk is transposed through a padded
    // per-warp shared-memory tile so that the global side is coalesced: thread -> (row = i*8 + lane/4, 4 columns at
    // (lane%4)*4), i.e. 64 contiguous bytes per row and 8 rows per instruction for the residual read and the store.
    using T = EpiTraits<EPI>;
    const int ew = warp;
    const int quarter = warp & 3;                  // TMEM lane quarter this warp may read
    const int sub = ew >> 2;                       // which 16-column chunks: sub, sub+NSUB, ...
    constexpr int NSUB = EW / 4;
    constexpr int NCHUNK = BN / 16;
    float* stg = staging + ew * (32 * STG_PITCH);
    const int trow = lane >> 2;
    const int tcol = (lane & 3) * 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t empty0 = (CG == 2) ? mapa_u32(smem_u32(&tmem_empty[0]), 0) : 0u;
    const uint32_t empty1 = (CG == 2) ? mapa_u32(smem_u32(&tmem_empty[1]), 0) : 0u;
    for (int round = 0, tile; (tile = tile_at(round)) < num_tiles; ++round) {
      const int m_blk = tile / p.num_n_blocks;
      const int n_blk = tile - m_blk * p.num_n_blocks;
      const int row_base = (m_blk * CG + (int)rank) * BM + quarter * 32;
      const int col0 = n_blk * BN;
      long long orow[4];
      int rbrow[4];
      bool rok[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = row_base + i * 8 + trow;
        rok[i] = row < p.M;
        orow[i] = row;
        rbrow[i] = 0;
        if (EPI == TGB_EPI_F32_ROWBIAS) {
          const int grp = row / p.period;
          rbrow[i] = row - grp * p.period;
          orow[i] = (long long)grp * (p.period + p.skip) + p.skip + rbrow[i];
        }
      }
      // residual prefetch, one chunk ahead (independent of the MMA: the first one is issued before the tile is ready)
      float4 rnext[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) rnext[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (T::RESID) {
        const int c = col0 + sub * 16 + tcol;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (rok[i] && c < p.N) rnext[i] = *reinterpret_cast<const float4*>(p.resid + (long long)(row_base + i * 8 + trow) * p.ldr + c);
      }
      // LayerNorm fold: (mean, rstd) of this thread's four rows (consumer) / running partial statistics (producer)
      float rmean[4], rrstd[4], ssum[4], ssq[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        rmean[i] = 0.f; rrstd[i] = 0.f; ssum[i] = 0.f; ssq[i] = 0.f;
        if (T::LN && rok[i]) {
          const float2 st = __ldg(reinterpret_cast<const float2*>(p.rowstats) + (row_base + i * 8 + trow));
          rmean[i] = -st.x;
          rrstd[i] = st.y;
        }
      }
      mbar_wait_parked(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * BN);
#pragma unroll 1
      for (int j = sub; j < NCHUNK; j += NSUB) {
        const int col = col0 + j * 16;
        if (col >= p.N) break;                       // warp-uniform
        uint32_t v[16];
        tmem_ld_32x32b_x16_nomem(t_row + j * 16, v);
        float4 rcur[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) rcur[i] = rnext[i];
        if (T::RESID) {
          const int cn = col + NSUB * 16 + tcol;
          const bool more = (j + NSUB < NCHUNK) && (cn < p.N);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (more && rok[i]) rnext[i] = *reinterpret_cast<const float4*>(p.resid + (long long)(row_base + i * 8 + trow) * p.ldr + cn);
        }
        tmem_ld_wait_dep(v);
        float4* srow = reinterpret_cast<float4*>(stg + lane * STG_PITCH);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          srow[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                                __uint_as_float(v[4 * q + 3]));
        __syncwarp();
        const int c = col + tcol;
        const bool cok = c < p.N;
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (EPI != TGB_EPI_F32_ROWBIAS && p.bias != nullptr && cok) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + c));
        float4 cs4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (T::LN && cok) cs4 = __ldg(reinterpret_cast<const float4*>(p.colsum + c));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 a = *reinterpret_cast<const float4*>(stg + (i * 8 + trow) * STG_PITCH + tcol);
          if (rok[i] && cok) {
            if (EPI == TGB_EPI_F32_ROWBIAS) bias4 = __ldg(reinterpret_cast<const float4*>(p.rowbias + (long long)rbrow[i] * p.N + c));
            const float4 r = epilogue_store4<EPI>(p, a, bias4, rcur[i], cs4, rmean[i], rrstd[i], orow[i], c);
            if (T::STATS) {
              ssum[i] += (r.x + r.y) + (r.z + r.w);
              ssq[i] += fmaf(r.x, r.x, r.y * r.y) + fmaf(r.z, r.z, r.w * r.w);
            }
          }
        }
        __syncwarp();
      }
      if (T::STATS) {
        // this warp's share of the tile (chunks sub, sub+NSUB, ...) -> one (sum, sum of squares) slot per row; every
        // slot of every row is written exactly once per launch, in a fixed order: deterministic, no atomics
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          ssum[i] += __shfl_xor_sync(0xffffffffu, ssum[i], 1);
          ssq[i] += __shfl_xor_sync(0xffffffffu, ssq[i], 1);
          ssum[i] += __shfl_xor_sync(0xffffffffu, ssum[i], 2);
          ssq[i] += __shfl_xor_sync(0xffffffffu, ssq[i], 2);
          if ((lane & 3) == 0 && rok[i])
            p.pstats[(long long)(n_blk * NSUB + sub) * p.M + row_base + i * 8 + trow] = make_float2(ssum[i], ssq[i]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 1) mbar_arrive_relaxed(&tmem_empty[acc]);
        else mbar_arrive_cluster_relaxed(acc == 0 ? empty0 : empty1);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == W_ALLOC) {
    tc_fence_after();
    tmem_dealloc_cg<CG>(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------ host side
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
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

// 2-D bf16 row-major [rows, cols] (leading dim ld elements) -> tensor map with a [box_rows x 64] SWIZZLE_128B box.
// A tensor map is a pure function of (base, extents, leading dimension, box / type / swizzle): the ~780 GEMM launches of a
// step use a few dozen distinct ones (the activation buffers are reused by every layer), so each host thread keeps the
// encoded descriptors in a small direct-mapped cache instead of calling cuTensorMapEncodeTiled up to five times per
// launch.  `kind` distinguishes the builders and their box parameters.
struct TmapSlot {
  const void* base;
  long long rows, cols, ld;
  int kind;
  bool used;
  CUtensorMap map;
};
constexpr int TMAP_CACHE_SLOTS = 256;
static TmapSlot* tmap_slot(const void* base, long long rows, long long cols, long long ld, int kind, bool* hit) {
  static thread_local TmapSlot cache[TMAP_CACHE_SLOTS] = {};
  unsigned long long h = reinterpret_cast<unsigned long long>(base) >> 4;
  h ^= (unsigned long long)rows * 0x9E3779B97F4A7C15ull;
  h ^= (unsigned long long)cols * 0xC2B2AE3D27D4EB4Full + (unsigned long long)ld * 0x165667B19E3779F9ull + (unsigned)kind * 0x27D4EB2Full;
  h ^= h >> 29;
  TmapSlot* s = &cache[h % TMAP_CACHE_SLOTS];
  *hit = s->used && s->base == base && s->rows == rows && s->cols == cols && s->ld == ld && s->kind == kind;
  return s;
}
template <typename Build>
static int cached_tmap(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int kind, Build&& build) {
  bool hit;
  TmapSlot* s = tmap_slot(base, rows, cols, ld, kind, &hit);
  if (hit) {
    *map = s->map;
    return TGB_OK;
  }
  const int st = build(map);
  if (st != TGB_OK) return st;
  s->base = base; s->rows = rows; s->cols = cols; s->ld = ld; s->kind = kind; s->used = true;
  s->map = *map;
  return TGB_OK;
}

static int make_tmap_uncached(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows) {
  auto fn = get_encode_fn();
  if (!fn) {
    set_last_error("cuTensorMapEncodeTiled not available from the driver");
    return TGB_ECUDA;
  }
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed (CUresult %d) rows=%lld cols=%lld ld=%lld box_rows=%d", (int)r, rows,
                   cols, ld, box_rows);
    return TGB_ECUDA;
  }
  return TGB_OK;
}

static int make_tmap(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows) {
  return cached_tmap(map, base, rows, cols, ld, 0x1000 + box_rows,
                     [&](CUtensorMap* m) { return make_tmap_uncached(m, base, rows, cols, ld, box_rows); });
}

// bf16 output [M, N] (leading dim ldo) -> tensor map with the epilogue's [32 rows x 64 columns] SWIZZLE_128B store box
static int make_tmap_out_uncached(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld) {
  auto fn = get_encode_fn();
  if (!fn) {
    set_last_error("cuTensorMapEncodeTiled not available from the driver");
    return TGB_ECUDA;
  }
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {64u, 32u};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled (output) failed (CUresult %d) rows=%lld cols=%lld ld=%lld", (int)r, rows, cols, ld);
    return TGB_ECUDA;
  }
  return TGB_OK;
}

static int make_tmap_out(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld) {
  return cached_tmap(map, base, rows, cols, ld, 0x2000,
                     [&](CUtensorMap* m) { return make_tmap_out_uncached(m, base, rows, cols, ld); });
}

// the fp32 residual epilogue's boxes: [32 rows x pw columns], fp32 or bf16 copy (row bytes 32: no swizzle, 64 / 128: swizzled)
static int make_tmap_box16_uncached(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, bool f32, int pw) {
  auto fn = get_encode_fn();
  if (!fn) {
    set_last_error("cuTensorMapEncodeTiled not available from the driver");
    return TGB_ECUDA;
  }
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * (f32 ? 4 : 2)};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(pw), 32u};
  cuuint32_t estr[2] = {1, 1};
  const int row_bytes = pw * (f32 ? 4 : 2);           // 32 -> no swizzle, 64 -> SWIZZLE_64B, 128 -> SWIZZLE_128B
  const CUtensorMapSwizzle swz = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE);
  CUresult r = fn(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base),
                  dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled (residual box) failed (CUresult %d) rows=%lld cols=%lld ld=%lld", (int)r, rows, cols, ld);
    return TGB_ECUDA;
  }
  return TGB_OK;
}

static int make_tmap_box16(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, bool f32, int pw = 16) {
  return cached_tmap(map, base, rows, cols, ld, 0x3000 + pw * 2 + (f32 ? 1 : 0),
                     [&](CUtensorMap* m) { return make_tmap_box16_uncached(m, base, rows, cols, ld, f32, pw); });
}

// the bf16 epilogues read their bias unconditionally: a call without one gets this block of zeros
constexpr int ZERO_BIAS_N = 8192;
__device__ float g_zero_bias[ZERO_BIAS_N];

// Shift of the persistent schedule (tile_at in the kernel) that minimises the busiest cluster's work.  Only a half-width
// last N tile makes tiles unequal (it costs ~0.65 of a full one: half the MMA time, the same A traffic); without one every
// shift is equally good.  Exhaustive search over the shifts with that cost model, cached per problem shape.
static int balanced_shift(int clusters, int m_blocks, int n_blocks, bool half_last) {
  const long long tiles = (long long)m_blocks * n_blocks;
  if (!half_last || tiles <= clusters || clusters <= 1) return 0;
  struct Key { int c, m, n; };
  static Key keys[16];
  static int vals[16];
  static int used = 0;
  static std::mutex mu;                                     // the cache is shared by every host thread / stream
  std::lock_guard<std::mutex> lock(mu);
  for (int i = 0; i < used; ++i)
    if (keys[i].c == clusters && keys[i].m == m_blocks && keys[i].n == n_blocks) return vals[i];
  const int rounds = (int)((tiles + clusters - 1) / clusters);
  int best = 0;
  long long best_cost = -1;
  for (int s = 0; s < clusters; ++s) {
    long long worst = 0;
    for (int c = 0; c < clusters && (best_cost < 0 || worst < best_cost); ++c) {
      long long cost = 0;                                   // in units of 1/20 tile: full = 20, half = 13
      for (int r = 0; r < rounds; ++r) {
        const long long t = (long long)r * clusters + (c + (long long)r * s) % clusters;
        if (t >= tiles) break;
        cost += (t % n_blocks == n_blocks - 1) ? 13 : 20;
      }
      if (cost > worst) worst = cost;
    }
    if (best_cost < 0 || worst < best_cost) { best_cost = worst; best = s; }
  }
  const int slot = used < 16 ? used++ : 15;
  keys[slot] = Key{clusters, m_blocks, n_blocks};
  vals[slot] = best;
  return best;
}

static std::atomic<int> g_force_cg{0};     // 0 = choose, 1 / 2 = force (tests exercise both paths)

static std::atomic<int> g_force_ew{0};     // 0 = choose per epilogue, 8 / 16 = force
static std::atomic<int> g_force_resid_piece{0};   // 0 = choose by K, 16 / 32 = force the residual piece width (A/B knob)
// 1 = balanced tile schedule (default); 0 = plain round-robin; 2 = row-block schedule when M fills whole rounds of row
// blocks, else 1 (A/B knob).  Measured (tools/gemm_ab.py, profiles/r2_gemm_ab_rowblocks.txt): the row-block schedule
// removes the A re-reads between clusters but puts 74 A row blocks (230 MB for K = 6144) in flight at once, more than the
// L2 holds, so each cluster's own five re-reads of its block miss instead: fc2 -5..8 %, qkv -6 %, proj -2 %.  Not default.
static std::atomic<int> g_sched_balanced{1};

// Row-block schedule: ceil(m_blocks / clusters) passes of whole row blocks; taken when the last pass leaves at most 3 %
// of the cluster-passes idle (config 2: 514 row blocks on 74 clusters = 6.95 passes, 99.2 %).
static bool row_block_schedule_ok(int clusters, int m_blocks) {
  if (clusters <= 0 || m_blocks < clusters) return false;
  const long long passes = (m_blocks + clusters - 1) / clusters;
  return (long long)m_blocks * 100 >= passes * clusters * 97;
}

template <int BN, int EPI, int CG, int EW, int PW = 16>
static int launch_gemm_ew(const void* A, long long lda, const void* W, long long ldw, GemmParams p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, CG, EW, EpiTraits<EPI>::FAM, PW>;
  p.num_m_blocks = (p.M + BM * CG - 1) / (BM * CG);
  p.num_n_blocks = (p.N + BN - 1) / BN;
  p.num_k_blocks = (p.K + BK - 1) / BK;
  CUtensorMap ta, tb;
  int st = make_tmap(&ta, A, p.M, p.K, lda, BM);
  if (st != TGB_OK) return st;
  st = make_tmap(&tb, W, p.N, p.K, ldw, Cfg::B_ROWS);
  if (st != TGB_OK) return st;
  CUtensorMap tc = ta, tr = ta, tc2 = ta;      // placeholders where an epilogue family does not use them
  if (EpiTraits<EPI>::OUT_BF16) {
    st = make_tmap_out(&tc, p.out, p.M, p.N, p.ldo);
    if (st != TGB_OK) return st;
  }
  if (EpiTraits<EPI>::RESID) {
    st = make_tmap_box16(&tc, p.out, p.M, p.N, p.ldo, true, PW);
    if (st != TGB_OK) return st;
    st = make_tmap_box16(&tr, p.resid, p.M, p.N, p.ldr, true, PW);
    if (st != TGB_OK) return st;
    if (EpiTraits<EPI>::STATS) {
      st = make_tmap_box16(&tc2, p.out2, p.M, p.N, p.ldo2, false, PW);
      if (st != TGB_OK) return st;
    }
  }
  auto kern = gemm_bf16_tcgen05_kernel<BN, EPI, CG, EW, PW>;
  TGB_ONCE_PER_DEVICE(TGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES)));
  const int tiles = p.num_m_blocks * p.num_n_blocks;
  const int max_clusters = sm_count() / CG;
  int clusters = tiles < max_clusters ? tiles : max_clusters;
  const int sched = g_sched_balanced.load();
  p.sched_shift = sched ? balanced_shift(clusters, p.num_m_blocks, p.num_n_blocks,
                                         BN == 256 && p.N - (p.num_n_blocks - 1) * BN <= BN / 2) : 0;
  if (sched == 2 && row_block_schedule_ok(clusters, p.num_m_blocks)) p.sched_shift = -1;
  if (sched == 3 && p.num_n_blocks >= 2 && clusters >= p.num_n_blocks && tiles >= 4 * clusters) {
    clusters = clusters / p.num_n_blocks * p.num_n_blocks;
    p.sched_shift = -2;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * CG);
  cfg.blockDim = dim3(Cfg::THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  TGB_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, tc, tr, tc2, p));
  return TGB_OK;
}

template <int BN, int EPI, int CG>
static int launch_gemm(const void* A, long long lda, const void* W, long long ldw, const GemmParams& p, cudaStream_t stream) {
  // 8 epilogue warps by default (measured with tools/gemm_ab.py: the bias / GELU epilogues are marginally better with 8
  // warps and the extra pipeline stage than with 16)
  const int force_ew = g_force_ew.load();
  int ew = force_ew ? force_ew : 8;
  if constexpr (EpiTraits<EPI>::RESID) {       // 8 epilogue warps: the statistics slot layout (2 per 128-column block of N) is fixed
    // wide residual pieces (whole 128-byte lines) for short main loops, where the epilogue traffic sets the time
    const int force_pw = g_force_resid_piece.load();
    const int pw = force_pw ? force_pw : (p.K <= 2048 ? 32 : 16);
    if constexpr (BN >= 128) {       // the piece width must not depend on the tile width: it decides which columns a slot sums
      if (pw == 32) return launch_gemm_ew<BN, EPI, CG, 8, 32>(A, lda, W, ldw, p, stream);
    }
    return launch_gemm_ew<BN, EPI, CG, 8, 16>(A, lda, W, ldw, p, stream);
  } else {
    if (ew == 16) return launch_gemm_ew<BN, EPI, CG, 16>(A, lda, W, ldw, p, stream);
    return launch_gemm_ew<BN, EPI, CG, 8>(A, lda, W, ldw, p, stream);
  }
}

template <int BN, int CG>
static int dispatch_epi(const void* A, long long lda, const void* W, long long ldw, const GemmParams& p, cudaStream_t s) {
  switch (p.epi) {
    case TGB_EPI_BF16: return launch_gemm<BN, TGB_EPI_BF16, CG>(A, lda, W, ldw, p, s);
    case TGB_EPI_BF16_GELU: return launch_gemm<BN, TGB_EPI_BF16_GELU, CG>(A, lda, W, ldw, p, s);
    case TGB_EPI_F32: return launch_gemm<BN, TGB_EPI_F32, CG>(A, lda, W, ldw, p, s);
    case TGB_EPI_F32_RESID: return launch_gemm<BN, TGB_EPI_F32_RESID, CG>(A, lda, W, ldw, p, s);
    case TGB_EPI_F32_ROWBIAS: return launch_gemm<BN, TGB_EPI_F32_ROWBIAS, CG>(A, lda, W, ldw, p, s);
    case TGB_EPI_BF16_RELU: return launch_gemm<BN, TGB_EPI_BF16_RELU, CG>(A, lda, W, ldw, p, s);
    case TGB_EPI_BF16_LN: return launch_gemm<BN, TGB_EPI_BF16_LN, CG>(A, lda, W, ldw, p, s);
    case TGB_EPI_BF16_GELU_LN: return launch_gemm<BN, TGB_EPI_BF16_GELU_LN, CG>(A, lda, W, ldw, p, s);
    case TGB_EPI_F32_RESID_STATS: return launch_gemm<BN, TGB_EPI_F32_RESID_STATS, CG>(A, lda, W, ldw, p, s);
  }
  set_last_error("unknown epilogue %d", p.epi);
  return TGB_EINVAL;
}

template <int CG>
static int dispatch_bn(int bn, const void* A, long long lda, const void* W, long long ldw, const GemmParams& p,
                       cudaStream_t s) {
  switch (bn) {
    case 256: return dispatch_epi<256, CG>(A, lda, W, ldw, p, s);
    case 128: return dispatch_epi<128, CG>(A, lda, W, ldw, p, s);
    case 64: return dispatch_epi<64, CG>(A, lda, W, ldw, p, s);
  }
  set_last_error("gemm: unsupported N tile %d (256/128/64)", bn);
  return TGB_EINVAL;
}

// Tile choice: minimise (waves of tiles) x (time of one wave).  Measured (tools/gemm_small_m.py,
// profiles/r2_gemm_small_m.txt; qkv / fc1 shapes, M = 1028 .. 16448): a wave of 128-wide tiles takes 0.86 of a wave of
// 256-wide ones -- half the MMA time, but the same A traffic from the L2, which is what a narrow tile is bound by -- so a
// narrow tile only pays when it saves whole waves; single-CTA 128-row tiles cost 5-10 % more per wave than CTA pairs and win
// exactly when they need fewer waves (M = 2056: 2 instead of 3).  The choice never changes a result: every output
// element accumulates over K in the same order whatever the tile.
static void choose_tile(int M, int N, int* bn_out, int* cg_out, int min_bn = 64) {
  // measured on B200: a tcgen05.mma with N=176 costs as much as N=256 (N is processed in 128-wide passes), so the
  // tiles are 256 / 128 / 64 wide; prefer 256, then 128.
  const int cands[4] = {256, 128, 64, 64};
  int best_bn = 256, best_cg = 2;
  double best_cost = -1.0;
  for (int cg = 2; cg >= 1; --cg) {
    const int force_cg = g_force_cg.load();
    if (force_cg && cg != force_cg) continue;
    const long long mblocks = (M + BM * cg - 1) / (BM * cg);
    for (int i = 0; i < 4; ++i) {
      const int bn = cands[i];
      if (bn < min_bn) continue;
      const long long nblocks = (N + bn - 1) / bn;
      const long long tiles = nblocks * mblocks;
      const long long slots = sm_count() / cg;
      const long long waves = (tiles + slots - 1) / slots;
      double eff = bn >= 256 ? 1.0 : (bn == 128 ? 0.58 : 0.27);
      if (cg == 1) eff *= 0.93;
      const double cost = (double)waves * bn / eff;
      if (best_cost < 0 || cost < best_cost * 0.999) { best_cost = cost; best_bn = bn; best_cg = cg; }
    }
  }
  *bn_out = best_bn;
  *cg_out = best_cg;
}

}  // namespace tgb

using namespace tgb;

extern "C" int tgb_gemm_set_resid_piece(int width) {
  if (width != 0 && width != 16 && width != 32) {
    set_last_error("residual piece width must be 0 (auto), 16 or 32");
    return TGB_EINVAL;
  }
  g_force_resid_piece = width;
  return TGB_OK;
}

extern "C" int tgb_gemm_set_epilogue_warps(int warps) {
  if (warps != 0 && warps != 8 && warps != 16) {
    set_last_error("epilogue warps must be 0 (auto), 8 or 16");
    return TGB_EINVAL;
  }
  g_force_ew = warps;
  return TGB_OK;
}

extern "C" int tgb_gemm_set_schedule(int mode) {
  if (mode < 0 || mode > 3) {
    set_last_error("schedule must be 0 (round-robin), 1 (balanced tiles, default), 2 (row blocks when M allows) or 3 (fixed groups)");
    return TGB_EINVAL;
  }
  g_sched_balanced = mode;
  return TGB_OK;
}

extern "C" int tgb_gemm_set_cta_group(int cta_group) {
  if (cta_group < 0 || cta_group > 2) {
    set_last_error("cta_group must be 0 (auto), 1 or 2");
    return TGB_EINVAL;
  }
  g_force_cg = cta_group;
  return TGB_OK;
}

static int gemm_entry(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K, int epilogue,
                      const float* bias, const float* resid, int64_t ldr, const float* rowbias, int period, int skip,
                      void* out, int64_t ldo, int bn, const float* rowstats, const float* colsum, void* out_bf16,
                      int64_t ldo2, float* partial_stats, tgb_stream_t stream) {
  int st = check_device();
  if (st != TGB_OK) return st;
  TGB_CHECK_ARG(A && W && out, "gemm: null pointer");
  TGB_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: bad shape M=%d N=%d K=%d", M, N, K);
  TGB_CHECK_ARG(N % 8 == 0 && K % 8 == 0, "gemm: N (%d) and K (%d) must be multiples of 8", N, K);
  TGB_CHECK_ARG(lda % 8 == 0 && ldw % 8 == 0 && lda >= K && ldw >= K, "gemm: lda/ldw must be multiples of 8 and >= K");
  TGB_CHECK_ARG(((uintptr_t)A % 16) == 0 && ((uintptr_t)W % 16) == 0 && ((uintptr_t)out % 16) == 0,
                "gemm: A, W, out must be 16-byte aligned");
  TGB_CHECK_ARG(ldo % 8 == 0 && ldo >= N, "gemm: ldo must be a multiple of 8 and >= N");
  TGB_CHECK_ARG(bias == nullptr || ((uintptr_t)bias % 16) == 0, "gemm: bias must be 16-byte aligned");
  if (epilogue == TGB_EPI_F32_RESID || epilogue == TGB_EPI_F32_RESID_STATS)
    TGB_CHECK_ARG(resid && ldr % 4 == 0 && ldr >= N && ((uintptr_t)resid % 16) == 0, "gemm: bad residual");
  if (epilogue == TGB_EPI_F32_ROWBIAS)
    TGB_CHECK_ARG(rowbias && period > 0 && skip >= 0 && ((uintptr_t)rowbias % 16) == 0, "gemm: bad rowbias");
  if (epilogue == TGB_EPI_BF16_LN || epilogue == TGB_EPI_BF16_GELU_LN)
    TGB_CHECK_ARG(rowstats && colsum && ((uintptr_t)rowstats % 8) == 0 && ((uintptr_t)colsum % 16) == 0,
                  "gemm: the LayerNorm-fold epilogues need rowstats [M,2] and colsum [N]");
  if (epilogue == TGB_EPI_F32_RESID_STATS)
    TGB_CHECK_ARG(out_bf16 && partial_stats && ldo2 % 8 == 0 && ldo2 >= N && ((uintptr_t)out_bf16 % 16) == 0 &&
                      ((uintptr_t)partial_stats % 8) == 0,
                  "gemm: the statistics epilogue needs out_bf16 (ld %% 8 == 0) and partial_stats");
  int auto_bn, cg;
  // the statistics slots are (128-column block of N, epilogue warp) pairs whatever the tile: that partition -- and with it
  // a row's statistics and the whole encoder -- is bit-identical whatever batch the row is encoded in, and a small pass may
  // still take 128-wide tiles to spread over more SMs (64-wide ones would split a slot's pieces between tiles)
  choose_tile(M, N, &auto_bn, &cg, epilogue == TGB_EPI_F32_RESID_STATS ? 128 : 64);
  if (bn == 0) bn = auto_bn;
  if (epilogue == TGB_EPI_F32_RESID_STATS && bn < 128) bn = 128;
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.epi = epilogue;
  p.num_m_blocks = p.num_n_blocks = p.num_k_blocks = 0;
  p.period = period > 0 ? period : 1; p.skip = skip;
  p.bias = bias; p.resid = resid; p.rowbias = rowbias; p.out = out; p.ldo = ldo; p.ldr = ldr;
  p.rowstats = rowstats; p.colsum = colsum; p.out2 = reinterpret_cast<__nv_bfloat16*>(out_bf16); p.ldo2 = ldo2;
  p.pstats = reinterpret_cast<float2*>(partial_stats);
  p.nslots = ((N + 127) / 128) * 2;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (cg == 2) return dispatch_bn<2>(bn, A, lda, W, ldw, p, s);
  return dispatch_bn<1>(bn, A, lda, W, ldw, p, s);
}

extern "C" int tgb_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K, int epilogue,
                             const float* bias, const float* resid, int64_t ldr, const float* rowbias, int period,
                             int skip, void* out, int64_t ldo, int bn, tgb_stream_t stream) {
  if (epilogue == TGB_EPI_BF16_LN || epilogue == TGB_EPI_BF16_GELU_LN || epilogue == TGB_EPI_F32_RESID_STATS) {
    set_last_error("gemm: epilogue %d takes the extra operands of tgb_gemm_bf16_ln", epilogue);
    return TGB_EINVAL;
  }
  return gemm_entry(A, lda, W, ldw, M, N, K, epilogue, bias, resid, ldr, rowbias, period, skip, out, ldo, bn, nullptr,
                    nullptr, nullptr, 0, nullptr, stream);
}

extern "C" int tgb_gemm_ln_stats_slots(int M, int N) {
  if (M <= 0 || N <= 0) return 0;
  return ((N + 127) / 128) * 2;
}

extern "C" int tgb_gemm_bf16_ln(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K, int epilogue,
                                const float* bias, const float* resid, int64_t ldr, const float* rowstats,
                                const float* colsum, void* out, int64_t ldo, void* out_bf16, int64_t ldo2,
                                float* partial_stats, tgb_stream_t stream) {
  if (epilogue != TGB_EPI_BF16_LN && epilogue != TGB_EPI_BF16_GELU_LN && epilogue != TGB_EPI_F32_RESID_STATS) {
    set_last_error("gemm_ln: epilogue must be TGB_EPI_BF16_LN, TGB_EPI_BF16_GELU_LN or TGB_EPI_F32_RESID_STATS");
    return TGB_EINVAL;
  }
  return gemm_entry(A, lda, W, ldw, M, N, K, epilogue, bias, resid, ldr, nullptr, 0, 0, out, ldo, 0, rowstats, colsum,
                    out_bf16, ldo2, partial_stats, stream);
}
