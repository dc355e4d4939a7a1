// bf16 GEMM with fused epilogues on the 5th-gen tensor cores:  out = epi(A[M,K] * W[N,K]^T)
//
// Persistent, warp-specialised kernel, one CTA per SM:
//   warp 0     TMA producer   cp.async.bulk.tensor (SWIZZLE_128B) of A [128 x 64] and W [BN x 64] bf16 tiles into a
//                             STAGES-deep shared-memory ring, completion on `full` mbarriers
//   warp 1     MMA issuer     one thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16) x4 per
//                             stage, accumulating in TMEM; tcgen05.commit frees the stage / publishes the tile
//   warp 2     TMEM allocator 512 columns = two BN-column fp32 accumulators (double buffered)
//   warps 4-7  epilogue       tcgen05.ld the finished accumulator (lane == output row), fuse bias / GELU /
//                             residual / row-bias, store bf16 or fp32; overlaps the next tile's main loop
//
// Both operands are K-major (row-major activations [M,K] and nn.Linear weights [N,K]), so the same descriptor
// recipe serves A and B.  Tails: TMA zero-fills out-of-bounds rows/columns, the epilogue masks stores.
#include "common.cuh"

#include <cuda.h>
#include <cudaTypedefs.h>

namespace tgb {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 256;
constexpr uint32_t TMEM_COLS = 512;

struct GemmParams {
  int M, N, K;
  int epi;
  int num_m_blocks, num_n_blocks, num_k_blocks;
  int period, skip;
  const float* bias;
  const float* resid;
  const float* rowbias;
  void* out;
  long long ldo, ldr;
};

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (196 * 1024) / STAGE_BYTES < 8 ? (196 * 1024) / STAGE_BYTES : 8;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + BAR_BYTES;
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N constraint for M=128");
  static_assert(STAGE_BYTES % 1024 == 0, "stages must stay 1024-B aligned for SWIZZLE_128B");
  static_assert(2 * BN <= (int)TMEM_COLS, "two accumulators must fit TMEM");
};

template <int EPI>
__device__ __forceinline__ void epilogue_store16(const GemmParams& p, const uint32_t (&acc)[16], long long out_row,
                                                 int row, int rb_row, int col, int ncols) {
  // ncols in {8,16}: number of valid columns in this 16-wide chunk (N % 8 == 0).
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[j]);
  if (EPI == TGB_EPI_F32_ROWBIAS) {
    const float4* rb = reinterpret_cast<const float4*>(p.rowbias + (long long)rb_row * p.N + col);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j * 4 < ncols) {
        const float4 b = __ldg(rb + j);
        v[4 * j + 0] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
      }
    }
  } else if (p.bias != nullptr) {
    const float4* bp = reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j * 4 < ncols) {
        const float4 b = __ldg(bp + j);
        v[4 * j + 0] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
      }
    }
  }
  if (EPI == TGB_EPI_BF16_GELU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = gelu_erf(v[j]);
  }
  if (EPI == TGB_EPI_F32_RESID) {
    const float4* rp = reinterpret_cast<const float4*>(p.resid + (long long)row * p.ldr + col);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j * 4 < ncols) {
        const float4 r = rp[j];
        v[4 * j + 0] += r.x; v[4 * j + 1] += r.y; v[4 * j + 2] += r.z; v[4 * j + 3] += r.w;
      }
    }
  }
  if (EPI == TGB_EPI_BF16 || EPI == TGB_EPI_BF16_GELU) {
    uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + out_row * p.ldo + col);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (j * 8 < ncols) {
        uint4 o;
        o.x = pack_bf16(v[8 * j + 0], v[8 * j + 1]);
        o.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
        o.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]);
        o.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
        op[j] = o;
      }
    }
  } else {
    float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + out_row * p.ldo + col);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j * 4 < ncols) op[j] = make_float4(v[4 * j + 0], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
  }
}

template <int BN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                         const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_blocks * p.num_n_blocks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 128);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / p.num_n_blocks;
        const int n_blk = tile - m_blk * p.num_n_blocks;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          tma_load_2d(sa, &tma_a, &full_bar[stage], kb * BK, m_blk * BM);
          tma_load_2d(sa + Cfg::A_BYTES, &tma_b, &full_bar[stage], kb * BK, n_blk * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------- MMA issuer (single thread)
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
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
            umma_bf16_ss(d_tmem, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);          // stage reusable once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);              // accumulator complete
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ------------------------------------------------------------- epilogue (4 warps, TMEM lane == row)
    const int ew = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / p.num_n_blocks;
      const int n_blk = tile - m_blk * p.num_n_blocks;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int row = m_blk * BM + ew * 32 + lane;
      const bool row_ok = row < p.M;
      long long out_row = row;
      int rb_row = 0;
      if (EPI == TGB_EPI_F32_ROWBIAS) {
        const int grp = row / p.period;
        rb_row = row - grp * p.period;
        out_row = (long long)grp * (p.period + p.skip) + p.skip + rb_row;
      }
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + static_cast<uint32_t>(acc * BN);
#pragma unroll 2
      for (int c = 0; c < BN; c += 16) {
        const int col = n_blk * BN + c;
        if (col >= p.N) break;                      // warp-uniform
        uint32_t v[16];
        tmem_ld_32x32b_x16(t_row + c, v);
        tmem_ld_wait();
        const int ncols = (p.N - col) >= 16 ? 16 : 8;
        if (row_ok) epilogue_store16<EPI>(p, v, out_row, row, rb_row, col, ncols);
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
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
static int make_tmap(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows) {
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

template <int BN, int EPI>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_bf16_tcgen05_kernel<BN, EPI>;
  static bool configured = false;
  if (!configured) {
    TGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  const int tiles = p.num_m_blocks * p.num_n_blocks;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(ta, tb, p);
  TGB_CUDA(cudaGetLastError());
  return TGB_OK;
}

template <int BN>
static int dispatch_epi(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t s) {
  switch (p.epi) {
    case TGB_EPI_BF16: return launch_gemm<BN, TGB_EPI_BF16>(ta, tb, p, s);
    case TGB_EPI_BF16_GELU: return launch_gemm<BN, TGB_EPI_BF16_GELU>(ta, tb, p, s);
    case TGB_EPI_F32: return launch_gemm<BN, TGB_EPI_F32>(ta, tb, p, s);
    case TGB_EPI_F32_RESID: return launch_gemm<BN, TGB_EPI_F32_RESID>(ta, tb, p, s);
    case TGB_EPI_F32_ROWBIAS: return launch_gemm<BN, TGB_EPI_F32_ROWBIAS>(ta, tb, p, s);
  }
  set_last_error("unknown epilogue %d", p.epi);
  return TGB_EINVAL;
}

static int choose_bn(int M, int N) {
  const int cands[4] = {256, 176, 128, 64};
  const int mblocks = (M + BM - 1) / BM;
  int best = 256;
  long long best_cost = -1;
  for (int i = 0; i < 4; ++i) {
    const int bn = cands[i];
    const long long nblocks = (N + bn - 1) / bn;
    const long long tiles = nblocks * mblocks;
    const long long waves = (tiles + sm_count() - 1) / sm_count();
    // cost ~ MMA cycles of the busiest SM; narrower tiles pay more shared-memory traffic per flop
    long long cost = waves * bn * (bn >= 176 ? 100 : (bn == 128 ? 108 : 125));
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

}  // namespace tgb

using namespace tgb;

extern "C" int tgb_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K, int epilogue,
                             const float* bias, const float* resid, int64_t ldr, const float* rowbias, int period,
                             int skip, void* out, int64_t ldo, int bn, tgb_stream_t stream) {
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
  if (epilogue == TGB_EPI_F32_RESID)
    TGB_CHECK_ARG(resid && ldr % 4 == 0 && ldr >= N && ((uintptr_t)resid % 16) == 0, "gemm: bad residual");
  if (epilogue == TGB_EPI_F32_ROWBIAS)
    TGB_CHECK_ARG(rowbias && period > 0 && skip >= 0 && ((uintptr_t)rowbias % 16) == 0, "gemm: bad rowbias");
  if (bn == 0) bn = choose_bn(M, N);
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.epi = epilogue;
  p.num_m_blocks = (M + BM - 1) / BM;
  p.num_n_blocks = (N + bn - 1) / bn;
  p.num_k_blocks = (K + BK - 1) / BK;
  p.period = period > 0 ? period : 1; p.skip = skip;
  p.bias = bias; p.resid = resid; p.rowbias = rowbias; p.out = out; p.ldo = ldo; p.ldr = ldr;
  CUtensorMap ta, tb;
  st = make_tmap(&ta, A, M, K, lda, BM);
  if (st != TGB_OK) return st;
  st = make_tmap(&tb, W, N, K, ldw, bn);
  if (st != TGB_OK) return st;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  switch (bn) {
    case 256: return dispatch_epi<256>(ta, tb, p, s);
    case 176: return dispatch_epi<176>(ta, tb, p, s);
    case 128: return dispatch_epi<128>(ta, tb, p, s);
    case 64: return dispatch_epi<64>(ta, tb, p, s);
  }
  set_last_error("gemm: unsupported N tile %d (256/176/128/64)", bn);
  return TGB_EINVAL;
}
