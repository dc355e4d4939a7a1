// Shared device helpers for the sm_100a kernels: mbarrier / TMA / tcgen05 PTX wrappers,
// bf16 packing, warp reductions.  Everything here is inline PTX for sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <cstdio>
#include <atomic>

#include "../../include/videotgb_b200.h"
#include "tgb_debug.h"

namespace tgb {

// ------------------------------------------------------------------ error plumbing
void set_last_error(const char* fmt, ...);
int cuda_status(cudaError_t e, const char* what);

#define TGB_CHECK_ARG(cond, ...)                       \
  do {                                                 \
    if (!(cond)) {                                     \
      ::tgb::set_last_error(__VA_ARGS__);              \
      return TGB_EINVAL;                               \
    }                                                  \
  } while (0)

#define TGB_CUDA(call)                                              \
  do {                                                              \
    cudaError_t _e = (call);                                        \
    if (_e != cudaSuccess) return ::tgb::cuda_status(_e, #call);    \
  } while (0)

// Per-function attributes (cudaFuncSetAttribute: the > 48 KB dynamic shared memory opt-in) apply to the CURRENT device
// only, so "already configured" is remembered per call site AND per device.  Re-running `stmt` is idempotent, so a race
// between two host threads costs one redundant driver call; devices beyond 63 simply reconfigure on every launch.
#define TGB_ONCE_PER_DEVICE(stmt)                                                                   \
  do {                                                                                              \
    static std::atomic<unsigned long long> tgb_once_mask_{0ull};                                    \
    int tgb_dev_ = 0;                                                                               \
    TGB_CUDA(cudaGetDevice(&tgb_dev_));                                                             \
    const unsigned long long tgb_bit_ = tgb_dev_ < 64 ? (1ull << tgb_dev_) : 0ull;                  \
    if (!(tgb_once_mask_.load(std::memory_order_acquire) & tgb_bit_)) {                             \
      stmt;                                                                                         \
      tgb_once_mask_.fetch_or(tgb_bit_, std::memory_order_release);                                 \
    }                                                                                               \
  } while (0)

int check_device();      // TGB_OK if the current device is sm_100, else TGB_EARCH
int sm_count();

// attention_qf_tc.cu: tcgen05 form of tgb_attention_bf16 for the Q-Former cross-attention shape (32 queries, d = 64,
// H % 4 == 0, 64 < Sk <= 257); tgb_attention_bf16 dispatches to it when `eligible` says so.
bool qf_cross_attention_tc_eligible(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                    const void* out, int64_t ldo, int B, int H, int Sq, int Sk, int D, float scale);
int qf_cross_attention_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out,
                          int64_t ldo, const float* mask, int B, int H, int Sk, float scale, cudaStream_t stream);
// ... and for the Q-Former self-attention shape (Sq = Sk = 32, d = 64, H % 4 == 0)
bool qf_self_attention_tc_eligible(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                   const void* out, int64_t ldo, int B, int H, int Sq, int Sk, int D, float scale);
int qf_self_attention_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out,
                         int64_t ldo, const float* mask, int B, int H, float scale, cudaStream_t stream);

// ------------------------------------------------------------------ small device utils
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Exact-erf GELU (HF ACT2FN["gelu"]):  gelu(x) = x * Phi(x),  Phi(x) = 1 - 0.5*erfc(x/sqrt2).
// erfc via Abramowitz-Stegun 7.1.26 (|erf error| <= 1.5e-7) with the -0.5 folded into the coefficients:
//   h(|x|) = 0.5*erfc(|x|/sqrt2) = t*(a1 + t*(a2 + t*(a3 + t*(a4 + t*a5)))) * exp(-x^2/2) / 2,  t = 1/(1 + p*|x|/sqrt2)
//   gelu(x) = max(x, 0) - |x|*h        (branch-free; equals x*(1-h) for x >= 0 and x*h for x < 0)
// Evaluated on PAIRS with Blackwell's packed fp32 pipe (FFMA2 / FMUL2): 10 packed FP32 + 4 ALU + 4 MUFU per two
// elements (the scalar form was 24 FP32 + 4 MUFU); the epilogue's instruction issue is what the GELU GEMM pays for.
__device__ __forceinline__ float2 gelu_erf2(float2 x) {
  const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
  constexpr float K0 = 0.3275911f * 0.70710678118654752f;
  const float2 d = __ffma2_rn(ax, make_float2(K0, K0), make_float2(1.0f, 1.0f));
  const float2 t = make_float2(rcp_approx(d.x), rcp_approx(d.y));
  constexpr float A1 = -0.5f * 0.254829592f, A2 = -0.5f * -0.284496736f, A3 = -0.5f * 1.421413741f,
                  A4 = -0.5f * -1.453152027f, A5 = -0.5f * 1.061405429f;
  float2 poly = __ffma2_rn(t, make_float2(A5, A5), make_float2(A4, A4));
  poly = __ffma2_rn(poly, t, make_float2(A3, A3));
  poly = __ffma2_rn(poly, t, make_float2(A2, A2));
  poly = __ffma2_rn(poly, t, make_float2(A1, A1));
  poly = __fmul2_rn(poly, t);
  constexpr float KE = -0.72134752044448170f;                         // exp(-x^2/2) = 2^(KE * x^2)
  const float2 arg = __fmul2_rn(__fmul2_rn(x, x), make_float2(KE, KE));
  const float2 e = make_float2(ex2_approx(arg.x), ex2_approx(arg.y));
  const float2 nh = __fmul2_rn(poly, e);                              // -h
  return __ffma2_rn(ax, nh, make_float2(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f)));
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// try_wait with a suspend-time hint: the hardware parks the thread until the phase completes or the (system-capped) hint
// expires, so a long wait costs a handful of wake-ups instead of a polling loop.
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
      : "memory");
  return ok != 0;
}
// Bounded waits: a pipeline bug must surface as a trapped kernel (cudaErrorLaunchFailure), never as a hung GPU.
// The polling loop is kept to three instructions (measured: with the timer read inside it, waiting warps issued half of
// all instructions of the bias GEMM -- on a power-capped part that is throughput taken from the tensor cores); the clock
// is only consulted every 4096 polls.
//   mbar_wait        plain try_wait polling: for the latency-critical single-thread roles (TMA producer, MMA issuer)
//   mbar_wait_parked try_wait with a suspend-time hint (NANOSLEEP.SYNCS until the barrier fires): for whole warps that
//                    wait a long time (the epilogue warps waiting for an accumulator)
template <bool PARKED>
__device__ __forceinline__ void mbar_wait_impl(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint64_t t0 = 0;
  for (;;) {
#pragma unroll 1
    for (int it = 0; it < 4096; ++it)
      if (PARKED ? mbar_try_wait_hint(bar, parity, 1000000u) : mbar_try_wait(bar, parity)) return;
    uint64_t t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t0 == 0) t0 = t1;
    if (t1 - t0 > 4000000000ull) {   // 4 s
      printf("tgb: mbarrier wait timed out (block %d thread %d parity %u)\n", blockIdx.x, threadIdx.x, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { mbar_wait_impl<false>(bar, parity); }
__device__ __forceinline__ void mbar_wait_parked(uint64_t* bar, uint32_t parity) { mbar_wait_impl<true>(bar, parity); }

// ------------------------------------------------------------------ TMA
__device__ __forceinline__ void tma_prefetch_desc(const void* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// ------------------------------------------------------------------ tcgen05 / TMEM
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {   // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {        // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// K-major operand tile in shared memory, rows of 64 bf16 (128 B) written by TMA with SWIZZLE_128B:
// 8-row x 128-B swizzle atoms stacked every 1024 B (SBO), LBO unused for swizzled K-major layouts.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);       // [0,14)  start address >> 4
  d |= static_cast<uint64_t>(1) << 16;                           // [16,30) leading byte offset >> 4 (ignored)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                   // [32,46) stride byte offset >> 4
  d |= static_cast<uint64_t>(1) << 46;                           // [46,48) descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;                           // [61,64) SWIZZLE_128B
  return d;
}

// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, dense.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrives on `bar` once every tcgen05.mma previously issued BY THIS THREAD has completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

}  // namespace tgb

// ------------------------------------------------------------------ clusters / CTA pairs (cta_group::2)
namespace tgb {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_addr` in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Relaxed arrivals for hand-offs that publish no memory -- "this warp has finished READING the accumulator" (the reads
// have completed once tcgen05.wait::ld returns).  The default .release form costs a MEMBAR.ALL.GPU that first drains every
// global store the thread has in flight (measured: 7 % of the residual GEMM's stall samples sat on it).
__device__ __forceinline__ void mbar_arrive_relaxed(uint64_t* bar) {
  asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)));
}
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr));
}

template <int CG>
__device__ __forceinline__ void tmem_alloc_cg(uint32_t* smem_slot, uint32_t ncols) {   // whole warp (same warp id in both CTAs)
  if (CG == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc_cg(uint32_t taddr, uint32_t ncols) {
  if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// TMA load whose completion bytes are posted to the LEADER CTA's mbarrier (cluster address `mbar_cluster_addr`).
__device__ __forceinline__ void tma_load_2d_cg2(void* smem_dst, const void* map, uint32_t mbar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(mbar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}

template <int CG>
__device__ __forceinline__ void umma_bf16_ss_cg(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                uint32_t accumulate) {
  if (CG == 1) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
  }
}
// commit -> arrive on `bar` (same smem offset) in every CTA of the pair
template <int CG>
__device__ __forceinline__ void umma_commit_cg(uint64_t* bar) {
  if (CG == 1) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  } else {
    const uint16_t mask = 0x3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
  }
}

// tcgen05.wait::ld with the destination registers threaded through as in/out operands, so the compiler cannot
// schedule a use of the loaded values above the wait (and ordinary global loads may still move across it).
__device__ __forceinline__ void tmem_ld_wait_dep(uint32_t (&v)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]));
}
// 32 columns per lane in one tcgen05.ld (the bf16 epilogue keeps a lane's whole 32-column row piece in registers)
__device__ __forceinline__ void tmem_ld_32x32b_x32_nomem(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"
      "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait_dep32(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                 "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                 "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31]));
}
// bulk tensor store shared -> global (2-D), bulk-group completion
__device__ __forceinline__ void tma_store_2d(const void* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld_32x32b_x16_nomem(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}

}  // namespace tgb
