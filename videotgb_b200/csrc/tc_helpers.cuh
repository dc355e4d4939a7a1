// tcgen05 / TMA helpers shared by the tensor-core attention kernels (attention_tc.cu, attention_qf_tc.cu).
#pragma once
#include "common.cuh"

namespace tgb {

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// L2 prefetch of a box (no shared-memory destination, no completion to wait for)
__device__ __forceinline__ void tma_prefetch_3d(const void* map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// MN-major SWIZZLE_128B operand: 64 MN elements (128 B) per K row, 8-row groups every `sbo` bytes, 64-wide MN blocks
// every `lbo` bytes.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

#define TGB_R8(a, o) "=r"(a[o]), "=r"(a[o + 1]), "=r"(a[o + 2]), "=r"(a[o + 3]), "=r"(a[o + 4]), "=r"(a[o + 5]), "=r"(a[o + 6]), "=r"(a[o + 7])
#define TGB_W8(a, o) "r"(a[o]), "r"(a[o + 1]), "r"(a[o + 2]), "r"(a[o + 3]), "r"(a[o + 4]), "r"(a[o + 5]), "r"(a[o + 6]), "r"(a[o + 7])

template <int OFF, int NREG>
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[NREG]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
      "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : TGB_R8(v, OFF), TGB_R8(v, OFF + 8), TGB_R8(v, OFF + 16), TGB_R8(v, OFF + 24)
      : "r"(taddr));
}
template <int OFF, int NREG>
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[NREG]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : TGB_R8(v, OFF), TGB_R8(v, OFF + 8)
      : "r"(taddr));
}
template <int OFF, int NREG>
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[NREG]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : TGB_R8(v, OFF) : "r"(taddr));
}
template <int OFF, int NREG>
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[NREG]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
      "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), TGB_W8(v, OFF), TGB_W8(v, OFF + 8), TGB_W8(v, OFF + 16), TGB_W8(v, OFF + 24)
      : "memory");
}
template <int OFF, int NREG>
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&v)[NREG]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(v[OFF]), "r"(v[OFF + 1]),
               "r"(v[OFF + 2]), "r"(v[OFF + 3])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int OFF, int NREG>
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[NREG]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), TGB_W8(v, OFF) : "memory");
}

template <int OFF, int NREG>
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[NREG]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               ::"r"(taddr), TGB_W8(v, OFF), TGB_W8(v, OFF + 8) : "memory");
}

}  // namespace tgb
