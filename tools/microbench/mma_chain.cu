// Micro-benchmark: cost of tcgen05.mma (cta_group::1, kind::f16, M = 128, K = 16) per instruction as a function of N, of the
// number of INDEPENDENT accumulators the instructions alternate between, and of the A operand's source (shared memory /
// tensor memory).  One CTA, one issuing thread; clocks are SM cycles (clock64) from the first issue to the completion of
// the last MMA (tcgen05.commit -> mbarrier).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_chain tools/microbench/mma_chain.cu
#include <vector>
#include <utility>
#include "../../videotgb_b200/csrc/common.cuh"

using namespace tgb;

__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}

struct Cfg { int n, chains, ts, count; };

// 48 MMAs, fully unrolled: descriptors and accumulator addresses are compile-time offsets from a few registers, so the
// issuing thread spends a handful of instructions per MMA (a first version with run-time indices measured its own loop:
// ~207 clk per MMA whatever the shape).
template <int N, int CHAINS, bool TS>
__device__ __forceinline__ void run(uint32_t tmem, uint64_t da, uint64_t db, uint64_t* bar, uint32_t& phase, long long* out) {
  constexpr uint32_t idesc = umma_idesc_bf16(128, N);
  constexpr int stride = CHAINS > 1 ? (448 / CHAINS) / 16 * 16 : 0;
  constexpr int COUNT = 48;
  for (int rep = 0; rep < 2; ++rep) {
    const long long t0 = clock64();
#pragma unroll
    for (int i = 0; i < COUNT; ++i) {
      const uint32_t d = tmem + (i % CHAINS) * stride;
      if (TS) umma_ts(d, tmem + 448 + (i & 3) * 8, db + ((i & 3) * 2), idesc, i >= CHAINS ? 1u : 0u);
      else umma_bf16_ss(d, da + ((i & 3) * 2), db + ((i & 3) * 2), idesc, i >= CHAINS ? 1u : 0u);
    }
    const long long t1 = clock64();
    umma_commit(bar);
    mbar_wait(bar, phase);
    phase ^= 1;
    const long long t2 = clock64();
    if (rep == 1) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
}

__global__ void __launch_bounds__(128, 1) bench_kernel(const Cfg* cfgs, int ncfg, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                    // 128 x 128 B
  uint8_t* sB = smem + 16384;            // 256 x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  fence_proxy_async();
  if (threadIdx.x < 32) tmem_alloc(slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    uint32_t phase = 0;
    const uint64_t da = umma_desc_sw128(smem_u32(sA));
    const uint64_t db = umma_desc_sw128(smem_u32(sB));
    int c = 0;
#define RUN(N, CH) run<N, CH, false>(tmem, da, db, bar, phase, out + 2 * (c++));
#define RUNT(N, CH) run<N, CH, true>(tmem, da, db, bar, phase, out + 2 * (c++));
    RUN(48, 1) RUN(48, 2) RUN(48, 4) RUN(96, 1) RUN(96, 2) RUN(96, 4) RUN(144, 1) RUN(144, 2) RUN(144, 3) RUN(256, 1)
    RUNT(48, 1) RUNT(48, 2) RUNT(48, 4) RUNT(96, 1) RUNT(96, 2) RUNT(96, 4) RUNT(144, 1) RUNT(144, 2) RUNT(144, 3) RUNT(256, 1)
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

namespace tgb {   // the two symbols common.cuh expects from runtime.cu
void set_last_error(const char*, ...) {}
int cuda_status(cudaError_t, const char*) { return -3; }
}

int main() {
  std::vector<Cfg> cfgs;
  for (int ts = 0; ts < 2; ++ts)
    for (auto nc : {std::pair<int, int>{48, 1}, {48, 2}, {48, 4}, {96, 1}, {96, 2}, {96, 4}, {144, 1}, {144, 2}, {144, 3}, {256, 1}})
      cfgs.push_back(Cfg{nc.first, nc.second, ts, 48});
  Cfg* d_cfg; long long* d_out;
  cudaMalloc(&d_cfg, cfgs.size() * sizeof(Cfg));
  cudaMalloc(&d_out, cfgs.size() * 2 * sizeof(long long));
  cudaMemcpy(d_cfg, cfgs.data(), cfgs.size() * sizeof(Cfg), cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  bench_kernel<<<1, 128, 64 * 1024>>>(d_cfg, (int)cfgs.size(), d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<long long> out(cfgs.size() * 2);
  cudaMemcpy(out.data(), d_out, out.size() * sizeof(long long), cudaMemcpyDeviceToHost);
  printf("A-src  N  chains  MMAs  issue_clk/MMA  total_clk/MMA  floor(128*N/256)\n");
  for (size_t c = 0; c < cfgs.size(); ++c)
    printf("%-5s %4d %5d %6d %12.1f %14.1f %10d\n", cfgs[c].ts ? "tmem" : "smem", cfgs[c].n, cfgs[c].chains, cfgs[c].count,
           (double)out[2 * c] / cfgs[c].count, (double)out[2 * c + 1] / cfgs[c].count, 128 * cfgs[c].n / 256);
  return 0;
}
