// Row-wise LayerNorm and small elementwise helpers (HBM-bound; one warp per row, 16-byte accesses,
// warp-shuffle reductions, two-pass statistics held in registers).
#include "common.cuh"

namespace tgb {

constexpr int LN_MAX_VEC = 16;   // float4 per lane -> N <= 32*4*16 = 2048 in registers; larger rows re-read

template <int VEC>   // VEC = number of float4 per lane held in registers (N <= 128*VEC)
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, long long ldx,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float eps, int M, int N, __nv_bfloat16* __restrict__ ob,
                                                        long long ldb, float* __restrict__ of, long long ldf) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nvec = N >> 2;
  for (long long row = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < M;
       row += (long long)gridDim.x * warps_per_block) {
    const float4* xr = reinterpret_cast<const float4*>(x + row * ldx);
    float4 v[VEC];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int c = lane + 32 * i;
      v[i] = (c < nvec) ? xr[c] : make_float4(0.f, 0.f, 0.f, 0.f);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) / static_cast<float>(N);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
        q += (a * a + b * b) + (cc * cc + d * d);
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(N) + eps);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const int c = lane + 32 * i;
      if (c < nvec) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c);
        float4 o;
        o.x = (v[i].x - mean) * rstd * g.x + b.x;
        o.y = (v[i].y - mean) * rstd * g.y + b.y;
        o.z = (v[i].z - mean) * rstd * g.z + b.z;
        o.w = (v[i].w - mean) * rstd * g.w + b.w;
        if (of) reinterpret_cast<float4*>(of + row * ldf)[c] = o;
        if (ob) {
          uint2 pk;
          pk.x = pack_bf16(o.x, o.y);
          pk.y = pack_bf16(o.z, o.w);
          reinterpret_cast<uint2*>(ob + row * ldb)[c] = pk;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ x, long long ldx,
                                                            __nv_bfloat16* __restrict__ out, long long ldo, int M, int nvec) {
  const long long total = (long long)M * nvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / nvec;
    const int c = static_cast<int>(i - r * nvec);
    const float4 a = reinterpret_cast<const float4*>(x + r * ldx)[c];
    uint2 pk;
    pk.x = pack_bf16(a.x, a.y);
    pk.y = pack_bf16(a.z, a.w);
    reinterpret_cast<uint2*>(out + r * ldo)[c] = pk;
  }
}

// out[g, i] = mean_w x[g*width + w, i]; sequential fp32 accumulation over w (torch .mean(0) order for small width).
__global__ void __launch_bounds__(256) group_mean_kernel(const float* __restrict__ x, float* __restrict__ out, int G,
                                                         int width, long long inner) {
  const long long total = (long long)G * inner;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long g = i / inner;
    const long long c = i - g * inner;
    float s = 0.f;
    for (int w = 0; w < width; ++w) s += x[(g * width + w) * inner + c];
    out[i] = s / static_cast<float>(width);
  }
}

}  // namespace tgb

using namespace tgb;

extern "C" int tgb_layernorm_f32(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps, int M,
                                 int N, void* out_bf16, int64_t ld_bf16, float* out_f32, int64_t ld_f32,
                                 tgb_stream_t stream) {
  int st = check_device();
  if (st != TGB_OK) return st;
  TGB_CHECK_ARG(x && gamma && beta && (out_bf16 || out_f32), "layernorm: null pointer");
  TGB_CHECK_ARG(M >= 0 && N > 0 && N % 4 == 0 && N <= 128 * LN_MAX_VEC, "layernorm: N=%d must be a multiple of 4 and <= %d", N, 128 * LN_MAX_VEC);
  TGB_CHECK_ARG(ldx % 4 == 0 && ldx >= N && ((uintptr_t)x % 16) == 0 && ((uintptr_t)gamma % 16) == 0 && ((uintptr_t)beta % 16) == 0,
                "layernorm: misaligned input");
  TGB_CHECK_ARG(!out_bf16 || (ld_bf16 % 4 == 0 && ld_bf16 >= N && ((uintptr_t)out_bf16 % 8) == 0), "layernorm: bad bf16 output");
  TGB_CHECK_ARG(!out_f32 || (ld_f32 % 4 == 0 && ld_f32 >= N && ((uintptr_t)out_f32 % 16) == 0), "layernorm: bad f32 output");
  if (M == 0) return TGB_OK;
  const int warps = 8;
  long long blocks = ((long long)M + warps - 1) / warps;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  const int nvec_lane = (N / 4 + 31) / 32;
  cudaStream_t s = (cudaStream_t)stream;
  auto ob = (__nv_bfloat16*)out_bf16;
#define TGB_LN(V) layernorm_kernel<V><<<(int)blocks, 256, 0, s>>>(x, ldx, gamma, beta, eps, M, N, ob, ld_bf16, out_f32, ld_f32)
  if (nvec_lane <= 2) TGB_LN(2);
  else if (nvec_lane <= 4) TGB_LN(4);
  else if (nvec_lane <= 6) TGB_LN(6);
  else if (nvec_lane <= 8) TGB_LN(8);
  else if (nvec_lane <= 11) TGB_LN(11);
  else TGB_LN(16);
#undef TGB_LN
  TGB_CUDA(cudaGetLastError());
  return TGB_OK;
}

extern "C" int tgb_cast_f32_bf16(const float* x, int64_t ldx, void* out, int64_t ldo, int M, int N, tgb_stream_t stream) {
  int st = check_device();
  if (st != TGB_OK) return st;
  TGB_CHECK_ARG(x && out && M >= 0 && N > 0 && N % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0, "cast: N, ldx, ldo must be multiples of 4");
  TGB_CHECK_ARG(((uintptr_t)x % 16) == 0 && ((uintptr_t)out % 8) == 0, "cast: misaligned");
  if (M == 0) return TGB_OK;
  const long long total = (long long)M * (N / 4);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  cast_f32_bf16_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, ldx, (__nv_bfloat16*)out, ldo, M, N / 4);
  TGB_CUDA(cudaGetLastError());
  return TGB_OK;
}

extern "C" int tgb_group_mean_f32(const float* x, float* out, int G, int width, int64_t inner, tgb_stream_t stream) {
  int st = check_device();
  if (st != TGB_OK) return st;
  TGB_CHECK_ARG(x && out && G >= 0 && width > 0 && inner > 0, "group_mean: bad arguments");
  if (G == 0) return TGB_OK;
  const long long total = (long long)G * inner;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  group_mean_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, out, G, width, inner);
  TGB_CUDA(cudaGetLastError());
  return TGB_OK;
}
