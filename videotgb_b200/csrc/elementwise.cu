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

// LayerNorm fold, first layer: rows that did not come out of a statistics-producing GEMM.  One warp per row; same
// two-pass statistics as layernorm_kernel; writes the bf16 copy of the raw row and (mean, rstd).
template <int VEC>
__global__ void __launch_bounds__(256) rowstats_cast_kernel(const float* __restrict__ x, long long ldx, float eps, int M,
                                                            int N, __nv_bfloat16* __restrict__ ob, long long ldb,
                                                            float2* __restrict__ stats) {
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
        uint2 pk;
        pk.x = pack_bf16(v[i].x, v[i].y);
        pk.y = pack_bf16(v[i].z, v[i].w);
        reinterpret_cast<uint2*>(ob + row * ldb)[c] = pk;
      }
    }
    const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(N) + eps);
    if (lane == 0) stats[row] = make_float2(mean, rstd);
  }
}

// LayerNorm fold: per-row (sum, sum of squares) slots written by the TGB_EPI_F32_RESID_STATS epilogue -> (mean, rstd).
// One thread per row, slots summed in index order in float64 (deterministic; the subtraction E[x^2] - mean^2 is exact
// to fp32 partial-sum accuracy).
// max_ratio (optional): running maximum of |mean| * rstd over all rows, kept as the bit pattern of a non-negative float --
// the fold's error grows with that ratio (see tgb_gemm_bf16_ln), the host checks it once after loading weights.
__global__ void __launch_bounds__(256) ln_stats_finalize_kernel(const float2* __restrict__ part, int M, int slots, int N,
                                                                float eps, float2* __restrict__ stats,
                                                                float* __restrict__ max_ratio) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  float ratio = 0.f;                       // rows past M take part in the warp reduction with a neutral value (no early exit)
  if (row < M) {
    // slot-major layout [slots][M]: a warp reads 32 consecutive rows of one slot = 256 contiguous bytes per load
    // (eleven loads in flight per thread: the kernel moves 8 * slots bytes per row and is latency-bound otherwise)
    double s = 0.0, q = 0.0;
    for (int j0 = 0; j0 < slots; j0 += 11) {
      float2 v[11];
#pragma unroll
      for (int k = 0; k < 11; ++k) v[k] = j0 + k < slots ? __ldg(part + (long long)(j0 + k) * M + row) : make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 11; ++k) {          // slots past the end add +0.0: the sums are unchanged
        s += static_cast<double>(v[k].x);
        q += static_cast<double>(v[k].y);
      }
    }
    const double mean = s / N;
    double var = q / N - mean * mean;
    var = var > 0.0 ? var : 0.0;
    const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    stats[row] = make_float2(static_cast<float>(mean), rstd);
    ratio = fabsf(static_cast<float>(mean)) * rstd;
  }
  if (max_ratio != nullptr) {
    // one atomic per CTA: thousands of same-address atomics serialise in the L2 and were the kernel's floor
    __shared__ float wmax[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ratio = fmaxf(ratio, __shfl_xor_sync(0xffffffffu, ratio, o));
    if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = ratio;
    __syncthreads();
    if (threadIdx.x == 0) {
      float r = wmax[0];
#pragma unroll
      for (int w = 1; w < 8; ++w) r = fmaxf(r, wmax[w]);
      if (r > 0.f) atomicMax(reinterpret_cast<int*>(max_ratio), __float_as_int(r));
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


// N4: LLM hand-off.  One pass builds what the reference makes with an embedding lookup and two
// torch.cat calls (src/models/LSTP_SF_blip2_module.py:482-495): inputs_embeds[b] = [visual tokens of video b | word
// embeddings of its question], attention_mask[b] = [ones | question mask].  Pure copies (bit-exact); one warp per row.
__global__ void __launch_bounds__(256) llm_inputs_kernel(const float* __restrict__ visual, const float* __restrict__ table,
                                                         const long long* __restrict__ ids, const long long* __restrict__ qmask,
                                                         long long vocab, int B, int Nv, int Tq, int nvec,
                                                         float* __restrict__ embeds, long long* __restrict__ mask) {
  const int lane = threadIdx.x & 31;
  const long long rows = (long long)B * (Nv + Tq);
  for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows;
       r += (long long)gridDim.x * (blockDim.x >> 5)) {
    const long long b = r / (Nv + Tq);
    const int j = static_cast<int>(r - b * (Nv + Tq));
    const float4* src;
    long long m = 1;
    if (j < Nv) {
      src = reinterpret_cast<const float4*>(visual) + (b * Nv + j) * nvec;
    } else {
      long long id = ids[b * Tq + (j - Nv)];
      id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);      // host validated; clamp keeps a bad id from faulting
      src = reinterpret_cast<const float4*>(table) + id * nvec;
      m = qmask[b * Tq + (j - Nv)];
    }
    float4* dst = reinterpret_cast<float4*>(embeds) + r * nvec;
    for (int c = lane; c < nvec; c += 32) __stcs(dst + c, __ldg(src + c));
    if (lane == 0) mask[r] = m;
  }
}

// outputs[outputs == from] = to   (LSTP_SF_blip2_module.py:503-504: generated id 0 -> 2 for LLaMA tokenizers)
__global__ void replace_ids_kernel(long long* __restrict__ ids, long long n, long long from, long long to) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (ids[i] == from) ids[i] = to;
}

}  // namespace tgb

using namespace tgb;

extern "C" int tgb_ln_stats_finalize(const float* partial_stats, int M, int slots, int N, float eps, float* rowstats,
                                     float* max_ratio, tgb_stream_t stream) {
  int st = check_device();
  if (st != TGB_OK) return st;
  TGB_CHECK_ARG(partial_stats && rowstats && M >= 0 && slots > 0 && N > 0, "ln_stats_finalize: bad arguments");
  TGB_CHECK_ARG(((uintptr_t)partial_stats % 8) == 0 && ((uintptr_t)rowstats % 8) == 0, "ln_stats_finalize: misaligned");
  if (M == 0) return TGB_OK;
  ln_stats_finalize_kernel<<<(M + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float2*>(partial_stats), M, slots, N, eps, reinterpret_cast<float2*>(rowstats), max_ratio);
  TGB_CUDA(cudaGetLastError());
  return TGB_OK;
}

extern "C" int tgb_rowstats_cast_f32(const float* x, int64_t ldx, int M, int N, float eps, void* out_bf16, int64_t ldo,
                                     float* rowstats, tgb_stream_t stream) {
  int st = check_device();
  if (st != TGB_OK) return st;
  TGB_CHECK_ARG(x && out_bf16 && rowstats, "rowstats_cast: null pointer");
  TGB_CHECK_ARG(M >= 0 && N > 0 && N % 4 == 0 && N <= 128 * LN_MAX_VEC, "rowstats_cast: N=%d must be a multiple of 4 and <= %d", N, 128 * LN_MAX_VEC);
  TGB_CHECK_ARG(ldx % 4 == 0 && ldx >= N && ((uintptr_t)x % 16) == 0, "rowstats_cast: misaligned input");
  TGB_CHECK_ARG(ldo % 4 == 0 && ldo >= N && ((uintptr_t)out_bf16 % 8) == 0 && ((uintptr_t)rowstats % 8) == 0, "rowstats_cast: bad output");
  if (M == 0) return TGB_OK;
  const int warps = 8;
  long long blocks = ((long long)M + warps - 1) / warps;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  const int nvec_lane = (N / 4 + 31) / 32;
  cudaStream_t s = (cudaStream_t)stream;
  auto ob = (__nv_bfloat16*)out_bf16;
  auto rs = reinterpret_cast<float2*>(rowstats);
#define TGB_RS(V) rowstats_cast_kernel<V><<<(int)blocks, 256, 0, s>>>(x, ldx, eps, M, N, ob, ldo, rs)
  if (nvec_lane <= 2) TGB_RS(2);
  else if (nvec_lane <= 6) TGB_RS(6);
  else if (nvec_lane <= 11) TGB_RS(11);
  else TGB_RS(16);
#undef TGB_RS
  TGB_CUDA(cudaGetLastError());
  return TGB_OK;
}

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

extern "C" int tgb_llm_inputs_f32(const float* visual, const float* table, const int64_t* ids, const int64_t* qmask,
                                  int64_t vocab, int B, int Nv, int Tq, int H, float* embeds, int64_t* mask,
                                  tgb_stream_t stream) {
  int st = check_device();
  if (st != TGB_OK) return st;
  TGB_CHECK_ARG(B >= 0 && Nv >= 0 && Tq >= 0 && H > 0 && H % 4 == 0 && vocab > 0, "llm_inputs: bad shape B=%d Nv=%d Tq=%d H=%d", B, Nv, Tq, H);
  if ((long long)B * (Nv + Tq) == 0) return TGB_OK;
  TGB_CHECK_ARG((visual || Nv == 0) && ((table && ids && qmask) || Tq == 0) && embeds && mask, "llm_inputs: null pointer");
  TGB_CHECK_ARG(((uintptr_t)visual % 16) == 0 && ((uintptr_t)table % 16) == 0 && ((uintptr_t)embeds % 16) == 0, "llm_inputs: misaligned");
  const long long rows = (long long)B * (Nv + Tq);
  if (rows == 0) return TGB_OK;
  long long blocks = (rows + 7) / 8;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  llm_inputs_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(visual, table, (const long long*)ids, (const long long*)qmask,
                                                                   vocab, B, Nv, Tq, H / 4, embeds, (long long*)mask);
  TGB_CUDA(cudaGetLastError());
  return TGB_OK;
}

extern "C" int tgb_replace_ids(int64_t* ids, int64_t n, int64_t from, int64_t to, tgb_stream_t stream) {
  int st = check_device();
  if (st != TGB_OK) return st;
  TGB_CHECK_ARG(ids && n >= 0, "replace_ids: bad arguments");
  if (n == 0) return TGB_OK;
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)sm_count() * 4;
  if (blocks > cap) blocks = cap;
  replace_ids_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((long long*)ids, n, from, to);
  TGB_CUDA(cudaGetLastError());
  return TGB_OK;
}
