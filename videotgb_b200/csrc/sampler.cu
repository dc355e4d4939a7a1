// Temporal-sampler specific kernels:
//   of_patch_reduce   the HBM-bound half of TemporalOFEmbedding: z[f,(c,ky,kx)] = sum_p fc_w[p] * flow[f,c,py*16+ky,px*16+kx]
//                     (conv(2->768,k16,s16) followed by Linear(196->1) over the patch axis reassociates exactly to
//                      this weighted patch reduction followed by one [T,512]x[512,768] GEMM)
//   sampler_assemble  bos / eos / zero-row placement + frame_pos_embed + LayerNorm(eps)
//   embed_layernorm   LayerNorm(table[ids] + pos[t] + addvec)   (question / instruction token embeddings)
//   rope_bf16         rotary embedding with an opaque [pos,64] = [sin(32)|cos(32)] table, in place on bf16 heads
#include "common.cuh"

namespace tgb {

// ---------------------------------------------------------------------------------------------------------
// One CTA per (frame, channel) plane.  256 threads = 4 patch-row groups x (16 ky x 4 float4-of-kx).
// Every flow element is read exactly once, as 16-byte coalesced loads (64 contiguous bytes per 4 lanes).
// ---------------------------------------------------------------------------------------------------------
constexpr int OFR_THREADS = 256;

__global__ void __launch_bounds__(OFR_THREADS) of_patch_reduce_kernel(const float* __restrict__ flow,
                                                                      const float* __restrict__ fcw,
                                                                      __nv_bfloat16* __restrict__ z, int S, int P,
                                                                      int C, int ldz) {
  // generic in S (image) and P (patch) as long as P == 16 (float4 x 4 per patch row); checked on the host
  extern __shared__ float sh[];                 // fcw [g*g] then partial sums [4][P*P]
  const int g = S / P;
  float* sw = sh;
  float* part = sh + g * g;
  const int plane = blockIdx.x;                 // f*C + c
  const int tid = threadIdx.x;
  for (int i = tid; i < g * g; i += OFR_THREADS) sw[i] = fcw[i];
  __syncthreads();
  const int grp = tid >> 6;                     // 0..3: handles patch rows py = grp, grp+4, ...
  const int ky = (tid >> 2) & 15;
  const int kx4 = tid & 3;                      // float4 index within the 16-wide patch row
  const float* base = flow + (long long)plane * S * S;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int py = grp; py < g; py += 4) {
    const float4* rowp = reinterpret_cast<const float4*>(base + (long long)(py * P + ky) * S) + kx4;
    const float* wrow = sw + py * g;
#pragma unroll 7
    for (int px = 0; px < g; ++px) {
      const float4 v = __ldg(rowp + px * (P / 4));
      const float w = wrow[px];
      acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y);
      acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
    }
  }
  reinterpret_cast<float4*>(part)[tid] = acc;   // part[grp][ky][kx4*4..]
  __syncthreads();
  if (tid < 64) {
    float4 s = reinterpret_cast<float4*>(part)[tid];
#pragma unroll
    for (int gidx = 1; gidx < 4; ++gidx) {
      const float4 t = reinterpret_cast<float4*>(part)[gidx * 64 + tid];
      s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
    const int f = plane / C, c = plane - f * C;
    uint2 pk;
    pk.x = pack_bf16(s.x, s.y);
    pk.y = pack_bf16(s.z, s.w);
    // column (c*P + ky)*P + kx  == conv weight [768, C, P, P] flattened over (C,P,P)
    *reinterpret_cast<uint2*>(z + (long long)f * ldz + (c * P + (tid >> 2)) * P + (tid & 3) * 4) = pk;
  }
}

// ---------------------------------------------------------------------------------------------------------
// LayerNorm helpers for N = 768-class rows: one warp per row, NV float4 per lane.
// ---------------------------------------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ void warp_ln_store(float4 (&v)[NV], int nvec, int lane, int N, const float* gamma,
                                              const float* beta, float eps, float* of, __nv_bfloat16* ob) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) / static_cast<float>(N);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) {
      const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (cc * cc + d * d);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(N) + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + 32 * i;
    if (c < nvec) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c);
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x + b.x;
      o.y = (v[i].y - mean) * rstd * g.y + b.y;
      o.z = (v[i].z - mean) * rstd * g.z + b.z;
      o.w = (v[i].w - mean) * rstd * g.w + b.w;
      if (of) reinterpret_cast<float4*>(of)[c] = o;
      if (ob) {
        uint2 pk;
        pk.x = pack_bf16(o.x, o.y);
        pk.y = pack_bf16(o.z, o.w);
        reinterpret_cast<uint2*>(ob)[c] = pk;
      }
    }
  }
}

constexpr int ROW_NV = 8;    // N <= 1024

// rows (b, t), t in [0, L+2):  t==0 bos | 1..L y[b,t-1] | L+1 zeros ; row ends[b] := eos ; + pos[t] ; LayerNorm
__global__ void __launch_bounds__(256) sampler_assemble_kernel(const float* __restrict__ y, const float* __restrict__ bos,
                                                               const float* __restrict__ eos, const float* __restrict__ pos,
                                                               const int* __restrict__ ends, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, float eps, int B, int L, int N,
                                                               float* __restrict__ of, __nv_bfloat16* __restrict__ ob) {
  const int lane = threadIdx.x & 31;
  const int nvec = N >> 2;
  const long long rows = (long long)B * (L + 2);
  for (long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); r < rows; r += (long long)gridDim.x * 8) {
    const int b = static_cast<int>(r / (L + 2));
    const int t = static_cast<int>(r - (long long)b * (L + 2));
    const float* src = nullptr;
    if (t == ends[b]) src = eos;
    else if (t == 0) src = bos;
    else if (t <= L) src = y + ((long long)b * L + (t - 1)) * N;
    float4 v[ROW_NV];
#pragma unroll
    for (int i = 0; i < ROW_NV; ++i) {
      const int c = lane + 32 * i;
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < nvec) {
        const float4 p = __ldg(reinterpret_cast<const float4*>(pos + (long long)t * N) + c);
        float4 a = src ? __ldg(reinterpret_cast<const float4*>(src) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[i] = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
      }
    }
    warp_ln_store<ROW_NV>(v, nvec, lane, N, gamma, beta, eps, of ? of + r * N : nullptr, ob ? ob + r * N : nullptr);
  }
}

// out[r] = LayerNorm(table[ids[r]] + (pos ? pos[r % period] : 0) + (addvec ? addvec : 0))
__global__ void __launch_bounds__(256) embed_layernorm_kernel(const long long* __restrict__ ids, const float* __restrict__ table,
                                                              long long vocab, const float* __restrict__ pos, int period,
                                                              const float* __restrict__ addvec, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, float eps, long long R, int N,
                                                              float* __restrict__ of, long long ldf,
                                                              __nv_bfloat16* __restrict__ ob, long long ldb) {
  const int lane = threadIdx.x & 31;
  const int nvec = N >> 2;
  for (long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); r < R; r += (long long)gridDim.x * 8) {
    long long id = ids[r];
    // nn.Embedding raises (CPU) / trips a device-side assert (CUDA) on an id outside the table.  No exceptions or host
    // round trips here: the row comes out as NaN, which every consumer propagates -- loud, not silently clamped.
    const bool bad_id = id < 0 || id >= vocab;
    if (bad_id) id = 0;
    const float4* src = reinterpret_cast<const float4*>(table + id * N);
    float4 v[ROW_NV];
#pragma unroll
    for (int i = 0; i < ROW_NV; ++i) {
      const int c = lane + 32 * i;
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < nvec) {
        float4 a = __ldg(src + c);
        if (pos) {
          const float4 p = __ldg(reinterpret_cast<const float4*>(pos + (long long)(r % period) * N) + c);
          a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
        }
        if (addvec) {
          const float4 p = __ldg(reinterpret_cast<const float4*>(addvec) + c);
          a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
        }
        if (bad_id) a.x = a.y = a.z = a.w = __int_as_float(0x7fc00000);
        v[i] = a;
      }
    }
    warp_ln_store<ROW_NV>(v, nvec, lane, N, gamma, beta, eps, of ? of + r * ldf : nullptr, ob ? ob + r * ldb : nullptr);
  }
}

// In-place RoPE on bf16 head vectors: x[r, h*64 + 2i], x[r, h*64 + 2i+1] with position t = r % S:
//   x' = x * cos_pos + rotate_half(x) * sin_pos, sin_pos/cos_pos = table[t, i] / table[t, 32 + i] repeated pairwise,
//   rotate_half(x) = [-x1, x0, -x3, x2, ...]            (xropebert.py:334-377)
__global__ void __launch_bounds__(256) rope_bf16_kernel(__nv_bfloat16* __restrict__ x, long long ldx,
                                                        const float* __restrict__ table, long long rows, int S, int H,
                                                        int HD) {
  const int half = HD / 2;
  const long long pairs = rows * H * half;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < pairs; i += (long long)gridDim.x * blockDim.x) {
    const int pi = static_cast<int>(i % half);
    const long long rh = i / half;
    const int h = static_cast<int>(rh % H);
    const long long r = rh / H;
    const int t = static_cast<int>(r % S);
    const float sn = __ldg(table + (long long)t * HD + pi);
    const float cs = __ldg(table + (long long)t * HD + half + pi);
    __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(x + r * ldx + h * HD + 2 * pi);
    const float2 v = __bfloat1622float2(*p);
    const float o0 = v.x * cs + (-v.y) * sn;
    const float o1 = v.y * cs + v.x * sn;
    *p = __floats2bfloat162_rn(o0, o1);
  }
}

// A3: additive[b, i] = (1 - mask[b, i]) * value in fp32 (the reference's two statements), and the index of the row's
// last set position, sum(mask[b, :]) - 1, that TemporalOFEmbedding uses for the eos row.  One warp per row.
__global__ void __launch_bounds__(256) additive_mask_kernel(const long long* __restrict__ mask, int B, int S, float value,
                                                            float* __restrict__ out, int* __restrict__ last_index) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const long long* mr = mask + (long long)row * S;
  long long sum = 0;
  for (int i = lane; i < S; i += 32) {
    const long long m = mr[i];
    sum += m;
    if (out) out[(long long)row * S + i] = __fmul_rn(__fsub_rn(1.0f, (float)m), value);
  }
  if (last_index) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) last_index[row] = (int)(sum - 1);
  }
}

}  // namespace tgb

using namespace tgb;

extern "C" int tgb_additive_mask(const int64_t* mask, int B, int S, float value, float* out, int32_t* last_index,
                                 tgb_stream_t stream) {
  int st = check_device();
  if (st != TGB_OK) return st;
  TGB_CHECK_ARG(mask && (out || last_index), "additive_mask: null pointer");
  TGB_CHECK_ARG(B >= 0 && S > 0, "additive_mask: bad shape B=%d S=%d", B, S);
  if (B == 0) return TGB_OK;
  additive_mask_kernel<<<(B + 7) / 8, 256, 0, (cudaStream_t)stream>>>((const long long*)mask, B, S, value, out, last_index);
  TGB_CUDA(cudaGetLastError());
  return TGB_OK;
}

static int blocks_for(long long items, int per_block, int per_sm) {
  long long b = (items + per_block - 1) / per_block;
  const long long cap = (long long)sm_count() * per_sm;
  if (b > cap) b = cap;
  return b < 1 ? 1 : (int)b;
}

extern "C" int tgb_of_patch_reduce(const float* flow, const float* fc_weight, void* z_bf16, int frames, int C, int S,
                                   int P, int ldz, tgb_stream_t stream) {
  int st = check_device();
  if (st != TGB_OK) return st;
  TGB_CHECK_ARG(flow && fc_weight && z_bf16, "of_patch_reduce: null pointer");
  TGB_CHECK_ARG(P == 16 && S % P == 0 && frames >= 0 && C > 0, "of_patch_reduce: needs patch 16 and S %% 16 == 0");
  TGB_CHECK_ARG(ldz >= C * P * P && ldz % 4 == 0 && ((uintptr_t)flow % 16) == 0 && ((uintptr_t)z_bf16 % 8) == 0,
                "of_patch_reduce: bad ldz / alignment");
  if (frames == 0) return TGB_OK;
  const int g = S / P;
  const size_t smem = (size_t)(g * g + 4 * P * P) * sizeof(float);
  of_patch_reduce_kernel<<<frames * C, OFR_THREADS, smem, (cudaStream_t)stream>>>(flow, fc_weight, (__nv_bfloat16*)z_bf16, S, P, C, ldz);
  TGB_CUDA(cudaGetLastError());
  return TGB_OK;
}

extern "C" int tgb_sampler_assemble(const float* y, const float* bos, const float* eos, const float* pos,
                                    const int32_t* ends, const float* gamma, const float* beta, float eps, int B, int L,
                                    int N, float* out_f32, void* out_bf16, tgb_stream_t stream) {
  int st = check_device();
  if (st != TGB_OK) return st;
  TGB_CHECK_ARG(y && bos && eos && pos && ends && gamma && beta && (out_f32 || out_bf16), "sampler_assemble: null pointer");
  TGB_CHECK_ARG(B >= 0 && L > 0 && N % 4 == 0 && N <= 128 * ROW_NV, "sampler_assemble: bad shape");
  if (B == 0) return TGB_OK;
  sampler_assemble_kernel<<<blocks_for((long long)B * (L + 2), 8, 8), 256, 0, (cudaStream_t)stream>>>(
      y, bos, eos, pos, ends, gamma, beta, eps, B, L, N, out_f32, (__nv_bfloat16*)out_bf16);
  TGB_CUDA(cudaGetLastError());
  return TGB_OK;
}

extern "C" int tgb_embed_layernorm(const int64_t* ids, const float* table, int64_t vocab, const float* pos, int period,
                                   const float* addvec, const float* gamma, const float* beta, float eps, int64_t rows,
                                   int N, float* out_f32, int64_t ld_f32, void* out_bf16, int64_t ld_bf16,
                                   tgb_stream_t stream) {
  int st = check_device();
  if (st != TGB_OK) return st;
  TGB_CHECK_ARG(ids && table && gamma && beta && (out_f32 || out_bf16), "embed_layernorm: null pointer");
  TGB_CHECK_ARG(rows >= 0 && vocab > 0 && N % 4 == 0 && N <= 128 * ROW_NV && (!pos || period > 0), "embed_layernorm: bad shape");
  TGB_CHECK_ARG((!out_f32 || ld_f32 % 4 == 0) && (!out_bf16 || ld_bf16 % 4 == 0), "embed_layernorm: bad leading dims");
  if (rows == 0) return TGB_OK;
  embed_layernorm_kernel<<<blocks_for(rows, 8, 8), 256, 0, (cudaStream_t)stream>>>(
      (const long long*)ids, table, vocab, pos, period, addvec, gamma, beta, eps, rows, N, out_f32, ld_f32,
      (__nv_bfloat16*)out_bf16, ld_bf16);
  TGB_CUDA(cudaGetLastError());
  return TGB_OK;
}

extern "C" int tgb_rope_bf16(void* x, int64_t ldx, const float* table, int64_t rows, int S, int H, int HD,
                             tgb_stream_t stream) {
  int st = check_device();
  if (st != TGB_OK) return st;
  TGB_CHECK_ARG(x && table && rows >= 0 && S > 0 && H > 0 && HD % 2 == 0 && ldx % 2 == 0 && ((uintptr_t)x % 4) == 0,
                "rope: bad arguments");
  if (rows == 0) return TGB_OK;
  rope_bf16_kernel<<<blocks_for(rows * H * (HD / 2), 256, 8), 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)x, ldx, table, rows, S, H, HD);
  TGB_CUDA(cudaGetLastError());
  return TGB_OK;
}

namespace tgb {
// out[r, n] = bias[n] + sum_k x[r,k] * w[n,k] in fp32 for tiny N (the sampler's mrc_head = Linear(768, 2),
// xropebert.py:953,1164).  One warp per row; fp32 end to end so the frame scores keep full precision.
template <int NMAX>
__global__ void __launch_bounds__(256) linear_small_kernel(const float* __restrict__ x, long long ldx,
                                                           const float* __restrict__ w, const float* __restrict__ bias,
                                                           long long M, int N, int K, float* __restrict__ out, long long ldo) {
  const int lane = threadIdx.x & 31;
  for (long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); r < M; r += (long long)gridDim.x * 8) {
    float acc[NMAX];
#pragma unroll
    for (int n = 0; n < NMAX; ++n) acc[n] = 0.f;
    const float4* xr = reinterpret_cast<const float4*>(x + r * ldx);
    for (int c = lane; c < (K >> 2); c += 32) {
      const float4 a = xr[c];
#pragma unroll
      for (int n = 0; n < NMAX; ++n) {
        if (n < N) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(w + (long long)n * K) + c);
          acc[n] = fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, acc[n]))));
        }
      }
    }
#pragma unroll
    for (int n = 0; n < NMAX; ++n) {
      const float s = warp_sum(acc[n]);
      if (lane == 0 && n < N) out[r * ldo + n] = s + (bias ? bias[n] : 0.f);
    }
  }
}
}  // namespace tgb

extern "C" int tgb_linear_small_f32(const float* x, int64_t ldx, const float* w, const float* bias, int64_t M, int N,
                                    int K, float* out, int64_t ldo, tgb_stream_t stream) {
  int st = check_device();
  if (st != TGB_OK) return st;
  TGB_CHECK_ARG(x && w && out && M >= 0 && N > 0 && N <= 8 && K % 4 == 0 && ldx % 4 == 0 && ldo >= N,
                "linear_small: needs N <= 8, K %% 4 == 0");
  TGB_CHECK_ARG(((uintptr_t)x % 16) == 0 && ((uintptr_t)w % 16) == 0, "linear_small: misaligned");
  if (M == 0) return TGB_OK;
  linear_small_kernel<8><<<blocks_for(M, 8, 8), 256, 0, (cudaStream_t)stream>>>(x, ldx, w, bias, M, N, K, out, ldo);
  TGB_CUDA(cudaGetLastError());
  return TGB_OK;
}
