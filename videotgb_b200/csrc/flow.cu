// RAFT optical-flow extractor building blocks (src/models/components/xraft.py, raft_utils/).
// Activations are NHWC ("pixel rows"): [N*H*W, C] with the channel dimension contiguous, so every convolution is
// im2col (this file) + the tcgen05 GEMM (gemm_tcgen05.cu), whose output rows are directly the next layer's pixels.
// Everything here is memory-bound glue: 16-byte accesses wherever the channel count allows.
#include "common.cuh"

namespace tgb {

// ---------------------------------------------------------------------------------------------------------
// im2col: out[(n,oy,ox), (ky*kw + kx)*C + c] = x[(n, oy*sh - ph + ky, ox*sw - pw + kx), c]  (0 outside the image),
// columns [kh*kw*C, ldo) zero.  VEC path: C % 8 == 0 -> one 16-byte copy per 8 channels.
// ---------------------------------------------------------------------------------------------------------
// Vector path: one thread owns (output pixel, 8-channel chunk) and walks the kh*kw taps; consecutive lanes write
// consecutive 16-byte chunks of a tap segment.  blockIdx.y walks the output rows (n, oy), blockIdx.x the (ox, chunk) pairs
// of one row, so the only per-thread divisions are two 32-bit ones (the flat 64-bit index it replaces cost five 64-bit
// divisions per thread -- more instructions than the copies themselves).
__global__ void __launch_bounds__(256) im2col_nhwc_vec_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, int H, int W,
                                                              int C, int kh, int kw, int sh, int sw, int ph, int pw, int Ho,
                                                              int Wo, int NHo, __nv_bfloat16* __restrict__ out,
                                                              long long ldo) {
  const unsigned c8 = static_cast<unsigned>(C) >> 3;
  const int K = kh * kw * C;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= static_cast<unsigned>(Wo) * c8) return;
  const int ox = static_cast<int>(idx / c8);
  const int c = static_cast<int>(idx - static_cast<unsigned>(ox) * c8) << 3;
  const int ix0 = ox * sw - pw;
  for (int ry = blockIdx.y; ry < NHo; ry += gridDim.y) {
    const int n = ry / Ho;
    const int oy = ry - n * Ho;
    const int iy0 = oy * sh - ph;
    const long long r = (long long)ry * Wo + ox;
    __nv_bfloat16* o = out + r * ldo + c;
    for (int ky = 0; ky < kh; ++ky) {
      const int iy = iy0 + ky;
      const bool yok = iy >= 0 && iy < H;
      const __nv_bfloat16* srow = x + (((long long)n * H + iy) * W) * ldx + c;
      for (int kx = 0; kx < kw; ++kx) {
        const int ix = ix0 + kx;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (yok && ix >= 0 && ix < W) v = __ldg(reinterpret_cast<const uint4*>(srow + (long long)ix * ldx));
        *reinterpret_cast<uint4*>(o + (ky * kw + kx) * C) = v;
      }
    }
    if (c == 0)                                                 // zero the K padding of this row (up to ldo)
      for (int k = K; k < (int)ldo; ++k) out[r * ldo + k] = __float2bfloat16(0.f);
  }
}

// Scalar path (C = 2 / 3 input channels): one thread per (ox, output column) pair, walking the output rows (n, oy) its
// blockIdx.y is responsible for; the column decomposition (tap, channel) is done once per thread in 32-bit arithmetic.
__global__ void __launch_bounds__(256) im2col_nhwc_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, int H, int W,
                                                          int C, int kh, int kw, int sh, int sw, int ph, int pw, int Ho,
                                                          int Wo, int NHo, __nv_bfloat16* __restrict__ out,
                                                          long long ldo) {
  const int K = kh * kw * C;
  const unsigned units = static_cast<unsigned>(ldo);
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= static_cast<unsigned>(Wo) * units) return;
  const int ox = static_cast<int>(idx / units);
  const int col = static_cast<int>(idx - static_cast<unsigned>(ox) * units);
  const bool in_k = col < K;
  int c = 0, ky = 0, kx = 0;
  if (in_k) {
    const int tap = col / C;
    c = col - tap * C;
    ky = tap / kw;
    kx = tap - ky * kw;
  }
  const int ix = ox * sw - pw + kx;
  const bool xok = in_k && ix >= 0 && ix < W;
  for (int ry = blockIdx.y; ry < NHo; ry += gridDim.y) {
    const int n = ry / Ho;
    const int oy = ry - n * Ho;
    const int iy = oy * sh - ph + ky;
    const bool ok = xok && iy >= 0 && iy < H;
    const long long r = (long long)ry * Wo + ox;
    out[r * ldo + col] = ok ? x[(((long long)n * H + iy) * W + ix) * ldx + c] : __float2bfloat16(0.f);
  }
}

// ---------------------------------------------------------------------------------------------------------
// InstanceNorm2d (affine=False, eps) over H*W per (n, c) on fp32 NHWC, optional ReLU; fp32 and/or bf16 outputs.
// Two kernels: (1) per (n, 32-channel group, pixel split) a two-pass (mean, M2) over the split's pixels -> partials in
// a caller-provided workspace; (2) a fully parallel apply that merges the partials (Chan's formula) per channel.
// ---------------------------------------------------------------------------------------------------------
constexpr int IN_SPLITS = 16;

__global__ void __launch_bounds__(256) instnorm_stats_kernel(const float* __restrict__ x, int HW, int C,
                                                             float* __restrict__ ws) {      // ws [N][IN_SPLITS][C][2]
  __shared__ float s_a[8][32], s_mean[32];
  const int n = blockIdx.y, split = blockIdx.z, cl = threadIdx.x & 31, c = blockIdx.x * 32 + cl, pl = threadIdx.x >> 5;
  const int per = (HW + IN_SPLITS - 1) / IN_SPLITS;
  const int p0 = split * per, p1 = min(HW, p0 + per);
  const int cnt = max(p1 - p0, 0);
  const bool cok = c < C;
  const float* xb = x + (long long)n * HW * C;
  float s = 0.f;
  if (cok)
    for (int p = p0 + pl; p < p1; p += 8) s += xb[(long long)p * C + c];
  s_a[pl][cl] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += s_a[k][threadIdx.x];
    s_mean[threadIdx.x] = cnt > 0 ? t / (float)cnt : 0.f;
  }
  __syncthreads();
  const float mean = s_mean[cl];
  float q = 0.f;
  if (cok)
    for (int p = p0 + pl; p < p1; p += 8) {
      const float d = xb[(long long)p * C + c] - mean;
      q += d * d;
    }
  __syncthreads();
  s_a[pl][cl] = q;
  __syncthreads();
  if (threadIdx.x < 32 && cok) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += s_a[k][threadIdx.x];
    float* w = ws + (((long long)n * IN_SPLITS + split) * C + c) * 2;
    w[0] = mean;
    w[1] = t;                                                    // M2 of this split
  }
}

// merge the split partials (Chan's formula) once per (n, c): fin[(n*C + c)*2] = (mean, rstd)
__global__ void __launch_bounds__(256) instnorm_finalize_kernel(const float* __restrict__ ws, int HW, int C, float eps,
                                                                long long NC, float* __restrict__ fin) {
  const int per = (HW + IN_SPLITS - 1) / IN_SPLITS;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < NC; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / C;
    const int c = (int)(i - n * C);
    float mean = 0.f, m2 = 0.f, cnt = 0.f;
#pragma unroll
    for (int sidx = 0; sidx < IN_SPLITS; ++sidx) {
      const int p0 = sidx * per;
      const float nb = (float)max(min(HW, p0 + per) - p0, 0);
      if (nb > 0.f) {
        const float* w = ws + ((n * IN_SPLITS + sidx) * C + c) * 2;
        const float delta = w[0] - mean, tot = cnt + nb;
        mean += delta * nb / tot;
        m2 += w[1] + delta * delta * cnt * nb / tot;
        cnt = tot;
      }
    }
    fin[i * 2] = mean;
    fin[i * 2 + 1] = rsqrtf(m2 / (float)HW + eps);              // biased variance, as F.instance_norm
  }
}

// fully parallel apply: 4 channels per thread, 16-byte accesses; blockIdx.y = image, 32-bit index math inside it (the
// flat 64-bit index this replaces paid three 64-bit divisions per 16 bytes and ran at a quarter of HBM speed)
__global__ void __launch_bounds__(256) instnorm_apply_kernel(const float* __restrict__ x, const float* __restrict__ fin, int HW,
                                                             int C, int relu, float* __restrict__ of,
                                                             __nv_bfloat16* __restrict__ ob) {
  const unsigned c4n = static_cast<unsigned>(C) >> 2;
  const unsigned per_image = static_cast<unsigned>(HW) * c4n;
  const long long n = blockIdx.y;
  const float* xn = x + n * HW * C;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < per_image; i += gridDim.x * blockDim.x) {
    const unsigned p = i / c4n;
    const int c = static_cast<int>(i - p * c4n) << 2;
    const long long off = (long long)p * C + c;
    const float4 v = *reinterpret_cast<const float4*>(xn + off);
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(fin + (n * C + c) * 2));         // mean0 rstd0 mean1 rstd1
    const float4 s1 = __ldg(reinterpret_cast<const float4*>(fin + (n * C + c) * 2 + 4));
    float4 o;
    o.x = (v.x - s0.x) * s0.y; o.y = (v.y - s0.z) * s0.w; o.z = (v.z - s1.x) * s1.y; o.w = (v.w - s1.z) * s1.w;
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    if (of) *reinterpret_cast<float4*>(of + n * HW * C + off) = o;
    if (ob) {
      uint2 pk;
      pk.x = pack_bf16(o.x, o.y);
      pk.y = pack_bf16(o.z, o.w);
      *reinterpret_cast<uint2*>(ob + n * HW * C + off) = pk;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// out = act(scale * x (+ y)) on [M, N] with row pitches; act: 0 identity, 1 relu, 2 sigmoid, 3 tanh.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_apply(float v, int mode) {
  if (mode == 1) return fmaxf(v, 0.f);
  if (mode == 2) return 1.f / (1.f + __expf(-v));
  if (mode == 3) return tanhf(v);
  return v;
}

// 4-wide variant for N % 4 == 0 with 16-byte aligned rows (every hot use)
__global__ void __launch_bounds__(256) act4_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ y,
                                                   long long ldy, long long M, int N4, int mode, float scale,
                                                   float* __restrict__ of, long long ldf, __nv_bfloat16* __restrict__ ob,
                                                   long long ldb) {
  const long long total = M * N4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / N4;
    const int c = (int)(i - r * N4) * 4;
    float4 v = *reinterpret_cast<const float4*>(x + r * ldx + c);
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    if (y) {
      const float4 u = *reinterpret_cast<const float4*>(y + r * ldy + c);
      v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
    }
    v.x = act_apply(v.x, mode); v.y = act_apply(v.y, mode); v.z = act_apply(v.z, mode); v.w = act_apply(v.w, mode);
    if (of) *reinterpret_cast<float4*>(of + r * ldf + c) = v;
    if (ob) {
      uint2 pk;
      pk.x = pack_bf16(v.x, v.y);
      pk.y = pack_bf16(v.z, v.w);
      *reinterpret_cast<uint2*>(ob + r * ldb + c) = pk;
    }
  }
}

__global__ void __launch_bounds__(256) act_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ y,
                                                  long long ldy, long long M, int N, int mode, float scale,
                                                  float* __restrict__ of, long long ldf, __nv_bfloat16* __restrict__ ob,
                                                  long long ldb) {
  const long long total = M * N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / N;
    const int c = (int)(i - r * N);
    float v = x[r * ldx + c] * scale;
    if (y) v += y[r * ldy + c];
    v = act_apply(v, mode);
    if (of) of[r * ldf + c] = v;
    if (ob) ob[r * ldb + c] = __float2bfloat16(v);
  }
}

// SepConvGRU gates (raft_utils/update.py:42-59).  zr: fp32 [M, 2C] = [z_pre | r_pre] (the two convolutions share their
// input and run as one GEMM).  rh_out = sigmoid(r_pre) * h as bf16 into a wider (concat) buffer.
__global__ void __launch_bounds__(256) gru_rh_kernel(const float* __restrict__ zr, const float* __restrict__ h, long long M,
                                                     int C, __nv_bfloat16* __restrict__ rh, long long ld) {
  const long long total = M * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C;
    const int c = (int)(i - r * C);
    const float rg = 1.f / (1.f + __expf(-zr[r * 2 * C + C + c]));
    rh[r * ld + c] = __float2bfloat16(rg * h[i]);
  }
}
// h <- (1 - z) h + z tanh(q_pre);  also the bf16 copy into the concat buffer.
__global__ void __launch_bounds__(256) gru_update_kernel(const float* __restrict__ zr, const float* __restrict__ q,
                                                         float* __restrict__ h, long long M, int C,
                                                         __nv_bfloat16* __restrict__ hb, long long ld) {
  const long long total = M * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C;
    const int c = (int)(i - r * C);
    const float z = 1.f / (1.f + __expf(-zr[r * 2 * C + c]));
    const float hn = (1.f - z) * h[i] + z * tanhf(q[i]);
    h[i] = hn;
    hb[r * ld + c] = __float2bfloat16(hn);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Correlation pyramid: 2x2 average pooling of the last two dims of [R, h, w] (floor), and the 9x9 lookup.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) corr_pool_kernel(const float* __restrict__ in, float* __restrict__ out, long long R,
                                                        int h, int w) {
  const int ho = h >> 1, wo = w >> 1;
  const long long total = R * ho * wo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % wo);
    const long long t = i / wo;
    const int y = (int)(t % ho);
    const long long r = t / ho;
    const float* p = in + (r * h + 2 * y) * w + 2 * x;
    out[i] = (p[0] + p[1] + p[w] + p[w + 1]) * 0.25f;
  }
}

struct CorrPyr { const float* lvl[4]; int h[4]; int w[4]; };

// out[(n,p), l*81 + a*9 + b] = bilinear(corr_l[(n,p)], x/2^l + (a-4), y/2^l + (b-4)) with zero padding and
// align_corners=True pixel coordinates -- x takes the FIRST window index, as the reference (corr.py:37-46).
// (x, y) = pixel grid + flow.  One warp per (n, p); lanes sweep the 324 outputs.
__global__ void __launch_bounds__(256) corr_lookup_kernel(CorrPyr pyr, const float* __restrict__ flow, long long NP, int h1,
                                                          int w1, __nv_bfloat16* __restrict__ out, long long ldo) {
  const int lane = threadIdx.x & 31;
  for (long long np = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); np < NP;
       np += (long long)gridDim.x * (blockDim.x >> 5)) {
    const int p = (int)(np % ((long long)h1 * w1));
    const float cx = (float)(p % w1) + flow[np * 2], cy = (float)(p / w1) + flow[np * 2 + 1];
    for (int o = lane; o < (int)ldo; o += 32) {
      float v = 0.f;
      if (o < 324) {
        const int l = o / 81, k = o - l * 81, a = k / 9, b = k - a * 9;
        const int h = pyr.h[l], w = pyr.w[l];
        const float inv = 1.f / (float)(1 << l);
        // the reference normalises to [-1,1] and grid_sample maps back: ((2x/(w-1) - 1) + 1)/2 * (w-1)
        const float xs = cx * inv + (float)(a - 4), ys = cy * inv + (float)(b - 4);
        const float gx = ((2.f * xs / (float)(w - 1) - 1.f) + 1.f) * 0.5f * (float)(w - 1);
        const float gy = ((2.f * ys / (float)(h - 1) - 1.f) + 1.f) * 0.5f * (float)(h - 1);
        const float fx = floorf(gx), fy = floorf(gy);
        const int x0 = (int)fx, y0 = (int)fy;
        const float ax = gx - fx, ay = gy - fy;
        const float* c = pyr.lvl[l] + np * h * w;
        auto tap = [&](int yy, int xx) -> float { return (yy >= 0 && yy < h && xx >= 0 && xx < w) ? c[yy * w + xx] : 0.f; };
        v = tap(y0, x0) * (1.f - ax) * (1.f - ay) + tap(y0, x0 + 1) * ax * (1.f - ay) + tap(y0 + 1, x0) * (1.f - ax) * ay +
            tap(y0 + 1, x0 + 1) * ax * ay;
      }
      out[np * ldo + o] = __float2bfloat16(v);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Convex upsampling (xraft.py:88-99): mask fp32 [N*h*w, 576] (channel = k*64 + dy*8 + dx, already scaled by 0.25),
// flow fp32 [N*h*w, 2] -> out fp32 [N, 2, 8h, 8w] = sum_k softmax_k(mask) * 8*flow(neighbour k, zero padded).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) raft_upsample_kernel(const float* __restrict__ mask, long long ldm,
                                                            const float* __restrict__ flow, int N, int h, int w,
                                                            float* __restrict__ out) {
  const long long total = (long long)N * h * w * 64;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int sub = (int)(i & 63);
    const long long np = i >> 6;
    const int p = (int)(np % ((long long)h * w));
    const long long n = np / ((long long)h * w);
    const int y = p / w, x = p - y * w;
    const float* m = mask + np * ldm + sub;
    float mv[9], mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < 9; ++k) { mv[k] = m[k * 64]; mx = fmaxf(mx, mv[k]); }
    float den = 0.f, fx = 0.f, fy = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const float e = __expf(mv[k] - mx);
      den += e;
      const int yy = y + k / 3 - 1, xx = x + k % 3 - 1;
      if (yy >= 0 && yy < h && xx >= 0 && xx < w) {
        const float* f = flow + ((n * h + yy) * w + xx) * 2;
        fx += e * f[0];
        fy += e * f[1];
      }
    }
    const float s = 8.f / den;
    const int oy = y * 8 + (sub >> 3), ox = x * 8 + (sub & 7);
    out[((n * 2 + 0) * (8LL * h) + oy) * (8LL * w) + ox] = fx * s;
    out[((n * 2 + 1) * (8LL * h) + oy) * (8LL * w) + ox] = fy * s;
  }
}

// image fp32 NCHW (0..255) -> bf16 NHWC of 2*(x/255) - 1  (xraft.py:106-107)
__global__ void __launch_bounds__(256) raft_prep_kernel(const float* __restrict__ img, int C, int HW, long long N,
                                                        __nv_bfloat16* __restrict__ out, long long ldo) {
  const long long total = N * HW * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long np = i / C;
    const long long n = np / HW;
    const int p = (int)(np - n * HW);
    const float v = 2.f * (img[(n * C + c) * HW + p] / 255.0f) - 1.0f;
    out[np * ldo + c] = __float2bfloat16(v);
  }
}

static int ew_grid(long long work) {
  long long blocks = (work + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace tgb

using namespace tgb;

extern "C" int tgb_im2col_nhwc_bf16(const void* x, int64_t ldx, int N, int H, int W, int C, int kh, int kw, int sh, int sw,
                                    int ph, int pw, void* out, int64_t ldo, tgb_stream_t stream) {
  int st = check_device();
  if (st != TGB_OK) return st;
  TGB_CHECK_ARG(x && out, "im2col: null pointer");
  TGB_CHECK_ARG(N >= 0 && H > 0 && W > 0 && C > 0 && kh > 0 && kw > 0 && sh > 0 && sw > 0 && ph >= 0 && pw >= 0, "im2col: bad shape");
  TGB_CHECK_ARG(ldx >= C && ldo >= (int64_t)kh * kw * C && ldo % 8 == 0, "im2col: ldx >= C, ldo >= kh*kw*C and ldo %% 8 == 0 required");
  const int Ho = (H + 2 * ph - kh) / sh + 1, Wo = (W + 2 * pw - kw) / sw + 1;
  TGB_CHECK_ARG(Ho > 0 && Wo > 0, "im2col: empty output");
  const long long rows = (long long)N * Ho * Wo;
  if (rows == 0) return TGB_OK;
  const bool vec = C % 8 == 0 && ldx % 8 == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)out % 16) == 0;
  auto xb = (const __nv_bfloat16*)x;
  auto ob = (__nv_bfloat16*)out;
  if (vec) {
    const long long nho = (long long)N * Ho;
    TGB_CHECK_ARG(nho < 0x7fffffffLL && (long long)Wo * (C / 8) < 0x7fffffffLL, "im2col: problem too large");
    dim3 grid((unsigned)(((long long)Wo * (C / 8) + 255) / 256), (unsigned)(nho < 65535 ? nho : 65535));
    im2col_nhwc_vec_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(xb, ldx, H, W, C, kh, kw, sh, sw, ph, pw, Ho, Wo, (int)nho, ob, ldo);
  } else {
    const long long nho = (long long)N * Ho;
    TGB_CHECK_ARG(nho < 0x7fffffffLL && (long long)Wo * ldo < 0x7fffffffLL, "im2col: problem too large");
    // each thread walks NHo / gridDim.y rows: a few rows per thread keep the per-thread column decomposition amortised
    long long gy = nho < 65535 ? nho : 65535;
    const long long gx = ((long long)Wo * ldo + 255) / 256;
    while (gy > 1 && gx * gy > 64LL * sm_count()) gy = (gy + 1) / 2;
    im2col_nhwc_kernel<<<dim3((unsigned)gx, (unsigned)gy), 256, 0, (cudaStream_t)stream>>>(xb, ldx, H, W, C, kh, kw, sh, sw, ph, pw, Ho, Wo, (int)nho, ob, ldo);
  }
  TGB_CUDA(cudaGetLastError());
  return TGB_OK;
}

extern "C" int tgb_instance_norm_nhwc(const float* x, int N, int HW, int C, float eps, int relu, float* workspace,
                                      float* out_f32, void* out_bf16, tgb_stream_t stream) {
  int st = check_device();
  if (st != TGB_OK) return st;
  TGB_CHECK_ARG(x && workspace && (out_f32 || out_bf16) && N >= 0 && HW > 0 && C > 0 && C % 4 == 0 && N <= 65535,
                "instance_norm: bad arguments (C %% 4 == 0)");
  TGB_CHECK_ARG(((uintptr_t)x % 16) == 0 && ((uintptr_t)workspace % 16) == 0, "instance_norm: misaligned");
  if (N == 0) return TGB_OK;
  float* fin = workspace + (size_t)N * IN_SPLITS * C * 2;
  dim3 grid((C + 31) / 32, N, IN_SPLITS);
  instnorm_stats_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, HW, C, workspace);
  instnorm_finalize_kernel<<<ew_grid((long long)N * C), 256, 0, (cudaStream_t)stream>>>(workspace, HW, C, eps, (long long)N * C, fin);
  const long long per_image = (long long)HW * (C / 4);
  TGB_CHECK_ARG(per_image < 0x7fffffffLL && N <= 65535, "instance_norm: problem too large");
  long long bx = (per_image + 255) / 256;
  const long long cap = ((long long)sm_count() * 16 + N - 1) / N;                  // ~16 CTAs per SM over all images
  if (bx > cap) bx = cap > 0 ? cap : 1;
  instnorm_apply_kernel<<<dim3((unsigned)bx, (unsigned)N), 256, 0, (cudaStream_t)stream>>>(x, fin, HW, C, relu, out_f32,
                                                                                          (__nv_bfloat16*)out_bf16);
  TGB_CUDA(cudaGetLastError());
  return TGB_OK;
}

extern "C" int tgb_act_f32(const float* x, int64_t ldx, const float* y, int64_t ldy, int64_t M, int N, int mode, float scale,
                           float* out_f32, int64_t ld_f32, void* out_bf16, int64_t ld_bf16, tgb_stream_t stream) {
  int st = check_device();
  if (st != TGB_OK) return st;
  TGB_CHECK_ARG(x && (out_f32 || out_bf16) && M >= 0 && N > 0 && mode >= 0 && mode <= 3, "act: bad arguments");
  if (M == 0) return TGB_OK;
  const bool v4 = N % 4 == 0 && ldx % 4 == 0 && ((uintptr_t)x % 16) == 0 &&
                  (!y || (ldy % 4 == 0 && ((uintptr_t)y % 16) == 0)) &&
                  (!out_f32 || (ld_f32 % 4 == 0 && ((uintptr_t)out_f32 % 16) == 0)) &&
                  (!out_bf16 || (ld_bf16 % 4 == 0 && ((uintptr_t)out_bf16 % 8) == 0));
  if (v4)
    act4_kernel<<<ew_grid(M * (N / 4)), 256, 0, (cudaStream_t)stream>>>(x, ldx, y, ldy, M, N / 4, mode, scale, out_f32, ld_f32,
                                                                     (__nv_bfloat16*)out_bf16, ld_bf16);
  else
    act_kernel<<<ew_grid(M * N), 256, 0, (cudaStream_t)stream>>>(x, ldx, y, ldy, M, N, mode, scale, out_f32, ld_f32,
                                                                (__nv_bfloat16*)out_bf16, ld_bf16);
  TGB_CUDA(cudaGetLastError());
  return TGB_OK;
}

extern "C" int tgb_gru_rh(const float* zr, const float* h, int64_t M, int C, void* rh_bf16, int64_t ld, tgb_stream_t stream) {
  int st = check_device();
  if (st != TGB_OK) return st;
  TGB_CHECK_ARG(zr && h && rh_bf16 && M >= 0 && C > 0 && ld >= C, "gru_rh: bad arguments");
  if (M == 0) return TGB_OK;
  gru_rh_kernel<<<ew_grid(M * C), 256, 0, (cudaStream_t)stream>>>(zr, h, M, C, (__nv_bfloat16*)rh_bf16, ld);
  TGB_CUDA(cudaGetLastError());
  return TGB_OK;
}

extern "C" int tgb_gru_update(const float* zr, const float* q, float* h, int64_t M, int C, void* h_bf16, int64_t ld,
                              tgb_stream_t stream) {
  int st = check_device();
  if (st != TGB_OK) return st;
  TGB_CHECK_ARG(zr && q && h && h_bf16 && M >= 0 && C > 0 && ld >= C, "gru_update: bad arguments");
  if (M == 0) return TGB_OK;
  gru_update_kernel<<<ew_grid(M * C), 256, 0, (cudaStream_t)stream>>>(zr, q, h, M, C, (__nv_bfloat16*)h_bf16, ld);
  TGB_CUDA(cudaGetLastError());
  return TGB_OK;
}

extern "C" int tgb_corr_pool(const float* in, float* out, int64_t R, int h, int w, tgb_stream_t stream) {
  int st = check_device();
  if (st != TGB_OK) return st;
  TGB_CHECK_ARG(in && out && R >= 0 && h >= 2 && w >= 2, "corr_pool: bad arguments");
  if (R == 0) return TGB_OK;
  corr_pool_kernel<<<ew_grid(R * (h / 2) * (w / 2)), 256, 0, (cudaStream_t)stream>>>(in, out, R, h, w);
  TGB_CUDA(cudaGetLastError());
  return TGB_OK;
}

extern "C" int tgb_corr_lookup(const float* l0, const float* l1, const float* l2, const float* l3, const float* flow, int N,
                               int h, int w, void* out_bf16, int64_t ldo, tgb_stream_t stream) {
  int st = check_device();
  if (st != TGB_OK) return st;
  TGB_CHECK_ARG(l0 && l1 && l2 && l3 && flow && out_bf16, "corr_lookup: null pointer");
  TGB_CHECK_ARG(N >= 0 && h >= 16 && w >= 16 && ldo >= 324, "corr_lookup: feature map must be at least 16x16 (4 pyramid levels), ldo >= 324");
  if (N == 0) return TGB_OK;
  CorrPyr p;
  const float* ls[4] = {l0, l1, l2, l3};
  int hh = h, ww = w;
  for (int i = 0; i < 4; ++i) { p.lvl[i] = ls[i]; p.h[i] = hh; p.w[i] = ww; hh >>= 1; ww >>= 1; }
  const long long NP = (long long)N * h * w;
  long long blocks = (NP + 7) / 8;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  corr_lookup_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(p, flow, NP, h, w, (__nv_bfloat16*)out_bf16, ldo);
  TGB_CUDA(cudaGetLastError());
  return TGB_OK;
}

extern "C" int tgb_raft_upsample(const float* mask, int64_t ldm, const float* flow, int N, int h, int w, float* out,
                                 tgb_stream_t stream) {
  int st = check_device();
  if (st != TGB_OK) return st;
  TGB_CHECK_ARG(mask && flow && out && N >= 0 && h > 0 && w > 0 && ldm >= 576, "raft_upsample: bad arguments");
  if (N == 0) return TGB_OK;
  raft_upsample_kernel<<<ew_grid((long long)N * h * w * 64), 256, 0, (cudaStream_t)stream>>>(mask, ldm, flow, N, h, w, out);
  TGB_CUDA(cudaGetLastError());
  return TGB_OK;
}

extern "C" int tgb_raft_prep_image(const float* img_nchw, int N, int C, int HW, void* out_bf16, int64_t ldo, tgb_stream_t stream) {
  int st = check_device();
  if (st != TGB_OK) return st;
  TGB_CHECK_ARG(img_nchw && out_bf16 && N >= 0 && C > 0 && HW > 0 && ldo >= C, "raft_prep_image: bad arguments");
  if (N == 0) return TGB_OK;
  raft_prep_kernel<<<ew_grid((long long)N * C * HW), 256, 0, (cudaStream_t)stream>>>(img_nchw, C, HW, N, (__nv_bfloat16*)out_bf16, ldo);
  TGB_CUDA(cudaGetLastError());
  return TGB_OK;
}
