// HBM-bound frame kernels: normalise (A0), patchify / im2col for the patch-embed GEMM (A10), gather (A9),
// CLS-row fill.  All are pure streaming kernels: 16-byte coalesced accesses, grid-stride over a grid sized as a
// multiple of the SM count.
#include "common.cuh"
#include <cstring>

namespace tgb {

struct Norm3 { float mean[3]; float stdv[3]; };

// (u/255 - mean)/std with the reference's op order and IEEE fp32 rounding at every step.
__device__ __forceinline__ float norm_px(uint8_t u, float mean, float stdv) {
  return __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(u), 255.0f), mean), stdv);
}

// Division-free evaluation of norm_px: q = a*(1/255) and p = t*(1/std), each corrected by one FMA residual step.  For
// divisors that are compile/run-time constants this reproduces the correctly rounded IEEE quotient for (almost) every
// input; the HOST verifies all 3 x 256 possible inputs against real IEEE division before a kernel is allowed to use it
// (norm_fast_is_exact), otherwise the shared-memory table of exact results is used.
struct NormFast { float r255; float mean[3]; float stdv[3]; float rstd[3]; };
__host__ __device__ __forceinline__ float norm_px_fast(float a, float r255, float mean, float stdv, float rstd) {
#ifdef __CUDA_ARCH__
  float q = __fmul_rn(a, r255);
  q = __fmaf_rn(__fmaf_rn(-q, 255.0f, a), r255, q);
  const float t = __fsub_rn(q, mean);
  float p = __fmul_rn(t, rstd);
  return __fmaf_rn(__fmaf_rn(-p, stdv, t), rstd, p);
#else
  volatile float q0 = a * r255;
  volatile float q = fmaf(fmaf(-q0, 255.0f, a), r255, q0);
  volatile float t = q - mean;
  volatile float p0 = t * rstd;
  return fmaf(fmaf(-p0, stdv, t), rstd, p0);
#endif
}

// u8 [T,H,W,3] -> f32 [T,3,H,W].  The 3 x 256 possible results are computed once per CTA with the exact reference
// arithmetic (IEEE division is ~10 instructions; doing it per pixel made the kernel ALU-bound at 55 % of HBM) and
// looked up from shared memory, so the result stays bit-exact.  One thread handles 4 consecutive pixels: reads 12
// bytes (three 4-byte words, coalesced across the warp) and writes one float4 per channel plane.
__global__ void __launch_bounds__(256) normalize_u8_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst,
                                                           int T, int hw, Norm3 nm) {
  __shared__ float lut[3][256];
  for (int i = threadIdx.x; i < 768; i += blockDim.x) lut[i >> 8][i & 255] = norm_px((uint8_t)(i & 255), nm.mean[i >> 8], nm.stdv[i >> 8]);
  __syncthreads();
  // blockIdx.y walks the frames, blockIdx.x the pixel quads of a frame: 32-bit index math only (the flat 64-bit pixel
  // index this replaces cost one 64-bit division per 60 bytes moved)
  const int quads = hw >> 2;                     // hw % 4 == 0
  for (int t = blockIdx.y; t < T; t += gridDim.y) {
    const uint32_t* s = reinterpret_cast<const uint32_t*>(src + (long long)t * hw * 3);
    float* d = dst + (long long)t * 3 * hw;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += gridDim.x * blockDim.x) {
      const uint32_t w0 = __ldg(s + 3 * q), w1 = __ldg(s + 3 * q + 1), w2 = __ldg(s + 3 * q + 2);
      uint8_t b[12];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        b[i] = (w0 >> (8 * i)) & 0xff;
        b[4 + i] = (w1 >> (8 * i)) & 0xff;
        b[8 + i] = (w2 >> (8 * i)) & 0xff;
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float4 o;
        o.x = lut[c][b[0 + c]];
        o.y = lut[c][b[3 + c]];
        o.z = lut[c][b[6 + c]];
        o.w = lut[c][b[9 + c]];
        __stcs(reinterpret_cast<float4*>(d + c * hw + 4 * q), o);
      }
    }
  }
}

// Patchify (im2col for the 14x14/14 patch-embed GEMM).  Persistent CTAs walk (frame, band of P image rows) items with a
// two-deep cp.async pipeline: the band (u8: P*S*3 contiguous bytes; f32: 3 planes x P*S floats) is fetched with 16-byte
// asynchronous copies while the previous band is re-laid out in shared memory as the S/P im2col rows it produces
// (column = (c*P + ky)*P + kx, zero padded to ld; u8 goes through a 3x256 bf16 table of the exact reference
// arithmetic) and streamed out with 16-byte coalesced stores (the rows of a band are contiguous in the output).
__device__ __forceinline__ void cp_async16_g2s(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}

template <bool U8, int PC, bool FAST>   // PC: compile-time patch size (unrolled loops) or 0; FAST: verified division-free normalise
__global__ void __launch_bounds__(256) patchify_kernel(const void* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                       int n_items, int S, int P_rt, int ld, Norm3 nm, NormFast nf) {
  extern __shared__ __align__(16) uint8_t psm[];
  const int P = PC > 0 ? PC : P_rt;
  const int g = S / P;
  const int band = P * S;                                   // pixels per channel in a band
  const int raw_bytes = U8 ? ((band * 3 + 15) & ~15) : 3 * band * 4;
  __nv_bfloat16* lut = reinterpret_cast<__nv_bfloat16*>(psm);                 // [3][256] (u8 only)
  uint8_t* raw0 = psm + 1536;
  __nv_bfloat16* outs = reinterpret_cast<__nv_bfloat16*>(psm + 1536 + 2 * raw_bytes);
  if (U8 && !FAST)
    for (int i = threadIdx.x; i < 768; i += blockDim.x)
      lut[i] = __float2bfloat16(norm_px((uint8_t)(i & 255), nm.mean[i >> 8], nm.stdv[i >> 8]));
  const uint16_t* lut16 = reinterpret_cast<const uint16_t*>(lut);
  const int kdim = 3 * P * P;

  auto fetch = [&](int item, int buf) {
    const int f = item / g, py = item - f * g;
    uint8_t* rb = raw0 + buf * raw_bytes;
    if (U8) {
      const uint8_t* sp = reinterpret_cast<const uint8_t*>(src) + ((long long)f * S + (long long)py * P) * S * 3;
      for (int i = threadIdx.x; i < band * 3 / 16; i += blockDim.x) cp_async16_g2s(rb + i * 16, sp + i * 16);
    } else {
      for (int c = 0; c < 3; ++c) {
        const uint8_t* sp = reinterpret_cast<const uint8_t*>(reinterpret_cast<const float*>(src) +
                                                             (((long long)f * 3 + c) * S + (long long)py * P) * S);
        for (int i = threadIdx.x; i < band / 4; i += blockDim.x) cp_async16_g2s(rb + c * band * 4 + i * 16, sp + i * 16);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int item = blockIdx.x;
  if (item < n_items) fetch(item, 0);
  int it = 0;
  for (; item < n_items; item += gridDim.x, ++it) {
    const int nxt = item + gridDim.x;
    if (nxt < n_items) {
      fetch(nxt, (it + 1) & 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const uint8_t* raw = raw0 + (it & 1) * raw_bytes;
    const float* rawf = reinterpret_cast<const float*>(raw);
    // re-layout: one task = the P contiguous kx of one (channel, ky, patch)
    const int ntask = 3 * P * g;
    for (int t = threadIdx.x; t < ntask; t += blockDim.x) {
      const int c = t / (P * g);
      const int r = t - c * (P * g);
      const int ky = r / g, px = r - ky * g;
      const int pix0 = ky * S + px * P;
      uint32_t* d32 = reinterpret_cast<uint32_t*>(outs + px * ld + (c * P + ky) * P);
      if (PC > 0) {
        uint32_t w[PC > 0 ? PC / 2 : 1];
        if (U8) {
          uint8_t bytes[PC > 0 ? PC : 1];
#pragma unroll
          for (int kx = 0; kx < PC; ++kx) bytes[kx] = raw[(pix0 + kx) * 3 + c];
          if (FAST) {
            const float mean = nf.mean[c], stdv = nf.stdv[c], rstd = nf.rstd[c];
#pragma unroll
            for (int kx = 0; kx < PC; kx += 2)
              // byte -> float through the 2^23 mantissa trick (ALU/FMA pipes; I2F would go to the quarter-rate XU pipe)
              w[kx >> 1] = pack_bf16(
                  norm_px_fast(__uint_as_float(0x4B000000u | bytes[kx]) - 8388608.0f, nf.r255, mean, stdv, rstd),
                  norm_px_fast(__uint_as_float(0x4B000000u | bytes[kx + 1]) - 8388608.0f, nf.r255, mean, stdv, rstd));
          } else {
#pragma unroll
            for (int kx = 0; kx < PC; kx += 2)
              w[kx >> 1] = (uint32_t)lut16[c * 256 + bytes[kx]] | ((uint32_t)lut16[c * 256 + bytes[kx + 1]] << 16);
          }
        } else {
#pragma unroll
          for (int kx = 0; kx < PC; kx += 2) {
            const float2 v = *reinterpret_cast<const float2*>(rawf + c * band + pix0 + kx);
            w[kx >> 1] = pack_bf16(v.x, v.y);
          }
        }
#pragma unroll
        for (int q = 0; q < PC / 2; ++q) d32[q] = w[q];
      } else {
        for (int kx = 0; kx < P; kx += 2) {
          if (U8) {
            const uint32_t lo = lut16[c * 256 + raw[(pix0 + kx) * 3 + c]];
            const uint32_t hi = lut16[c * 256 + raw[(pix0 + kx + 1) * 3 + c]];
            d32[kx >> 1] = lo | (hi << 16);
          } else {
            const float2 v = *reinterpret_cast<const float2*>(rawf + c * band + pix0 + kx);
            d32[kx >> 1] = pack_bf16(v.x, v.y);
          }
        }
      }
    }
    for (int i = threadIdx.x; i < g * (ld - kdim); i += blockDim.x) {
      const int px = i / (ld - kdim);
      outs[px * ld + kdim + (i - px * (ld - kdim))] = __float2bfloat16(0.f);
    }
    __syncthreads();
    const int f = item / g, py = item - f * g;
    uint4* out4 = reinterpret_cast<uint4*>(dst + ((long long)f * g + py) * g * ld);
    const uint4* in4 = reinterpret_cast<const uint4*>(outs);
    for (int i = threadIdx.x; i < g * ld / 8; i += blockDim.x) __stcs(out4 + i, in4[i]);
    __syncthreads();
  }
}

// u8 patchify, P = 14 fast path ("pair" kernel).  One task = one image row ky of a PAIR of horizontally adjacent patches:
// 2*P*3 = 84 contiguous source bytes = 21 aligned 32-bit words, read with conflict-poor LDS.32 (lane stride 21 words
// across pairs), every byte pulled out with one PRMT straight into the mantissa of 2^23 (byte -> float without the
// quarter-rate I2F), normalised with ONE fma (y = fma(u, 1/(255 std), -mean/std)) and packed to bf16.  The host checks
// exhaustively (3 x 256 inputs) that bf16(fma) equals bf16 of the reference's ((u/255) - mean)/std chain before this
// kernel may be used, so the output stays bit-identical.  Lanes run along ky (7-word stride in the staged im2col rows:
// conflict-free stores); the staged band is streamed out with 16-byte coalesced stores as before.
struct NormFma { float scale[3]; float bias[3]; };

template <int P>
__global__ void __launch_bounds__(128) patchify_u8_pair_kernel(const uint8_t* __restrict__ src,
                                                               __nv_bfloat16* __restrict__ dst, int n_items, int S,
                                                               int ld, NormFma nf) {
  extern __shared__ __align__(16) uint8_t psm[];
  constexpr int NW = 2 * P * 3 / 4;                         // words per task
  const int g = S / P;
  const int band_bytes = P * S * 3;
  const int raw_bytes = (band_bytes + 15) & ~15;
  uint8_t* raw0 = psm;
  __nv_bfloat16* outs = reinterpret_cast<__nv_bfloat16*>(psm + 2 * raw_bytes);
  constexpr int kdim = 3 * P * P;
  for (int i = threadIdx.x; i < g * (ld - kdim); i += blockDim.x) {       // zero the K padding once: never overwritten
    const int px = i / (ld - kdim);
    outs[px * ld + kdim + (i - px * (ld - kdim))] = __float2bfloat16(0.f);
  }

  auto fetch = [&](int item, int buf) {
    const int f = item / g, py = item - f * g;
    uint8_t* rb = raw0 + buf * raw_bytes;
    const uint8_t* sp = src + ((long long)f * S + (long long)py * P) * S * 3;
    for (int i = threadIdx.x; i < band_bytes / 16; i += blockDim.x) cp_async16_g2s(rb + i * 16, sp + i * 16);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int item = blockIdx.x;
  if (item < n_items) fetch(item, 0);
  int it = 0;
  const int ntask = P * (g / 2);
  for (; item < n_items; item += gridDim.x, ++it) {
    const int nxt = item + gridDim.x;
    if (nxt < n_items) {
      fetch(nxt, (it + 1) & 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const uint32_t* raw32 = reinterpret_cast<const uint32_t*>(raw0 + (it & 1) * raw_bytes);
    for (int t = threadIdx.x; t < ntask; t += blockDim.x) {
      const int pair = t / P, ky = t - pair * P;
      const uint32_t* rp = raw32 + (ky * S * 3 + pair * 2 * P * 3) / 4;
      uint32_t w[NW];
#pragma unroll
      for (int i = 0; i < NW; ++i) w[i] = rp[i];
#pragma unroll
      for (int pp = 0; pp < 2; ++pp) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          uint32_t* d32 = reinterpret_cast<uint32_t*>(outs + (2 * pair + pp) * ld + (c * P + ky) * P);
          const float sc = nf.scale[c], bi = nf.bias[c];
#pragma unroll
          for (int kx = 0; kx < P; kx += 2) {
            const int i0 = (pp * P + kx) * 3 + c, i1 = i0 + 3;
            const float u0 = __uint_as_float(__byte_perm(w[i0 >> 2], 0x4B000000u, 0x7650 | (i0 & 3))) - 8388608.0f;
            const float u1 = __uint_as_float(__byte_perm(w[i1 >> 2], 0x4B000000u, 0x7650 | (i1 & 3))) - 8388608.0f;
            d32[kx >> 1] = pack_bf16(__fmaf_rn(u0, sc, bi), __fmaf_rn(u1, sc, bi));
          }
        }
      }
    }
    __syncthreads();
    const int f = item / g, py = item - f * g;
    uint4* out4 = reinterpret_cast<uint4*>(dst + ((long long)f * g + py) * g * ld);
    const uint4* in4 = reinterpret_cast<const uint4*>(outs);
    for (int i = threadIdx.x; i < g * ld / 8; i += blockDim.x) __stcs(out4 + i, in4[i]);
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------
// N3: ResizeVideo -> ToUint8 on decoded frames.  u8 [T,H0,W0,3] -> u8 [T,S,S,3], bilinear with
// align_corners=False, no antialias, non-aspect-preserving (src/gadgets/functional_video.py:33-40), then truncation to
// uint8 (transforms.py:213-220).  Truncation makes flat regions sensitive to the last fp32 bit (200*w0 + 200*w1 may
// round to 199.99998), so the kernel follows ATen's CPU arithmetic operation for operation -- source index through one
// fma, horizontal pass first, each pass fma(x0, w0, round(x1*w1)) -- all with explicit _rn intrinsics.
// One thread = 4 consecutive output pixels of one row (12 output bytes = three aligned words).
// ---------------------------------------------------------------------------------------------------------
struct BilinearTap { int i0, i1; float w0, w1; };

__device__ __forceinline__ BilinearTap bilinear_tap(int o, int in_size, int out_size, float scale) {
  BilinearTap t;
  if (in_size == out_size) { t.i0 = t.i1 = o; t.w0 = 1.f; t.w1 = 0.f; return t; }
  float real = __fmaf_rn(scale, __fadd_rn((float)o, 0.5f), -0.5f);
  real = real < 0.f ? 0.f : real;
  int i0 = (int)real;                                 // real >= 0: truncation == floor
  i0 = i0 < in_size - 1 ? i0 : in_size - 1;
  float l1 = __fsub_rn(real, (float)i0);
  l1 = fminf(fmaxf(l1, 0.f), 1.f);
  t.i0 = i0;
  t.i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  t.w1 = l1;
  t.w0 = __fsub_rn(1.f, l1);
  return t;
}

__global__ void __launch_bounds__(256) resize_bilinear_u8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                                 int H0, int W0, int S, float scale_h, float scale_w) {
  const int xq = blockIdx.x * blockDim.x + threadIdx.x;       // quad of output columns
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int t = blockIdx.z;
  if (xq * 4 >= S || y >= S) return;
  const BilinearTap ty = bilinear_tap(y, H0, S, scale_h);
  const uint8_t* r0 = src + ((long long)t * H0 + ty.i0) * W0 * 3;
  const uint8_t* r1 = src + ((long long)t * H0 + ty.i1) * W0 * 3;
  uint8_t o[12];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const BilinearTap tx = bilinear_tap(xq * 4 + q, W0, S, scale_w);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      // byte -> float through the 2^23 mantissa trick (I2F would run on the quarter-rate XU pipe, which bound this kernel)
      const float s00 = __uint_as_float(0x4B000000u | __ldg(r0 + tx.i0 * 3 + c)) - 8388608.0f;
      const float s01 = __uint_as_float(0x4B000000u | __ldg(r0 + tx.i1 * 3 + c)) - 8388608.0f;
      const float s10 = __uint_as_float(0x4B000000u | __ldg(r1 + tx.i0 * 3 + c)) - 8388608.0f;
      const float s11 = __uint_as_float(0x4B000000u | __ldg(r1 + tx.i1 * 3 + c)) - 8388608.0f;
      const float top = __fmaf_rn(s00, tx.w0, __fmul_rn(s01, tx.w1));
      const float bot = __fmaf_rn(s10, tx.w0, __fmul_rn(s11, tx.w1));
      const float v = __fmaf_rn(top, ty.w0, __fmul_rn(bot, ty.w1));
      o[q * 3 + c] = (uint8_t)(int)v;                      // v in [0,255]: truncation toward zero like .to(torch.uint8)
    }
  }
  uint32_t* d = reinterpret_cast<uint32_t*>(dst + (((long long)t * S + y) * S + xq * 4) * 3);
#pragma unroll
  for (int k = 0; k < 3; ++k)
    d[k] = (uint32_t)o[4 * k] | ((uint32_t)o[4 * k + 1] << 8) | ((uint32_t)o[4 * k + 2] << 16) | ((uint32_t)o[4 * k + 3] << 24);
}

// Fast variant (base pointer and total size multiples of 4).  The generic kernel above spends 38 instructions per output
// byte, a third of them 64-bit address arithmetic for its 48 single-byte loads per thread, and is issue bound at half the
// HBM roofline.  Here:
//   * the two horizontal taps of a pixel are adjacent pixels = 6 contiguous bytes, fetched as three aligned 32-bit words
//     (one address computation, immediate offsets) and aligned with two funnel shifts; at the right edge (i1 == i0) the
//     first pixel's bytes are replicated into the second pixel's position with two predicated PRMTs;
//   * byte -> float is ONE PRMT that drops the byte into the mantissa of 2^23 (0x4B0000bb) plus the exact subtraction;
//   * float -> byte is an add of 2^23 with round-toward-zero (v >= 0: truncation == floor) and PRMT packing, instead of
//     F2I on the quarter-rate XU pipe;
//   * taps are computed once per CTA (column tap in registers, row taps in shared memory) for all the rows it walks.
// Same arithmetic, operation for operation, as the generic kernel (bit-exact with ATen's CPU path).
__device__ __forceinline__ void resize_load8(const uint32_t* __restrict__ fa, const uint8_t* __restrict__ fbytes, int o,
                                             int limit, uint32_t& lo, uint32_t& hi) {
  if (o + 12 <= limit) {                       // three whole words are inside the buffer (always, except at its very end)
    const uint32_t* w = fa + (o >> 2);
    const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
    const uint32_t sh = (static_cast<uint32_t>(o) & 3u) << 3;
    lo = __funnelshift_r(w0, w1, sh);
    hi = __funnelshift_r(w1, w2, sh);
  } else {                                     // last pixels of the last frame: byte loads, nothing read past the end
    uint32_t b[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) b[k] = (o + k < limit) ? __ldg(fbytes + o + k) : 0u;
    lo = b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24);
    hi = b[4] | (b[5] << 8);
  }
}

struct XTap { int xoff; uint32_t edge; float w0, w1; };

// One thread = 4 consecutive output pixels of a row (12 output bytes = three aligned words); a CTA of 64 x 4 threads walks
// `rows_per_cta` output rows of its 256 columns.  (A one-lane-per-pixel mapping, whose word loads touch 3 cache lines per
// warp instead of 9, measured slower: 3.4 vs 3.9 TB/s -- the kernel is bound by instruction issue, not by L1 wavefronts.)
struct YTap { int off0, off1; float w0, w1; };
constexpr int RESIZE_MAX_ROWS = 32;

__global__ void __launch_bounds__(256) resize_bilinear_u8_fast_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                                      int T, int H0, int W0, int S, float scale_h,
                                                                      float scale_w, int rows_per_cta) {
  __shared__ XTap xt[4][64];                   // [pixel of the quad][thread column]: conflict-free 16-byte reads
  __shared__ YTap yt[RESIZE_MAX_ROWS];         // row taps of the CTA's output rows (byte offsets inside the frame)
  const int tid = threadIdx.y * 64 + threadIdx.x;
  const int t = blockIdx.z;
  const long long frame_bytes = (long long)H0 * W0 * 3;
  const uint8_t* fb = src + (long long)t * frame_bytes;
  const int fm = static_cast<int>(reinterpret_cast<uintptr_t>(fb) & 3u);
  const int y_base = blockIdx.y * rows_per_cta;
  {
    const int px = blockIdx.x * 256 + tid;     // tid = 4 * column + q
    if (px < S) {
      const BilinearTap tp = bilinear_tap(px, W0, S, scale_w);
      XTap e;
      e.xoff = tp.i0 * 3;
      e.edge = tp.i1 == tp.i0 ? 1u : 0u;
      e.w0 = tp.w0;
      e.w1 = tp.w1;
      xt[tid & 3][tid >> 2] = e;
    }
    if (tid < rows_per_cta && y_base + tid < S) {
      const BilinearTap tp = bilinear_tap(y_base + tid, H0, S, scale_h);
      YTap e;
      e.off0 = fm + tp.i0 * W0 * 3;
      e.off1 = fm + tp.i1 * W0 * 3;
      e.w0 = tp.w0;
      e.w1 = tp.w1;
      yt[tid] = e;
    }
  }
  __syncthreads();
  const int xq = blockIdx.x * 64 + threadIdx.x;
  if (xq * 4 >= S) return;
  const uint32_t* fa = reinterpret_cast<const uint32_t*>(fb - fm);
  const uint8_t* fbytes = fb - fm;
  const int limit = (t == T - 1) ? static_cast<int>(frame_bytes) + fm : 0x7fffffff;
  constexpr uint32_t MAGIC = 0x4B000000u;      // 2^23: PRMT selector 0x754k puts byte k into its mantissa
  // the right-edge case (second tap == first tap) only exists in the last quad(s) of a row: one branch per thread
  const bool any_edge = (xt[0][threadIdx.x].edge | xt[1][threadIdx.x].edge | xt[2][threadIdx.x].edge | xt[3][threadIdx.x].edge) != 0u;
  const int y_end = min(S - y_base, rows_per_cta);
  for (int yr = threadIdx.y; yr < y_end; yr += 4) {
    const YTap ty = yt[yr];
    const int row0 = ty.off0, row1 = ty.off1;
    // all 24 word loads of the thread are issued before the first use (the taps stay in shared memory until needed)
    uint32_t lo[8], hi[8];
    const int last_off = max(row0, row1) + xt[3][threadIdx.x].xoff;
    if (last_off + 12 <= limit) {                // always, except for the last pixels of the last frame
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int xo = xt[q][threadIdx.x].xoff;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int o = (r == 0 ? row0 : row1) + xo;
          const uint32_t* w = fa + (o >> 2);
          const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
          const uint32_t sh = (static_cast<uint32_t>(o) & 3u) << 3;
          lo[2 * q + r] = __funnelshift_r(w0, w1, sh);
          hi[2 * q + r] = __funnelshift_r(w1, w2, sh);
        }
      }
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int r = 0; r < 2; ++r)
          resize_load8(fa, fbytes, (r == 0 ? row0 : row1) + xt[q][threadIdx.x].xoff, limit, lo[2 * q + r], hi[2 * q + r]);
    }
    if (any_edge) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (xt[q][threadIdx.x].edge) {           // second tap = the same pixel: bytes 3,4,5 := bytes 0,1,2
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            hi[2 * q + r] = __byte_perm(lo[2 * q + r], 0u, 0x4421);
            lo[2 * q + r] = __byte_perm(lo[2 * q + r], 0u, 0x0210);
          }
        }
    }
    uint32_t m[12];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const XTap x = xt[q][threadIdx.x];
      float top[3], bot[3];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const uint32_t l = lo[2 * q + r], h = hi[2 * q + r];
        // byte -> float: PRMT into the 2^23 mantissa + exact subtraction on the FMA pipe for four of the six samples, the
        // XU pipe's I2F.U8 (byte select built in, quarter rate, otherwise idle) for the other two
        const float a0 = __uint_as_float(__byte_perm(l, MAGIC, 0x7540)) - 8388608.0f;
        const float a1 = __uint_as_float(__byte_perm(l, MAGIC, 0x7541)) - 8388608.0f;
        const float a2 = static_cast<float>((l >> 16) & 0xffu);
        const float b0 = __uint_as_float(__byte_perm(l, MAGIC, 0x7543)) - 8388608.0f;
        const float b1 = __uint_as_float(__byte_perm(h, MAGIC, 0x7540)) - 8388608.0f;
        const float b2 = static_cast<float>((h >> 8) & 0xffu);
        float* dstv = r == 0 ? top : bot;
        dstv[0] = __fmaf_rn(a0, x.w0, __fmul_rn(b0, x.w1));
        dstv[1] = __fmaf_rn(a1, x.w0, __fmul_rn(b1, x.w1));
        dstv[2] = __fmaf_rn(a2, x.w0, __fmul_rn(b2, x.w1));
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float v = __fmaf_rn(top[c], ty.w0, __fmul_rn(bot[c], ty.w1));
        m[q * 3 + c] = __float_as_uint(__fadd_rz(v, 8388608.0f));     // low mantissa byte = trunc(v), v in [0, 255]
      }
    }
    uint32_t* d = reinterpret_cast<uint32_t*>(dst + (((long long)t * S + y_base + yr) * S + xq * 4) * 3);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const uint32_t p01 = __byte_perm(m[4 * k], m[4 * k + 1], 0x0040);
      const uint32_t p23 = __byte_perm(m[4 * k + 2], m[4 * k + 3], 0x0040);
      d[k] = __byte_perm(p01, p23, 0x5410);
    }
  }
}

// Gather: dst[(b*k + i)] = src[b, idx[b,i]] as 16-byte words; blockIdx.x = output frame (no 65535 limit), the blockIdx.y
// CTAs of a frame grid-stride over its words.
__global__ void __launch_bounds__(256) gather_frames_kernel(const uint4* __restrict__ src, const long long* __restrict__ idx,
                                                            uint4* __restrict__ dst, int N, int k, long long words) {
  const int of = blockIdx.x;                    // output frame b*k + i
  const int b = of / k;
  long long j = idx[of];
  j = j < 0 ? 0 : (j >= N ? N - 1 : j);         // host validated; clamp keeps a bad index from faulting
  const uint4* s = src + ((long long)b * N + j) * words;
  uint4* d = dst + (long long)of * words;
  const long long stride = (long long)gridDim.y * blockDim.x;
  long long w = (long long)blockIdx.y * blockDim.x + threadIdx.x;
  for (; w + 3 * stride < words; w += 4 * stride) {       // 4 independent 16-B loads in flight per thread
    const uint4 a0 = __ldg(s + w), a1 = __ldg(s + w + stride), a2 = __ldg(s + w + 2 * stride), a3 = __ldg(s + w + 3 * stride);
    d[w] = a0; d[w + stride] = a1; d[w + 2 * stride] = a2; d[w + 3 * stride] = a3;
  }
  for (; w < words; w += stride) d[w] = __ldg(s + w);
}

__global__ void fill_cls_kernel(float* __restrict__ hidden, const float* __restrict__ cls, const float* __restrict__ pos0,
                                int F, long long frame_stride, int D) {
  const long long n = (long long)F * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long f = i / D;
    const int d = static_cast<int>(i - f * D);
    hidden[f * frame_stride + d] = cls[d] + pos0[d];
  }
}

static int grid_for(long long work_items, int threads, int per_sm) {
  long long blocks = (work_items + threads - 1) / threads;
  const long long cap = (long long)sm_count() * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace tgb

using namespace tgb;

// round-to-nearest-even fp32 -> bf16 bit pattern on the host (finite inputs)
static uint16_t bf16_bits_rn(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

static std::atomic<int> g_patchify_variant{0};   // 0 = choose, 1 = force the table / residual-corrected kernels (tests)

static Norm3 make_norm(const float* mean, const float* stdv) {
  Norm3 n;
  for (int i = 0; i < 3; ++i) { n.mean[i] = mean[i]; n.stdv[i] = stdv[i]; }
  return n;
}

extern "C" int tgb_normalize_frames_u8(const uint8_t* frames, float* out, int T, int H, int W, const float* mean,
                                       const float* stdv, tgb_stream_t stream) {
  int st = check_device();
  if (st != TGB_OK) return st;
  TGB_CHECK_ARG(frames && out && mean && stdv, "normalize: null pointer");
  TGB_CHECK_ARG(T >= 0 && H > 0 && W > 0 && (H * W) % 4 == 0, "normalize: H*W must be a multiple of 4");
  TGB_CHECK_ARG(((uintptr_t)frames % 4) == 0 && ((uintptr_t)out % 16) == 0, "normalize: misaligned buffers");
  if (T == 0) return TGB_OK;
  TGB_CHECK_ARG((long long)H * W * 3 < 0x7fffffffLL, "normalize: frame too large");
  const int quads = H * W / 4;
  int gx = (quads + 1023) / 1024;                              // ~4 quads per thread and frame
  if (gx > 64) gx = 64;
  int gy = sm_count() * 8 / gx;
  if (gy < 1) gy = 1;
  if (gy > T) gy = T;
  if (gy > 65535) gy = 65535;
  normalize_u8_kernel<<<dim3(gx, gy), 256, 0, (cudaStream_t)stream>>>(frames, out, T, H * W, make_norm(mean, stdv));
  TGB_CUDA(cudaGetLastError());
  return TGB_OK;
}

static int patchify_common(const void* src, void* out, int F, int S, int P, int ld, bool u8, Norm3 nm, tgb_stream_t stream) {
  int st = check_device();
  if (st != TGB_OK) return st;
  TGB_CHECK_ARG(src && out, "patchify: null pointer");
  TGB_CHECK_ARG(F >= 0 && S > 0 && P > 0 && S % P == 0, "patchify: image %d not divisible by patch %d", S, P);
  TGB_CHECK_ARG(ld % 8 == 0 && ld >= 3 * P * P && ((uintptr_t)out % 16) == 0, "patchify: ld must be a multiple of 8 >= 3*P*P");
  if (F == 0) return TGB_OK;
  TGB_CHECK_ARG(((uintptr_t)src % 16) == 0 && (u8 ? ((long long)P * S * 3) % 16 == 0 : ((long long)P * S) % 4 == 0),
                "patchify: source band must be 16-byte aligned (P*S*3 %% 16 == 0 for u8 frames)");
  TGB_CHECK_ARG(P % 2 == 0 && (!u8 || ((long long)P * S * 3) % 16 == 0), "patchify: patch size must be even (and P*S*3 %% 16 == 0 for u8)");
  const size_t raw_bytes = u8 ? (((size_t)P * S * 3 + 15) & ~(size_t)15) : (size_t)3 * P * S * 4;
  const size_t smem = 1536 + 2 * raw_bytes + (size_t)(S / P) * ld * 2;
  TGB_CHECK_ARG(smem <= 200 * 1024, "patchify: band of %d x %d pixels does not fit in shared memory", P, S);
  const int items = F * (S / P);
  const int per_sm = smem > 100 * 1024 ? 1 : (smem > 56 * 1024 ? 2 : 4);
  const int grid = items < sm_count() * per_sm ? items : sm_count() * per_sm;
  cudaStream_t st_ = (cudaStream_t)stream;
#define TGB_PATCHIFY(U, PCV, FST)                                                                                      \
  do {                                                                                                                \
    TGB_ONCE_PER_DEVICE(TGB_CUDA(cudaFuncSetAttribute(patchify_kernel<U, PCV, FST>,                                  \
                                                      cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)));     \
    patchify_kernel<U, PCV, FST><<<grid, 256, smem, st_>>>(src, (__nv_bfloat16*)out, items, S, P, ld, nm, nf);        \
  } while (0)
  NormFast nf = {};
  bool fast = false;
  if (u8) {                       // exhaustive host check of the division-free path against IEEE division
    nf.r255 = 1.0f / 255.0f;
    fast = true;
    for (int c = 0; c < 3; ++c) {
      nf.mean[c] = nm.mean[c]; nf.stdv[c] = nm.stdv[c]; nf.rstd[c] = 1.0f / nm.stdv[c];
      for (int v = 0; v < 256 && fast; ++v) {
        volatile float x = (float)v / 255.0f;
        volatile float y = x - nm.mean[c];
        volatile float z = y / nm.stdv[c];
        if (norm_px_fast((float)v, nf.r255, nf.mean[c], nf.stdv[c], nf.rstd[c]) != z) fast = false;
      }
    }
  }
  // single-fma normalise: allowed only if bf16(fma(u, scale, bias)) == bf16(reference chain) for all 3 x 256 inputs
  NormFma nfma = {};
  bool fma_ok = u8 && P == 14 && (S / P) % 2 == 0 && S % 4 == 0 && g_patchify_variant != 1;
  for (int c = 0; c < 3 && fma_ok; ++c) {
    nfma.scale[c] = (float)(1.0 / (255.0 * (double)nm.stdv[c]));
    nfma.bias[c] = (float)(-(double)nm.mean[c] / (double)nm.stdv[c]);
    for (int v = 0; v < 256 && fma_ok; ++v) {
      volatile float x = (float)v / 255.0f;
      volatile float y = x - nm.mean[c];
      volatile float z = y / nm.stdv[c];
      volatile float fz = fmaf((float)v, nfma.scale[c], nfma.bias[c]);
      if (bf16_bits_rn(z) != bf16_bits_rn(fz)) fma_ok = false;
    }
  }
  if (fma_ok) {
    const size_t rb = ((size_t)P * S * 3 + 15) & ~(size_t)15;
    const size_t sm = 2 * rb + (size_t)(S / P) * ld * 2;
    TGB_ONCE_PER_DEVICE(TGB_CUDA(cudaFuncSetAttribute(patchify_u8_pair_kernel<14>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                      200 * 1024)));
    int per = (int)((227 * 1024) / (sm + 1024));
    if (per > 8) per = 8;
    if (per < 1) per = 1;
    const int gr = items < sm_count() * per ? items : sm_count() * per;
    patchify_u8_pair_kernel<14><<<gr, 128, sm, st_>>>((const uint8_t*)src, (__nv_bfloat16*)out, items, S, ld, nfma);
  } else if (u8) {
    if (P == 14 && fast) TGB_PATCHIFY(true, 14, true);
    else if (P == 14) TGB_PATCHIFY(true, 14, false);
    else TGB_PATCHIFY(true, 0, false);
  } else {
    if (P == 14) TGB_PATCHIFY(false, 14, false); else TGB_PATCHIFY(false, 0, false);
  }
#undef TGB_PATCHIFY
  TGB_CUDA(cudaGetLastError());
  return TGB_OK;
}

extern "C" int tgb_patchify_f32(const float* pixels, void* out, int F, int S, int P, int ld, tgb_stream_t stream) {
  Norm3 nm = {};
  return patchify_common(pixels, out, F, S, P, ld, false, nm, stream);
}

extern "C" int tgb_patchify_u8(const uint8_t* frames, void* out, int F, int S, int P, int ld, const float* mean,
                               const float* stdv, tgb_stream_t stream) {
  TGB_CHECK_ARG(mean && stdv, "patchify_u8: null mean/std");
  return patchify_common(frames, out, F, S, P, ld, true, make_norm(mean, stdv), stream);
}

extern "C" int tgb_gather_frames(const void* src, const int64_t* indices, void* dst, int B, int N, int k,
                                 size_t frame_bytes, tgb_stream_t stream) {
  int st = check_device();
  if (st != TGB_OK) return st;
  TGB_CHECK_ARG(src && indices && dst, "gather: null pointer");
  TGB_CHECK_ARG(B >= 0 && N > 0 && k > 0, "gather: bad shape B=%d N=%d k=%d", B, N, k);
  TGB_CHECK_ARG(frame_bytes % 16 == 0 && ((uintptr_t)src % 16) == 0 && ((uintptr_t)dst % 16) == 0,
                "gather: frame_bytes and pointers must be multiples of 16");
  TGB_CHECK_ARG((long long)B * k <= 2147483647LL, "gather: B*k = %lld output frames", (long long)B * k);
  if (B == 0) return TGB_OK;
  const long long words = (long long)(frame_bytes / 16);
  // enough CTAs per frame to keep >= 4 x SMs CTAs in flight, each thread moving >= 4 words
  int per_frame = (4 * sm_count() + B * k - 1) / (B * k);
  const long long max_useful = (words + 256 * 4 - 1) / (256 * 4);
  if (per_frame > max_useful) per_frame = (int)max_useful;
  if (per_frame < 1) per_frame = 1;
  dim3 grid(B * k, per_frame);
  gather_frames_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const uint4*)src, (const long long*)indices, (uint4*)dst, N, k, words);
  TGB_CUDA(cudaGetLastError());
  return TGB_OK;
}

extern "C" int tgb_fill_cls_rows(float* hidden, const float* cls, const float* pos0, int F, int tokens, int D,
                                 tgb_stream_t stream) {
  int st = check_device();
  if (st != TGB_OK) return st;
  TGB_CHECK_ARG(hidden && cls && pos0 && F >= 0 && tokens > 0 && D > 0, "fill_cls: bad arguments");
  if (F == 0) return TGB_OK;
  fill_cls_kernel<<<grid_for((long long)F * D, 256, 4), 256, 0, (cudaStream_t)stream>>>(hidden, cls, pos0, F, (long long)tokens * D, D);
  TGB_CUDA(cudaGetLastError());
  return TGB_OK;
}

extern "C" int tgb_patchify_set_variant(int variant) {
  TGB_CHECK_ARG(variant == 0 || variant == 1, "patchify_set_variant: 0 (choose) or 1 (generic kernel)");
  g_patchify_variant = variant;
  return TGB_OK;
}

static std::atomic<int> g_resize_variant{0};   // 0 = choose (fast kernel when the input is 4-byte aligned), 1 = generic byte-load kernel

extern "C" int tgb_resize_set_variant(int variant) {
  if (variant != 0 && variant != 1) {
    set_last_error("resize variant must be 0 (auto) or 1 (generic byte-load kernel)");
    return TGB_EINVAL;
  }
  g_resize_variant = variant;
  return TGB_OK;
}

extern "C" int tgb_resize_bilinear_u8(const uint8_t* frames_thwc, uint8_t* out_thwc, int T, int H0, int W0, int S,
                                      tgb_stream_t stream) {
  int st = check_device();
  if (st != TGB_OK) return st;
  TGB_CHECK_ARG(frames_thwc && out_thwc, "resize: null pointer");
  TGB_CHECK_ARG(T >= 0 && H0 > 0 && W0 > 0 && S > 0 && S % 4 == 0, "resize: bad shape T=%d H0=%d W0=%d S=%d (S %% 4 == 0)", T, H0, W0, S);
  TGB_CHECK_ARG(T <= 65535 && ((uintptr_t)out_thwc % 4) == 0, "resize: at most 65535 frames per call, 4-byte aligned output");
  if (T == 0) return TGB_OK;
  // ATen area_pixel_compute_scale (align_corners=False, no scale_factor): (float)input_size / output_size
  volatile float sh = (float)H0 / (float)S, sw = (float)W0 / (float)S;
  // (A row-staged variant -- source rows copied to shared memory with 16-byte loads, taps precomputed per CTA -- was
  //  measured slower, 2.2 vs 3.3 TB/s: the kernel is instruction-issue bound at ~42 instructions per output byte, not
  //  load bound; profiles/r1_hbm_kernels.jsonl.)
  dim3 block(64, 4);
  const long long frame_bytes = (long long)H0 * W0 * 3;
  if (g_resize_variant == 0 && ((uintptr_t)frames_thwc % 4) == 0 && frame_bytes + 16 < 0x7fffffffLL) {
    // fast kernel: CTAs walk `rows` output rows so that the per-CTA column taps are amortised; enough CTAs for every SM
    int rows = 32;
    while (rows > 4 && (long long)((S + rows - 1) / rows) * ((S / 4 + 63) / 64) * T < 4LL * sm_count()) rows >>= 1;
    dim3 grid((S / 4 + 63) / 64, (S + rows - 1) / rows, T);
    resize_bilinear_u8_fast_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(frames_thwc, out_thwc, T, H0, W0, S, sh, sw, rows);
  } else {
    dim3 grid((S / 4 + 63) / 64, (S + 3) / 4, T);
    resize_bilinear_u8_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(frames_thwc, out_thwc, H0, W0, S, sh, sw);
  }
  TGB_CUDA(cudaGetLastError());
  return TGB_OK;
}
