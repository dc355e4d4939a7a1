// HBM-bound frame kernels: normalise (A0), patchify / im2col for the patch-embed GEMM (A10), gather (A9),
// CLS-row fill.  All are pure streaming kernels: 16-byte coalesced accesses, grid-stride over a grid sized as a
// multiple of the SM count.
#include "common.cuh"

namespace tgb {

struct Norm3 { float mean[3]; float stdv[3]; };

// (u/255 - mean)/std with the reference's op order and IEEE fp32 rounding at every step.
__device__ __forceinline__ float norm_px(uint8_t u, float mean, float stdv) {
  return __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(u), 255.0f), mean), stdv);
}

// u8 [T,H,W,3] -> f32 [T,3,H,W].  One thread handles 4 consecutive pixels of one frame: reads 12 bytes (three
// 4-byte words, coalesced across the warp) and writes one float4 per channel plane.
__global__ void __launch_bounds__(256) normalize_u8_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst,
                                                           long long n_quads, int hw, Norm3 nm) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n_quads; q += stride) {
    const long long pix = q * 4;                 // global pixel index (t*hw + p), hw % 4 == 0
    const long long t = pix / hw;
    const int p = static_cast<int>(pix - t * hw);
    const uint32_t* s = reinterpret_cast<const uint32_t*>(src + pix * 3);
    const uint32_t w0 = __ldg(s), w1 = __ldg(s + 1), w2 = __ldg(s + 2);
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
      o.x = norm_px(b[0 + c], nm.mean[c], nm.stdv[c]);
      o.y = norm_px(b[3 + c], nm.mean[c], nm.stdv[c]);
      o.z = norm_px(b[6 + c], nm.mean[c], nm.stdv[c]);
      o.w = norm_px(b[9 + c], nm.mean[c], nm.stdv[c]);
      *reinterpret_cast<float4*>(dst + (t * 3 + c) * hw + p) = o;
    }
  }
}

// Patchify: one thread produces 8 consecutive bf16 of one im2col row (16-byte store).  Row = (frame, patch),
// column = (c*P + ky)*P + kx.  P is even so an 8-group never needs more than two source segments... it may cross
// a ky boundary, so elements are fetched individually through the read-only path (the source frame is re-read
// from L1/L2; HBM sees each byte once).
template <bool U8>
__global__ void __launch_bounds__(256) patchify_kernel(const void* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                       long long n_groups, int S, int P, int ld, Norm3 nm) {
  const int g = S / P;
  const int kdim = 3 * P * P;
  const int groups_per_row = ld / 8;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_groups; i += stride) {
    const long long row = i / groups_per_row;
    const int col0 = static_cast<int>(i - row * groups_per_row) * 8;
    const long long f = row / (g * g);
    const int pidx = static_cast<int>(row - f * (g * g));
    const int py = pidx / g, px = pidx - py * g;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = col0 + j;
      float x = 0.0f;
      if (col < kdim) {
        const int c = col / (P * P);
        const int r = col - c * P * P;
        const int ky = r / P, kx = r - ky * P;
        const int y = py * P + ky, xx = px * P + kx;
        if (U8) {
          const uint8_t u = __ldg(reinterpret_cast<const uint8_t*>(src) + ((f * S + y) * S + xx) * 3 + c);
          x = norm_px(u, nm.mean[c], nm.stdv[c]);
        } else {
          x = __ldg(reinterpret_cast<const float*>(src) + ((f * 3 + c) * S + y) * (long long)S + xx);
        }
      }
      v[j] = x;
    }
    uint4 o;
    o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]);
    o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(dst + row * ld + col0) = o;
  }
}

// Gather: dst[(b*k + i)] = src[b, idx[b,i]] as 16-byte words; blockIdx.y = output frame, grid-stride over words.
__global__ void __launch_bounds__(256) gather_frames_kernel(const uint4* __restrict__ src, const long long* __restrict__ idx,
                                                            uint4* __restrict__ dst, int N, int k, long long words) {
  const int of = blockIdx.y;                    // output frame b*k + i
  const int b = of / k;
  long long j = idx[of];
  j = j < 0 ? 0 : (j >= N ? N - 1 : j);         // host validated; clamp keeps a bad index from faulting
  const uint4* s = src + ((long long)b * N + j) * words;
  uint4* d = dst + (long long)of * words;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
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
  const long long quads = (long long)T * H * W / 4;
  normalize_u8_kernel<<<grid_for(quads, 256, 8), 256, 0, (cudaStream_t)stream>>>(frames, out, quads, H * W,
                                                                                make_norm(mean, stdv));
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
  const long long groups = (long long)F * (S / P) * (S / P) * (ld / 8);
  if (u8)
    patchify_kernel<true><<<grid_for(groups, 256, 8), 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)out, groups, S, P, ld, nm);
  else
    patchify_kernel<false><<<grid_for(groups, 256, 8), 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)out, groups, S, P, ld, nm);
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
  TGB_CHECK_ARG((long long)B * k <= 65535, "gather: B*k = %lld exceeds 65535 output frames per call", (long long)B * k);
  if (B == 0) return TGB_OK;
  const long long words = (long long)(frame_bytes / 16);
  // enough CTAs per frame to keep >= 4 x SMs CTAs in flight, each thread moving >= 4 words
  int per_frame = (4 * sm_count() + B * k - 1) / (B * k);
  const long long max_useful = (words + 256 * 4 - 1) / (256 * 4);
  if (per_frame > max_useful) per_frame = (int)max_useful;
  if (per_frame < 1) per_frame = 1;
  dim3 grid(per_frame, B * k);
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
