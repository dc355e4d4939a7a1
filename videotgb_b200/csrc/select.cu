// Span draw and span -> frame-index selection.  Integer / IEEE-exact work: this
// file must NOT be compiled with fast-math; every float op that decides an index uses an explicit _rn intrinsic.
#include "common.cuh"

namespace tgb {

// ---------------------------------------------------------------------------------------------------------
// A7: index[r] = first argmax_i softmax_i((logits[r,i] + gumbel[r,i]) / tau), one CTA per row.
// Mirrors the reference op order: add, divide by tau, softmax (max-subtract, exp, sum, divide), argmax.
// ---------------------------------------------------------------------------------------------------------
constexpr int GA_THREADS = 128;

// Rows [0, split) read `logits_a` and write `index_a`, rows [split, R) read `logits_b` / write `index_b` (the start and
// end logits of the reference's cat(), without the cat).  EXPO: `noise` holds the Exp(1) draws of
// torch.Tensor.exponential_ and the Gumbel noise is formed here as -log(e), the statement F.gumbel_softmax makes.
template <bool EXPO>
__global__ void __launch_bounds__(GA_THREADS) gumbel_argmax_kernel(const float* __restrict__ logits_a,
                                                                  const float* __restrict__ logits_b, int split,
                                                                  const float* __restrict__ noise, float tau, int L,
                                                                  long long* __restrict__ index_a,
                                                                  long long* __restrict__ index_b) {
  extern __shared__ float xs[];                 // [L]
  __shared__ float red[GA_THREADS / 32];
  __shared__ int redi[GA_THREADS / 32];
  const int r = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* lr = r < split ? logits_a + (long long)r * L : logits_b + (long long)(r - split) * L;
  const float* gr = noise + (long long)r * L;
  float m = -INFINITY;
  for (int i = tid; i < L; i += GA_THREADS) {
    const float g = EXPO ? -logf(gr[i]) : gr[i];
    const float x = __fdiv_rn(__fadd_rn(lr[i], g), tau);
    xs[i] = x;
    m = fmaxf(m, x);
  }
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int w = 1; w < GA_THREADS / 32; ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  float s = 0.f;
  for (int i = tid; i < L; i += GA_THREADS) {
    const float e = expf(__fsub_rn(xs[i], m));
    xs[i] = e;
    s += e;
  }
  s = warp_sum(s);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  s = red[0];
#pragma unroll
  for (int w = 1; w < GA_THREADS / 32; ++w) s += red[w];
  float best = -1.f;
  int besti = 0x7fffffff;
  for (int i = tid; i < L; i += GA_THREADS) {      // increasing i per thread -> strict '>' keeps the first
    const float y = __fdiv_rn(xs[i], s);
    if (y > best) { best = y; besti = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
  }
  __syncthreads();
  if (lane == 0) { red[warp] = best; redi[warp] = besti; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < GA_THREADS / 32; ++w)
      if (red[w] > best || (red[w] == best && redi[w] < besti)) { best = red[w]; besti = redi[w]; }
    if (r < split) index_a[r] = besti;
    else index_b[r - split] = besti;
  }
}

// ---------------------------------------------------------------------------------------------------------
// A8: one warp per video.  bitmap union of the mapped spans -> sorted candidate list -> doubling ->
// numpy-linspace midpoint subsampling.
// ---------------------------------------------------------------------------------------------------------
constexpr int SEL_MAX_FRAMES = 4096;

__device__ __forceinline__ long long map_index(long long s, int N, int L, int mapping, bool as_double) {
  if (mapping == TGB_MAP_SCALED) {
    const long long num = s * (long long)(N - 1);
    if (as_double) return (long long)__ddiv_rn((double)num, (double)(L - 1));
    return (long long)__fdiv_rn(__ll2float_rn(num), (float)(L - 1));
  }
  if (as_double) return (long long)__dmul_rn(__ddiv_rn((double)s, (double)L), (double)N);
  return (long long)__fmul_rn(__fdiv_rn(__ll2float_rn(s), (float)L), (float)N);
}

__global__ void __launch_bounds__(32) spans_to_indices_kernel(const long long* __restrict__ cand_start,
                                                              const long long* __restrict__ cand_end, int top_k,
                                                              const int* __restrict__ of_lengths, int B, int N,
                                                              int nframe, int mapping, int inclusive, int host_ints,
                                                              long long* __restrict__ indices) {
  __shared__ uint32_t bits[SEL_MAX_FRAMES / 32];
  __shared__ short cand[SEL_MAX_FRAMES];
  const int j = blockIdx.x, lane = threadIdx.x;
  // the reference divides by video_length - 1 (ZeroDivisionError below 2; the host mirror raises it for host-side lengths).
  // Device-side lengths cannot raise: a value below 2 is treated as 2, so the arithmetic below stays defined.
  const int L = of_lengths[j] < 2 ? 2 : of_lengths[j];
  const int nwords = (N + 31) / 32;
  for (int w = lane; w < nwords; w += 32) bits[w] = 0u;
  __syncwarp();
  for (int d = 0; d < top_k; ++d) {
    long long s = cand_start[(long long)d * B + j], e = cand_end[(long long)d * B + j];
    bool dbl = host_ints != 0;
    if (s >= L || e >= L || (s == 0 && e == 0)) { s = 0; e = L - 1; dbl = true; }
    long long a = map_index(s, N, L, mapping, dbl);
    long long b = map_index(e, N, L, mapping, dbl) + (inclusive ? 1 : 0);
    if (a < 0) a = 0;
    if (b > N) b = N;
    for (int w = lane; w < nwords; w += 32) {           // set bits [a, b) of word w
      const long long lo = (long long)w * 32, hi = lo + 32;
      const long long x0 = a > lo ? a : lo, x1 = b < hi ? b : hi;
      if (x1 > x0) {
        const uint32_t n = (uint32_t)(x1 - x0);
        const uint32_t mask = (n == 32 ? 0xffffffffu : ((1u << n) - 1u)) << (uint32_t)(x0 - lo);
        bits[w] |= mask;
      }
    }
    __syncwarp();
  }
  // count + fallback to all frames
  int cnt = 0;
  for (int w = lane; w < nwords; w += 32) cnt += __popc(bits[w]);
  cnt = (int)__reduce_add_sync(0xffffffffu, (unsigned)cnt);
  if (cnt == 0) {
    for (int w = lane; w < nwords; w += 32) {
      const int rem = N - w * 32;
      bits[w] = rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
    }
    cnt = N;
    __syncwarp();
  }
  // ordered compaction: 32 words per step, exclusive warp scan of the popcounts
  int base = 0;
  for (int w0 = 0; w0 < nwords; w0 += 32) {
    const int w = w0 + lane;
    uint32_t word = w < nwords ? bits[w] : 0u;
    const int c = __popc(word);
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    int pos = base + incl - c;
    while (word) {
      const int bit = __ffs(word) - 1;
      cand[pos++] = (short)(w * 32 + bit);
      word &= word - 1;
    }
    base += __shfl_sync(0xffffffffu, incl, 31);
  }
  __syncwarp();
  // doubling: [a,b] -> [a,a,b,b] until len >= nframe  ==> element i of the doubled list is cand[i >> d]
  int d = 0;
  long long len = cnt;
  while (len < nframe) { len <<= 1; ++d; }
  long long* out = indices + (long long)j * nframe;
  if (len > nframe) {
    const double step = __ddiv_rn((double)len, (double)nframe);     // numpy.linspace(0, len, nframe+1)
    for (int x = lane; x < nframe; x += 32) {
      const long long i0 = (long long)__dmul_rn((double)x, step);
      const long long i1 = (x + 1 == nframe) ? len : (long long)__dmul_rn((double)(x + 1), step);
      const long long mid = (i0 + i1 - 1) >> 1;                     // floor division, operands >= 0
      out[x] = cand[mid >> d];
    }
  } else {
    for (int x = lane; x < nframe; x += 32) out[x] = cand[x >> d];
  }
}


// ---------------------------------------------------------------------------------------------------------
// N2: self-refinement pseudo labels, src/models/LSTP_SF_blip2_module.py:195-222.
// Largest-rectangle monotone stack over the min-shifted, zero-padded per-frame scores, then the rescale to flow
// positions.  The reference works on `tensor.tolist()` values, i.e. Python float64: every float op below is an
// explicit double _rn op on the exactly-converted float32 inputs.  One warp per video; the warp finds the minimum,
// lane 0 runs the (inherently sequential) stack.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) pseudo_span_kernel(const float* __restrict__ scores,
                                                         const int* __restrict__ of_lengths, int N,
                                                         long long* __restrict__ start_targets,
                                                         long long* __restrict__ end_targets) {
  extern __shared__ double ps_sm[];                  // [N + 2] shifted scores, then [N + 2] int stack
  double* sc = ps_sm;
  int* stack = reinterpret_cast<int*>(ps_sm + (N + 2));
  const int j = blockIdx.x, lane = threadIdx.x;
  const float* row = scores + (long long)j * N;
  float lo = INFINITY;
  for (int i = lane; i < N; i += 32) lo = fminf(lo, row[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
  for (int i = lane; i < N; i += 32) sc[i + 1] = __dsub_rn((double)row[i], (double)lo);
  if (lane == 0) { sc[0] = 0.0; sc[N + 1] = 0.0; }
  __syncwarp();
  if (lane != 0) return;
  double best = 0.0;
  int st = 0, en = N - 1, top = 0;
  for (int i = 0; i < N + 2; ++i) {
    const double cur = sc[i];
    while (top > 0 && sc[stack[top - 1]] > cur) {
      const int tmp = stack[--top];
      // (a non-negative padded sequence never empties the stack here: sc[0] == 0 is never popped)
      const double area = __dmul_rn((double)(i - stack[top - 1] - 1), sc[tmp]);
      if (area > best) { best = area; st = stack[top - 1]; en = i - 2; }
    }
    stack[top++] = i;
  }
  const double den = (double)(N - 1), mul = (double)(of_lengths[j] - 1);
  start_targets[j] = (long long)__dmul_rn(__ddiv_rn((double)st, den), mul);
  end_targets[j] = (long long)__dmul_rn(__ddiv_rn((double)en, den), mul);
}

// ---------------------------------------------------------------------------------------------------------
// N2: span cross-entropy, LSTP_SF_blip2_module.py:238-253.  logits f32 [B, L, 2] (start | end interleaved, as the
// mrc head emits them); CrossEntropyLoss(ignore_index = L), targets clamped to [0, L]; loss = (CE_start + CE_end)/2
// with each CE the mean over its non-ignored rows (0/0 = NaN like torch).  One CTA (B warps round-robin).
// ---------------------------------------------------------------------------------------------------------
constexpr int CE_THREADS = 256;

__global__ void __launch_bounds__(CE_THREADS) span_ce_loss_kernel(const float* __restrict__ logits,
                                                                 const long long* __restrict__ start_targets,
                                                                 const long long* __restrict__ end_targets, int B,
                                                                 int L, float* __restrict__ loss) {
  __shared__ float part[2][CE_THREADS / 32];
  __shared__ int cnt[2][CE_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc[2] = {0.f, 0.f};
  int n[2] = {0, 0};
  for (int r = warp; r < 2 * B; r += CE_THREADS / 32) {      // row r: video r>>1, head r&1
    const int b = r >> 1, h = r & 1;
    long long t = (h ? end_targets : start_targets)[b];
    t = t < 0 ? 0 : (t > L ? L : t);
    if (t == L) continue;                                      // ignore_index (warp-uniform)
    const float* x = logits + (long long)b * L * 2 + h;
    float m = -INFINITY;
    for (int i = lane; i < L; i += 32) m = fmaxf(m, x[2 * i]);
    m = warp_max(m);
    float s = 0.f;
    for (int i = lane; i < L; i += 32) s += expf(x[2 * i] - m);
    s = warp_sum(s);
    acc[h] += (m + logf(s)) - x[2 * t];                        // -log_softmax(x)[t]
    n[h] += 1;
  }
  if (lane == 0) { part[0][warp] = acc[0]; part[1][warp] = acc[1]; cnt[0][warp] = n[0]; cnt[1][warp] = n[1]; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot[2] = {0.f, 0.f};
    int c[2] = {0, 0};
    for (int w = 0; w < CE_THREADS / 32; ++w) { tot[0] += part[0][w]; tot[1] += part[1][w]; c[0] += cnt[0][w]; c[1] += cnt[1][w]; }
    loss[0] = (tot[0] / (float)c[0] + tot[1] / (float)c[1]) * 0.5f;
  }
}

}  // namespace tgb

using namespace tgb;

extern "C" int tgb_gumbel_argmax(const float* logits, const float* gumbel, float tau, int R, int L, int64_t* index,
                                 tgb_stream_t stream) {
  int st = check_device();
  if (st != TGB_OK) return st;
  TGB_CHECK_ARG(logits && gumbel && index, "gumbel_argmax: null pointer");
  TGB_CHECK_ARG(R >= 0 && L > 0 && L <= 40960 && tau > 0.f, "gumbel_argmax: bad shape R=%d L=%d tau=%f", R, L, tau);
  if (R == 0) return TGB_OK;
  const size_t smem = (size_t)L * sizeof(float);
  if (smem > 48 * 1024) {
    TGB_ONCE_PER_DEVICE(TGB_CUDA(cudaFuncSetAttribute(gumbel_argmax_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)));
  }
  gumbel_argmax_kernel<false><<<R, GA_THREADS, smem, (cudaStream_t)stream>>>(logits, logits, R, gumbel, tau, L,
                                                                             (long long*)index, (long long*)index);
  TGB_CUDA(cudaGetLastError());
  return TGB_OK;
}

extern "C" int tgb_span_draw(const float* start_logits, const float* end_logits, const float* noise, int noise_is_exponential,
                             float tau, int B, int L, int64_t* cand_start, int64_t* cand_end, tgb_stream_t stream) {
  int st = check_device();
  if (st != TGB_OK) return st;
  TGB_CHECK_ARG(start_logits && end_logits && noise && cand_start && cand_end, "span_draw: null pointer");
  TGB_CHECK_ARG(B >= 0 && L > 0 && L <= 40960 && tau > 0.f, "span_draw: bad shape B=%d L=%d tau=%f", B, L, tau);
  if (B == 0) return TGB_OK;
  const size_t smem = (size_t)L * sizeof(float);
  if (smem > 48 * 1024) {
    TGB_ONCE_PER_DEVICE({
      TGB_CUDA(cudaFuncSetAttribute(gumbel_argmax_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      TGB_CUDA(cudaFuncSetAttribute(gumbel_argmax_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    });
  }
  if (noise_is_exponential)
    gumbel_argmax_kernel<true><<<2 * B, GA_THREADS, smem, (cudaStream_t)stream>>>(start_logits, end_logits, B, noise, tau, L,
                                                                                 (long long*)cand_start, (long long*)cand_end);
  else
    gumbel_argmax_kernel<false><<<2 * B, GA_THREADS, smem, (cudaStream_t)stream>>>(start_logits, end_logits, B, noise, tau, L,
                                                                                  (long long*)cand_start, (long long*)cand_end);
  TGB_CUDA(cudaGetLastError());
  return TGB_OK;
}

extern "C" int tgb_spans_to_indices(const int64_t* cand_start, const int64_t* cand_end, int top_k,
                                    const int32_t* of_lengths, int B, int num_frames, int nframe, int mapping,
                                    int inclusive, int host_ints, int64_t* indices, tgb_stream_t stream) {
  int st = check_device();
  if (st != TGB_OK) return st;
  TGB_CHECK_ARG(cand_start && cand_end && of_lengths && indices, "spans_to_indices: null pointer");
  TGB_CHECK_ARG(B >= 0 && top_k > 0 && nframe > 0, "spans_to_indices: bad shape B=%d top_k=%d nframe=%d", B, top_k, nframe);
  TGB_CHECK_ARG(num_frames > 0 && num_frames <= SEL_MAX_FRAMES, "spans_to_indices: num_frames=%d must be in [1,%d]", num_frames, SEL_MAX_FRAMES);
  TGB_CHECK_ARG(mapping == TGB_MAP_SCALED || mapping == TGB_MAP_RATIO, "spans_to_indices: unknown mapping %d", mapping);
  if (B == 0) return TGB_OK;
  spans_to_indices_kernel<<<B, 32, 0, (cudaStream_t)stream>>>((const long long*)cand_start, (const long long*)cand_end, top_k,
                                                              of_lengths, B, num_frames, nframe, mapping, inclusive, host_ints,
                                                              (long long*)indices);
  TGB_CUDA(cudaGetLastError());
  return TGB_OK;
}

extern "C" int tgb_pseudo_span(const float* scores, const int32_t* of_lengths, int B, int num_frames,
                               int64_t* start_targets, int64_t* end_targets, tgb_stream_t stream) {
  int st = check_device();
  if (st != TGB_OK) return st;
  TGB_CHECK_ARG(scores && of_lengths && start_targets && end_targets, "pseudo_span: null pointer");
  TGB_CHECK_ARG(B >= 0 && num_frames >= 2 && num_frames <= 3000, "pseudo_span: bad shape B=%d num_frames=%d (2..3000)", B, num_frames);
  if (B == 0) return TGB_OK;
  const size_t smem = (size_t)(num_frames + 2) * (sizeof(double) + sizeof(int));
  pseudo_span_kernel<<<B, 32, smem, (cudaStream_t)stream>>>(scores, of_lengths, num_frames, (long long*)start_targets,
                                                           (long long*)end_targets);
  TGB_CUDA(cudaGetLastError());
  return TGB_OK;
}

extern "C" int tgb_span_ce_loss(const float* logits, const int64_t* start_targets, const int64_t* end_targets, int B, int L,
                                float* loss, tgb_stream_t stream) {
  int st = check_device();
  if (st != TGB_OK) return st;
  TGB_CHECK_ARG(logits && start_targets && end_targets && loss, "span_ce_loss: null pointer");
  TGB_CHECK_ARG(B > 0 && L > 0, "span_ce_loss: bad shape B=%d L=%d", B, L);
  span_ce_loss_kernel<<<1, CE_THREADS, 0, (cudaStream_t)stream>>>(logits, (const long long*)start_targets,
                                                                  (const long long*)end_targets, B, L, loss);
  TGB_CUDA(cudaGetLastError());
  return TGB_OK;
}
