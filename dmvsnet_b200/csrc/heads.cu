// E1 / E2 / S1: the pointwise kernels around the regularisation nets (sm_100a).
//
//   E1 depth_head      <- DepthNet.forward   networks/mvsnet.py:15-66  (+ depth_regression module.py:454-460)
//   E2 refine_head     <- DepthNet.refine    networks/mvsnet.py:67-100
//   S1 hypotheses_*    <- get_depth_range_samples networks/module.py:476-649 (+ F.interpolate mvsnet.py:232-233)
//
// All three are bandwidth kernels: one thread per pixel, x fastest so every load/store of a
// [.., h, w] plane is a coalesced 128-byte line per warp.
#include "common.cuh"

namespace dmvs {

__device__ __forceinline__ float confidence_of(const float d[4], float interval) {
  // 2 * (sigmoid(interval / (population-std over the 4 regressed depths + 1e-5)) - 0.5)   mvsnet.py:61-62
  const float mean = (d[0] + d[1] + d[2] + d[3]) * 0.25f;
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) var += (d[i] - mean) * (d[i] - mean);
  var *= 0.25f;
  const float z = interval / (sqrtf(var) + 1e-5f);
  const float sig = 1.0f / (1.0f + expf(-z));
  return 2.0f * (sig - 0.5f);
}

// ------------------------------------------------------------------------------------------ E1
// Softmax over D for each of the 4 logit channels, expectation against the per-pixel hypotheses, then the dual-depth
// bookkeeping (row class, extrapolation stack, window selection, confidence).
__device__ __forceinline__ void dual_depth_tail(const float (&d4)[4], int x, int y, int b, long long hw, long long pix,
                                                float* __restrict__ hyp_c, float* __restrict__ conf, float interval) {
  // row class: 0 small, 1 huge, 2 small with doubled range, 3 huge with doubled range   mvsnet.py:25-28,33-56
  const int r = y & 3;
  float lo = (r & 1) ? fminf(d4[2], d4[3]) : fminf(d4[0], d4[1]);
  float hi = (r & 1) ? fmaxf(d4[2], d4[3]) : fmaxf(d4[0], d4[1]);
  if (r & 2) {
    const float lo2 = 2.f * lo - hi, hi2 = 2.f * hi - lo;
    lo = lo2;
    hi = hi2;
  }
  const float s6[6] = {3.f * lo - 2.f * hi, 2.f * lo - hi, lo, hi, 2.f * hi - lo, 3.f * hi - 2.f * lo};
  const bool low_window = ((x & 1) == 0) == ((r & 1) == 0);
  const int off = low_window ? 0 : 2;
#pragma unroll
  for (int j = 0; j < 4; ++j) hyp_c[(long long)(b * 4 + j) * hw + pix] = s6[off + j];
  conf[(long long)b * hw + pix] = confidence_of(d4, interval);
}

// Block = PXB pixels of one row x the 4 logit channels (threadIdx.y = channel): each thread owns one (pixel, channel)
// softmax column.  With D known at compile time the D logits live in registers - read from memory exactly once with D
// independent loads in flight, exponentiated once, probability volume written in the same pass.  The four regressed
// depths of a pixel meet in shared memory for the dual-depth tail.
// PXB pixels per block: the block's loads of one (channel, plane) row are PXB*4 contiguous bytes - with 32 they are single
// 128-byte lines scattered over 4*D planes 2-8 MB apart (DRAM row misses), 64-128 keep a DRAM row open for the block.
template <int DT, int PXB>  // DT > 0: compile-time D (registers); DT == 0: any D, three streaming passes (2nd / 3rd hit L2)
__global__ void __launch_bounds__(PXB * 4) depth_head_kernel(const float* __restrict__ logits, const float* __restrict__ hyp,
                                                         const float* __restrict__ interval_p, float* __restrict__ prob,
                                                         float* __restrict__ d4o, float* __restrict__ hyp_c,
                                                         float* __restrict__ conf, int Drt, int h, int w) {
  __shared__ float d4s[4][PXB];
  const int D = (DT > 0) ? DT : Drt;
  const int x = blockIdx.x * PXB + threadIdx.x;
  const int y = blockIdx.y;
  const int c = threadIdx.y;
  const int b = blockIdx.z;
  const long long hw = (long long)h * w;
  const long long pix = (long long)y * w + x;
  if (x < w) {
    const float* hp = hyp + (long long)b * D * hw + pix;
    const float* lp = logits + ((long long)(b * 4 + c) * D) * hw + pix;
    float* pp = prob ? prob + ((long long)(b * 4 + c) * D) * hw + pix : nullptr;
    float acc = 0.f;
    if (DT > 0) {
      constexpr int DN = DT > 0 ? DT : 1;
      float v[DN];
#pragma unroll
      for (int k = 0; k < DN; ++k) v[k] = __ldg(lp + k * hw);
      float mx = v[0];
#pragma unroll
      for (int k = 1; k < DN; ++k) mx = fmaxf(mx, v[k]);
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < DN; ++k) {
        v[k] = expf(v[k] - mx);
        sum += v[k];
      }
      // one IEEE division per softmax column, then multiplications: p differs from exp/sum by <= 1 ulp, and the per-element
      // division (its slow path fires for the denormal exponentials of a peaked softmax) was 2/3 of the kernel's instructions
      const float inv_sum = 1.0f / sum;
      // hypotheses in batches of 8 independent loads (a load -> fma -> load chain would pay D memory latencies in a row)
#pragma unroll
      for (int k0 = 0; k0 < DN; k0 += 8) {
        float hh[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) hh[i] = (k0 + i < DN) ? __ldg(hp + (k0 + i) * hw) : 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (k0 + i < DN) {
            const float pr = v[k0 + i] * inv_sum;
            if (pp) pp[(k0 + i) * hw] = pr;
            acc += pr * hh[i];
          }
        }
      }
    } else {
      float mx = -INFINITY;
      for (int k = 0; k < D; ++k) mx = fmaxf(mx, __ldg(lp + k * hw));
      float sum = 0.f;
      for (int k = 0; k < D; ++k) sum += expf(__ldg(lp + k * hw) - mx);
      const float inv_sum = 1.0f / sum;
      for (int k = 0; k < D; ++k) {
        const float pr = expf(__ldg(lp + k * hw) - mx) * inv_sum;
        if (pp) pp[k * hw] = pr;
        acc += pr * __ldg(hp + k * hw);
      }
    }
    d4o[(long long)(b * 4 + c) * hw + pix] = acc;
    d4s[c][threadIdx.x] = acc;
  }
  __syncthreads();
  if (c == 0 && x < w) {
    const float d4[4] = {d4s[0][threadIdx.x], d4s[1][threadIdx.x], d4s[2][threadIdx.x], d4s[3][threadIdx.x]};
    dual_depth_tail(d4, x, y, b, hw, pix, hyp_c, conf, __ldg(interval_p));
  }
}

// ------------------------------------------------------------------------------------------ E2
__global__ void __launch_bounds__(128) refine_head_kernel(const float* __restrict__ logits, const float* __restrict__ hyp_c,
                                                          const float* __restrict__ interval_p, float alpha,
                                                          float* __restrict__ depth, float* __restrict__ conf,
                                                          float* __restrict__ d4o, int h, int w) {
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = blockIdx.y * 4 + threadIdx.y;
  if (x >= w || y >= h) return;
  const int b = blockIdx.z;
  const long long hw = (long long)h * w;
  const long long pix = (long long)y * w + x;
  float hv[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) hv[k] = __ldg(hyp_c + (long long)(b * 4 + k) * hw + pix);
  float d4[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) l[k] = __ldg(logits + ((long long)(b * 4 + c) * 4 + k) * hw + pix) * alpha;
    const float mx = fmaxf(fmaxf(l[0], l[1]), fmaxf(l[2], l[3]));
    float e[4], sum = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      e[k] = expf(l[k] - mx);
      sum += e[k];
    }
    float acc = 0.f;
    const float inv_sum = 1.0f / sum;
#pragma unroll
    for (int k = 0; k < 4; ++k) acc += (e[k] * inv_sum) * hv[k];
    d4[c] = acc;
    d4o[(long long)(b * 4 + c) * hw + pix] = acc;
  }
  // (row%2, col%2): 00 small-min, 01 small-max, 10 huge-max, 11 huge-min       mvsnet.py:80-91
  float out;
  if ((y & 1) == 0)
    out = ((x & 1) == 0) ? fminf(d4[0], d4[1]) : fmaxf(d4[0], d4[1]);
  else
    out = ((x & 1) == 0) ? fmaxf(d4[2], d4[3]) : fminf(d4[2], d4[3]);
  depth[(long long)b * hw + pix] = out;
  conf[(long long)b * hw + pix] = confidence_of(d4, __ldg(interval_p));
}

// ------------------------------------------------------------------------------------------ S1, stage 0
// torch.linspace(start, end, n): start + step*i for i < n/2, end - step*(n-1-i) otherwise.
__device__ __forceinline__ float linspace_at(float start, float end, int n, int i) {
  const float step = __fdiv_rn(__fsub_rn(end, start), (float)(n - 1));
  return (i < n / 2) ? __fadd_rn(start, __fmul_rn(step, (float)i)) : __fsub_rn(end, __fmul_rn(step, (float)(n - 1 - i)));
}

__global__ void __launch_bounds__(128) hypotheses_first_kernel(const float* __restrict__ depth_values, int Nd,
                                                               float* __restrict__ hyp, float* __restrict__ interval_out,
                                                               int D, int h, int w, int inverse) {
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = blockIdx.y * 4 + threadIdx.y;
  const int b = blockIdx.z;
  const float lo = __ldg(depth_values + (long long)b * Nd), hi = __ldg(depth_values + (long long)b * Nd + Nd - 1);
  // the shift uses batch 0's interval (module.py:564,603)
  const float lo0 = __ldg(depth_values), hi0 = __ldg(depth_values + Nd - 1);
  const float si = __fdiv_rn(__fsub_rn(hi0, lo0), (float)(D - 1));
  if (interval_out && b == 0 && x == 0 && y == 0) {
    float out_si = si;
    if (inverse) {  // recomputed from the shifted ends twice (module.py:606-621); equal up to rounding
      const float s2 = __fdiv_rn(__fsub_rn(__fsub_rn(hi0, si), __fsub_rn(lo0, si)), (float)(D - 1));
      out_si = __fdiv_rn(__fsub_rn(__fadd_rn(hi0, s2), __fadd_rn(lo0, s2)), (float)(D - 1));
    }
    *interval_out = out_si;
  }
  if (x >= w || y >= h) return;
  const bool even = ((x + y) & 1) == 0;
  const long long hw = (long long)h * w;
  float* op = hyp + (long long)b * D * hw + (long long)y * w + x;
  if (!inverse) {
    const float step = __fdiv_rn(__fsub_rn(hi, lo), (float)(D - 1));
    for (int k = 0; k < D; ++k) {
      const float plane = __fadd_rn(lo, __fmul_rn((float)k, step));
      op[k * hw] = even ? __fsub_rn(plane, si) : __fadd_rn(plane, si);
    }
  } else {
    const float s2 = __fdiv_rn(__fsub_rn(__fsub_rn(hi0, si), __fsub_rn(lo0, si)), (float)(D - 1));
    const float a = even ? __fsub_rn(lo, si) : __fadd_rn(lo, s2);
    const float e = even ? __fsub_rn(hi, si) : __fadd_rn(hi, s2);
    const float ia = __fdiv_rn(1.0f, a), ie = __fdiv_rn(1.0f, e);
    for (int k = 0; k < D; ++k) op[k * hw] = __fdiv_rn(1.0f, linspace_at(ia, ie, D, k));
  }
}

// ------------------------------------------------------------------------------------------ S1, stages > 0
struct RangeSample {
  float base, step;  // linear: lo, (hi-lo)/(D-1);  inverse: 1/lo, (1/hi-1/lo)/(D-1)
};

__device__ __forceinline__ RangeSample make_range(float last, bool even, int D, float ip, int inverse) {
  // even pixels: "_n" range [last-(D+2)/2*ip, last+(D-2)/2*ip]; odd: "_p" (module.py:476-507,525-554)
  const float big = (float)((D + 2) * 0.5), small = (float)((D - 2) * 0.5);
  const float lo = __fsub_rn(last, __fmul_rn(even ? big : small, ip));
  const float hi = __fadd_rn(last, __fmul_rn(even ? small : big, ip));
  RangeSample r;
  if (inverse) {
    const float ilo = __fdiv_rn(1.0f, lo), ihi = __fdiv_rn(1.0f, hi);
    r.base = ilo;
    r.step = __fdiv_rn(__fsub_rn(ihi, ilo), (float)(D - 1));
  } else {
    r.base = lo;
    r.step = __fdiv_rn(__fsub_rn(hi, lo), (float)(D - 1));
  }
  return r;
}

__device__ __forceinline__ float eval_range(const RangeSample& r, int k, int inverse) {
  const float v = __fadd_rn(r.base, __fmul_rn((float)k, r.step));
  return inverse ? __fdiv_rn(1.0f, v) : v;
}

// Each output pixel blends the plane values of four pixels of the coarser map, and (for the x2 upsample of the cascade) every
// coarse pixel feeds ~16 output pixels: the plane values - one IEEE division each in inverse-depth mode - are evaluated once
// per coarse pixel of the block's footprint into shared memory, KP planes at a time, and the 128 output pixels blend from
// there.  Same operations on the same operands as evaluating them per output pixel: bit-identical results.
constexpr int kHypRows = 6, kHypCols = 34, kHypKP = 4;  // footprint of a 32x4 output block for scale <= 1: <= 5 x 33 coarse pixels

__global__ void __launch_bounds__(128) hypotheses_next_kernel(const float* __restrict__ last_depth,
                                                              const float* __restrict__ interval_pixel,
                                                              float* __restrict__ hyp, float* __restrict__ interval_out,
                                                              int D, int h0, int w0, int h, int w, int inverse) {
  __shared__ float s_val[kHypKP][kHypRows][kHypCols];
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = blockIdx.y * 4 + threadIdx.y;
  const int b = blockIdx.z;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const float ip = __ldg(interval_pixel);
  if (interval_out && b == 0 && x == 0 && y == 0) *interval_out = __fdiv_rn(__fmul_rn((float)D, ip), (float)(D - 1));
  const float scy = (float)h0 / (float)h, scx = (float)w0 / (float)w;
  // F.interpolate(bilinear, align_corners=False): src = scale*(dst+0.5)-0.5 clamped at 0
  const float sy = fmaxf(__fsub_rn(__fmul_rn(scy, (float)y + 0.5f), 0.5f), 0.f);
  const float sx = fmaxf(__fsub_rn(__fmul_rn(scx, (float)x + 0.5f), 0.5f), 0.f);
  const int y0 = (int)sy, x0 = (int)sx;
  const int y1 = y0 + ((y0 < h0 - 1) ? 1 : 0), x1 = x0 + ((x0 < w0 - 1) ? 1 : 0);
  const float ly1 = sy - (float)y0, ly0 = 1.0f - ly1;
  const float lx1 = sx - (float)x0, lx0 = 1.0f - lx1;
  // origin of the block's coarse footprint = (y0, x0) of its first output pixel (both are monotonic in y / x)
  const int oy = (int)fmaxf(__fsub_rn(__fmul_rn(scy, (float)(blockIdx.y * 4) + 0.5f), 0.5f), 0.f);
  const int ox = (int)fmaxf(__fsub_rn(__fmul_rn(scx, (float)(blockIdx.x * 32) + 0.5f), 0.5f), 0.f);
  const bool valid = x < w && y < h;
  const float* lp = last_depth + (long long)b * h0 * w0;
  // this thread's (up to two) coarse pixels of the footprint
  RangeSample rs[2];
  int sr[2], sc[2];
  bool have[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int idx = tid + j * 128;
    sr[j] = idx / kHypCols;
    sc[j] = idx - sr[j] * kHypCols;
    const int gy = oy + sr[j], gx = ox + sc[j];
    have[j] = idx < kHypRows * kHypCols && gy < h0 && gx < w0;
    if (have[j]) rs[j] = make_range(__ldg(lp + gy * w0 + gx), ((gy + gx) & 1) == 0, D, ip, inverse);
  }
  const int r0 = y0 - oy, r1 = y1 - oy, c0 = x0 - ox, c1 = x1 - ox;
  const long long hw = (long long)h * w;
  float* op = hyp + (long long)b * D * hw + (long long)y * w + x;
  for (int k0 = 0; k0 < D; k0 += kHypKP) {
#pragma unroll
    for (int j = 0; j < 2; ++j)
      if (have[j]) {
#pragma unroll
        for (int kk = 0; kk < kHypKP; ++kk)
          if (k0 + kk < D) s_val[kk][sr[j]][sc[j]] = eval_range(rs[j], k0 + kk, inverse);
      }
    __syncthreads();
    if (valid) {
#pragma unroll
      for (int kk = 0; kk < kHypKP; ++kk)
        if (k0 + kk < D) {
          const float top = __fadd_rn(__fmul_rn(lx0, s_val[kk][r0][c0]), __fmul_rn(lx1, s_val[kk][r0][c1]));
          const float bot = __fadd_rn(__fmul_rn(lx0, s_val[kk][r1][c0]), __fmul_rn(lx1, s_val[kk][r1][c1]));
          op[(long long)(k0 + kk) * hw] = __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
        }
    }
    __syncthreads();
  }
}

int g_head_px = 32;  // pixels per block of the depth head (dmvs_debug_set("head_px", 32 | 64 | 128))

static inline dim3 pixel_grid(int B, int h, int w) { return dim3(ceil_div(w, 32), ceil_div(h, 4), B); }

}  // namespace dmvs

using namespace dmvs;

extern "C" int dmvs_depth_head_f32(const float* logits, const float* hyp, const float* interval, float* prob, float* d4,
                                   float* hyp_c, float* conf, int B, int D, int h, int w, void* stream) {
  DMVS_REQUIRE(logits && hyp && interval && d4 && hyp_c && conf, DMVS_ERR_BAD_POINTER, "depth_head: null pointer");
  DMVS_REQUIRE(B >= 1 && B <= 65535 && D >= 1 && h >= 1 && w >= 1, DMVS_ERR_BAD_SHAPE, "depth_head: bad dims");
  DMVS_REQUIRE(h <= 65535, DMVS_ERR_BAD_SHAPE, "depth_head: h=%d too large", h);
  cudaStream_t st = (cudaStream_t)stream;
  const int pxb = g_head_px;
#define DMVS_HEAD_LAUNCH(N, P)                                                                                              \
  depth_head_kernel<N, P><<<dim3(ceil_div(w, P), h, B), dim3(P, 4), 0, st>>>(logits, hyp, interval, prob, d4, hyp_c, conf, D, h, w)
#define DMVS_HEAD_CASE(N)                                 \
  case N:                                                 \
    if (pxb == 32) DMVS_HEAD_LAUNCH(N, 32);               \
    else if (pxb == 128) DMVS_HEAD_LAUNCH(N, 128);        \
    else DMVS_HEAD_LAUNCH(N, 64);                         \
    break;
  switch (D) {
    DMVS_HEAD_CASE(8)
    DMVS_HEAD_CASE(16)
    DMVS_HEAD_CASE(32)
    DMVS_HEAD_CASE(48)
    DMVS_HEAD_CASE(64)
    default: DMVS_HEAD_LAUNCH(0, 64);
  }
#undef DMVS_HEAD_CASE
#undef DMVS_HEAD_LAUNCH
  return check_launch("depth_head");
}

extern "C" int dmvs_refine_head_f32(const float* logits_c, const float* hyp_c, const float* interval, float alpha,
                                    float* depth, float* conf, float* d4, int B, int h, int w, void* stream) {
  DMVS_REQUIRE(logits_c && hyp_c && interval && depth && conf && d4, DMVS_ERR_BAD_POINTER, "refine_head: null pointer");
  DMVS_REQUIRE(B >= 1 && B <= 65535 && h >= 1 && w >= 1, DMVS_ERR_BAD_SHAPE, "refine_head: bad dims");
  refine_head_kernel<<<pixel_grid(B, h, w), dim3(32, 4), 0, (cudaStream_t)stream>>>(logits_c, hyp_c, interval, alpha, depth,
                                                                                  conf, d4, h, w);
  return check_launch("refine_head");
}

extern "C" int dmvs_hypotheses_first_f32(const float* depth_values, int Nd, float* hyp, float* interval_out, int B, int D,
                                         int h, int w, int inverse, void* stream) {
  DMVS_REQUIRE(depth_values && hyp, DMVS_ERR_BAD_POINTER, "hypotheses_first: null pointer");
  DMVS_REQUIRE(B >= 1 && B <= 65535 && D >= 2 && Nd >= 2 && h >= 1 && w >= 1, DMVS_ERR_BAD_SHAPE, "hypotheses_first: bad dims");
  hypotheses_first_kernel<<<pixel_grid(B, h, w), dim3(32, 4), 0, (cudaStream_t)stream>>>(depth_values, Nd, hyp, interval_out,
                                                                                       D, h, w, inverse);
  return check_launch("hypotheses_first");
}

extern "C" int dmvs_hypotheses_next_f32(const float* last_depth, const float* interval_pixel, float* hyp,
                                        float* interval_out, int B, int D, int h0, int w0, int h, int w, int inverse,
                                        void* stream) {
  DMVS_REQUIRE(last_depth && interval_pixel && hyp, DMVS_ERR_BAD_POINTER, "hypotheses_next: null pointer");
  DMVS_REQUIRE(B >= 1 && B <= 65535 && D >= 2 && h0 >= 1 && w0 >= 1 && h >= 1 && w >= 1, DMVS_ERR_BAD_SHAPE,
               "hypotheses_next: bad dims");
  DMVS_REQUIRE(h0 <= h && w0 <= w, DMVS_ERR_BAD_SHAPE, "hypotheses_next: the hypotheses are upsampled, never reduced (%dx%d -> %dx%d)", h0, w0, h, w);
  hypotheses_next_kernel<<<pixel_grid(B, h, w), dim3(32, 4), 0, (cudaStream_t)stream>>>(last_depth, interval_pixel, hyp,
                                                                                      interval_out, D, h0, w0, h, w, inverse);
  return check_launch("hypotheses_next");
}
