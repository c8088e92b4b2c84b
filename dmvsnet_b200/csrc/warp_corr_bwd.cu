// N2 (W1 part): backward of the fused homography warp + 2-group correlation w.r.t. the feature maps (sm_100a).
//
// Replaces what autograd records for reference networks/mvsnet.py:137-146 (the product with the reference features, the
// mean over C/2 channels, the sum over views) and networks/module.py:247-249 (F.grid_sample): the gradient reaches the
// reference features and, through the bilinear weights, the source features.  The sampling grid is built under
// torch.no_grad() in the reference (module.py:222), so there is no gradient w.r.t. the hypotheses or the cameras.
//
//   d ref[c, p]        = 2/C * sum_{s,d} g[c&1, d, p] * sum_k wk(s,d,p) * src_s[c, corner_k(s,d,p)]
//   d src_s[c, q]     += 2/C * g[c&1, d, p] * wk(s,d,p) * ref[c, p]        for the four corners q of every sample
//
// Mapping: one thread = one reference pixel x 4 consecutive channels x DP depth planes; the C/4 threads of a pixel are
// neighbouring lanes.  Everything is channel-last, so a corner of a sample is C contiguous floats: the lanes of a pixel
// gather it with one 16-byte load each (for d ref) and scatter into it with one 16-byte vector reduction each
// (red.global.add.v4.f32, resolved in L2; no return value travels back).  Zero-weight corners (zero padding) issue no
// reduction.  d ref is accumulated in registers over the thread's planes and views and leaves with one vector reduction.
//
// The position arithmetic is the forward kernels' (warp_corr.cu), so forward and backward agree on every corner and weight.
// A non-finite sample position contributes nothing (ATen's CUDA grid_sampler drops such samples in both directions).
// Summation order across threads is not fixed (atomics): results are reproducible to fp32 rounding, not bit for bit.
#include "common.cuh"

namespace dmvs {

struct W1BwdParams {
  const float* ref;                // channel-last, pixel stride ref_ps, batch stride ref_bs
  const float* src[DMVS_MAX_SRC];  // channel-last, pixel stride src_ps, batch stride src_bs
  float* gsrc[DMVS_MAX_SRC];       // dense channel-last [B,h,w,C], zeroed before the launch
  float* gref;                     // dense channel-last [B,h,w,C], zeroed before the launch
  const float* rt;
  const float* hyp;
  const float* gcost;  // [B,2,D,h,w]
  long long ref_bs, src_bs;
  int ref_ps, src_ps;
  int B, D, h, w, n_src, n_chunks;
  float half_w, half_h;
};

__device__ __forceinline__ void route_axis_b(float pos, int n, int& base, float& c0, float& c1) {
  const float f0 = floorf(pos);
  const float w1 = pos - f0;
  const float w0 = 1.0f - w1;
  const int i0 = __float2int_rd(fminf(fmaxf(f0, -4.0f), (float)n + 4.0f));
  base = min(max(i0, 0), n - 2);
  const int rel = i0 - base;
  c0 = (rel == 0) ? w0 : ((rel == -1) ? w1 : 0.0f);
  c1 = (rel == 0) ? w1 : ((rel == 1) ? w0 : 0.0f);
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

constexpr int kBwdThreads = 256;

template <int C, int DP>
__global__ void __launch_bounds__(kBwdThreads) warp_corr_bwd_kernel(const __grid_constant__ W1BwdParams p) {
  constexpr int LG = C / 4;            // lanes per pixel
  constexpr int PX = kBwdThreads / LG;  // pixels per block, along x
  __shared__ float s_rt[DMVS_MAX_SRC * 12];
  const int b = blockIdx.z;
  for (int i = threadIdx.x; i < p.n_src * 12; i += kBwdThreads) s_rt[i] = p.rt[(long long)b * p.n_src * 12 + i];
  __syncthreads();

  const int tile_x = blockIdx.x / p.n_chunks;
  const int chunk = blockIdx.x - tile_x * p.n_chunks;
  const int sl = threadIdx.x % LG;
  const int x = tile_x * PX + threadIdx.x / LG;
  const int y = blockIdx.y;
  if (x >= p.w) return;
  const int hw = p.h * p.w;
  const int pix = y * p.w + x;
  const float fx = (float)x, fy = (float)y;
  const float inv_half = 2.0f / (float)C;

  const float4 r = ldg4(p.ref + (long long)b * p.ref_bs + (long long)pix * p.ref_ps + sl * 4);
  float4 gr = make_float4(0.f, 0.f, 0.f, 0.f);

#pragma unroll 1
  for (int dd = 0; dd < DP; ++dd) {
    const int d = chunk * DP + dd;
    if (d >= p.D) break;
    const float dep = __ldg(p.hyp + ((long long)(b * p.D + d) * hw) + pix);
    const float* gp = p.gcost + ((long long)(b * 2) * p.D + d) * hw + pix;
    const float g0 = __ldg(gp) * inv_half, g1 = __ldg(gp + (long long)p.D * hw) * inv_half;
    // what every corner of this sample receives per unit weight
    const float t0 = g0 * r.x, t1 = g1 * r.y, t2 = g0 * r.z, t3 = g1 * r.w;
#pragma unroll 1
    for (int s = 0; s < p.n_src; ++s) {
      const float* m = s_rt + s * 12;
      const float rx = __fadd_rn(__fmaf_rn(m[1], fy, __fmul_rn(m[0], fx)), m[2]);
      const float ry = __fadd_rn(__fmaf_rn(m[4], fy, __fmul_rn(m[3], fx)), m[5]);
      const float rz = __fadd_rn(__fmaf_rn(m[7], fy, __fmul_rn(m[6], fx)), m[8]);
      const float X = __fadd_rn(__fmul_rn(rx, dep), m[9]);
      const float Y = __fadd_rn(__fmul_rn(ry, dep), m[10]);
      float Z = __fadd_rn(__fmul_rn(rz, dep), m[11]);
      if (Z == 0.0f) Z += 1e-5f;
      const float u = __fdiv_rn(X, Z), v = __fdiv_rn(Y, Z);
      const float ix = __fmul_rn(__fadd_rn(__fsub_rn(__fdiv_rn(u, p.half_w), 1.0f), 1.0f), p.half_w);
      const float iy = __fmul_rn(__fadd_rn(__fsub_rn(__fdiv_rn(v, p.half_h), 1.0f), 1.0f), p.half_h);
      if (!(fabsf(ix) <= 3.0e38f) || !(fabsf(iy) <= 3.0e38f)) continue;
      int xb, yb;
      float cx0, cx1, cy0, cy1;
      route_axis_b(ix, p.w, xb, cx0, cx1);
      route_axis_b(iy, p.h, yb, cy0, cy1);
      const float w00 = cy0 * cx0, w01 = cy0 * cx1, w10 = cy1 * cx0, w11 = cy1 * cx1;
      if (w00 == 0.0f && w01 == 0.0f && w10 == 0.0f && w11 == 0.0f) continue;  // footprint entirely in the zero padding

      const int q = yb * p.w + xb;
      const float* sp = p.src[s] + (long long)b * p.src_bs + (long long)q * p.src_ps + sl * 4;
      const float4 v00 = ldg4(sp), v01 = ldg4(sp + p.src_ps);
      const float4 v10 = ldg4(sp + (long long)p.w * p.src_ps), v11 = ldg4(sp + (long long)(p.w + 1) * p.src_ps);
      gr.x = fmaf(g0, w00 * v00.x + w01 * v01.x + w10 * v10.x + w11 * v11.x, gr.x);
      gr.y = fmaf(g1, w00 * v00.y + w01 * v01.y + w10 * v10.y + w11 * v11.y, gr.y);
      gr.z = fmaf(g0, w00 * v00.z + w01 * v01.z + w10 * v10.z + w11 * v11.z, gr.z);
      gr.w = fmaf(g1, w00 * v00.w + w01 * v01.w + w10 * v10.w + w11 * v11.w, gr.w);

      float* gq = p.gsrc[s] + ((long long)b * hw + q) * C + sl * 4;
      if (w00 != 0.0f) red_add_v4(gq, w00 * t0, w00 * t1, w00 * t2, w00 * t3);
      if (w01 != 0.0f) red_add_v4(gq + C, w01 * t0, w01 * t1, w01 * t2, w01 * t3);
      if (w10 != 0.0f) red_add_v4(gq + (long long)p.w * C, w10 * t0, w10 * t1, w10 * t2, w10 * t3);
      if (w11 != 0.0f) red_add_v4(gq + (long long)(p.w + 1) * C, w11 * t0, w11 * t1, w11 * t2, w11 * t3);
    }
  }
  red_add_v4(p.gref + ((long long)b * hw + pix) * C + sl * 4, gr.x, gr.y, gr.z, gr.w);
}

template <int C>
static int launch_w1_bwd(W1BwdParams p, cudaStream_t st) {
  constexpr int PX = kBwdThreads / (C / 4);
  const long long threads = (long long)p.B * p.h * p.w * (C / 4);
  // planes per thread: more planes amortise the reference load and the d ref reduction; keep >= 4 waves in flight
  int dp = 8;
  while (dp > 1 && threads * ceil_div(p.D, dp) < 4LL * kNumSMs * 2048) dp >>= 1;
  p.n_chunks = ceil_div(p.D, dp);
  dim3 grid(ceil_div(p.w, PX) * p.n_chunks, p.h, p.B);
  DMVS_REQUIRE(grid.z <= 65535 && grid.y <= 65535, DMVS_ERR_BAD_SHAPE, "warp_corr_backward: grid too large (h=%d, B=%d)", p.h, p.B);
  switch (dp) {
    case 8: warp_corr_bwd_kernel<C, 8><<<grid, kBwdThreads, 0, st>>>(p); break;
    case 4: warp_corr_bwd_kernel<C, 4><<<grid, kBwdThreads, 0, st>>>(p); break;
    case 2: warp_corr_bwd_kernel<C, 2><<<grid, kBwdThreads, 0, st>>>(p); break;
    default: warp_corr_bwd_kernel<C, 1><<<grid, kBwdThreads, 0, st>>>(p); break;
  }
  return check_launch("warp_corr_backward");
}

}  // namespace dmvs

extern "C" int dmvs_warp_corr_backward_f32(const float* ref, long long ref_bstride, int ref_pixstride, const float* const* src,
                                           long long src_bstride, int src_pixstride, int n_src, const float* rt, const float* hyp,
                                           const float* grad_cost, float* grad_ref, float* const* grad_src, int B, int C, int D,
                                           int h, int w, void* stream) {
  using namespace dmvs;
  DMVS_REQUIRE(ref && src && rt && hyp && grad_cost && grad_ref && grad_src, DMVS_ERR_BAD_POINTER, "warp_corr_backward: null pointer");
  DMVS_REQUIRE(n_src >= 1 && n_src <= DMVS_MAX_SRC, DMVS_ERR_BAD_SHAPE, "warp_corr_backward: n_src=%d not in [1,%d]", n_src, DMVS_MAX_SRC);
  DMVS_REQUIRE(B >= 1 && D >= 1 && h >= 2 && w >= 2, DMVS_ERR_BAD_SHAPE, "warp_corr_backward: bad dims B=%d D=%d h=%d w=%d", B, D, h, w);
  DMVS_REQUIRE(C == 8 || C == 16 || C == 32, DMVS_ERR_BAD_SHAPE, "warp_corr_backward: C=%d unsupported (8, 16, 32)", C);
  DMVS_REQUIRE(ref_pixstride >= C && src_pixstride >= C && ref_pixstride % 4 == 0 && src_pixstride % 4 == 0 && ref_bstride % 4 == 0 &&
                   src_bstride % 4 == 0,
               DMVS_ERR_BAD_SHAPE, "warp_corr_backward: channel-last maps need pixel strides >= C and multiples of 4 floats");
  DMVS_REQUIRE((long long)h * w * (src_pixstride > ref_pixstride ? src_pixstride : ref_pixstride) < (1LL << 31), DMVS_ERR_BAD_SHAPE,
               "warp_corr_backward: feature map too large for 32-bit offsets");
  DMVS_REQUIRE(aligned16(ref) && aligned16(grad_ref), DMVS_ERR_BAD_POINTER, "warp_corr_backward: maps must be 16-byte aligned");
  W1BwdParams p;
  p.ref = ref;
  p.gref = grad_ref;
  for (int i = 0; i < DMVS_MAX_SRC; ++i) {
    p.src[i] = nullptr;
    p.gsrc[i] = nullptr;
  }
  for (int i = 0; i < n_src; ++i) {
    DMVS_REQUIRE(src[i] && grad_src[i] && aligned16(src[i]) && aligned16(grad_src[i]), DMVS_ERR_BAD_POINTER,
                 "warp_corr_backward: src[%d] / grad_src[%d] null or not 16-byte aligned", i, i);
    p.src[i] = src[i];
    p.gsrc[i] = grad_src[i];
  }
  p.rt = rt;
  p.hyp = hyp;
  p.gcost = grad_cost;
  p.ref_bs = ref_bstride;
  p.src_bs = src_bstride;
  p.ref_ps = ref_pixstride;
  p.src_ps = src_pixstride;
  p.B = B; p.D = D; p.h = h; p.w = w; p.n_src = n_src; p.n_chunks = 1;
  p.half_w = (float)((double)(w - 1) / 2.0);
  p.half_h = (float)((double)(h - 1) / 2.0);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t bytes = (size_t)B * h * w * C * sizeof(float);
  cudaError_t e = cudaMemsetAsync(grad_ref, 0, bytes, st);
  for (int i = 0; i < n_src && e == cudaSuccess; ++i) e = cudaMemsetAsync(grad_src[i], 0, bytes, st);
  if (e != cudaSuccess) {
    set_error("warp_corr_backward: memset failed: %s", cudaGetErrorString(e));
    return DMVS_ERR_CUDA;
  }
  switch (C) {
    case 8: return launch_w1_bwd<8>(p, st);
    case 16: return launch_w1_bwd<16>(p, st);
    default: return launch_w1_bwd<32>(p, st);
  }
}
