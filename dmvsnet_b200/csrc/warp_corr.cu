// W1: fused homography warp + 2-group correlation over source views (sm_100a).
//
// Replaces reference networks/mvsnet.py:111-153 (CostAgg.forward) and networks/module.py:212-251
// (homo_warping): the per-source sampling grid, F.grid_sample, the [B,C,D,h,w] warped volume, the
// product volume and the running sum over views are all fused into one pass that reads the features
// and the hypotheses once and writes the [B,2,D,h,w] cost volume once.
//
// Arithmetic kept in the reference's order where it is observable (module.py:233-241 and ATen's
// vectorised CPU grid_sampler): rotate the pixel, scale by the depth, translate, patch Z == 0,
// divide, normalise by (size-1)/2, un-normalise, floor, (1-frac)/frac weights, per-corner zero padding.
// The one re-association: the channel dot product is taken per corner and the four bilinear weights are
// applied to the four dot products (4 FMA per gathered value instead of 8 flops), see DESIGN.md.
//
// Mapping: one thread = one reference pixel x DP consecutive depth planes.  A warp covers 32
// consecutive x, so for every channel and corner the 32 lanes read ~32 consecutive floats of one
// source row (NCHW, coalesced through L1); the reference pixel's C features stay in registers.
#include <cuda_fp16.h>

#include "common.cuh"

namespace dmvs {

struct WarpCorrParams {
  const float* ref;
  const float* src[DMVS_MAX_SRC];
  const float* rt;
  const float* hyp;
  float* cost;    // [B,2,D,h,w] fp32, nullable
  uint2* cells;   // conv0 input cells (DMVS_FMT_COST2), nullable: [B][D][h][w+1] x 16 B, cell x = [voxel x-1 | voxel x]
  long long ref_bs, src_bs;
  int B, D, h, w, n_src, d_begin, d_end, n_chunks;
  float half_w, half_h;  // (w-1)/2, (h-1)/2 rounded to fp32 like the reference's python-float divisor
};

// Column/row bookkeeping for zero padding with always-in-range addresses: we load rows/cols
// (b, b+1) with b = clamp(i0, 0, n-2) and route the two bilinear weights onto those two loads.
__device__ __forceinline__ void route_axis(float pos, int n, int& base, float& c0, float& c1) {
  const float f0 = floorf(pos);
  const float w1 = pos - f0;    // weight of the far (east / south) neighbour
  const float w0 = 1.0f - w1;   // weight of the near (west / north) neighbour
  // saturating float->int keeps +-inf / huge values out of range; NaN is handled by the caller
  const int i0 = __float2int_rd(fminf(fmaxf(f0, -4.0f), (float)n + 4.0f));
  base = min(max(i0, 0), n - 2);
  const int rel = i0 - base;
  c0 = (rel == 0) ? w0 : ((rel == -1) ? w1 : 0.0f);
  c1 = (rel == 0) ? w1 : ((rel == 1) ? w0 : 0.0f);
}

template <int C, int DP>
__global__ void __launch_bounds__(128, 4) warp_corr_kernel(const __grid_constant__ WarpCorrParams p) {
  // plane chunks are the fastest-varying block index: the blocks that share a pixel tile (same reference
  // features, neighbouring source footprints) are co-resident, so the features cross L2->L1 while hot
  // instead of being re-streamed once per chunk sweep.
  const int tile_x = blockIdx.x / p.n_chunks;
  const int chunk = blockIdx.x - tile_x * p.n_chunks;
  const int x = tile_x * 32 + threadIdx.x;
  const int y = blockIdx.y * 4 + threadIdx.y;
  if (x >= p.w || y >= p.h) return;
  const int b = blockIdx.z;
  const int d0 = p.d_begin + chunk * DP;
  const int hw = p.h * p.w;
  const int pix = y * p.w + x;

  float refv[C];
  {
    const float* rp = p.ref + (long long)b * p.ref_bs + pix;
#pragma unroll
    for (int c = 0; c < C; ++c) refv[c] = __ldg(rp + (long long)c * hw);
  }
  const float fx = (float)x, fy = (float)y;
  const float inv_half = 2.0f / (float)C;  // mean over C/2 channels; C/2 is a power of two -> exact

#pragma unroll 1
  for (int dd = 0; dd < DP; ++dd) {
    const int d = d0 + dd;
    if (d >= p.d_end) break;
    const float dep = __ldg(p.hyp + ((long long)(b * p.D + d) * hw) + pix);
    float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll 1
    for (int s = 0; s < p.n_src; ++s) {
      const float* m = p.rt + (b * p.n_src + s) * 12;
      // rot @ (x, y, 1)
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

      int xb, yb;
      float cx0, cx1, cy0, cy1;
      route_axis(ix, p.w, xb, cx0, cx1);
      route_axis(iy, p.h, yb, cy0, cy1);

      const float* q0 = p.src[s] + (long long)b * p.src_bs + (yb * p.w + xb);
      const float* q1 = q0 + p.w;
      float s00[2] = {0.f, 0.f}, s01[2] = {0.f, 0.f}, s10[2] = {0.f, 0.f}, s11[2] = {0.f, 0.f};
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float v00 = __ldg(q0), v01 = __ldg(q0 + 1), v10 = __ldg(q1), v11 = __ldg(q1 + 1);
        s00[c & 1] = fmaf(refv[c], v00, s00[c & 1]);
        s01[c & 1] = fmaf(refv[c], v01, s01[c & 1]);
        s10[c & 1] = fmaf(refv[c], v10, s10[c & 1]);
        s11[c & 1] = fmaf(refv[c], v11, s11[c & 1]);
        q0 += hw;
        q1 += hw;
      }
      const float w00 = cy0 * cx0, w01 = cy0 * cx1, w10 = cy1 * cx0, w11 = cy1 * cx1;
      float g0 = w00 * s00[0] + w01 * s01[0] + w10 * s10[0] + w11 * s11[0];
      float g1 = w00 * s00[1] + w01 * s01[1] + w10 * s10[1] + w11 * s11[1];
      // the reference multiplies (masked) zeros by NaN weights when the sample position is not finite
      if (!(fabsf(ix) <= 3.0e38f) || !(fabsf(iy) <= 3.0e38f)) g0 = g1 = __int_as_float(0x7fc00000);
      acc0 += g0 * inv_half;
      acc1 += g1 * inv_half;
    }
    if (p.cost) {
      float* cp = p.cost + ((long long)(b * 2) * p.D + d) * hw + pix;
      cp[0] = acc0;
      cp[(long long)p.D * hw] = acc1;
    }
    if (p.cells) {
      // the tensor path's conv0 consumes the two cost channels as hi/lo fp16 pairs, two voxels per 16-byte cell:
      // this voxel is the second half of cell x and the first half of cell x+1 (16 contiguous bytes at 16x + 8)
      const __half h0 = __float2half_rn(acc0), h1 = __float2half_rn(acc1);
      const __half2 hi = __halves2half2(h0, h1);
      const __half2 lo = __halves2half2(__float2half_rn(acc0 - __half2float(h0)), __float2half_rn(acc1 - __half2float(h1)));
      const uint2 v = make_uint2(*reinterpret_cast<const uint32_t*>(&hi), *reinterpret_cast<const uint32_t*>(&lo));
      uint2* row = p.cells + (((long long)(b * p.D + d) * p.h + y) * (p.w + 1)) * 2;  // 2 x uint2 per cell
      row[2 * x + 1] = v;
      row[2 * x + 2] = v;
      if (x == 0) row[0] = make_uint2(0u, 0u);                      // voxel -1
      if (x == p.w - 1) row[2 * p.w + 1] = make_uint2(0u, 0u);      // voxel w
    }
  }
}

template <int C>
static int launch_warp_corr(const WarpCorrParams& p0, cudaStream_t st) {
  WarpCorrParams p = p0;
  const int nd = p.d_end - p.d_begin;
  const long long pixels = (long long)p.B * p.h * p.w;
  // planes per thread: keep >= ~4 waves of 2048 threads/SM in flight, otherwise favour reuse of the ref registers
  int dp = 4;
  while (dp > 1 && pixels * ceil_div(nd, dp) < 4LL * kNumSMs * 2048) dp >>= 1;
  p.n_chunks = ceil_div(nd, dp);
  dim3 block(32, 4, 1), grid(ceil_div(p.w, 32) * p.n_chunks, ceil_div(p.h, 4), p.B);
  DMVS_REQUIRE(grid.z <= 65535 && grid.y <= 65535, DMVS_ERR_BAD_SHAPE, "warp_corr: grid too large (h=%d, B=%d)", p.h, p.B);
  if (dp == 4)
    warp_corr_kernel<C, 4><<<grid, block, 0, st>>>(p);
  else if (dp == 2)
    warp_corr_kernel<C, 2><<<grid, block, 0, st>>>(p);
  else
    warp_corr_kernel<C, 1><<<grid, block, 0, st>>>(p);
  return check_launch("warp_corr");
}

}  // namespace dmvs

extern "C" int dmvs_warp_corr_f32(const float* ref, long long ref_bstride, const float* const* src, long long src_bstride,
                                  int n_src, const float* rt, const float* hyp, float* cost, void* cost_cells, int B, int C,
                                  int D, int h, int w, int d_begin, int d_end, void* stream) {
  using namespace dmvs;
  DMVS_REQUIRE(ref && src && rt && hyp && (cost || cost_cells), DMVS_ERR_BAD_POINTER, "warp_corr: null pointer");
  DMVS_REQUIRE(!cost_cells || aligned16(cost_cells), DMVS_ERR_BAD_POINTER, "warp_corr: cost_cells must be 16-byte aligned");
  DMVS_REQUIRE(n_src >= 1 && n_src <= DMVS_MAX_SRC, DMVS_ERR_BAD_SHAPE, "warp_corr: n_src=%d not in [1,%d]", n_src, DMVS_MAX_SRC);
  DMVS_REQUIRE(B >= 1 && D >= 1 && h >= 2 && w >= 2, DMVS_ERR_BAD_SHAPE, "warp_corr: bad dims B=%d D=%d h=%d w=%d", B, D, h, w);
  DMVS_REQUIRE(0 <= d_begin && d_begin <= d_end && d_end <= D, DMVS_ERR_BAD_SHAPE, "warp_corr: bad plane range [%d,%d) of %d",
               d_begin, d_end, D);
  DMVS_REQUIRE((long long)C * h * w < (1LL << 31), DMVS_ERR_BAD_SHAPE, "warp_corr: feature map too large for 32-bit offsets");
  if (d_begin == d_end) return DMVS_OK;
  WarpCorrParams p;
  p.ref = ref;
  for (int i = 0; i < DMVS_MAX_SRC; ++i) p.src[i] = (i < n_src) ? src[i] : nullptr;
  for (int i = 0; i < n_src; ++i) DMVS_REQUIRE(src[i] != nullptr, DMVS_ERR_BAD_POINTER, "warp_corr: src[%d] is null", i);
  p.rt = rt;
  p.hyp = hyp;
  p.cost = cost;
  p.cells = reinterpret_cast<uint2*>(cost_cells);
  p.ref_bs = ref_bstride;
  p.src_bs = src_bstride;
  p.B = B; p.D = D; p.h = h; p.w = w; p.n_src = n_src; p.d_begin = d_begin; p.d_end = d_end; p.n_chunks = 1;
  p.half_w = (float)((double)(w - 1) / 2.0);
  p.half_h = (float)((double)(h - 1) / 2.0);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (C) {
    case 8: return launch_warp_corr<8>(p, st);
    case 16: return launch_warp_corr<16>(p, st);
    case 32: return launch_warp_corr<32>(p, st);
    default: set_error("warp_corr: C=%d unsupported (8, 16, 32)", C); return DMVS_ERR_BAD_SHAPE;
  }
}
