// W1, channel-last gather variant: fused homography warp + 2-group correlation (sm_100a).
//
// Same contract as warp_corr.cu (reference networks/mvsnet.py:111-153 CostAgg.forward +
// networks/module.py:212-251 homo_warping) but the SOURCE feature maps are channel-last ([B,h,w,C], any pixel
// stride), so the two x-neighbours of a bilinear footprint are ONE contiguous run of 2*C floats.
//
// Why: the L1 data path serves one 128-byte line per cycle, whatever fraction of it the warp uses.  With NCHW
// sources a warp-wide load of one (corner, channel) touches as many lines as its 32 lanes have distinct source rows
// / 32-float segments: 2-3 when neighbouring pixels sample neighbouring positions, up to 32 when the hypotheses are
// rough (per-pixel depth from the previous stage), and every lane needs 4*C such loads.  Channel-last, a footprint
// row is 64 / 128 / 256 contiguous bytes (C = 8 / 16 / 32): G = C/2 lanes fetch it with one LDG.128 each, so a
// warp-wide load serves 32/G samples and touches ~1-3 lines per sample row in the worst case and shares lines
// between neighbouring samples in the best case.
//
// Mapping: a warp owns a 16x2 patch of reference pixels and DP consecutive planes.  Lane groups of G lanes own one
// pixel at a time (lane = corner dx x 16-byte channel chunk k, its 4 reference channels in registers); the patch is
// walked in zig-zag order ((x+y) parity classes first: the checkerboard hypotheses of the sampler,
// module.py:577-594, make same-parity neighbours sample neighbouring source positions).  Sample positions are
// computed ONCE per (pixel, plane, source) by one lane (phase A, the reference's op order, module.py:233-241 +
// ATen's grid_sampler un-normalisation) and broadcast through shared memory as 2 x 16-byte records
// {element offset, cx[dx], cy0, cy1} with the zero padding already routed into the weights; phase B is
// 1 LDS.128 + 2 LDG.128 + 12 FMA per lane and sample.  Partial sums stay per lane over all sources and are reduced
// across the group once per plane chunk by a transposing butterfly; results leave through shared memory as
// 64-byte row segments (and, optionally, as the fp16 hi/lo cells conv0 reads by TMA).
#include <cuda_fp16.h>

#include "common.cuh"

namespace dmvs {

struct W1cParams {
  const float* ref;                // [B,C,h,w] NCHW (ref_ps == 0) or channel-last with pixel stride ref_ps; batch stride ref_bs
  const float* src[DMVS_MAX_SRC];  // [B,h,w,C] channel-last, pixel stride src_ps, batch stride src_bs
  const float* rt;
  const float* hyp;
  float* cost;   // [B,2,D,h,w] fp32, nullable
  uint2* cells;  // conv0 input cells (DMVS_FMT_COST2), nullable
  const unsigned char* flags;  // nullable, [B][tiles_y][tiles_x][D]: compute and write only the flagged (tile, plane) pairs
  long long ref_bs, src_bs;
  int src_ps, ref_ps;
  int src_cs;  // floats between the two x-corners of a footprint: src_ps for a plain channel-last map, C for the pair layout
  int B, D, h, w, n_src, d_begin, d_end, n_chunks;
  float half_w, half_h;
};

// zero padding with always-in-range addresses: load rows/cols (b, b+1), b = clamp(i0, 0, n-2), and route the two
// bilinear weights onto those two loads (same scheme as warp_corr.cu)
__device__ __forceinline__ void route_axis_c(float pos, int n, int& base, float& c0, float& c1) {
  const float f0 = floorf(pos);
  const float w1 = pos - f0;
  const float w0 = 1.0f - w1;
  const int i0 = __float2int_rd(fminf(fmaxf(f0, -4.0f), (float)n + 4.0f));
  base = min(max(i0, 0), n - 2);
  const int rel = i0 - base;
  c0 = (rel == 0) ? w0 : ((rel == -1) ? w1 : 0.0f);
  c1 = (rel == 0) ? w1 : ((rel == 1) ? w0 : 0.0f);
}

constexpr int kW1Warps = 8;  // warps per block; block tile = 32 x 8 pixels (2 x 4 patches of 16 x 2)

template <int C, int DP>
struct W1cCfg {
  static constexpr int CH = C / 4;        // 16-byte chunks per pixel
  static constexpr int G = 2 * CH;        // lanes per sample: both x corners
  static constexpr int S = 32 / G;        // pixels per warp step
  static constexpr int NSTEP = 32 / S;    // steps per 16x2 patch
  static constexpr int PB = (DP < 32 / S) ? DP : 32 / S;  // planes per record batch
  static constexpr int V = 2 * DP;        // partial sums per lane
  static constexpr int kWarpFloats = DP * 32 + V * 32 + 32 * 8;  // hyp | out | records (32 x 2 x uint4)
  static constexpr size_t kSmem = (size_t)DMVS_MAX_SRC * 12 * 4 + (size_t)kW1Warps * kWarpFloats * 4;
};

template <int C, int DP>
__global__ void __launch_bounds__(kW1Warps * 32) warp_corr_nhwc_kernel(const __grid_constant__ W1cParams p) {
  using Cfg = W1cCfg<C, DP>;
  constexpr int CH = Cfg::CH, G = Cfg::G, S = Cfg::S, NSTEP = Cfg::NSTEP, PB = Cfg::PB, V = Cfg::V;
  extern __shared__ __align__(16) float smem[];
  float* s_rt = smem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* s_hyp = smem + DMVS_MAX_SRC * 12 + warp * Cfg::kWarpFloats;
  float* s_out = s_hyp + DP * 32;
  uint4* s_rec = reinterpret_cast<uint4*>(s_out + V * 32);

  // plane chunks are the fastest-varying block index (co-resident blocks share source footprints in L2)
  const int tile_x = blockIdx.x / p.n_chunks;
  const int chunk = blockIdx.x - tile_x * p.n_chunks;
  const int b = blockIdx.z;
  const int d0 = p.d_begin + chunk * DP;
  const int X0 = tile_x * 32 + (warp & 1) * 16;
  const int Y0 = blockIdx.y * 8 + (warp >> 1) * 2;
  const int hw = p.h * p.w;

  // pass 2 of the staged variant: only the (tile, plane) pairs the staged kernel left behind
  const unsigned char* tflags = nullptr;
  if (p.flags) {
    const int tiles_x = gridDim.x / p.n_chunks;
    tflags = p.flags + (((long long)b * gridDim.y + blockIdx.y) * tiles_x + tile_x) * p.D;
    bool any = false;
    for (int j = 0; j < DP; ++j)
      if (d0 + j < p.d_end && tflags[d0 + j]) any = true;
    if (!any) return;  // block-uniform, before any barrier
  }
  for (int i = threadIdx.x; i < p.n_src * 12; i += blockDim.x) s_rt[i] = p.rt[b * p.n_src * 12 + i];

  // natural lane <-> pixel mapping for the coalesced loads / stores: 16 x-consecutive pixels in each of 2 rows
  const int nx = X0 + (lane & 15), ny = Y0 + (lane >> 4);
  const bool n_ok = nx < p.w && ny < p.h;
  {
    const int pix = min(ny, p.h - 1) * p.w + min(nx, p.w - 1);
#pragma unroll
    for (int j = 0; j < DP; ++j)
      s_hyp[j * 32 + lane] = (d0 + j < p.d_end) ? __ldg(p.hyp + ((long long)(b * p.D + d0 + j) * hw) + pix) : 1.0f;
  }
  __syncthreads();
  if (X0 >= p.w || Y0 >= p.h) return;  // whole patch outside the image (no block-level sync below)

  const int pg = lane / G, lig = lane % G;
  const int dx = lig / CH, k = lig % CH;
  const float inv_half = 2.0f / (float)C;  // mean over C/2 channels; a power of two -> exact
  const long long row_stride = (long long)p.w * p.src_ps;

#pragma unroll 1
  for (int t = 0; t < NSTEP; ++t) {
    // this group's pixel in zig-zag order: z < 16: (x = z, y = z & 1), z >= 16: (x = z - 16, y = ~x & 1)
    const int z = t * S + pg;
    const int zx = z & 15, zy = (z < 16) ? (zx & 1) : ((zx & 1) ^ 1);
    const int zn = zy * 16 + zx;
    float r[4];
    {
      const int pix = min(Y0 + zy, p.h - 1) * p.w + min(X0 + zx, p.w - 1);
      if (p.ref_ps > 0) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(p.ref + (long long)b * p.ref_bs + (long long)pix * p.ref_ps + 4 * k));
        r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
      } else {
        const float* rp = p.ref + (long long)b * p.ref_bs + (long long)(4 * k) * hw + pix;
#pragma unroll
        for (int i = 0; i < 4; ++i) r[i] = __ldg(rp + (long long)i * hw);
      }
    }
    float acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = 0.0f;

    // phase A bookkeeping: record `lane` belongs to pixel rp of this step and plane rpb of the batch
    const int a_rp = lane % S, a_pb = lane / S;
    const int az = t * S + a_rp;
    const int azx = az & 15, azy = (az < 16) ? (azx & 1) : ((azx & 1) ^ 1);
    const int azn = azy * 16 + azx;
    const float fx = (float)min(X0 + azx, p.w - 1), fy = (float)min(Y0 + azy, p.h - 1);

#pragma unroll 1
    for (int s = 0; s < p.n_src; ++s) {
      const float* m = s_rt + s * 12;
      const float* sp = p.src[s] + (long long)b * p.src_bs + dx * p.src_cs + k * 4;  // dense maps and pairs: lig * 4
      // rot @ (x, y, 1), shared by all planes of this source
      const float rx = __fadd_rn(__fmaf_rn(m[1], fy, __fmul_rn(m[0], fx)), m[2]);
      const float ry = __fadd_rn(__fmaf_rn(m[4], fy, __fmul_rn(m[3], fx)), m[5]);
      const float rz = __fadd_rn(__fmaf_rn(m[7], fy, __fmul_rn(m[6], fx)), m[8]);
#pragma unroll
      for (int j0 = 0; j0 < DP; j0 += PB) {
        // ---- phase A: one lane per (pixel, plane) computes the sample record
        if (a_pb < PB) {
          const float dep = s_hyp[(j0 + a_pb) * 32 + azn];
          const float X = __fadd_rn(__fmul_rn(rx, dep), m[9]);
          const float Y = __fadd_rn(__fmul_rn(ry, dep), m[10]);
          float Z = __fadd_rn(__fmul_rn(rz, dep), m[11]);
          if (Z == 0.0f) Z += 1e-5f;
          const float u = __fdiv_rn(X, Z), v = __fdiv_rn(Y, Z);
          const float ix = __fmul_rn(__fadd_rn(__fsub_rn(__fdiv_rn(u, p.half_w), 1.0f), 1.0f), p.half_w);
          const float iy = __fmul_rn(__fadd_rn(__fsub_rn(__fdiv_rn(v, p.half_h), 1.0f), 1.0f), p.half_h);
          int xb, yb;
          float cx0, cx1, cy0, cy1;
          route_axis_c(ix, p.w, xb, cx0, cx1);
          route_axis_c(iy, p.h, yb, cy0, cy1);
          // the reference multiplies (masked) zeros by NaN weights when the sample position is not finite
          if (!(fabsf(ix) <= 3.0e38f) || !(fabsf(iy) <= 3.0e38f)) cx0 = cx1 = cy0 = cy1 = __int_as_float(0x7fc00000);
          const int off = (yb * p.w + xb) * p.src_ps;
          uint4* rec = s_rec + (a_pb * S + a_rp) * 2;
          rec[0] = make_uint4((uint32_t)off, __float_as_uint(cx0), __float_as_uint(cy0), __float_as_uint(cy1));
          rec[1] = make_uint4((uint32_t)off, __float_as_uint(cx1), __float_as_uint(cy0), __float_as_uint(cy1));
        }
        __syncwarp();
        // ---- phase B: every lane gathers its 16 bytes of both footprint rows
#pragma unroll
        for (int pb = 0; pb < PB; ++pb) {
          const int j = j0 + pb;
          if (d0 + j < p.d_end) {
            const uint4 rec = s_rec[(pb * S + pg) * 2 + dx];
            const float* q = sp + (int)rec.x;
            const float4 v0 = __ldg(reinterpret_cast<const float4*>(q));
            const float4 v1 = __ldg(reinterpret_cast<const float4*>(q + row_stride));
            const float cx = __uint_as_float(rec.y);
            const float w0 = __uint_as_float(rec.z) * cx, w1 = __uint_as_float(rec.w) * cx;
            const float a0 = fmaf(r[2], v0.z, r[0] * v0.x), b0 = fmaf(r[3], v0.w, r[1] * v0.y);
            const float a1 = fmaf(r[2], v1.z, r[0] * v1.x), b1 = fmaf(r[3], v1.w, r[1] * v1.y);
            acc[2 * j] = fmaf(w1, a1, fmaf(w0, a0, acc[2 * j]));
            acc[2 * j + 1] = fmaf(w1, b1, fmaf(w0, b0, acc[2 * j + 1]));
          }
        }
        __syncwarp();
      }
    }

    // ---- reduce the V partial sums over the group's G lanes: transposing butterfly while more than one value is
    // left per lane (each stage halves the values a lane carries), plain butterfly afterwards
    int vbase = 0;
    bool writer = true;
#pragma unroll
    for (int st = 0; (G / 2) >> st >= 1; ++st) {
      const int mk = (G / 2) >> st;
      if ((V >> st) > 1) {
        const int len = V >> (st + 1);
        const bool up = (lig & mk) != 0;
#pragma unroll
        for (int i = 0; i < len; ++i) {
          const float send = up ? acc[i] : acc[i + len];
          const float keep = up ? acc[i + len] : acc[i];
          acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, mk);
        }
        vbase += up ? len : 0;
      } else {
        acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], mk);
        writer = writer && ((lig & mk) == 0);
      }
    }
    constexpr int kLeft = (V / G) > 1 ? (V / G) : 1;
    if (writer) {
#pragma unroll
      for (int i = 0; i < kLeft; ++i) s_out[(vbase + i) * 32 + zn] = acc[i] * inv_half;
    }
  }
  __syncwarp();

  // ---- coalesced output: lane <-> natural pixel, 2 x 64-byte row segments per (plane, group)
  if (!n_ok) return;
  const int pix = ny * p.w + nx;
#pragma unroll
  for (int j = 0; j < DP; ++j) {
    const int d = d0 + j;
    if (d >= p.d_end) break;
    if (tflags && !tflags[d]) continue;
    const float acc0 = s_out[(2 * j) * 32 + lane], acc1 = s_out[(2 * j + 1) * 32 + lane];
    if (p.cost) {
      float* cp = p.cost + ((long long)(b * 2) * p.D + d) * hw + pix;
      cp[0] = acc0;
      cp[(long long)p.D * hw] = acc1;
    }
    if (p.cells) {
      // fp16 hi/lo pairs, two voxels per 16-byte cell: this voxel is the second half of cell x and the first half of cell x+1
      const __half h0 = __float2half_rn(acc0), h1 = __float2half_rn(acc1);
      const __half2 hi = __halves2half2(h0, h1);
      const __half2 lo = __halves2half2(__float2half_rn(acc0 - __half2float(h0)), __float2half_rn(acc1 - __half2float(h1)));
      const uint2 v = make_uint2(*reinterpret_cast<const uint32_t*>(&hi), *reinterpret_cast<const uint32_t*>(&lo));
      uint2* row = p.cells + (((long long)(b * p.D + d) * p.h + ny) * (p.w + 1)) * 2;
      row[2 * nx + 1] = v;
      row[2 * nx + 2] = v;
      if (nx == 0) row[0] = make_uint2(0u, 0u);
      if (nx == p.w - 1) row[2 * p.w + 1] = make_uint2(0u, 0u);
    }
  }
}

template <int C, int DP>
static int launch_w1c(const W1cParams& p0, cudaStream_t st) {
  using Cfg = W1cCfg<C, DP>;
  W1cParams p = p0;
  p.n_chunks = ceil_div(p.d_end - p.d_begin, DP);
  static PerDevice state;  // per template instance; the opt-in is a per-device attribute
  const int slot = current_device_slot();
  if (slot < 0 || !state.configured[slot]) {
    cudaError_t e = cudaFuncSetAttribute(warp_corr_nhwc_kernel<C, DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem);
    if (e != cudaSuccess) {
      set_error("warp_corr_nhwc: cannot reserve %zu bytes of shared memory: %s", Cfg::kSmem, cudaGetErrorString(e));
      return DMVS_ERR_CUDA;
    }
    if (slot >= 0) state.configured[slot] = true;
  }
  dim3 block(kW1Warps * 32, 1, 1), grid(ceil_div(p.w, 32) * p.n_chunks, ceil_div(p.h, 8), p.B);
  DMVS_REQUIRE(grid.z <= 65535 && grid.y <= 65535, DMVS_ERR_BAD_SHAPE, "warp_corr_nhwc: grid too large (h=%d, B=%d)", p.h, p.B);
  warp_corr_nhwc_kernel<C, DP><<<grid, block, Cfg::kSmem, st>>>(p);
  return check_launch("warp_corr_nhwc");
}

template <int C>
static int dispatch_w1c(const W1cParams& p, cudaStream_t st) {
  const int nd = p.d_end - p.d_begin;
  if (nd <= 4) return launch_w1c<C, 4>(p, st);
  if (C == 32 && nd > 8) return launch_w1c<C, 16>(p, st);
  return launch_w1c<C, 8>(p, st);
}

// ------------------------------------------------------------------------------------------------ NCHW -> NHWC
// [B,C,h,w] (batch stride xbs) -> [B,h,w,C] dense.  A block transposes 32 pixels x C channels through shared memory:
// coalesced 128-byte reads per channel, 16-byte writes that together cover the pixels' contiguous C*4 bytes.
template <int C>
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ x, long long xbs, float* __restrict__ y,
                                                           long long hw) {
  __shared__ float tile[8][32 * (C + 1)];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long pix0 = ((long long)blockIdx.x * 8 + warp) * 32;
  if (pix0 >= hw) return;
  const int b = blockIdx.y;
  const float* xp = x + (long long)b * xbs;
  float* t = tile[warp];
  const long long pix = pix0 + lane;
#pragma unroll
  for (int c = 0; c < C; ++c) t[lane * (C + 1) + c] = (pix < hw) ? __ldg(xp + (long long)c * hw + pix) : 0.0f;
  __syncwarp();
  float* yp = y + ((long long)b * hw + pix0) * C;
  const int n_valid = (int)min((long long)32, hw - pix0) * C;
#pragma unroll
  for (int i = 0; i < C; i += 4) {
    // element e of the 32 x C block, 4 consecutive floats per lane
    const int e = (i / 4) * 128 + lane * 4;
    if (e < n_valid) {
      const int px = e / C, c = e % C;
      const float* q = t + px * (C + 1) + c;
      *reinterpret_cast<float4*>(yp + e) = make_float4(q[0], q[1], q[2], q[3]);
    }
  }
}

// pass 2 of dmvs_warp_corr_staged_f32 (arguments already validated there)
int launch_w1_pass2(const float* ref, long long ref_bstride, int ref_pixstride, const float* const* src, long long src_bstride,
                    int src_pixstride, int src_cornerstride, int n_src, const float* rt, const float* hyp, float* cost, void* cost_cells,
                    const unsigned char* flags, int B, int C, int D, int h, int w, int d_begin, int d_end, cudaStream_t st) {
  W1cParams p;
  p.ref = ref;
  for (int i = 0; i < DMVS_MAX_SRC; ++i) p.src[i] = (i < n_src) ? src[i] : nullptr;
  p.rt = rt; p.hyp = hyp; p.cost = cost; p.cells = reinterpret_cast<uint2*>(cost_cells); p.flags = flags;
  p.ref_bs = ref_bstride; p.src_bs = src_bstride; p.src_ps = src_pixstride; p.ref_ps = ref_pixstride;
  p.src_cs = src_cornerstride ? src_cornerstride : src_pixstride;
  p.B = B; p.D = D; p.h = h; p.w = w; p.n_src = n_src; p.d_begin = d_begin; p.d_end = d_end; p.n_chunks = 1;
  p.half_w = (float)((double)(w - 1) / 2.0);
  p.half_h = (float)((double)(h - 1) / 2.0);
  switch (C) {
    case 8: return dispatch_w1c<8>(p, st);
    case 16: return dispatch_w1c<16>(p, st);
    default: return dispatch_w1c<32>(p, st);
  }
}

}  // namespace dmvs

extern "C" int dmvs_features_nhwc_f32(const float* x, long long x_bstride, float* y, int B, int C, int h, int w, void* stream) {
  using namespace dmvs;
  DMVS_REQUIRE(x && y, DMVS_ERR_BAD_POINTER, "features_nhwc: null pointer");
  DMVS_REQUIRE(aligned16(y), DMVS_ERR_BAD_POINTER, "features_nhwc: output must be 16-byte aligned");
  DMVS_REQUIRE(B >= 1 && B <= 65535 && h >= 1 && w >= 1, DMVS_ERR_BAD_SHAPE, "features_nhwc: bad dims B=%d h=%d w=%d", B, h, w);
  const long long hw = (long long)h * w;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid((unsigned)((hw + 255) / 256), B, 1);
  switch (C) {
    case 8: nchw_to_nhwc_kernel<8><<<grid, 256, 0, st>>>(x, x_bstride, y, hw); break;
    case 16: nchw_to_nhwc_kernel<16><<<grid, 256, 0, st>>>(x, x_bstride, y, hw); break;
    case 32: nchw_to_nhwc_kernel<32><<<grid, 256, 0, st>>>(x, x_bstride, y, hw); break;
    default: set_error("features_nhwc: C=%d unsupported (8, 16, 32)", C); return DMVS_ERR_BAD_SHAPE;
  }
  return check_launch("features_nhwc");
}

extern "C" int dmvs_warp_corr_nhwc_f32(const float* ref, long long ref_bstride, int ref_pixstride, const float* const* src,
                                       long long src_bstride, int src_pixstride, int src_cornerstride, int n_src, const float* rt, const float* hyp, float* cost,
                                       void* cost_cells, int B, int C, int D, int h, int w, int d_begin, int d_end, void* stream) {
  using namespace dmvs;
  DMVS_REQUIRE(ref && src && rt && hyp && (cost || cost_cells), DMVS_ERR_BAD_POINTER, "warp_corr_nhwc: null pointer");
  DMVS_REQUIRE(!cost_cells || aligned16(cost_cells), DMVS_ERR_BAD_POINTER, "warp_corr_nhwc: cost_cells must be 16-byte aligned");
  DMVS_REQUIRE(n_src >= 1 && n_src <= DMVS_MAX_SRC, DMVS_ERR_BAD_SHAPE, "warp_corr_nhwc: n_src=%d not in [1,%d]", n_src, DMVS_MAX_SRC);
  DMVS_REQUIRE(B >= 1 && D >= 1 && h >= 2 && w >= 2, DMVS_ERR_BAD_SHAPE, "warp_corr_nhwc: bad dims B=%d D=%d h=%d w=%d", B, D, h, w);
  DMVS_REQUIRE(0 <= d_begin && d_begin <= d_end && d_end <= D, DMVS_ERR_BAD_SHAPE, "warp_corr_nhwc: bad plane range [%d,%d) of %d",
               d_begin, d_end, D);
  DMVS_REQUIRE(src_pixstride >= C && src_pixstride % 4 == 0 && src_bstride % 4 == 0, DMVS_ERR_BAD_SHAPE,
               "warp_corr_nhwc: pixel stride %d / batch stride %lld must be multiples of 4 floats and >= C", src_pixstride, src_bstride);
  DMVS_REQUIRE(src_cornerstride == 0 || (src_cornerstride >= C && src_cornerstride % 4 == 0), DMVS_ERR_BAD_SHAPE,
               "warp_corr_nhwc: corner stride %d must be 0 or a multiple of 4 floats >= C", src_cornerstride);
  DMVS_REQUIRE(ref_pixstride == 0 || (ref_pixstride >= C && ref_pixstride % 4 == 0 && ref_bstride % 4 == 0 && aligned16(ref)),
               DMVS_ERR_BAD_SHAPE, "warp_corr_nhwc: channel-last reference needs pixel stride %d >= C, strides multiples of 4 floats, 16-byte alignment",
               ref_pixstride);
  DMVS_REQUIRE((long long)src_pixstride * h * w < (1LL << 31) && (long long)C * h * w < (1LL << 31), DMVS_ERR_BAD_SHAPE,
               "warp_corr_nhwc: feature map too large for 32-bit offsets");
  if (d_begin == d_end) return DMVS_OK;
  W1cParams p;
  p.ref = ref;
  for (int i = 0; i < n_src; ++i) {
    DMVS_REQUIRE(src[i] != nullptr && aligned16(src[i]), DMVS_ERR_BAD_POINTER, "warp_corr_nhwc: src[%d] is null or not 16-byte aligned", i);
  }
  for (int i = 0; i < DMVS_MAX_SRC; ++i) p.src[i] = (i < n_src) ? src[i] : nullptr;
  p.rt = rt;
  p.hyp = hyp;
  p.cost = cost;
  p.cells = reinterpret_cast<uint2*>(cost_cells);
  p.flags = nullptr;
  p.ref_bs = ref_bstride;
  p.src_bs = src_bstride;
  p.src_ps = src_pixstride;
  p.src_cs = src_cornerstride ? src_cornerstride : src_pixstride;
  p.ref_ps = ref_pixstride;
  p.B = B; p.D = D; p.h = h; p.w = w; p.n_src = n_src; p.d_begin = d_begin; p.d_end = d_end; p.n_chunks = 1;
  p.half_w = (float)((double)(w - 1) / 2.0);
  p.half_h = (float)((double)(h - 1) / 2.0);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (C) {
    case 8: return dispatch_w1c<8>(p, st);
    case 16: return dispatch_w1c<16>(p, st);
    case 32: return dispatch_w1c<32>(p, st);
    default: set_error("warp_corr_nhwc: C=%d unsupported (8, 16, 32)", C); return DMVS_ERR_BAD_SHAPE;
  }
}
