// W1, TMA-staged variant: fused homography warp + 2-group correlation with the source footprint of a whole pixel
// tile staged in shared memory (sm_100a).
//
// Same contract as warp_corr.cu / warp_corr_nhwc.cu (reference networks/mvsnet.py:111-153 CostAgg.forward +
// networks/module.py:212-251 homo_warping), channel-last sources.
//
// Why: a bilinear gather through L1 pays one data-pipe wavefront per 128-byte LINE a quarter warp touches, so even
// the channel-last kernel spends 2.5-4 wavefronts per sample (profiles/r1k_w1c_ncu_full.csv).  From shared memory a
// wavefront is any 8 conflict-free 16-byte accesses: with the pixel's channels contiguous and the 16-byte chunks
// XOR-swizzled by the pixel index (exactly what TMA's SWIZZLE_32B/64B/128B modes write for C = 8/16/32), neighbouring
// pixels land in different bank groups and every LDS.128 of a warp is 4 full wavefronts = 1/2/4 per sample, the
// floor of the smem data path.  When the hypotheses of a 32x8 pixel tile are spatially coherent (sampler planes, or
// per-pixel depths regressed from a piecewise-smooth surface), the source positions of tile x DP planes fit a small
// box; one thread fetches it with ONE cp.async.bulk.tensor (zero fill outside the image = grid_sample's zeros padding)
// and it is reused by DP planes x 4 corners.
//
// Pass 1 (this kernel): block = 32x8 pixels x DP planes, thread = one pixel; per source: sample positions in the
// reference's op order -> block-wide bounding box (redux.sync + 32 smem words) -> box fits one of two shapes
// (wide / tall: the epipolar direction is either mostly horizontal or mostly vertical) -> TMA -> gather.  A tile
// whose box does not fit (depth discontinuities, rough hypotheses) writes nothing and raises flags[tile][plane].
// Pass 2: the channel-last gather kernel (warp_corr_nhwc.cu) recomputes exactly the flagged (tile, plane) pairs.
// Both passes are deterministic and independent of how the planes are sharded.
#include <cuda.h>
#include <cuda_fp16.h>
#include <limits.h>
#include <string.h>

#include "tc_common.cuh"

namespace dmvs {

struct alignas(64) W1sParams {
  CUtensorMap tm[2][DMVS_MAX_SRC];  // [box shape][source]
  const float* ref;
  const float* rt;
  const float* hyp;
  float* cost;
  uint2* cells;
  unsigned char* flags;  // [B][tiles_y][tiles_x][D]
  long long ref_bs;
  int ref_ps;
  int B, D, h, w, n_src, d_begin, d_end, n_chunks, chunk0, tiles_x;
  float half_w, half_h;
};

template <int C>
struct W1sBox;  // wide (BW0 x BH0) and tall (BW1 x BH1) boxes, in pixels; widths are multiples of 8 (swizzle phase = x & 7)
template <> struct W1sBox<8> { static constexpr int DP = 8, BW0 = 72, BH0 = 16, BW1 = 40, BH1 = 40; };
template <> struct W1sBox<16> { static constexpr int DP = 4, BW0 = 56, BH0 = 16, BW1 = 40, BH1 = 24; };
template <> struct W1sBox<32> { static constexpr int DP = 2, BW0 = 48, BH0 = 12, BW1 = 40, BH1 = 20; };

template <int C>
struct W1sCfg {
  using Box = W1sBox<C>;
  static constexpr int PIX0 = Box::BW0 * Box::BH0, PIX1 = Box::BW1 * Box::BH1;
  static constexpr int BOX_BYTES = (PIX0 > PIX1 ? PIX0 : PIX1) * C * 4;
  static constexpr size_t kSmem = 1024 /* alignment slack */ + BOX_BYTES + 2 * 8 * 16 /* red */ + 16 /* mbarrier */ + DMVS_MAX_SRC * 12 * 4 /* rt */;
};

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

template <int C, int DP>
__global__ void __launch_bounds__(256, (C == 32 ? 2 : 3)) warp_corr_staged_kernel(const __grid_constant__ W1sParams p) {
  using Cfg = W1sCfg<C>;
  using Box = W1sBox<C>;
  constexpr int CH = C / 4;
  extern __shared__ uint8_t w1s_raw[];
  const uint32_t raw = smem_u32(w1s_raw);
  const uint32_t box = (raw + 1023u) & ~1023u;
  uint8_t* aligned = w1s_raw + (box - raw);
  int4* s_red = reinterpret_cast<int4*>(aligned + Cfg::BOX_BYTES);        // [2][8]
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(aligned + Cfg::BOX_BYTES + 2 * 8 * 16);
  float* s_rt = reinterpret_cast<float*>(aligned + Cfg::BOX_BYTES + 2 * 8 * 16 + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_x = blockIdx.x / p.n_chunks;
  const int chunk = blockIdx.x - tile_x * p.n_chunks;
  const int b = blockIdx.z;
  // chunks are aligned to multiples of DP in absolute plane index and the bounding box is taken over ALL planes of the
  // chunk, so whether a (tile, chunk) is staged here or left to pass 2 does not depend on how the planes are sharded
  const int d0 = (p.chunk0 + chunk) * DP;
  const int hw = p.h * p.w;
  // zig-zag pixel mapping: a quarter warp = 8 same-parity pixels with consecutive x (rows alternate), so its source
  // positions advance ~1 px per lane and hit 8 different swizzle phases
  const int li = lane & 15, zig = lane >> 4;
  const int x = tile_x * 32 + (warp & 1) * 16 + li;
  const int y = blockIdx.y * 8 + (warp >> 1) * 2 + ((li & 1) ^ zig);
  const bool px_ok = x < p.w && y < p.h;
  const int xc = min(x, p.w - 1), yc = min(y, p.h - 1);
  const int pix = yc * p.w + xc;

  for (int i = threadIdx.x; i < p.n_src * 12; i += 256) s_rt[i] = __ldg(p.rt + b * p.n_src * 12 + i);
  if (threadIdx.x == 0) {
    mbar_init(s_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  float refv[C];
  if (p.ref_ps > 0) {
    const float4* rp = reinterpret_cast<const float4*>(p.ref + (long long)b * p.ref_bs + (long long)pix * p.ref_ps);
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      const float4 v = __ldg(rp + k);
      refv[4 * k] = v.x; refv[4 * k + 1] = v.y; refv[4 * k + 2] = v.z; refv[4 * k + 3] = v.w;
    }
  } else {
    const float* rp = p.ref + (long long)b * p.ref_bs + pix;
#pragma unroll
    for (int c = 0; c < C; ++c) refv[c] = __ldg(rp + (long long)c * hw);
  }
  float dep[DP];
#pragma unroll
  for (int j = 0; j < DP; ++j) dep[j] = (d0 + j < p.D) ? __ldg(p.hyp + ((long long)(b * p.D + d0 + j) * hw) + pix) : 1.0f;
  float acc[DP][2];
#pragma unroll
  for (int j = 0; j < DP; ++j) acc[j][0] = acc[j][1] = 0.0f;

  const float fx = (float)xc, fy = (float)yc;
  const float inv_half = 2.0f / (float)C;
  uint32_t phase = 0;
  bool fits = true;
  __syncthreads();

#pragma unroll 1
  for (int s = 0; s < p.n_src; ++s) {
    const float* m = s_rt + s * 12;
    const float rx = __fadd_rn(__fmaf_rn(m[1], fy, __fmul_rn(m[0], fx)), m[2]);
    const float ry = __fadd_rn(__fmaf_rn(m[4], fy, __fmul_rn(m[3], fx)), m[5]);
    const float rz = __fadd_rn(__fmaf_rn(m[7], fy, __fmul_rn(m[6], fx)), m[8]);
    const float tx = m[9], ty = m[10], tz = m[11];
    // ---- sample positions of this pixel's DP planes (module.py:233-241 + ATen's un-normalisation, same op order)
    int sx[DP], sy[DP];   // integer corner (x0, y0); sx = INT_MIN: no gather (contributes 0, or NaN when wx is NaN)
    float wx[DP], wy[DP];
    int mnx = INT_MAX, mxx = INT_MIN, mny = INT_MAX, mxy = INT_MIN;
#pragma unroll
    for (int j = 0; j < DP; ++j) {
      const float X = __fadd_rn(__fmul_rn(rx, dep[j]), tx);
      const float Y = __fadd_rn(__fmul_rn(ry, dep[j]), ty);
      float Z = __fadd_rn(__fmul_rn(rz, dep[j]), tz);
      if (Z == 0.0f) Z += 1e-5f;
      const float u = __fdiv_rn(X, Z), v = __fdiv_rn(Y, Z);
      const float ix = __fmul_rn(__fadd_rn(__fsub_rn(__fdiv_rn(u, p.half_w), 1.0f), 1.0f), p.half_w);
      const float iy = __fmul_rn(__fadd_rn(__fsub_rn(__fdiv_rn(v, p.half_h), 1.0f), 1.0f), p.half_h);
      const float f0x = floorf(ix), f0y = floorf(iy);
      const bool finite = (fabsf(ix) <= 3.0e38f) && (fabsf(iy) <= 3.0e38f);
      wx[j] = finite ? ix - f0x : __int_as_float(0x7fc00000);
      wy[j] = iy - f0y;
      // a footprint with at least one corner inside the image: x0 in [-1, w-1], y0 in [-1, h-1]
      const bool inside = finite && f0x >= -1.0f && f0x <= (float)(p.w - 1) && f0y >= -1.0f && f0y <= (float)(p.h - 1);
      const bool wanted = px_ok && d0 + j >= p.d_begin && d0 + j < p.d_end;   // this call writes the plane
      if (inside && px_ok && d0 + j < p.D) {
        mnx = min(mnx, (int)f0x); mxx = max(mxx, (int)f0x);
        mny = min(mny, (int)f0y); mxy = max(mxy, (int)f0y);
      }
      sx[j] = (inside && wanted) ? (int)f0x : INT_MIN;
      sy[j] = (inside && wanted) ? (int)f0y : 0;
      if (!wanted) wx[j] = 0.0f;  // dead lanes never produce NaN
    }
    // ---- block-wide bounding box
    mnx = __reduce_min_sync(0xffffffffu, mnx); mxx = __reduce_max_sync(0xffffffffu, mxx);
    mny = __reduce_min_sync(0xffffffffu, mny); mxy = __reduce_max_sync(0xffffffffu, mxy);
    int4* red = s_red + (s & 1) * 8;
    if (lane == 0) red[warp] = make_int4(mnx, mxx, mny, mxy);
    __syncthreads();  // also: every thread has finished gathering from the box of the previous source
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int4 r = red[i];
      mnx = min(mnx, r.x); mxx = max(mxx, r.y); mny = min(mny, r.z); mxy = max(mxy, r.w);
    }
    int bw = Box::BW0;
    if (mxx >= mnx) {  // at least one live sample in the block
      const int needw = mxx - mnx + 2, needh = mxy - mny + 2;
      int shape;
      if (needw <= Box::BW0 && needh <= Box::BH0) shape = 0;
      else if (needw <= Box::BW1 && needh <= Box::BH1) shape = 1;
      else { fits = false; break; }
      bw = shape ? Box::BW1 : Box::BW0;
      if (threadIdx.x == 0) {
        mbar_expect_tx(s_bar, (uint32_t)((shape ? Cfg::PIX1 : Cfg::PIX0) * C * 4));
        tma_load_4d(aligned, &p.tm[shape][s], s_bar, 0, mnx, mny, b);
      }
      mbar_wait(s_bar, phase);
      phase ^= 1u;
    }
    // ---- gather: 4 corners x C/4 conflict-free LDS.128 per sample
#pragma unroll
    for (int j = 0; j < DP; ++j) {
      float g0, g1;
      if (sx[j] != INT_MIN) {
        const int p00 = (sy[j] - mny) * bw + (sx[j] - mnx);
        float s00[2] = {0.f, 0.f}, s01[2] = {0.f, 0.f}, s10[2] = {0.f, 0.f}, s11[2] = {0.f, 0.f};
        const int pc[4] = {p00, p00 + 1, p00 + bw, p00 + bw + 1};
        uint32_t a[4], z[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          a[q] = box + (uint32_t)pc[q] * (C * 4);
          z[q] = (C == 32) ? (pc[q] & 7) : (C == 16) ? ((pc[q] >> 1) & 3) : ((pc[q] >> 2) & 1);
        }
#pragma unroll
        for (int k = 0; k < CH; ++k) {
          const float4 v00 = lds128(a[0] + (((uint32_t)k ^ z[0]) << 4));
          const float4 v01 = lds128(a[1] + (((uint32_t)k ^ z[1]) << 4));
          const float4 v10 = lds128(a[2] + (((uint32_t)k ^ z[2]) << 4));
          const float4 v11 = lds128(a[3] + (((uint32_t)k ^ z[3]) << 4));
          const float r0 = refv[4 * k], r1 = refv[4 * k + 1], r2 = refv[4 * k + 2], r3 = refv[4 * k + 3];
          s00[0] = fmaf(r2, v00.z, fmaf(r0, v00.x, s00[0])); s00[1] = fmaf(r3, v00.w, fmaf(r1, v00.y, s00[1]));
          s01[0] = fmaf(r2, v01.z, fmaf(r0, v01.x, s01[0])); s01[1] = fmaf(r3, v01.w, fmaf(r1, v01.y, s01[1]));
          s10[0] = fmaf(r2, v10.z, fmaf(r0, v10.x, s10[0])); s10[1] = fmaf(r3, v10.w, fmaf(r1, v10.y, s10[1]));
          s11[0] = fmaf(r2, v11.z, fmaf(r0, v11.x, s11[0])); s11[1] = fmaf(r3, v11.w, fmaf(r1, v11.y, s11[1]));
        }
        const float cx1 = wx[j], cx0 = 1.0f - cx1, cy1 = wy[j], cy0 = 1.0f - cy1;
        const float w00 = cy0 * cx0, w01 = cy0 * cx1, w10 = cy1 * cx0, w11 = cy1 * cx1;
        g0 = w00 * s00[0] + w01 * s01[0] + w10 * s10[0] + w11 * s11[0];
        g1 = w00 * s00[1] + w01 * s01[1] + w10 * s10[1] + w11 * s11[1];
      } else {
        g0 = g1 = (wx[j] != wx[j]) ? wx[j] : 0.0f;  // non-finite position: NaN like the reference; outside: zeros padding
      }
      acc[j][0] += g0 * inv_half;
      acc[j][1] += g1 * inv_half;
    }
  }

  // ---- flags: 1 = this (tile, plane) is left to pass 2
  if (p.flags && threadIdx.x < DP && d0 + threadIdx.x >= p.d_begin && d0 + threadIdx.x < p.d_end)
    p.flags[(((long long)b * gridDim.y + blockIdx.y) * p.tiles_x + tile_x) * p.D + d0 + threadIdx.x] = fits ? 0 : 1;
  if (!fits || !px_ok) return;
  const int opix = y * p.w + x;
#pragma unroll
  for (int j = 0; j < DP; ++j) {
    const int d = d0 + j;
    if (d < p.d_begin || d >= p.d_end) continue;
    const float acc0 = acc[j][0], acc1 = acc[j][1];
    if (p.cost) {
      float* cp = p.cost + ((long long)(b * 2) * p.D + d) * hw + opix;
      cp[0] = acc0;
      cp[(long long)p.D * hw] = acc1;
    }
    if (p.cells) {
      const __half h0 = __float2half_rn(acc0), h1 = __float2half_rn(acc1);
      const __half2 hi = __halves2half2(h0, h1);
      const __half2 lo = __halves2half2(__float2half_rn(acc0 - __half2float(h0)), __float2half_rn(acc1 - __half2float(h1)));
      const uint2 v = make_uint2(*reinterpret_cast<const uint32_t*>(&hi), *reinterpret_cast<const uint32_t*>(&lo));
      uint2* row = p.cells + (((long long)(b * p.D + d) * p.h + y) * (p.w + 1)) * 2;
      row[2 * x + 1] = v;
      row[2 * x + 2] = v;
      if (x == 0) row[0] = make_uint2(0u, 0u);
      if (x == p.w - 1) row[2 * p.w + 1] = make_uint2(0u, 0u);
    }
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFnW1)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFnW1 encode_fn_w1() {
  static EncodeTiledFnW1 fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFnW1>(f);
  }
  return fn;
}

int launch_w1_pass2(const float* ref, long long ref_bstride, int ref_pixstride, const float* const* src, long long src_bstride,
                    int src_pixstride, int src_cornerstride, int n_src, const float* rt, const float* hyp, float* cost, void* cost_cells,
                    const unsigned char* flags, int B, int C, int D, int h, int w, int d_begin, int d_end, cudaStream_t st);

template <int C, int DP>
static int launch_w1s_dp(W1sParams& p, cudaStream_t st) {
  using Cfg = W1sCfg<C>;
  static PerDevice state;  // per template instance; the opt-in is a per-device attribute
  const int slot = current_device_slot();
  if (slot < 0 || !state.configured[slot]) {
    cudaError_t e = cudaFuncSetAttribute(warp_corr_staged_kernel<C, DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem);
    if (e != cudaSuccess) {
      set_error("warp_corr_staged: cannot reserve %zu bytes of shared memory: %s", Cfg::kSmem, cudaGetErrorString(e));
      return DMVS_ERR_CUDA;
    }
    if (slot >= 0) state.configured[slot] = true;
  }
  p.chunk0 = p.d_begin / DP;
  p.n_chunks = ceil_div(p.d_end, DP) - p.chunk0;
  p.tiles_x = ceil_div(p.w, 32);
  dim3 grid(p.tiles_x * p.n_chunks, ceil_div(p.h, 8), p.B);
  DMVS_REQUIRE(grid.z <= 65535 && grid.y <= 65535, DMVS_ERR_BAD_SHAPE, "warp_corr_staged: grid too large (h=%d, B=%d)", p.h, p.B);
  warp_corr_staged_kernel<C, DP><<<grid, 256, Cfg::kSmem, st>>>(p);
  return check_launch("warp_corr_staged");
}

template <int C>
static int launch_w1s(W1sParams& p, const float* const* src, long long src_bs, int src_ps, cudaStream_t st) {
  using Box = W1sBox<C>;
  EncodeTiledFnW1 enc = encode_fn_w1();
  DMVS_REQUIRE(enc != nullptr, DMVS_ERR_CUDA, "warp_corr_staged: cuTensorMapEncodeTiled is not available from the driver");
  const CUtensorMapSwizzle swz = (C == 32) ? CU_TENSOR_MAP_SWIZZLE_128B : (C == 16) ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  for (int s = 0; s < p.n_src; ++s) {
    for (int shape = 0; shape < 2; ++shape) {
      const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)p.w, (cuuint64_t)p.h, (cuuint64_t)p.B};
      const cuuint64_t strides[3] = {(cuuint64_t)src_ps * 4, (cuuint64_t)p.w * src_ps * 4, (cuuint64_t)src_bs * 4};
      const cuuint32_t bx[4] = {(cuuint32_t)C, (cuuint32_t)(shape ? Box::BW1 : Box::BW0), (cuuint32_t)(shape ? Box::BH1 : Box::BH0), 1};
      const cuuint32_t es[4] = {1, 1, 1, 1};
      CUresult r = enc(&p.tm[shape][s], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(src[s]), dims, strides, bx, es,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      DMVS_REQUIRE(r == CUDA_SUCCESS, DMVS_ERR_CUDA, "warp_corr_staged: cuTensorMapEncodeTiled failed (%d) for C=%d h=%d w=%d stride=%d", (int)r,
                   C, p.h, p.w, src_ps);
    }
  }
  // planes per thread: the box shapes are sized for Box::DP planes; volumes with few planes (the D = 4 refine passes) use
  // a smaller chunk so no thread computes positions for planes that do not exist.  The choice depends on D only.
  if (p.D <= 4 && Box::DP > 4) return launch_w1s_dp<C, (Box::DP > 4 ? 4 : Box::DP)>(p, st);
  return launch_w1s_dp<C, Box::DP>(p, st);
}

}  // namespace dmvs

extern "C" size_t dmvs_warp_corr_flag_bytes(int B, int D, int h, int w) {
  if (B < 1 || D < 1 || h < 1 || w < 1) return 0;
  return (size_t)B * (size_t)((h + 7) / 8) * (size_t)((w + 31) / 32) * (size_t)D;
}

extern "C" int dmvs_warp_corr_staged_f32(const float* ref, long long ref_bstride, int ref_pixstride, const float* const* src,
                                         long long src_bstride, int src_pixstride, int src_cornerstride, int n_src, const float* rt, const float* hyp,
                                         float* cost, void* cost_cells, void* flags, int B, int C, int D, int h, int w, int d_begin,
                                         int d_end, void* stream) {
  using namespace dmvs;
  DMVS_REQUIRE(ref && src && rt && hyp && flags && (cost || cost_cells), DMVS_ERR_BAD_POINTER, "warp_corr_staged: null pointer");
  DMVS_REQUIRE(!cost_cells || aligned16(cost_cells), DMVS_ERR_BAD_POINTER, "warp_corr_staged: cost_cells must be 16-byte aligned");
  DMVS_REQUIRE(n_src >= 1 && n_src <= DMVS_MAX_SRC, DMVS_ERR_BAD_SHAPE, "warp_corr_staged: n_src=%d not in [1,%d]", n_src, DMVS_MAX_SRC);
  DMVS_REQUIRE(B >= 1 && D >= 1 && h >= 2 && w >= 2, DMVS_ERR_BAD_SHAPE, "warp_corr_staged: bad dims B=%d D=%d h=%d w=%d", B, D, h, w);
  DMVS_REQUIRE(0 <= d_begin && d_begin <= d_end && d_end <= D, DMVS_ERR_BAD_SHAPE, "warp_corr_staged: bad plane range [%d,%d) of %d",
               d_begin, d_end, D);
  DMVS_REQUIRE(src_pixstride >= C && src_pixstride % 4 == 0 && src_bstride % 4 == 0, DMVS_ERR_BAD_SHAPE,
               "warp_corr_staged: pixel stride %d / batch stride %lld must be multiples of 4 floats and >= C", src_pixstride, src_bstride);
  DMVS_REQUIRE(ref_pixstride == 0 || (ref_pixstride >= C && ref_pixstride % 4 == 0 && ref_bstride % 4 == 0 && aligned16(ref)),
               DMVS_ERR_BAD_SHAPE, "warp_corr_staged: bad channel-last reference stride %d", ref_pixstride);
  DMVS_REQUIRE((long long)src_pixstride * h * w < (1LL << 31) && (long long)C * h * w < (1LL << 31), DMVS_ERR_BAD_SHAPE,
               "warp_corr_staged: feature map too large for 32-bit offsets");
  for (int i = 0; i < n_src; ++i)
    DMVS_REQUIRE(src[i] != nullptr && aligned16(src[i]), DMVS_ERR_BAD_POINTER, "warp_corr_staged: src[%d] is null or not 16-byte aligned", i);
  if (d_begin == d_end) return DMVS_OK;
  W1sParams p;
  memset(&p, 0, sizeof(p));
  p.ref = ref; p.rt = rt; p.hyp = hyp; p.cost = cost; p.cells = reinterpret_cast<uint2*>(cost_cells);
  p.flags = static_cast<unsigned char*>(flags);
  p.ref_bs = ref_bstride; p.ref_ps = ref_pixstride;
  p.B = B; p.D = D; p.h = h; p.w = w; p.n_src = n_src; p.d_begin = d_begin; p.d_end = d_end;
  p.half_w = (float)((double)(w - 1) / 2.0);
  p.half_h = (float)((double)(h - 1) / 2.0);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc;
  switch (C) {
    case 8: rc = launch_w1s<8>(p, src, src_bstride, src_pixstride, st); break;
    case 16: rc = launch_w1s<16>(p, src, src_bstride, src_pixstride, st); break;
    case 32: rc = launch_w1s<32>(p, src, src_bstride, src_pixstride, st); break;
    default: set_error("warp_corr_staged: C=%d unsupported (8, 16, 32)", C); return DMVS_ERR_BAD_SHAPE;
  }
  if (rc != DMVS_OK) return rc;
  return launch_w1_pass2(ref, ref_bstride, ref_pixstride, src, src_bstride, src_pixstride, src_cornerstride, n_src, rt, hyp, cost, cost_cells,
                         static_cast<const unsigned char*>(flags), B, C, D, h, w, d_begin, d_end, st);
}
