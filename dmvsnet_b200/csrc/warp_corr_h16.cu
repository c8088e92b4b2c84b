// W1, half-precision staged variant: fused homography warp + 2-group correlation with the SOURCE feature maps stored as
// fp16 channel-last maps and staged in shared memory by TMA (sm_100a).
//
// Same contract as the other W1 kernels (reference networks/mvsnet.py:111-153 CostAgg.forward + networks/module.py:212-251
// homo_warping), except that the source features have been rounded to fp16 once (dmvs_features_nhwc_f16 / FeatureNet's
// epilogue); the reference view, the bilinear weights, the products and the sums stay fp32.
//
// Why: a fused bilinear gather is bound by the SMs' L1 / shared-memory data pipe (128 B per clock and SM), not by HBM
// (profiles/README.md): every sample pulls 4 corners x C values through it.  fp16 sources halve those bytes; staging the
// footprint of a whole pixel tile removes the 128-byte-line granularity of a global gather (a quarter warp reads any 8
// conflict-free 16-byte chunks per clock); and what the round-1 staged kernel paid beside the gather is gone:
//   * the box of every source is known BEFORE any sample position is: positions are coordinate-wise monotone in
//     (x, y, depth), so the footprint of tile x plane chunk lies in the bounding box of its 8 corner projections
//     (4 tile corners x {min, max} depth).  One warp works the boxes of all sources out in the prologue and all TMA loads are
//     in flight at once (ring of NBUF slots) instead of position -> block reduction -> TMA -> wait, source after source;
//   * no flag pass: a source whose box does not fit (rough hypotheses, depth discontinuity, wide baseline) - or a single
//     sample outside its box - is gathered straight from global memory by the same thread with the same arithmetic, so the
//     result does not depend on which path a sample took;
//   * one reciprocal + Newton step instead of four IEEE divisions per sample (positions within ~1.5 ulp of the
//     reference's; the normalise / un-normalise round trip of module.py:240-241 is kept, with multiplications).
//
// Block = 16x16 pixels x DP planes, thread = one pixel (zig-zag lane order: a quarter warp = 8 same-parity pixels with
// consecutive x, so its source positions advance ~1 px per lane and - box widths being multiples of 8 and the 16-byte chunks
// XOR-swizzled by TMA - hit 8 different bank groups).
#include <cuda.h>
#include <cuda_fp16.h>
#include <limits.h>
#include <string.h>

#include "tc_common.cuh"

namespace dmvs {

struct alignas(64) W1hParams {
  CUtensorMap tm[2][DMVS_MAX_SRC];  // [box shape][source]
  const __half* src[DMVS_MAX_SRC];  // the same maps for the direct path
  const float* ref;
  const float* rt;
  const float* hyp;
  float* cost;
  uint2* cells;
  long long ref_bs, src_bs;
  int ref_ps, src_ps;
  int B, D, h, w, n_src, d_begin, d_end, n_chunks, chunk0, tiles_x;
  int row0;  // absolute image row of row 0 of the reference band (ref / hyp / cost hold rows [row0, row0 + h) of the view)
  int hs;    // rows of the (full) source maps
};

// box shapes in pixels: wide (BW0 x BH0) and tall (BW1 x BH1); widths are multiples of 8 (bank phase = pixel index & 7)
template <int C> struct W1hBox;
// sized for ~2 px of disparity per plane (DTU-like 100 mm baselines) + the +-1 plane checkerboard of the sampler; a source whose
// footprint is larger (wide baseline, depth discontinuity in the tile) takes the direct path
template <> struct W1hBox<8> { static constexpr int DP = 8, BW0 = 40, BH0 = 24, BW1 = 24, BH1 = 40, NBUF = 4; };   // 15 KB per box
template <> struct W1hBox<16> { static constexpr int DP = 8, BW0 = 40, BH0 = 24, BW1 = 24, BH1 = 40, NBUF = 2; };  // 30 KB
template <> struct W1hBox<32> { static constexpr int DP = 4, BW0 = 40, BH0 = 20, BW1 = 24, BH1 = 32, NBUF = 2; };  // 50 KB

template <int C>
struct W1hCfg {
  using Box = W1hBox<C>;
  static constexpr int PIX0 = Box::BW0 * Box::BH0, PIX1 = Box::BW1 * Box::BH1;
  static constexpr int BOX_BYTES = (((PIX0 > PIX1 ? PIX0 : PIX1) * C * 2) + 1023) / 1024 * 1024;
  static constexpr int kTail = 8 * 16 /* red */ + DMVS_MAX_SRC * 16 /* boxes */ + Box::NBUF * 8 /* mbarriers */ + DMVS_MAX_SRC * 12 * 4 /* rt */;
  static constexpr size_t kSmem = 1024 /* alignment slack */ + (size_t)Box::NBUF * BOX_BYTES + kTail;
};

__device__ __forceinline__ uint4 lds128u(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

__device__ __forceinline__ float rcp_newton(float z) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(z));
  const float e = fmaf(-z, r, 1.0f);
  return fmaf(r, e, r);
}

// sample position of (rot @ (x,y,1)) * depth + trans (module.py:233-239): one fused multiply-add per coordinate and a refined
// reciprocal instead of the reference's mul / add / IEEE division.  The normalise / un-normalise round trip of module.py:240-241
// + ATen's grid_sampler is the identity up to ~1e-4 px and is skipped; the fp16 rounding of the sources
// is the larger term of this kernel's error budget.
struct W1hPos { float ix, iy; };
__device__ __forceinline__ W1hPos w1h_position(float rx, float ry, float rz, float tx, float ty, float tz, float dep) {
  const float X = fmaf(rx, dep, tx);
  const float Y = fmaf(ry, dep, ty);
  float Z = fmaf(rz, dep, tz);
  if (Z == 0.0f) Z = 1e-5f;
  const float iz = rcp_newton(Z);
  W1hPos o;
  o.ix = X * iz;
  o.iy = Y * iz;
  return o;
}

// 8 fp16 channels (one 16-byte chunk) against 8 reference channels: even channels -> group 0, odd -> group 1.
// FIRST: the sums start here (no zero initialisation)
template <bool FIRST>
__device__ __forceinline__ void dot8(const uint4& v, const float* r, float& s0, float& s1) {
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
  const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&v.z));
  const float2 d = __half22float2(*reinterpret_cast<const __half2*>(&v.w));
  const float i0 = FIRST ? r[0] * a.x : fmaf(r[0], a.x, s0);
  const float i1 = FIRST ? r[1] * a.y : fmaf(r[1], a.y, s1);
  s0 = fmaf(r[6], d.x, fmaf(r[4], c.x, fmaf(r[2], b.x, i0)));
  s1 = fmaf(r[7], d.y, fmaf(r[5], c.y, fmaf(r[3], b.y, i1)));
}

// one sample whose 2x2 footprint lies inside the staged box: a00 = shared-memory address of corner (y0, x0) BEFORE swizzling,
// row_bytes = box width in bytes.  TMA's swizzle XORs the 16-byte chunk index with address bits [8:7] (64B mode, C = 32) or
// bit 7 (32B mode, C = 16); folding that into the corner address makes chunk k of a pixel `addr ^ (k << 4)`.
template <int C>
__device__ __forceinline__ void gather_staged(uint32_t a00, uint32_t row_bytes, float cx1, float cy1, const float* refv, float& g0, float& g1) {
  constexpr int CH = C / 8;
  uint32_t a[4] = {a00, a00 + C * 2, a00 + row_bytes, a00 + row_bytes + C * 2};
  if (C == 32) {
#pragma unroll
    for (int c = 0; c < 4; ++c) a[c] ^= (a[c] >> 3) & 0x30u;
  } else if (C == 16) {
#pragma unroll
    for (int c = 0; c < 4; ++c) a[c] ^= (a[c] >> 3) & 0x10u;
  }
  float s[4][2];
#pragma unroll
  for (int k = 0; k < CH; ++k) {
    uint4 v[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) v[c] = lds128u(a[c] ^ ((uint32_t)k << 4));
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (k == 0) dot8<true>(v[c], refv, s[c][0], s[c][1]);
      else dot8<false>(v[c], refv + 8 * k, s[c][0], s[c][1]);
    }
  }
  const float cx0 = 1.0f - cx1, cy0 = 1.0f - cy1;
  const float w00 = cy0 * cx0, w01 = cy0 * cx1, w10 = cy1 * cx0, w11 = cy1 * cx1;
  g0 = fmaf(w11, s[3][0], fmaf(w10, s[2][0], fmaf(w01, s[1][0], w00 * s[0][0])));
  g1 = fmaf(w11, s[3][1], fmaf(w10, s[2][1], fmaf(w01, s[1][1], w00 * s[0][1])));
}

template <int C, int DP>
__global__ void __launch_bounds__(256, (C == 8 ? 3 : 2)) warp_corr_h16_kernel(const __grid_constant__ W1hParams p) {
  using Cfg = W1hCfg<C>;
  using Box = W1hBox<C>;
  constexpr int CH = C / 8;  // 16-byte chunks per pixel
  constexpr int NBUF = Box::NBUF;
  extern __shared__ uint8_t w1h_raw[];
  const uint32_t raw = smem_u32(w1h_raw);
  const uint32_t box0 = (raw + 1023u) & ~1023u;
  uint8_t* aligned = w1h_raw + (box0 - raw);
  uint8_t* tail = aligned + NBUF * Cfg::BOX_BYTES;
  int4* s_red = reinterpret_cast<int4*>(tail);                     // [8]
  int4* s_box = reinterpret_cast<int4*>(tail + 8 * 16);            // [n_src]: {ox, oy, shape (-1: direct), 0}
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(tail + 8 * 16 + DMVS_MAX_SRC * 16);
  float* s_rt = reinterpret_cast<float*>(tail + 8 * 16 + DMVS_MAX_SRC * 16 + NBUF * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_x = blockIdx.x / p.n_chunks;
  const int chunk = blockIdx.x - tile_x * p.n_chunks;
  const int b = blockIdx.z;
  const int d0 = (p.chunk0 + chunk) * DP;  // chunks are aligned in absolute plane index: the box of a (tile, chunk) does not depend on the shard
  const int hw = p.h * p.w;
  const int X0 = tile_x * 16, Y0 = blockIdx.y * 16;
  const int li = lane & 15, zig = lane >> 4;
  const int x = X0 + li;
  const int y = Y0 + warp * 2 + ((li & 1) ^ zig);
  const bool px_ok = x < p.w && y < p.h;
  const int xc = min(x, p.w - 1), yc = min(y, p.h - 1);
  const int pix = yc * p.w + xc;

  for (int i = threadIdx.x; i < p.n_src * 12; i += 256) s_rt[i] = __ldg(p.rt + b * p.n_src * 12 + i);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NBUF; ++i) mbar_init(s_bar + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  float dep[DP];
  int dlo = INT_MAX, dhi = INT_MIN;  // positive finite floats order like their bit patterns
  bool range_ok = true;
#pragma unroll
  for (int j = 0; j < DP; ++j) {
    // planes past the volume repeat the chunk's first plane: never stored, but inside the box like every real plane
    dep[j] = __ldg(p.hyp + ((long long)(b * p.D + min(d0 + j, p.D - 1)) * hw) + pix);
    if (d0 + j < p.D) {
      range_ok = range_ok && (dep[j] > 0.0f) && (dep[j] <= 3.0e38f);
      dlo = min(dlo, __float_as_int(dep[j]));
      dhi = max(dhi, __float_as_int(dep[j]));
    }
  }
  dlo = __reduce_min_sync(0xffffffffu, dlo);
  dhi = __reduce_max_sync(0xffffffffu, dhi);
  const int bad = __reduce_max_sync(0xffffffffu, range_ok ? 0 : 1);
  if (lane == 0) s_red[warp] = make_int4(dlo, dhi, bad, 0);

  float refv[C];
  if (p.ref_ps > 0) {
    const float4* rp = reinterpret_cast<const float4*>(p.ref + (long long)b * p.ref_bs + (long long)pix * p.ref_ps);
#pragma unroll
    for (int k = 0; k < C / 4; ++k) {
      const float4 v = __ldg(rp + k);
      refv[4 * k] = v.x; refv[4 * k + 1] = v.y; refv[4 * k + 2] = v.z; refv[4 * k + 3] = v.w;
    }
  } else {
    const float* rp = p.ref + (long long)b * p.ref_bs + pix;
#pragma unroll
    for (int c = 0; c < C; ++c) refv[c] = __ldg(rp + (long long)c * hw);
  }
  __syncthreads();

  // ---- prologue: the boxes of all sources (warp 0: lane = source-in-group x 8 vertices of the (x, y, depth) box)
  if (warp == 0) {
    int lo = INT_MAX, hi = INT_MIN, anybad = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int4 r = s_red[i];
      lo = min(lo, r.x); hi = max(hi, r.y); anybad |= r.z;
    }
    const float dmin = __int_as_float(lo), dmax = __int_as_float(hi);
    const int vtx = lane & 7;
    const float vx = (float)((vtx & 1) ? min(X0 + 15, p.w - 1) : X0);
    const float vy = (float)(((vtx & 2) ? min(Y0 + 15, p.h - 1) : Y0) + p.row0);
    const float vd = (vtx & 4) ? dmax : dmin;
    for (int sg = 0; sg < p.n_src; sg += 4) {
      const int s = sg + (lane >> 3);
      int mnx = INT_MAX, mxx = INT_MIN, mny = INT_MAX, mxy = INT_MIN, ok = 0;
      if (s < p.n_src) {
        const float* m = s_rt + s * 12;
        const float rx = __fadd_rn(__fmaf_rn(m[1], vy, __fmul_rn(m[0], vx)), m[2]);
        const float ry = __fadd_rn(__fmaf_rn(m[4], vy, __fmul_rn(m[3], vx)), m[5]);
        const float rz = __fadd_rn(__fmaf_rn(m[7], vy, __fmul_rn(m[6], vx)), m[8]);
        const float Z = __fadd_rn(__fmul_rn(rz, vd), m[11]);
        const W1hPos q = w1h_position(rx, ry, rz, m[9], m[10], m[11], vd);
        // monotonicity needs the denominator to keep its sign over the whole box: all 8 vertices in front of the camera
        ok = (Z > 1e-3f) && (fabsf(q.ix) < 1.0e6f) && (fabsf(q.iy) < 1.0e6f) && !anybad;
        if (ok) {
          mnx = mxx = (int)floorf(q.ix);
          mny = mxy = (int)floorf(q.iy);
        }
      }
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, o)); mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
        ok &= __shfl_xor_sync(0xffffffffu, ok, o);
      }
      if (vtx == 0 && s < p.n_src) {
        int shape = -1;
        // footprints with at least one corner inside the image have x0 in [-1, w-1], y0 in [-1, h-1]; one pixel of slack
        mnx = max(mnx - 1, -1); mxx = min(mxx + 1, p.w - 1);
        mny = max(mny - 1, -1); mxy = min(mxy + 1, p.hs - 1);
        if (ok && mxx >= mnx && mxy >= mny) {
          const int needw = mxx - mnx + 2, needh = mxy - mny + 2;
          if (needw <= Box::BW0 && needh <= Box::BH0) shape = 0;
          else if (needw <= Box::BW1 && needh <= Box::BH1) shape = 1;
        }
        s_box[s] = make_int4(mnx, mny, shape, 0);
      }
    }
    __syncwarp();
    if (lane == 0) {
      for (int s = 0; s < p.n_src && s < NBUF; ++s) {
        const int4 bx = s_box[s];
        if (bx.z >= 0) {
          mbar_expect_tx(s_bar + s, (uint32_t)((bx.z ? Cfg::PIX1 : Cfg::PIX0) * C * 2));
          tma_load_4d(aligned + s * Cfg::BOX_BYTES, &p.tm[bx.z][s], s_bar + s, 0, bx.x, bx.y, b);
        }
      }
    }
  }
  __syncthreads();

  float acc[DP][2];
#pragma unroll
  for (int j = 0; j < DP; ++j) acc[j][0] = acc[j][1] = 0.0f;
  const float fx = (float)xc, fy = (float)(yc + p.row0);
  const float inv_half = 2.0f / (float)C;
  uint32_t phases = 0;  // bit i: parity to wait for on slot i

#pragma unroll 1
  for (int s = 0; s < p.n_src; ++s) {
    const float* m = s_rt + s * 12;
    const float rx = __fadd_rn(__fmaf_rn(m[1], fy, __fmul_rn(m[0], fx)), m[2]);
    const float ry = __fadd_rn(__fmaf_rn(m[4], fy, __fmul_rn(m[3], fx)), m[5]);
    const float rz = __fadd_rn(__fmaf_rn(m[7], fy, __fmul_rn(m[6], fx)), m[8]);
    const float tx = m[9], ty = m[10], tz = m[11];
    const int4 bx = s_box[s];
    const int slot = s % NBUF;
    const int bw = bx.z ? Box::BW1 : Box::BW0, bh = bx.z ? Box::BH1 : Box::BH0;
    const uint32_t box = box0 + slot * Cfg::BOX_BYTES;
    if (bx.z >= 0) {
      mbar_wait(s_bar + slot, (phases >> slot) & 1u);
      phases ^= 1u << slot;
    }
    const __half* gsrc = p.src[s] + (long long)b * p.src_bs;
    // ---- positions of this pixel's DP planes; does every sample of the warp have its footprint inside the staged box?
    float pix_x[DP], pix_y[DP];
    const float bx_lo = (float)bx.x, bx_hi = (float)(bx.x + bw - 1), by_lo = (float)bx.y, by_hi = (float)(bx.y + bh - 1);
    bool in_box = bx.z >= 0;
#pragma unroll
    for (int j = 0; j < DP; ++j) {
      const W1hPos q = w1h_position(rx, ry, rz, tx, ty, tz, dep[j]);
      pix_x[j] = q.ix;
      pix_y[j] = q.iy;
      in_box = in_box && (q.ix >= bx_lo) && (q.ix < bx_hi) && (q.iy >= by_lo) && (q.iy < by_hi);  // false for NaN
    }
    if (__all_sync(0xffffffffu, in_box)) {
      // ---- fast path (warp-uniform): no per-sample tests; TMA's zero fill is grid_sample's zeros padding.  Planes outside
      // [d_begin, d_end) and pixels outside the image are computed and dropped at the store.
      const uint32_t row_bytes = (uint32_t)bw * (C * 2);
      const float boxf = (float)bx.x + (float)bx.y * (float)bw;  // pixel index of the box origin in box-row units (exact)
#pragma unroll
      for (int j = 0; j < DP; ++j) {
        const float f0x = floorf(pix_x[j]), f0y = floorf(pix_y[j]);
        const int pc = __float2int_rn(fmaf(f0y, (float)bw, f0x) - boxf);
        float g0, g1;
        gather_staged<C>(box + (uint32_t)pc * (C * 2), row_bytes, pix_x[j] - f0x, pix_y[j] - f0y, refv, g0, g1);
        acc[j][0] = fmaf(g0, inv_half, acc[j][0]);
        acc[j][1] = fmaf(g1, inv_half, acc[j][1]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < DP; ++j) {
        const bool wanted = px_ok && d0 + j >= p.d_begin && d0 + j < p.d_end;
        const float ix = pix_x[j], iy = pix_y[j];
        const float f0x = floorf(ix), f0y = floorf(iy);
        const bool finite = (fabsf(ix) <= 3.0e38f) && (fabsf(iy) <= 3.0e38f);
        const bool inside = finite && f0x >= -1.0f && f0x <= (float)(p.w - 1) && f0y >= -1.0f && f0y <= (float)(p.hs - 1);
        float g0 = 0.0f, g1 = 0.0f;
        if (wanted && !finite) g0 = g1 = __int_as_float(0x7fc00000);  // the reference multiplies zeros by NaN weights
        if (wanted && inside) {
          const int x0 = (int)f0x, y0 = (int)f0y;
          const int rxp = x0 - bx.x, ryp = y0 - bx.y;
          if (bx.z >= 0 && rxp >= 0 && rxp + 1 < bw && ryp >= 0 && ryp + 1 < bh) {
            gather_staged<C>(box + (uint32_t)(ryp * bw + rxp) * (C * 2), (uint32_t)bw * (C * 2), ix - f0x, iy - f0y, refv, g0, g1);
          } else {
            // ---- direct: the same corners from global memory, out-of-image corners contribute zero
            const bool xa = x0 >= 0, xb = x0 + 1 < p.w, ya = y0 >= 0, yb = y0 + 1 < p.hs;
            const int xl = max(x0, 0), xr = min(x0 + 1, p.w - 1), yt = max(y0, 0), yu = min(y0 + 1, p.hs - 1);
            const uint4* q4[4] = {reinterpret_cast<const uint4*>(gsrc + ((long long)yt * p.w + xl) * p.src_ps),
                                  reinterpret_cast<const uint4*>(gsrc + ((long long)yt * p.w + xr) * p.src_ps),
                                  reinterpret_cast<const uint4*>(gsrc + ((long long)yu * p.w + xl) * p.src_ps),
                                  reinterpret_cast<const uint4*>(gsrc + ((long long)yu * p.w + xr) * p.src_ps)};
            const bool ok4[4] = {xa && ya, xb && ya, xa && yb, xb && yb};
            const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
            float sd[4][2];
#pragma unroll
            for (int k = 0; k < CH; ++k) {
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const uint4 v = ok4[c] ? __ldg(q4[c] + k) : zero;
                if (k == 0) dot8<true>(v, refv, sd[c][0], sd[c][1]);
                else dot8<false>(v, refv + 8 * k, sd[c][0], sd[c][1]);
              }
            }
            const float cx1 = ix - f0x, cx0 = 1.0f - cx1, cy1 = iy - f0y, cy0 = 1.0f - cy1;
            const float w00 = cy0 * cx0, w01 = cy0 * cx1, w10 = cy1 * cx0, w11 = cy1 * cx1;
            g0 = fmaf(w11, sd[3][0], fmaf(w10, sd[2][0], fmaf(w01, sd[1][0], w00 * sd[0][0])));
            g1 = fmaf(w11, sd[3][1], fmaf(w10, sd[2][1], fmaf(w01, sd[1][1], w00 * sd[0][1])));
          }
        }
        acc[j][0] = fmaf(g0, inv_half, acc[j][0]);
        acc[j][1] = fmaf(g1, inv_half, acc[j][1]);
      }
    }
    // ---- refill this slot with the box of source s + NBUF once every thread is done with it
    if (s + NBUF < p.n_src) {
      __syncthreads();
      if (threadIdx.x == 0) {
        const int4 nb = s_box[s + NBUF];
        if (nb.z >= 0) {
          mbar_expect_tx(s_bar + slot, (uint32_t)((nb.z ? Cfg::PIX1 : Cfg::PIX0) * C * 2));
          tma_load_4d(aligned + slot * Cfg::BOX_BYTES, &p.tm[nb.z][s + NBUF], s_bar + slot, 0, nb.x, nb.y, b);
        }
      }
    }
  }

  if (!px_ok) return;
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

// ------------------------------------------------------------------------------------------------ fp32 -> fp16 channel-last
// x: fp32 channel-last [B,h,w,*] (pixel stride xps floats, batch stride xbs) -> y: fp16 dense [B,h,w,C].  8 channels per thread.
template <int C>
__global__ void __launch_bounds__(256) nhwc_f32_to_f16_kernel(const float* __restrict__ x, long long xbs, int xps, __half* __restrict__ y,
                                                              long long hw) {
  constexpr int CH = C / 8;
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= hw * CH) return;
  const long long pix = i / CH;
  const int k = (int)(i - pix * CH);
  const int b = blockIdx.y;
  const float4* xp = reinterpret_cast<const float4*>(x + (long long)b * xbs + pix * xps + 8 * k);
  const float4 a = __ldg(xp), c = __ldg(xp + 1);
  const __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(a.z, a.w), h2 = __floats2half2_rn(c.x, c.y), h3 = __floats2half2_rn(c.z, c.w);
  uint4 o;
  o.x = *reinterpret_cast<const uint32_t*>(&h0); o.y = *reinterpret_cast<const uint32_t*>(&h1);
  o.z = *reinterpret_cast<const uint32_t*>(&h2); o.w = *reinterpret_cast<const uint32_t*>(&h3);
  *reinterpret_cast<uint4*>(y + ((long long)b * hw + pix) * C + 8 * k) = o;
}

// x: fp32 NCHW (dense (c,h,w), batch stride xbs) -> y: fp16 dense [B,h,w,C]; a warp transposes 32 pixels through shared memory
template <int C>
__global__ void __launch_bounds__(256) nchw_f32_to_nhwc_f16_kernel(const float* __restrict__ x, long long xbs, __half* __restrict__ y, long long hw) {
  __shared__ __half tile[8][32 * (C + 2)];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long pix0 = ((long long)blockIdx.x * 8 + warp) * 32;
  if (pix0 >= hw) return;
  const int b = blockIdx.y;
  const float* xp = x + (long long)b * xbs;
  __half* t = tile[warp];
  const long long pix = pix0 + lane;
#pragma unroll
  for (int c = 0; c < C; ++c) t[lane * (C + 2) + c] = __float2half_rn((pix < hw) ? __ldg(xp + (long long)c * hw + pix) : 0.0f);
  __syncwarp();
  __half* yp = y + ((long long)b * hw + pix0) * C;
  const int n_valid = (int)min((long long)32, hw - pix0) * C;
#pragma unroll
  for (int i = 0; i < C; i += 2) {
    const int e = (i / 2) * 64 + lane * 2;  // 2 consecutive halfs per lane
    if (e < n_valid) {
      const int px = e / C, c = e % C;
      *reinterpret_cast<__half2*>(yp + e) = __halves2half2(t[px * (C + 2) + c], t[px * (C + 2) + c + 1]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFnW1h)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFnW1h encode_fn_w1h() {
  static EncodeTiledFnW1h fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFnW1h>(f);
  }
  return fn;
}

template <int C, int DP>
static int launch_w1h_dp(W1hParams& p, cudaStream_t st) {
  using Cfg = W1hCfg<C>;
  // the opt-in is per device: set it for whichever device the caller has made current (idempotent, a few hundred ns)
  cudaError_t e = cudaFuncSetAttribute(warp_corr_h16_kernel<C, DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem);
  if (e != cudaSuccess) {
    set_error("warp_corr_h16: cannot reserve %zu bytes of shared memory: %s", Cfg::kSmem, cudaGetErrorString(e));
    return DMVS_ERR_CUDA;
  }
  p.chunk0 = p.d_begin / DP;
  p.n_chunks = ceil_div(p.d_end, DP) - p.chunk0;
  p.tiles_x = ceil_div(p.w, 16);
  dim3 grid(p.tiles_x * p.n_chunks, ceil_div(p.h, 16), p.B);
  DMVS_REQUIRE(grid.z <= 65535 && grid.y <= 65535, DMVS_ERR_BAD_SHAPE, "warp_corr_h16: grid too large (h=%d, B=%d)", p.h, p.B);
  warp_corr_h16_kernel<C, DP><<<grid, 256, Cfg::kSmem, st>>>(p);
  return check_launch("warp_corr_h16");
}

template <int C>
static int launch_w1h(W1hParams& p, cudaStream_t st) {
  using Box = W1hBox<C>;
  EncodeTiledFnW1h enc = encode_fn_w1h();
  DMVS_REQUIRE(enc != nullptr, DMVS_ERR_CUDA, "warp_corr_h16: cuTensorMapEncodeTiled is not available from the driver");
  const CUtensorMapSwizzle swz = (C == 32) ? CU_TENSOR_MAP_SWIZZLE_64B : (C == 16) ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  for (int s = 0; s < p.n_src; ++s) {
    for (int shape = 0; shape < 2; ++shape) {
      const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)p.w, (cuuint64_t)p.hs, (cuuint64_t)p.B};
      const cuuint64_t strides[3] = {(cuuint64_t)p.src_ps * 2, (cuuint64_t)p.w * p.src_ps * 2, (cuuint64_t)p.src_bs * 2};
      const cuuint32_t bx[4] = {(cuuint32_t)C, (cuuint32_t)(shape ? Box::BW1 : Box::BW0), (cuuint32_t)(shape ? Box::BH1 : Box::BH0), 1};
      const cuuint32_t es[4] = {1, 1, 1, 1};
      CUresult r = enc(&p.tm[shape][s], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(p.src[s]), dims, strides, bx, es,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      DMVS_REQUIRE(r == CUDA_SUCCESS, DMVS_ERR_CUDA, "warp_corr_h16: cuTensorMapEncodeTiled failed (%d) for C=%d h=%d w=%d stride=%d", (int)r, C,
                   p.h, p.w, p.src_ps);
    }
  }
  if (p.D <= 4 && Box::DP > 4) return launch_w1h_dp<C, (Box::DP > 4 ? 4 : Box::DP)>(p, st);
  return launch_w1h_dp<C, Box::DP>(p, st);
}

}  // namespace dmvs

extern "C" int dmvs_features_nhwc_f16(const float* x, long long x_bstride, int x_pixstride, void* y, int B, int C, int h, int w,
                                      void* stream) {
  using namespace dmvs;
  DMVS_REQUIRE(x && y, DMVS_ERR_BAD_POINTER, "features_nhwc_f16: null pointer");
  DMVS_REQUIRE(aligned16(y), DMVS_ERR_BAD_POINTER, "features_nhwc_f16: output must be 16-byte aligned");
  DMVS_REQUIRE(B >= 1 && B <= 65535 && h >= 1 && w >= 1, DMVS_ERR_BAD_SHAPE, "features_nhwc_f16: bad dims B=%d h=%d w=%d", B, h, w);
  DMVS_REQUIRE(x_pixstride == 0 || (x_pixstride >= C && x_pixstride % 4 == 0 && x_bstride % 4 == 0 && aligned16(x)), DMVS_ERR_BAD_SHAPE,
               "features_nhwc_f16: channel-last input needs pixel stride %d >= C, strides multiples of 4 floats, 16-byte alignment", x_pixstride);
  const long long hw = (long long)h * w;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  __half* yp = static_cast<__half*>(y);
  if (x_pixstride > 0) {
    dim3 grid((unsigned)((hw * (C / 8) + 255) / 256), B, 1);
    switch (C) {
      case 8: nhwc_f32_to_f16_kernel<8><<<grid, 256, 0, st>>>(x, x_bstride, x_pixstride, yp, hw); break;
      case 16: nhwc_f32_to_f16_kernel<16><<<grid, 256, 0, st>>>(x, x_bstride, x_pixstride, yp, hw); break;
      case 32: nhwc_f32_to_f16_kernel<32><<<grid, 256, 0, st>>>(x, x_bstride, x_pixstride, yp, hw); break;
      default: set_error("features_nhwc_f16: C=%d unsupported (8, 16, 32)", C); return DMVS_ERR_BAD_SHAPE;
    }
  } else {
    dim3 grid((unsigned)((hw + 255) / 256), B, 1);
    switch (C) {
      case 8: nchw_f32_to_nhwc_f16_kernel<8><<<grid, 256, 0, st>>>(x, x_bstride, yp, hw); break;
      case 16: nchw_f32_to_nhwc_f16_kernel<16><<<grid, 256, 0, st>>>(x, x_bstride, yp, hw); break;
      case 32: nchw_f32_to_nhwc_f16_kernel<32><<<grid, 256, 0, st>>>(x, x_bstride, yp, hw); break;
      default: set_error("features_nhwc_f16: C=%d unsupported (8, 16, 32)", C); return DMVS_ERR_BAD_SHAPE;
    }
  }
  return check_launch("features_nhwc_f16");
}

extern "C" int dmvs_warp_corr_h16_f32(const float* ref, long long ref_bstride, int ref_pixstride, const void* const* src, long long src_bstride,
                                      int src_pixstride, int n_src, const float* rt, const float* hyp, float* cost, void* cost_cells, int B,
                                      int C, int D, int h, int w, int d_begin, int d_end, int ref_row0, int src_rows, void* stream) {
  using namespace dmvs;
  DMVS_REQUIRE(ref && src && rt && hyp && (cost || cost_cells), DMVS_ERR_BAD_POINTER, "warp_corr_h16: null pointer");
  DMVS_REQUIRE(!cost_cells || aligned16(cost_cells), DMVS_ERR_BAD_POINTER, "warp_corr_h16: cost_cells must be 16-byte aligned");
  DMVS_REQUIRE(n_src >= 1 && n_src <= DMVS_MAX_SRC, DMVS_ERR_BAD_SHAPE, "warp_corr_h16: n_src=%d not in [1,%d]", n_src, DMVS_MAX_SRC);
  DMVS_REQUIRE(B >= 1 && D >= 1 && h >= 2 && w >= 2, DMVS_ERR_BAD_SHAPE, "warp_corr_h16: bad dims B=%d D=%d h=%d w=%d", B, D, h, w);
  DMVS_REQUIRE(C == 8 || C == 16 || C == 32, DMVS_ERR_BAD_SHAPE, "warp_corr_h16: C=%d unsupported (8, 16, 32)", C);
  DMVS_REQUIRE(0 <= d_begin && d_begin <= d_end && d_end <= D, DMVS_ERR_BAD_SHAPE, "warp_corr_h16: bad plane range [%d,%d) of %d", d_begin,
               d_end, D);
  DMVS_REQUIRE(src_pixstride >= C && src_pixstride % 8 == 0 && src_bstride % 8 == 0, DMVS_ERR_BAD_SHAPE,
               "warp_corr_h16: pixel stride %d / batch stride %lld must be multiples of 8 halfs and >= C", src_pixstride, src_bstride);
  DMVS_REQUIRE(ref_pixstride == 0 || (ref_pixstride >= C && ref_pixstride % 4 == 0 && ref_bstride % 4 == 0 && aligned16(ref)),
               DMVS_ERR_BAD_SHAPE, "warp_corr_h16: bad channel-last reference stride %d", ref_pixstride);
  if (src_rows == 0) src_rows = h;
  DMVS_REQUIRE(ref_row0 >= 0 && src_rows >= 2 && ref_row0 + h <= src_rows, DMVS_ERR_BAD_SHAPE,
               "warp_corr_h16: reference band rows [%d, %d) do not lie inside the %d source rows", ref_row0, ref_row0 + h, src_rows);
  DMVS_REQUIRE((long long)src_pixstride * src_rows * w < (1LL << 31) && (long long)C * h * w < (1LL << 31), DMVS_ERR_BAD_SHAPE,
               "warp_corr_h16: feature map too large for 32-bit offsets");
  for (int i = 0; i < n_src; ++i)
    DMVS_REQUIRE(src[i] != nullptr && aligned16(src[i]), DMVS_ERR_BAD_POINTER, "warp_corr_h16: src[%d] is null or not 16-byte aligned", i);
  if (d_begin == d_end) return DMVS_OK;
  W1hParams p;
  memset(&p, 0, sizeof(p));
  for (int i = 0; i < n_src; ++i) p.src[i] = static_cast<const __half*>(src[i]);
  p.ref = ref; p.rt = rt; p.hyp = hyp; p.cost = cost; p.cells = reinterpret_cast<uint2*>(cost_cells);
  p.ref_bs = ref_bstride; p.ref_ps = ref_pixstride; p.src_bs = src_bstride; p.src_ps = src_pixstride;
  p.B = B; p.D = D; p.h = h; p.w = w; p.n_src = n_src; p.d_begin = d_begin; p.d_end = d_end;
  p.row0 = ref_row0; p.hs = src_rows;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (C) {
    case 8: return launch_w1h<8>(p, st);
    case 16: return launch_w1h<16>(p, st);
    default: return launch_w1h<32>(p, st);
  }
}
