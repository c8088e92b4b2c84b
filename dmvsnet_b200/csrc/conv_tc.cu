// R1 on 5th-generation tensor cores: implicit-GEMM 3-D convolution with tcgen05.mma (sm_100a).
//
// Replaces the same reference blocks as conv3d.cu (networks/module.py:120-208 inside CostRegNet_part,
// module.py:358-436) - this is the fast path, conv3d.cu the exact-fp32 one.
//
// GEMM view (per CTA):  D[M = 128 voxels, N = 2*Cout] += A[M, K = 16] * B[K, N]   for every (plane, tap, channel chunk)
//   * M: a 16(h) x 8(w) patch of one output depth plane.  UMMA row r = 8*hl + wl.
//   * A: the input halo tile staged ONCE in shared memory in the UMMA canonical no-swizzle K-major layout
//        "[16-byte channel chunk][voxel]": 8 consecutive voxels along w are 8 contiguous 16-byte rows (one core
//        matrix), the 16 h-rows of the patch are SBO = (row pitch) apart, the two K chunks LBO apart.  Because rows are
//        plain voxels, the A operand of tap (kd,kh,kw) is the SAME buffer behind a descriptor whose start address is
//        shifted by ((kd*SH + kh)*SW + kw) voxels: im2col happens in the descriptor, no data is moved per tap.
//   * precision: fp32 activations / weights are split x = hi + lo (two fp16).  One K = 16 step carries 8 channels as
//        [A_hi | A_lo] against B = [[W_hi | W_lo], [W_hi | 0]], i.e. hi*hi + lo*hi in columns [0,Cout) and hi*lo in
//        [Cout, 2*Cout); the epilogue adds the two halves.  Only lo*lo (2^-22 relative) is dropped: fp32-class accuracy
//        at fp16 tensor rate (single-pass TF32 / BF16 breaks the 1e-3 depth contract).
//   * D: fp32 accumulators in TMEM, one [128 x N] block per output plane of the tile; read back with tcgen05.ld
//        for the fused eval-BatchNorm + ReLU + skip epilogue, stored straight to NCDHW.
//
// One elected thread issues all MMAs of a pass and commits them to an mbarrier; the CTA's other threads fill /
// drain.  Two to four CTAs are co-resident per SM, so one CTA's fill and epilogue overlap another's MMAs.
//
// Two kernels share the staging / issue / epilogue code:
//   conv_tc_kernel   one tile per CTA, channel passes in sequence (layers whose weights do not fit next to a pipeline)
//   conv_tcp_kernel  persistent and warp-specialised (layers with resident weights): 8 producer warps stage tile k+1..k+S-1
//                    into a ring of S smem stages while one thread issues the MMAs of tile k into one of two TMEM
//                    accumulator sets and 4 epilogue warps drain tile k-1 from the other; full/empty and accfull/accempty
//                    mbarriers (tcgen05.commit arrives on them) are the only synchronisation in the steady state.
//
// Four staging modes:
//   S1  3x3x3 stride 1            halo box [TD+2][18][10], tap = descriptor shifted by (kd,kh,kw) voxels
//   S2  3x3x3 stride 2            box [2TD+1][33][17]; columns de-interleaved (8 even, 9 odd) so the 8 voxels of a core
//                                 matrix stay contiguous, rows 2 apart (SBO = 2 row pitches)
//   TR  ConvTranspose k3 s2 p1 op1 input box [TD+1][17][9]; the 8 output parity classes are 8 accumulators, each tap of the
//                                 transposed kernel belongs to exactly one class (even: k=1; odd: k=0 at +1, k=2 at 0)
//   C0  conv0 (Cin = 2)           K is packed along kw: a 16-byte chunk holds (hi,lo) x 2 channels of voxel x and x+1, the
//                                 K=16 step reads the chunks at x-1 and x+1 (LBO = 2 voxels) = taps kw 0,1,2 (+ a zero column)
#include "tc_common.cuh"

namespace dmvs {

// ------------------------------------------------------------------------------------------------ parameters
struct TcParams {
  const float* x;
  const uint4* wtc;  // packed fp16 weights: [chunk j][tap][kc][n][8 halfs]
  const float* scale;
  const float* shift;
  const float* skip;
  float* y;
  long long x_bs, y_bs, skip_bs;
  int B, Cin, Cout, Di, Hi, Wi, Do, Ho, Wo;
  int relu;
  int tiles_x, tiles_y, tiles_z, n_tiles;
};

constexpr int TILE_H = 16, TILE_W = 8;
enum { MODE_S1 = 0, MODE_S2 = 1, MODE_TR = 2, MODE_C0 = 3 };

// CIN_P: input channels per pass; NB: UMMA N = 2 * Cout_p; TD: planes per CTA (output planes; input planes for TR)
template <int MODE, int CIN_P, int NB, int TD>
struct TcCfg {
  static constexpr int SD = (MODE == MODE_S2) ? 2 * TD + 1 : (MODE == MODE_TR) ? TD + 1 : TD + 2;
  static constexpr int SH = (MODE == MODE_S2) ? 2 * TILE_H + 1 : (MODE == MODE_TR) ? TILE_H + 1 : TILE_H + 2;
  static constexpr int SW = (MODE == MODE_S2) ? 2 * TILE_W + 1 : (MODE == MODE_TR) ? TILE_W + 1 : TILE_W + 2;
  static constexpr int SV = SD * SH * SW;  // staged voxels (16-byte rows per chunk plane)
  static constexpr int CJ = (MODE == MODE_C0) ? 1 : CIN_P / 8;
  static constexpr int NPLANE = (MODE == MODE_C0) ? 1 : 2 * CJ;  // 16-byte chunk planes per pass
  static constexpr int TAPS = (MODE == MODE_C0) ? 9 : 27;
  static constexpr int A_PITCH = SV * 16;
  static constexpr int A_BYTES = (NPLANE * A_PITCH + 127) / 128 * 128;
  static constexpr int A_LBO = (MODE == MODE_C0) ? 32 : A_PITCH;
  static constexpr int A_SBO = (MODE == MODE_S2) ? 2 * SW * 16 : SW * 16;
  static constexpr int B_TILE = 2 * NB * 16;  // bytes per (chunk, tap): [kc][n][16B]
  static constexpr int B_BYTES = CJ * TAPS * B_TILE;
  static constexpr int NACC = (MODE == MODE_TR) ? 8 * TD : TD;
  static constexpr int COLS = NACC * NB;
  static_assert(COLS <= 512, "accumulators exceed TMEM");
};
__host__ __device__ constexpr int pow2_cols(int c) { return c <= 32 ? 32 : c <= 64 ? 64 : c <= 128 ? 128 : c <= 256 ? 256 : 512; }

struct TileCoord {
  int x0, y0, z0, b;  // tile origin: output coordinates for S1 / S2 / C0, input coordinates for TR
};
__device__ __forceinline__ TileCoord decode_tile(const TcParams& p, int lt, int td) {
  TileCoord t;
  t.x0 = (lt % p.tiles_x) * TILE_W;
  lt /= p.tiles_x;
  t.y0 = (lt % p.tiles_y) * TILE_H;
  lt /= p.tiles_y;
  t.z0 = (lt % p.tiles_z) * td;
  t.b = lt / p.tiles_z;
  return t;
}

// ---- stage the input box of one (tile, pass): fp32 NCDHW -> [chunk plane][voxel][8 x fp16].  nthr threads cooperate.
template <int MODE, int CIN_P, int NB, int TD>
__device__ __forceinline__ void fill_stage(const TcParams& p, uint8_t* sA, const TileCoord& tc, int pass, int tid, int nthr) {
  using Cfg = TcCfg<MODE, CIN_P, NB, TD>;
  const long long iplane = (long long)p.Hi * p.Wi;
  const long long cs = (long long)p.Di * iplane;  // channel stride
  for (int item = tid; item < Cfg::SV * Cfg::CJ; item += nthr) {
    const int j = item / Cfg::SV, sv = item - j * Cfg::SV;
    const int sx = sv % Cfg::SW, sy = (sv / Cfg::SW) % Cfg::SH, sz = sv / (Cfg::SW * Cfg::SH);
    int ix, iy, iz;
    if (MODE == MODE_S2) {
      ix = (sx < TILE_W) ? 2 * (tc.x0 + sx) : 2 * (tc.x0 + sx - TILE_W) - 1;
      iy = 2 * tc.y0 - 1 + sy;
      iz = 2 * tc.z0 - 1 + sz;
    } else if (MODE == MODE_TR) {
      ix = tc.x0 + sx; iy = tc.y0 + sy; iz = tc.z0 + sz;
    } else {
      ix = tc.x0 + sx - 1; iy = tc.y0 + sy - 1; iz = tc.z0 + sz - 1;
    }
    const bool vyz = (iy >= 0 && iy < p.Hi && iz >= 0 && iz < p.Di);
    if (MODE == MODE_C0) {
      // chunk = [hi c0, hi c1, lo c0, lo c1] of voxel ix, then the same of voxel ix + 1
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (vyz) {
        const float* src = p.x + (long long)tc.b * p.x_bs + (long long)iz * iplane + (long long)iy * p.Wi;
        if (ix >= 0 && ix < p.Wi) { v[0] = __ldg(src + ix); v[1] = __ldg(src + cs + ix); }
        if (ix + 1 >= 0 && ix + 1 < p.Wi) { v[2] = __ldg(src + ix + 1); v[3] = __ldg(src + cs + ix + 1); }
      }
      __half h[4], l[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        h[i] = __float2half_rn(v[i]);
        l[i] = __float2half_rn(v[i] - __half2float(h[i]));
      }
      const __half2 w0 = __halves2half2(h[0], h[1]), w1 = __halves2half2(l[0], l[1]);
      const __half2 w2 = __halves2half2(h[2], h[3]), w3 = __halves2half2(l[2], l[3]);
      *reinterpret_cast<uint4*>(sA + sv * 16) =
          make_uint4(*reinterpret_cast<const uint32_t*>(&w0), *reinterpret_cast<const uint32_t*>(&w1),
                     *reinterpret_cast<const uint32_t*>(&w2), *reinterpret_cast<const uint32_t*>(&w3));
    } else {
      float v[8];
      if (vyz && ix >= 0 && ix < p.Wi) {
        const float* src = p.x + (long long)tc.b * p.x_bs + ((long long)(pass * CIN_P + j * 8) * p.Di + iz) * iplane +
                           (long long)iy * p.Wi + ix;
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = __ldg(src + c * cs);
      } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = 0.f;
      }
      uint4 hi, lo;
      split_pack8(v, hi, lo);
      *reinterpret_cast<uint4*>(sA + (2 * j) * Cfg::A_PITCH + sv * 16) = hi;
      *reinterpret_cast<uint4*>(sA + (2 * j + 1) * Cfg::A_PITCH + sv * 16) = lo;
    }
  }
}

// ---- all MMAs of one (tile, pass); executed by ONE thread.  acc_base: TMEM address of accumulator 0.
template <int MODE, int CIN_P, int NB, int TD>
__device__ __forceinline__ void issue_tile(uint32_t a0, uint32_t b0, uint32_t acc_base, bool fresh) {
  using Cfg = TcCfg<MODE, CIN_P, NB, TD>;
  constexpr uint32_t idesc = make_idesc(NB);
  auto issue = [&](int acc, int tap, uint32_t voxel, bool first) {
#pragma unroll
    for (int j = 0; j < Cfg::CJ; ++j) {
      const uint64_t ad = make_desc(a0 + (MODE == MODE_C0 ? 0 : (2 * j) * Cfg::A_PITCH) + voxel * 16u, Cfg::A_LBO, Cfg::A_SBO);
      const uint64_t bd = make_desc(b0 + (j * Cfg::TAPS + tap) * Cfg::B_TILE, NB * 16, 128);
      umma_f16(acc_base + acc * NB, ad, bd, idesc, (!fresh || !first || j > 0) ? 1u : 0u);
    }
  };
  if (MODE == MODE_TR) {
#pragma unroll 1
    for (int t = 0; t < TD; ++t)
#pragma unroll 1
      for (int cls = 0; cls < 8; ++cls) {
        const int pz = cls >> 2, py = (cls >> 1) & 1, px = cls & 1;
        bool first = true;
        // per dimension: even output -> (k=1, offset 0); odd output -> (k=0, offset +1) and (k=2, offset 0)
        for (int a = 0; a <= pz; ++a)
          for (int bq = 0; bq <= py; ++bq)
            for (int c = 0; c <= px; ++c) {
              const int kz = pz ? (a ? 2 : 0) : 1, oz = pz ? (a ? 0 : 1) : 0;
              const int ky = py ? (bq ? 2 : 0) : 1, oy = py ? (bq ? 0 : 1) : 0;
              const int kx = px ? (c ? 2 : 0) : 1, ox = px ? (c ? 0 : 1) : 0;
              issue(t * 8 + cls, (kz * 3 + ky) * 3 + kx, (uint32_t)(((t + oz) * Cfg::SH + oy) * Cfg::SW + ox), first);
              first = false;
            }
      }
  } else {
#pragma unroll 1
    for (int t = 0; t < TD; ++t)
#pragma unroll 1
      for (int tap = 0; tap < Cfg::TAPS; ++tap) {
        uint32_t voxel;
        if (MODE == MODE_C0) {
          const int kd = tap / 3, kh = tap % 3;
          voxel = (uint32_t)(((t + kd) * Cfg::SH + kh) * Cfg::SW);
        } else {
          const int kd = tap / 9, kh = (tap % 9) / 3, kw = tap % 3;
          if (MODE == MODE_S2)
            voxel = (uint32_t)(((2 * t + kd) * Cfg::SH + kh) * Cfg::SW + (kw == 1 ? 0 : (kw == 0 ? TILE_W : TILE_W + 1)));
          else
            voxel = (uint32_t)(((t + kd) * Cfg::SH + kh) * Cfg::SW + kw);
        }
        issue(t, tap, voxel, tap == 0);
      }
  }
}

// ---- epilogue of one tile: TMEM -> registers -> BN / ReLU / skip -> NCDHW.  A warp reads TMEM lanes 32*(warp%4)..+31;
//      lane l of quadrant q is UMMA row 32q + l = patch position (hl = 4q + l/8, wl = l%8).
//      Planes t = t_begin, t_begin + t_step, ... are handled by the calling warp.
template <int MODE, int CIN_P, int NB, int TD>
__device__ __forceinline__ void epilogue_tile(const TcParams& p, const TileCoord& tc, uint32_t acc_base, int q, int lane, int t_begin,
                                              int t_step) {
  constexpr int COUT_P = NB / 2;
  const int hl = q * 4 + (lane >> 3), wl = lane & 7;
  const uint32_t lane_addr = acc_base + ((uint32_t)(q * 32) << 16);
  const long long oplane = (long long)p.Ho * p.Wo;
  if (MODE == MODE_TR) {
    const int iy = tc.y0 + hl, ix = tc.x0 + wl;
    const bool in_img = (iy < p.Hi) && (ix < p.Wi);
    for (int t = t_begin; t < TD; t += t_step) {
      if (tc.z0 + t >= p.Di) break;  // warp-uniform
#pragma unroll 1
      for (int pzy = 0; pzy < 4; ++pzy) {
        const int oz = 2 * (tc.z0 + t) + (pzy >> 1), oy = 2 * iy + (pzy & 1);
        const uint32_t te = lane_addr + (t * 8 + pzy * 2) * NB, to = te + NB;
#pragma unroll 1
        for (int c0 = 0; c0 < COUT_P; c0 += 8) {
          float he[8], le[8], ho[8], lo8[8];
          tmem_ld8(te + c0, he);
          tmem_ld8(te + COUT_P + c0, le);
          tmem_ld8(to + c0, ho);
          tmem_ld8(to + COUT_P + c0, lo8);
          if (!in_img) continue;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int co = c0 + c;
            if (co >= p.Cout) break;
            float e = he[c] + le[c], o = ho[c] + lo8[c];
            if (p.scale) {
              const float sc = __ldg(p.scale + co), sh = __ldg(p.shift + co);
              e = fmaf(e, sc, sh);
              o = fmaf(o, sc, sh);
            }
            if (p.relu) { e = fmaxf(e, 0.f); o = fmaxf(o, 0.f); }
            const long long off = ((long long)co * p.Do + oz) * oplane + (long long)oy * p.Wo + 2 * ix;
            if (p.skip) {
              const float2 sk = __ldg(reinterpret_cast<const float2*>(p.skip + (long long)tc.b * p.skip_bs + off));
              e += sk.x; o += sk.y;
            }
            *reinterpret_cast<float2*>(p.y + (long long)tc.b * p.y_bs + off) = make_float2(e, o);
          }
        }
      }
    }
  } else {
    const int oy = tc.y0 + hl, ox = tc.x0 + wl;
    const bool in_img = (oy < p.Ho) && (ox < p.Wo);
    for (int t = t_begin; t < TD; t += t_step) {
      const int oz = tc.z0 + t;
      if (oz >= p.Do) break;  // warp-uniform
      const uint32_t taddr = lane_addr + t * NB;
#pragma unroll 1
      for (int c0 = 0; c0 < COUT_P; c0 += 8) {
        float hi8[8], lo8[8];
        tmem_ld8(taddr + c0, hi8);
        tmem_ld8(taddr + COUT_P + c0, lo8);
        if (!in_img) continue;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int co = c0 + c;
          if (co >= p.Cout) break;
          float v = hi8[c] + lo8[c];
          if (p.scale) v = fmaf(v, __ldg(p.scale + co), __ldg(p.shift + co));
          if (p.relu) v = fmaxf(v, 0.f);
          const long long off = ((long long)co * p.Do + oz) * oplane + (long long)oy * p.Wo + ox;
          if (p.skip) v += __ldg(p.skip + (long long)tc.b * p.skip_bs + off);
          p.y[(long long)tc.b * p.y_bs + off] = v;
        }
      }
    }
  }
}

// ================================================================================================ one tile per CTA
constexpr int TC_THREADS = 256;

template <int MODE, int CIN_P, int NB, int TD>
__global__ void __launch_bounds__(TC_THREADS) conv_tc_kernel(const __grid_constant__ TcParams p) {
  using Cfg = TcCfg<MODE, CIN_P, NB, TD>;
  constexpr int TMEM_COLS = pow2_cols(Cfg::COLS);
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + Cfg::A_BYTES;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + Cfg::A_BYTES + Cfg::B_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const TileCoord tc = decode_tile(p, blockIdx.x, TD);

  if (warp == 0) tmem_alloc(tmem_slot, TMEM_COLS);
  if (tid == 32) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_pass = (MODE == MODE_C0) ? 1 : p.Cin / CIN_P;
  for (int pass = 0; pass < n_pass; ++pass) {
    if (pass > 0) mbar_wait(bar, (pass - 1) & 1);  // previous pass' MMAs have finished reading smem
    fill_stage<MODE, CIN_P, NB, TD>(p, sA, tc, pass, tid, TC_THREADS);
    {  // weights of this pass: a contiguous byte range of the packed image
      const uint4* wsrc = p.wtc + (size_t)pass * (Cfg::B_BYTES / 16);
      uint4* wdst = reinterpret_cast<uint4*>(sB);
      for (int i = tid; i < Cfg::B_BYTES / 16; i += TC_THREADS) wdst[i] = __ldg(wsrc + i);
    }
    fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      issue_tile<MODE, CIN_P, NB, TD>(smem_u32(sA), smem_u32(sB), tmem_base, pass == 0);
      umma_commit(bar);
    }
  }
  mbar_wait(bar, (n_pass - 1) & 1);
  tc_fence_after();
  epilogue_tile<MODE, CIN_P, NB, TD>(p, tc, tmem_base, warp & 3, lane, warp >> 2, 2);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ================================================================================================ persistent, pipelined
constexpr int P_PROD_WARPS = 8, P_EPI_WARPS = 4;
constexpr int P_THREADS = (P_PROD_WARPS + 1 + P_EPI_WARPS) * 32;

template <int MODE, int CIN_P, int NB, int TD, int STAGES>
struct TcpSmem {
  using Cfg = TcCfg<MODE, CIN_P, NB, TD>;
  static constexpr int ACC_SETS = (2 * Cfg::COLS <= 512) ? 2 : 1;
  static constexpr int TMEM_COLS = pow2_cols(ACC_SETS * Cfg::COLS);
  static constexpr int OFF_B = STAGES * Cfg::A_BYTES;
  static constexpr int OFF_BAR = OFF_B + (Cfg::B_BYTES + 127) / 128 * 128;
  static constexpr int BYTES = OFF_BAR + 8 * (2 * STAGES + 4) + 16;
  static_assert(BYTES <= 227 * 1024, "pipeline does not fit shared memory");
};

template <int MODE, int CIN_P, int NB, int TD, int STAGES>
__global__ void __launch_bounds__(P_THREADS, 1) conv_tcp_kernel(const __grid_constant__ TcParams p) {
  using Cfg = TcCfg<MODE, CIN_P, NB, TD>;
  using L = TcpSmem<MODE, CIN_P, NB, TD, STAGES>;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sB = smem + L::OFF_B;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* empty = full + STAGES;
  uint64_t* accfull = empty + STAGES;
  uint64_t* accempty = accfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accempty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int MMA_WARP = P_PROD_WARPS;

  if (warp == MMA_WARP) tmem_alloc(tmem_slot, L::TMEM_COLS);
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full + s, P_PROD_WARPS * 32);
      mbar_init(empty + s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(accfull + a, 1);
      mbar_init(accempty + a, P_EPI_WARPS * 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {  // the layer's weights stay resident for the life of the CTA
    const uint4* wsrc = p.wtc;
    uint4* wdst = reinterpret_cast<uint4*>(sB);
    for (int i = tid; i < Cfg::B_BYTES / 16; i += P_THREADS) wdst[i] = __ldg(wsrc + i);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < P_PROD_WARPS) {
    // ------------------------------------------------------------------ producers: global -> split fp16 -> smem ring
    int k = 0;
    for (int lt = blockIdx.x; lt < p.n_tiles; lt += gridDim.x, ++k) {
      const int s = k % STAGES, u = k / STAGES;
      mbar_wait(empty + s, (u & 1) ^ 1);  // passes immediately the first time round
      fill_stage<MODE, CIN_P, NB, TD>(p, smem + s * Cfg::A_BYTES, decode_tile(p, lt, TD), 0, tid, P_PROD_WARPS * 32);
      fence_proxy_async();
      mbar_arrive(full + s);
    }
  } else if (warp == MMA_WARP) {
    // ------------------------------------------------------------------ MMA issuer: one thread
    int k = 0;
    for (int lt = blockIdx.x; lt < p.n_tiles; lt += gridDim.x, ++k) {
      if (lane == 0) {
        const int s = k % STAGES, u = k / STAGES;
        const int a = (L::ACC_SETS == 2) ? (k & 1) : 0, v = (L::ACC_SETS == 2) ? (k >> 1) : k;
        mbar_wait(accempty + a, (v & 1) ^ 1);  // epilogue has drained this accumulator set
        mbar_wait(full + s, u & 1);            // producers have filled this stage
        tc_fence_after();
        issue_tile<MODE, CIN_P, NB, TD>(smem_u32(smem + s * Cfg::A_BYTES), smem_u32(sB), tmem_base + a * Cfg::COLS, true);
        umma_commit(empty + s);    // stage reusable once these MMAs have read it
        umma_commit(accfull + a);  // accumulators complete
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int q = warp & 3;
    int k = 0;
    for (int lt = blockIdx.x; lt < p.n_tiles; lt += gridDim.x, ++k) {
      const int a = (L::ACC_SETS == 2) ? (k & 1) : 0, v = (L::ACC_SETS == 2) ? (k >> 1) : k;
      mbar_wait(accfull + a, v & 1);
      tc_fence_after();
      epilogue_tile<MODE, CIN_P, NB, TD>(p, decode_tile(p, lt, TD), tmem_base + a * Cfg::COLS, q, lane, 0, 1);
      tc_fence_before();
      mbar_arrive(accempty + a);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tmem_base, L::TMEM_COLS);
}

// ================================================================================================ host side
static void tile_grid(TcParams& p, int mode, int td) {
  // tiles cover the output grid, except TR where they cover the input grid (each input voxel owns a 2x2x2 output block)
  const int gw = (mode == MODE_TR) ? p.Wi : p.Wo, gh = (mode == MODE_TR) ? p.Hi : p.Ho, gd = (mode == MODE_TR) ? p.Di : p.Do;
  p.tiles_x = ceil_div(gw, TILE_W);
  p.tiles_y = ceil_div(gh, TILE_H);
  p.tiles_z = ceil_div(gd, td);
  p.n_tiles = p.tiles_x * p.tiles_y * p.tiles_z * p.B;
}

template <typename K>
static int set_smem(K kern, int bytes, bool& configured) {
  if (configured) return DMVS_OK;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) {
    set_error("conv_tc: cudaFuncSetAttribute(%d bytes): %s", bytes, cudaGetErrorString(e));
    return DMVS_ERR_CUDA;
  }
  configured = true;
  return DMVS_OK;
}

template <int MODE, int CIN_P, int NB, int TD>
static int launch_tc(TcParams p, cudaStream_t st) {
  using Cfg = TcCfg<MODE, CIN_P, NB, TD>;
  constexpr int SMEM = Cfg::A_BYTES + Cfg::B_BYTES + 64;
  static_assert(SMEM <= 227 * 1024, "tile does not fit shared memory");
  tile_grid(p, MODE, TD);
  auto kern = conv_tc_kernel<MODE, CIN_P, NB, TD>;
  static bool configured = false;
  if (int rc = set_smem(kern, SMEM, configured)) return rc;
  kern<<<p.n_tiles, TC_THREADS, SMEM, st>>>(p);
  return check_launch("conv_tc");
}

template <int MODE, int CIN_P, int NB, int TD, int STAGES>
static int launch_tcp(TcParams p, cudaStream_t st) {
  using L = TcpSmem<MODE, CIN_P, NB, TD, STAGES>;
  tile_grid(p, MODE, TD);
  auto kern = conv_tcp_kernel<MODE, CIN_P, NB, TD, STAGES>;
  static bool configured = false;
  if (int rc = set_smem(kern, L::BYTES, configured)) return rc;
  const int grid = p.n_tiles < kNumSMs ? p.n_tiles : kNumSMs;  // one persistent CTA per SM
  kern<<<grid, P_THREADS, L::BYTES, st>>>(p);
  return check_launch("conv_tcp");
}

// returns DMVS_OK, an error, or +1 when this layer shape has no tensor-core specialisation (caller falls back to conv3d.cu)
int conv_layer_tc(const float* x, long long x_bs, const dmvs_conv_layer& L, const float* skip, long long skip_bs, float* y,
                  long long y_bs, int B, int Cin, int Cout, int Di, int Hi, int Wi, int kd, int stride, int transposed, int relu,
                  cudaStream_t st) {
  if (!L.w_tc || kd != 3) return 1;
  DMVS_REQUIRE(x && y, DMVS_ERR_BAD_POINTER, "conv_tc: null pointer");
  DMVS_REQUIRE(aligned16(L.w_tc), DMVS_ERR_BAD_POINTER, "conv_tc: packed weights must be 16-byte aligned");
  TcParams p;
  p.x = x; p.wtc = reinterpret_cast<const uint4*>(L.w_tc); p.scale = L.scale; p.shift = L.shift; p.skip = skip; p.y = y;
  p.x_bs = x_bs; p.y_bs = y_bs; p.skip_bs = skip_bs;
  p.B = B; p.Cin = Cin; p.Cout = Cout; p.Di = Di; p.Hi = Hi; p.Wi = Wi; p.relu = relu;
  p.tiles_x = p.tiles_y = p.tiles_z = p.n_tiles = 1;
  if (transposed) {
    p.Do = 2 * Di; p.Ho = 2 * Hi; p.Wo = 2 * Wi;
    DMVS_REQUIRE(aligned16(y) && (y_bs % 2 == 0) && (!skip || (aligned16(skip) && skip_bs % 2 == 0)), DMVS_ERR_BAD_POINTER,
                 "conv_tc: y/skip must be 16-byte aligned for the transposed conv");
    if (Cin == 16 && Cout == 8) return launch_tc<MODE_TR, 16, 16, 2>(p, st);      // conv11
    if (Cin == 32 && Cout == 16) return launch_tc<MODE_TR, 16, 32, 1>(p, st);     // conv9, two channel passes
    if (Cin == 64 && Cout == 32) return launch_tc<MODE_TR, 16, 64, 1>(p, st);     // conv7, four channel passes
    return 1;
  }
  if (stride == 2) {
    p.Do = (Di - 1) / 2 + 1; p.Ho = (Hi - 1) / 2 + 1; p.Wo = (Wi - 1) / 2 + 1;
    if (Cin == 8 && Cout == 16) return launch_tc<MODE_S2, 8, 32, 1>(p, st);     // conv1
    if (Cin == 16 && Cout == 32) return launch_tc<MODE_S2, 8, 64, 1>(p, st);     // conv3, two passes
    if (Cin == 32 && Cout == 64) return launch_tc<MODE_S2, 8, 128, 1>(p, st);    // conv5, four passes
    return 1;
  }
  p.Do = Di; p.Ho = Hi; p.Wo = Wi;
  if (Cin == 2 && Cout == 8) return launch_tc<MODE_C0, 2, 16, 4>(p, st);       // conv0, K packed along kw
  if (Cin == 8 && Cout <= 8) return launch_tc<MODE_S1, 8, 16, 4>(p, st);       // prob (8 -> 2)
  if (Cin == 16 && Cout == 16) return launch_tc<MODE_S1, 16, 32, 2>(p, st);    // conv2
  if (Cin == 32 && Cout == 32) return launch_tc<MODE_S1, 16, 64, 2>(p, st);     // conv4, two channel passes
  if (Cin == 64 && Cout == 64) return launch_tc<MODE_S1, 8, 128, 1>(p, st);     // conv6, eight channel passes
  return 1;
}

}  // namespace dmvs
