// N1: FeatureNet's 2-D convolutions (reference networks/module.py:274-340, Conv2d wrapper :28-69) as fp32 direct
// convolutions for small channel counts (3..32 in, 8..64 out), sm_100a.
//
// The FPN runs on N full-resolution views per reference view (230 GFLOP at DTU size, a third of the regularisation
// nets) and is FMA-bound once written for these shapes: cuDNN's fp32 kernels reach ~10 TFLOP/s on them.
// One kernel template covers every layer:
//   * block = 32x32 output pixels x CT output channels, 256 threads; thread = 4 x-consecutive pixels x CT channels
//     (64 / 32 accumulators), lanes 8 (x) x 4 (y) so a quarter warp reads 128 contiguous bytes of an input row;
//   * the input halo tile is staged through shared memory CC input channels at a time by double-buffered 16-byte
//     cp.async (zero-size copies = zero padding; the tile starts at a 16-byte aligned column left of the halo), the weights of the chunk as [ci][kh][kw][co] so four output channels arrive with one broadcast
//     LDS.128: per (ci, kh) a thread issues 2-3 input loads + K*CT/4 weight loads for K*4*CT FMAs (>= 12 FMA per load);
//   * epilogue: eval BatchNorm as scale/shift (or the conv bias as shift), ReLU, optional "+ nearest-x2-upsampled
//     coarser map" (the FPN top-down add, module.py:329,334), and stores in NCHW and/or channel-last split into the
//     two feature sets (`stageK` / `stageK_c` = the channel halves, module.py:326-336) so the W1 gather kernel reads
//     them in place.
// Arithmetic is plain fp32 FMA in (ci, kh, kw) order: differences to cuDNN are summation order only.
#include "tc_common.cuh"

namespace dmvs {

struct FeatConvParams {
  const float* x;       // [B,CIN,Hi,Wi]
  const float* w;       // [CIN][K][K][COUT]
  const float* scale;   // [COUT] or null
  const float* shift;   // [COUT] or null
  const float* up_add;  // [B,COUT,Ho/2,Wo/2] or null: added after the affine (no ReLU in that case in the FPN)
  float* y_nchw;        // [B,COUT,Ho,Wo] or null
  float* y_nhwc0;       // [B,Ho,Wo,COUT/2] channels [0,COUT/2) or null
  float* y_nhwc1;       // [B,Ho,Wo,COUT/2] channels [COUT/2,COUT) or null
  uint4* y_cells;       // CH16 cells [B][2*COUT/8 planes][Ho][Wo] (fp16 hi/lo, input format of the tensor-core 3x3 heads) or null
  int B, Hi, Wi, Ho, Wo, relu;
  int cells_s2d;        // y_cells holds the 2x2 pixel-unshuffled map: [B][2*(4*COUT)/8][Ho/2][Wo/2], channel (dy*2+dx)*COUT + c
};

template <int K, int S, int CIN, int COUT>
struct FeatCfg {
  static constexpr int CT = COUT < 16 ? COUT : 16;               // output channels per thread / block
  static constexpr int NZ = COUT / CT;                           // channel groups -> blockIdx.z
  static constexpr int TILE = 32;                                // output tile edge
  static constexpr int IN = (TILE - 1) * S + K;                  // input tile edge (rows)
  static constexpr int PADK = K / 2;
  // the staged tile starts OFF columns left of the halo so that smem column 0 is a 16-byte aligned global column
  // (tile origins are multiples of 32): rows are then staged with 16-byte cp.async
  static constexpr int OFF = (4 - PADK % 4) % 4;
  static constexpr int NIN = OFF + 3 * S + K;                    // input columns one thread touches per row, from its aligned base
  static constexpr int NLD = (NIN + 3) / 4;                      // as 16-byte loads
  static constexpr int PITCH = 4 * 7 * S + 4 * NLD;              // last thread's base + its loads
  static constexpr int CC_MAX = (44 * 1024) / (IN * PITCH * 4);  // input channels per chunk (about 44 KB of tile)
  static constexpr int CC = CIN < (CC_MAX < 1 ? 1 : CC_MAX) ? CIN : (CC_MAX < 1 ? 1 : (CC_MAX >= 8 ? 8 : (CC_MAX >= 4 ? 4 : (CC_MAX >= 2 ? 2 : 1))));
  static constexpr int NCHUNK = CIN / CC;
  static constexpr int NBUF = NCHUNK > 1 ? 2 : 1;                // double-buffered staging when there is something to overlap
  static constexpr size_t kSmem = (size_t)NBUF * (CC * IN * PITCH + CC * K * K * CT) * 4;
  static_assert(CIN % CC == 0, "chunking must divide CIN");
};

template <int K, int S, int CIN, int COUT>
__global__ void __launch_bounds__(256, 2) feat_conv_kernel(const __grid_constant__ FeatConvParams p) {
  using Cfg = FeatCfg<K, S, CIN, COUT>;
  constexpr int CT = Cfg::CT, NZ = Cfg::NZ, IN = Cfg::IN, PITCH = Cfg::PITCH, CC = Cfg::CC, NLD = Cfg::NLD, OFF = Cfg::OFF;
  constexpr int PAD = K / 2;
  extern __shared__ __align__(16) float fsm[];
  // fsm: [NBUF][CC][IN][PITCH] input tiles, then [NBUF][CC][K][K][CT] weights

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lx = lane & 7, ly = lane >> 3;
  const int b = blockIdx.z / NZ, cz = blockIdx.z - b * NZ;
  const int X0 = blockIdx.x * 32, Y0 = blockIdx.y * 32;
  const int ox = X0 + 4 * lx, oy = Y0 + 4 * warp + ly;        // this thread's first output pixel
  const int ix0 = X0 * S - PAD - OFF, iy0 = Y0 * S - PAD;     // input coordinates of smem (row 0, column 0); ix0 % 4 == 0

  float acc[4][CT];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < CT; ++j) acc[i][j] = 0.0f;

  const float* xb = p.x + (long long)b * CIN * p.Hi * p.Wi;
  constexpr int NBUF = Cfg::NBUF, IN_FLOATS = CC * IN * PITCH, W_FLOATS = CC * K * K * CT;
  float* s_wbase = fsm + NBUF * IN_FLOATS;

  // stage chunk c (CC input channels of the halo tile, zero padding via zero-size cp.async, and their weights)
  auto stage = [&](int c, int buf) {
    float* din = fsm + buf * IN_FLOATS;
    const int c0 = c * CC;
    if ((p.Wi & 3) == 0) {
      // 16-byte copies: a group of 4 columns is entirely inside or entirely outside the image (ix0 and Wi are multiples of 4)
      constexpr int G4 = PITCH / 4;
      for (int e = threadIdx.x; e < CC * IN * G4; e += 256) {
        const int g = e % G4, r = e / G4;
        const int ci = r / IN, row = r - ci * IN;
        const int gy = iy0 + row, gx = ix0 + 4 * g;
        const bool ok = gy >= 0 && gy < p.Hi && gx >= 0 && gx < p.Wi;
        const float* src = xb + ((long long)(c0 + ci) * p.Hi + (ok ? gy : 0)) * p.Wi + (ok ? gx : 0);
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(din + r * PITCH + 4 * g);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16 : 0) : "memory");
      }
    } else {
      for (int r = warp; r < CC * IN; r += 8) {
        const int ci = r / IN, row = r - ci * IN;
        const int gy = iy0 + row;
        const bool row_ok = gy >= 0 && gy < p.Hi;
        const float* src = xb + ((long long)(c0 + ci) * p.Hi + (row_ok ? gy : 0)) * p.Wi;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(din + r * PITCH);
#pragma unroll
        for (int col = lane; col < PITCH; col += 32) {
          const int gx = ix0 + col;
          const bool ok = row_ok && gx >= 0 && gx < p.Wi;
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + col * 4), "l"(src + (ok ? gx : 0)), "r"(ok ? 4 : 0) : "memory");
        }
      }
    }
    const uint32_t dw = (uint32_t)__cvta_generic_to_shared(s_wbase + buf * W_FLOATS);
    for (int e = threadIdx.x; e < W_FLOATS / 4; e += 256) {
      const int co4 = e % (CT / 4), t = e / (CT / 4);  // t = ci*K*K + tap
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dw + e * 16),
                   "l"(p.w + (long long)(c0 * K * K + t) * COUT + cz * CT + co4 * 4) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  stage(0, 0);
#pragma unroll 1
  for (int c = 0; c < Cfg::NCHUNK; ++c) {
    const int buf = (NBUF == 2) ? (c & 1) : 0;
    if (c + 1 < Cfg::NCHUNK) {
      stage(c + 1, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const float* s_in = fsm + buf * IN_FLOATS;
    const float* s_w = s_wbase + buf * W_FLOATS;
    // ---- FMA loop
#pragma unroll
    for (int ci = 0; ci < CC; ++ci) {
#pragma unroll
      for (int kh = 0; kh < K; ++kh) {
        float in[NLD * 4];
        const float* rp = s_in + (ci * IN + (4 * warp + ly) * S + kh) * PITCH + 4 * lx * S;
#pragma unroll
        for (int q = 0; q < NLD; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(rp + 4 * q);
          in[4 * q] = v.x; in[4 * q + 1] = v.y; in[4 * q + 2] = v.z; in[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int kw = 0; kw < K; ++kw) {
          const float* wp = s_w + ((ci * K + kh) * K + kw) * CT;
#pragma unroll
          for (int j4 = 0; j4 < CT / 4; ++j4) {
            const float4 wv = *reinterpret_cast<const float4*>(wp + 4 * j4);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float xv = in[OFF + i * S + kw];
              acc[i][4 * j4] = fmaf(xv, wv.x, acc[i][4 * j4]);
              acc[i][4 * j4 + 1] = fmaf(xv, wv.y, acc[i][4 * j4 + 1]);
              acc[i][4 * j4 + 2] = fmaf(xv, wv.z, acc[i][4 * j4 + 2]);
              acc[i][4 * j4 + 3] = fmaf(xv, wv.w, acc[i][4 * j4 + 3]);
            }
          }
        }
      }
    }
    __syncthreads();  // the next iteration's cp.async refills the other buffer only; this one is refilled one iteration later
  }

  // ---- epilogue
  if (oy >= p.Ho || ox >= p.Wo) return;
  const int nvalid = min(4, p.Wo - ox);
#pragma unroll
  for (int j = 0; j < CT; ++j) {
    const int co = cz * CT + j;
    const float sc = p.scale ? __ldg(p.scale + co) : 1.0f, sh = p.shift ? __ldg(p.shift + co) : 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float v = fmaf(acc[i][j], sc, sh);
      if (p.relu) v = fmaxf(v, 0.0f);
      acc[i][j] = v;
    }
    if (p.up_add) {
      const int hw2 = (p.Ho >> 1) * (p.Wo >> 1);
      const float* up = p.up_add + ((long long)b * COUT + co) * hw2 + (oy >> 1) * (p.Wo >> 1);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (i < nvalid) acc[i][j] += __ldg(up + ((ox + i) >> 1));
    }
  }
  if (p.y_nchw) {
    const long long hw = (long long)p.Ho * p.Wo;
    float* yp = p.y_nchw + ((long long)b * COUT + cz * CT) * hw + (long long)oy * p.Wo + ox;
    const bool vec = nvalid == 4 && ((p.Wo & 3) == 0);
#pragma unroll
    for (int j = 0; j < CT; ++j) {
      if (vec) {
        *reinterpret_cast<float4*>(yp + j * hw) = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (i < nvalid) yp[j * hw + i] = acc[i][j];
      }
    }
  }
  if (p.y_cells) {  // CH16 cells for a tensor-core consumer: 8 channels = one (hi, lo) cell pair per pixel
    const long long hw = (long long)p.Ho * p.Wo;
#pragma unroll
    for (int j8 = 0; j8 < CT / 8; ++j8) {
      uint4* cp = p.y_cells + ((long long)b * (COUT / 4) + (cz * CT + 8 * j8) / 4) * hw + (long long)oy * p.Wo + ox;
      const long long hw4 = (long long)(p.Ho >> 1) * (p.Wo >> 1);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (i < nvalid) {
          float v8[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) v8[c] = acc[i][8 * j8 + c];
          uint4 hi, lo;
          split_pack8(v8, hi, lo);
          if (p.cells_s2d) {
            // space-to-depth: pixel (oy, ox+i) is sub-position (dy,dx) of block (oy/2, (ox+i)/2); its 8 channels are chunk
            // (dy*2+dx)*COUT/8 + c/8 of the 4*COUT-channel unshuffled map
            const int sub = (oy & 1) * 2 + ((ox + i) & 1);
            const int chunk = sub * (COUT / 8) + (cz * CT + 8 * j8) / 8;
            uint4* q = p.y_cells + ((long long)b * COUT + 2 * chunk) * hw4 + (long long)(oy >> 1) * (p.Wo >> 1) + ((ox + i) >> 1);
            q[0] = hi;
            q[hw4] = lo;
          } else {
            cp[i] = hi;
            cp[hw + i] = lo;
          }
        }
      }
    }
  }
  if (p.y_nhwc0) {
    constexpr int HALF = COUT / 2;
#pragma unroll
    for (int j4 = 0; j4 < CT / 4; ++j4) {
      const int co = cz * CT + 4 * j4;
      float* base = (co < HALF) ? p.y_nhwc0 : p.y_nhwc1;
      const int cc = (co < HALF) ? co : co - HALF;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (i < nvalid)
          *reinterpret_cast<float4*>(base + (((long long)b * p.Ho + oy) * p.Wo + ox + i) * HALF + cc) =
              make_float4(acc[i][4 * j4], acc[i][4 * j4 + 1], acc[i][4 * j4 + 2], acc[i][4 * j4 + 3]);
    }
  }
}

template <int K, int S, int CIN, int COUT>
static int launch_feat(const FeatConvParams& p, cudaStream_t st) {
  using Cfg = FeatCfg<K, S, CIN, COUT>;
  static PerDevice state;  // per template instance; the opt-in is a per-device attribute
  const int slot = current_device_slot();
  if (slot < 0 || !state.configured[slot]) {
    cudaError_t e = cudaFuncSetAttribute(feat_conv_kernel<K, S, CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem);
    if (e != cudaSuccess) {
      set_error("conv2d: cannot reserve %zu bytes of shared memory: %s", Cfg::kSmem, cudaGetErrorString(e));
      return DMVS_ERR_CUDA;
    }
    if (slot >= 0) state.configured[slot] = true;
  }
  dim3 grid(ceil_div(p.Wo, 32), ceil_div(p.Ho, 32), p.B * Cfg::NZ);
  DMVS_REQUIRE(grid.y <= 65535 && grid.z <= 65535, DMVS_ERR_BAD_SHAPE, "conv2d: grid too large");
  feat_conv_kernel<K, S, CIN, COUT><<<grid, 256, Cfg::kSmem, st>>>(p);
  return check_launch("conv2d");
}

// 1x1 lateral with few input channels (inner2: 8 -> 32 at full resolution, + the upsampled coarser map): pure streaming,
// 192 bytes per pixel.  Thread = 4 x-consecutive pixels: CIN float4 loads up front, then COUT outputs in groups of 8
// (weights broadcast from shared memory), float4 stores.  No input tile: nothing is reused between threads.
template <int CIN, int COUT>
__global__ void __launch_bounds__(256) feat_pointwise_kernel(const __grid_constant__ FeatConvParams p) {
  __shared__ float s_w[CIN * COUT + COUT];
  for (int i = threadIdx.x; i < CIN * COUT; i += 256) s_w[i] = __ldg(p.w + i);  // [CIN][COUT]
  for (int i = threadIdx.x; i < COUT; i += 256) s_w[CIN * COUT + i] = p.shift ? __ldg(p.shift + i) : 0.0f;
  __syncthreads();
  const int w4 = p.Wo >> 2;  // Wo % 4 == 0 (checked by the launcher)
  const long long n4 = (long long)p.Ho * w4;
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t >= n4) return;
  const int b = blockIdx.y;
  const int oy = (int)(t / w4), ox = (int)(t - (long long)oy * w4) * 4;
  const long long hw = (long long)p.Ho * p.Wo;
  const float* xp = p.x + (long long)b * CIN * hw + (long long)oy * p.Wo + ox;
  float4 xin[CIN];
#pragma unroll
  for (int c = 0; c < CIN; ++c) xin[c] = __ldg(reinterpret_cast<const float4*>(xp + c * hw));
  const int hw2 = (p.Ho >> 1) * (p.Wo >> 1);
  const float* up = p.up_add ? p.up_add + (long long)b * COUT * hw2 + (long long)(oy >> 1) * (p.Wo >> 1) + (ox >> 1) : nullptr;
  float* yp = p.y_nchw ? p.y_nchw + (long long)b * COUT * hw + (long long)oy * p.Wo + ox : nullptr;
  uint4* cp = p.y_cells ? p.y_cells + (long long)b * (COUT / 4) * hw + (long long)oy * p.Wo + ox : nullptr;
#pragma unroll 1
  for (int g = 0; g < COUT; g += 8) {
    float2 u[8];
    if (up) {
#pragma unroll
      for (int j = 0; j < 8; ++j) u[j] = __ldg(reinterpret_cast<const float2*>(up + (long long)(g + j) * hw2));
    }
    float4 acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int c = 0; c < CIN; ++c) {
      const float4 wa = *reinterpret_cast<const float4*>(s_w + c * COUT + g), wb = *reinterpret_cast<const float4*>(s_w + c * COUT + g + 4);
      const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[j].x = fmaf(xin[c].x, wv[j], acc[j].x);
        acc[j].y = fmaf(xin[c].y, wv[j], acc[j].y);
        acc[j].z = fmaf(xin[c].z, wv[j], acc[j].z);
        acc[j].w = fmaf(xin[c].w, wv[j], acc[j].w);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float sc = p.scale ? __ldg(p.scale + g + j) : 1.0f, sh = s_w[CIN * COUT + g + j];
      float4 v = make_float4(fmaf(acc[j].x, sc, sh), fmaf(acc[j].y, sc, sh), fmaf(acc[j].z, sc, sh), fmaf(acc[j].w, sc, sh));
      if (p.relu) v = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
      if (up) { v.x += u[j].x; v.y += u[j].x; v.z += u[j].y; v.w += u[j].y; }
      if (yp) *reinterpret_cast<float4*>(yp + (long long)(g + j) * hw) = v;
      acc[j] = v;
    }
    if (cp) {  // the 8 channels of this group are one (hi, lo) cell pair per pixel; planes 2*(g/8) and 2*(g/8)+1
      uint4* c0p = cp + (long long)(g / 4) * hw;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float v8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v8[j] = (i == 0) ? acc[j].x : (i == 1) ? acc[j].y : (i == 2) ? acc[j].z : acc[j].w;
        uint4 hi, lo;
        split_pack8(v8, hi, lo);
        c0p[i] = hi;
        c0p[hw + i] = lo;
      }
    }
  }
}

template <int CIN, int COUT>
static int launch_pointwise(const FeatConvParams& p, cudaStream_t st) {
  const long long n4 = (long long)p.Ho * (p.Wo >> 2);
  dim3 grid((unsigned)((n4 + 255) / 256), p.B, 1);
  DMVS_REQUIRE(grid.y <= 65535, DMVS_ERR_BAD_SHAPE, "conv2d: batch too large");
  feat_pointwise_kernel<CIN, COUT><<<grid, 256, 0, st>>>(p);
  return check_launch("conv2d_pointwise");
}

// fp32 NCHW -> CH16 cells of the 2x2 pixel-unshuffled map (input of a 5x5 stride-2 layer run as a 3x3 stride-1 layer on
// 4*C channels): one thread per (b, 8-channel chunk, y, x)
__global__ void __launch_bounds__(256) f32_to_s2d_cells_kernel(const float* __restrict__ x, uint4* __restrict__ y, int C, int H, int W,
                                                               long long n) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const int xx = (int)(i % W);
  long long r = i / W;
  const int yy = (int)(r % H); r /= H;
  const int j = (int)(r % (C / 8));
  const int b = (int)(r / (C / 8));
  const long long hw = (long long)H * W, hw4 = (long long)(H >> 1) * (W >> 1);
  const float* src = x + ((long long)b * C + j * 8) * hw + (long long)yy * W + xx;
  float v[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) v[c] = __ldg(src + c * hw);
  uint4 hi, lo;
  split_pack8(v, hi, lo);
  const int chunk = ((yy & 1) * 2 + (xx & 1)) * (C / 8) + j;
  uint4* q = y + ((long long)b * C + 2 * chunk) * hw4 + (long long)(yy >> 1) * (W >> 1) + (xx >> 1);
  q[0] = hi;
  q[hw4] = lo;
}

}  // namespace dmvs

extern "C" int dmvs_features_s2d_cells_f32(const float* x, void* y_cells, int B, int C, int H, int W, void* stream) {
  using namespace dmvs;
  DMVS_REQUIRE(x && y_cells && aligned16(y_cells), DMVS_ERR_BAD_POINTER, "features_s2d_cells: null or misaligned pointer");
  DMVS_REQUIRE(B >= 1 && C >= 8 && C % 8 == 0 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0, DMVS_ERR_BAD_SHAPE,
               "features_s2d_cells: need C %% 8 == 0 and even H, W (B=%d C=%d H=%d W=%d)", B, C, H, W);
  const long long n = (long long)B * (C / 8) * H * W;
  f32_to_s2d_cells_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, reinterpret_cast<uint4*>(y_cells), C, H, W, n);
  return check_launch("features_s2d_cells");
}

extern "C" int dmvs_conv2d_f32(const float* x, const float* w, const float* scale, const float* shift, const float* up_add,
                               float* y_nchw, float* y_nhwc0, float* y_nhwc1, void* y_cells, int cells_s2d, int B, int Cin, int Cout, int Hi,
                               int Wi, int K, int stride, int relu, void* stream) {
  using namespace dmvs;
  DMVS_REQUIRE(x && w && (y_nchw || (y_nhwc0 && y_nhwc1) || y_cells), DMVS_ERR_BAD_POINTER, "conv2d: null pointer");
  DMVS_REQUIRE(!y_cells || aligned16(y_cells), DMVS_ERR_BAD_POINTER, "conv2d: y_cells must be 16-byte aligned");
  DMVS_REQUIRE((y_nhwc0 == nullptr) == (y_nhwc1 == nullptr), DMVS_ERR_BAD_POINTER, "conv2d: both channel-last outputs or none");
  DMVS_REQUIRE((!y_nhwc0 || aligned16(y_nhwc0)) && (!y_nhwc1 || aligned16(y_nhwc1)) && (!y_nchw || aligned16(y_nchw)),
               DMVS_ERR_BAD_POINTER, "conv2d: outputs must be 16-byte aligned");
  DMVS_REQUIRE(B >= 1 && Hi >= 1 && Wi >= 1, DMVS_ERR_BAD_SHAPE, "conv2d: bad dims B=%d Hi=%d Wi=%d", B, Hi, Wi);
  DMVS_REQUIRE((K == 1 || K == 3 || K == 5) && (stride == 1 || stride == 2), DMVS_ERR_BAD_SHAPE, "conv2d: K=%d stride=%d unsupported", K, stride);
  FeatConvParams p;
  p.x = x; p.w = w; p.scale = scale; p.shift = shift; p.up_add = up_add;
  p.y_nchw = y_nchw; p.y_nhwc0 = y_nhwc0; p.y_nhwc1 = y_nhwc1; p.y_cells = reinterpret_cast<uint4*>(y_cells);
  p.B = B; p.Hi = Hi; p.Wi = Wi; p.relu = relu; p.cells_s2d = cells_s2d;
  const int pad = K / 2;
  p.Ho = (Hi + 2 * pad - K) / stride + 1;
  p.Wo = (Wi + 2 * pad - K) / stride + 1;
  DMVS_REQUIRE(!cells_s2d || (y_cells && (p.Ho % 2) == 0 && (p.Wo % 2) == 0 && Cout % 8 == 0 && K != 1), DMVS_ERR_BAD_SHAPE,
               "conv2d: the pixel-unshuffled cell output needs y_cells, even output size and a tiled (K > 1) layer");
  DMVS_REQUIRE(!up_add || ((p.Ho % 2) == 0 && (p.Wo % 2) == 0), DMVS_ERR_BAD_SHAPE, "conv2d: up_add needs even output size, got %dx%d", p.Ho, p.Wo);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int key = ((K * 10 + stride) * 100 + Cin) * 100 + Cout;
  switch (key) {
    case 310308: return launch_feat<3, 1, 3, 8>(p, st);     // conv0.0
    case 310808: return launch_feat<3, 1, 8, 8>(p, st);     // conv0.1
    case 520816: return launch_feat<5, 2, 8, 16>(p, st);    // conv1.0
    case 311616: return launch_feat<3, 1, 16, 16>(p, st);   // conv1.1, conv1.2
    case 521632: return launch_feat<5, 2, 16, 32>(p, st);   // conv2.0
    case 313232: return launch_feat<3, 1, 32, 32>(p, st);   // conv2.1, conv2.2, out2
    case 313216: return launch_feat<3, 1, 32, 16>(p, st);   // out3
    case 113264: return launch_feat<1, 1, 32, 64>(p, st);   // out1
    case 111632:                                            // inner1
      if (!y_nhwc0 && (p.Wo & 3) == 0 && aligned16(x) && (!up_add || (((p.Wo >> 1) & 1) == 0 && (reinterpret_cast<uintptr_t>(up_add) & 7u) == 0)))
        return launch_pointwise<16, 32>(p, st);
      return launch_feat<1, 1, 16, 32>(p, st);
    case 110832:                                            // inner2
      // streaming kernel when it applies (NCHW output only, 4-pixel groups aligned); the tiled kernel otherwise
      if (!y_nhwc0 && (p.Wo & 3) == 0 && aligned16(x) && (!up_add || (((p.Wo >> 1) & 1) == 0 && (reinterpret_cast<uintptr_t>(up_add) & 7u) == 0)))
        return launch_pointwise<8, 32>(p, st);
      return launch_feat<1, 1, 8, 32>(p, st);
    default:
      set_error("conv2d: (K=%d, stride=%d, Cin=%d, Cout=%d) is not a FeatureNet layer shape", K, stride, Cin, Cout);
      return DMVS_ERR_BAD_SHAPE;
  }
}
