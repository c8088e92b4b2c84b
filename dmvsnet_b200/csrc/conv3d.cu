// R1 (fp32 CUDA-core path): 3x3x3 / 1x3x3 convolutions, stride 1|2, and the stride-2 transposed
// convolutions of the regularisation U-Nets, with eval-BatchNorm + ReLU + skip-add fused in the epilogue.
//
// Replaces reference networks/module.py:120-208 (Conv3d / Deconv3d / Conv2d / Deconv2d blocks) as used by
// CostRegNet_part (module.py:358-398) and CostRegNet_part_refine (module.py:400-436).
//
// Layout: activations stay NCDHW (the reference layout, so the cost volume and the logits need no
// transposition); weights are repacked by the host to [tap][Cin][Cout] so that one broadcast LDS.128
// feeds four output channels.  One thread owns one (x, y) column position, TD consecutive output planes
// and COUT_T output channels; a warp spans 32 consecutive x so every activation load/store is one
// coalesced line.  Weights are staged through shared memory in chunks of CI_T input channels.
//
// This is the exact-fp32 path (parity to ~1e-6); the tcgen05 implicit-GEMM path is layered on top of the
// same entry points.
#include "common.cuh"

namespace dmvs {

struct ConvParams {
  const float* x;
  const float* w;      // [taps][Cin][Cout]
  const float* scale;  // [Cout] or null
  const float* shift;  // [Cout] or null
  const float* skip;   // like y, or null
  float* y;
  long long x_bs, y_bs, skip_bs;  // batch strides in elements
  int B, Cin, Cout, Di, Hi, Wi, Do, Ho, Wo;
  int Cout_w;  // Cout rounded up to a multiple of 4: the channel stride of the packed weights
  int relu;
  int n_zblocks, n_cgroups;
};

constexpr int CI_T = 8;

template <int COUT_T>
__device__ __forceinline__ void stage_weights(float* sw, const ConvParams& p, int taps, int ci0, int co0) {
  // smem layout [ci][tap][COUT_T]; global [tap][Cin][Cout]
  const int n = CI_T * taps * COUT_T;
  for (int i = threadIdx.y * 32 + threadIdx.x; i < n; i += 128) {
    const int co = i % COUT_T;
    const int t = (i / COUT_T) % taps;
    const int ci = i / (COUT_T * taps);
    const int gci = ci0 + ci;
    sw[i] = (gci < p.Cin) ? __ldg(p.w + ((long long)t * p.Cin + gci) * p.Cout_w + co0 + co) : 0.f;
  }
}

template <int COUT_T, int TD>
__device__ __forceinline__ void epilogue(const ConvParams& p, float (&acc)[TD][COUT_T], int b, int co0, int z0, int y, int x) {
  const long long plane = (long long)p.Ho * p.Wo;
#pragma unroll
  for (int t = 0; t < TD; ++t) {
    const int z = z0 + t;
    if (z >= p.Do) break;
#pragma unroll
    for (int c = 0; c < COUT_T; ++c) {
      const int co = co0 + c;
      if (co >= p.Cout) break;
      float v = acc[t][c];
      if (p.scale) v = fmaf(v, __ldg(p.scale + co), __ldg(p.shift + co));
      if (p.relu) v = fmaxf(v, 0.f);
      const long long off = ((long long)co * p.Do + z) * plane + (long long)y * p.Wo + x;
      if (p.skip) v += __ldg(p.skip + (long long)b * p.skip_bs + off);
      p.y[(long long)b * p.y_bs + off] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// forward convolution, kernel KD x 3 x 3, padding (KD/2, 1, 1)
template <int COUT_T, int KD, int STRIDE, int TD>
__global__ void __launch_bounds__(128) conv_fwd_kernel(const __grid_constant__ ConvParams p) {
  static_assert(STRIDE == 1 || TD == 1, "plane blocking only for stride 1");
  constexpr int TAPS = KD * 9;
  constexpr int NPL = (KD == 3) ? (TD + 2) : TD;  // input planes touched by TD output planes (stride 1)
  extern __shared__ float sw[];
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = blockIdx.y * 4 + threadIdx.y;
  int bz = blockIdx.z;
  const int cg = bz % p.n_cgroups; bz /= p.n_cgroups;
  const int zb = bz % p.n_zblocks;
  const int b = bz / p.n_zblocks;
  const int z0 = zb * TD;
  const int co0 = cg * COUT_T;
  const bool active = (x < p.Wo) && (y < p.Ho);
  const long long iplane = (long long)p.Hi * p.Wi;

  float acc[TD][COUT_T];
#pragma unroll
  for (int t = 0; t < TD; ++t)
#pragma unroll
    for (int c = 0; c < COUT_T; ++c) acc[t][c] = 0.f;

  for (int ci0 = 0; ci0 < p.Cin; ci0 += CI_T) {
    __syncthreads();
    stage_weights<COUT_T>(sw, p, TAPS, ci0, co0);
    __syncthreads();
    if (!active) continue;
    const int nci = min(CI_T, p.Cin - ci0);
    for (int ci = 0; ci < nci; ++ci) {
      const float* xin = p.x + (long long)b * p.x_bs + (long long)(ci0 + ci) * p.Di * iplane;
      const float* wci = sw + ci * TAPS * COUT_T;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int iy = y * STRIDE + kh - 1;
        const bool vy = (iy >= 0) && (iy < p.Hi);
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int ix = x * STRIDE + kw - 1;
          const bool vxy = vy && (ix >= 0) && (ix < p.Wi);
          float v[NPL];
#pragma unroll
          for (int j = 0; j < NPL; ++j) {
            const int iz = (KD == 3) ? (z0 * STRIDE - 1 + j) : (z0 + j);
            v[j] = (vxy && iz >= 0 && iz < p.Di) ? __ldg(xin + (long long)iz * iplane + (long long)iy * p.Wi + ix) : 0.f;
          }
#pragma unroll
          for (int kd = 0; kd < KD; ++kd) {
            const float4* wp = reinterpret_cast<const float4*>(wci + ((kd * 3 + kh) * 3 + kw) * COUT_T);
#pragma unroll
            for (int q = 0; q < COUT_T / 4; ++q) {
              const float4 ww = wp[q];
#pragma unroll
              for (int t = 0; t < TD; ++t) {
                const float a = v[kd + t];
                acc[t][4 * q + 0] = fmaf(a, ww.x, acc[t][4 * q + 0]);
                acc[t][4 * q + 1] = fmaf(a, ww.y, acc[t][4 * q + 1]);
                acc[t][4 * q + 2] = fmaf(a, ww.z, acc[t][4 * q + 2]);
                acc[t][4 * q + 3] = fmaf(a, ww.w, acc[t][4 * q + 3]);
              }
            }
          }
        }
      }
    }
  }
  if (active) epilogue<COUT_T, TD>(p, acc, b, co0, z0, y, x);
}

// ---------------------------------------------------------------------------------------------
// transposed convolution k=3, s=2, p=1, output_padding=1 (output exactly 2x), gather form:
//   out[o] = sum_{k, i : 2i - 1 + k = o} in[i] * W[k]   ->   even o: k=1, i=o/2;  odd o: k=0, i=(o+1)/2 and k=2, i=(o-1)/2
// One thread owns input column i, i.e. the output pair (2i, 2i+1): all lanes of a warp then use the same
// taps (row / plane parity is warp-uniform) and the weight reads stay broadcasts.
template <int COUT_T, int KD>
__global__ void __launch_bounds__(128) conv_tr_kernel(const __grid_constant__ ConvParams p) {
  constexpr int TAPS = KD * 9;
  extern __shared__ float sw[];
  const int i = blockIdx.x * 32 + threadIdx.x;  // input column
  const int oy = blockIdx.y * 4 + threadIdx.y;
  int bz = blockIdx.z;
  const int cg = bz % p.n_cgroups; bz /= p.n_cgroups;
  const int oz = bz % p.n_zblocks;
  const int b = bz / p.n_zblocks;
  const int co0 = cg * COUT_T;
  const bool active = (i < p.Wi) && (oy < p.Ho);
  const long long iplane = (long long)p.Hi * p.Wi;

  // taps along y and z for this output row / plane (warp-uniform)
  int nky, ky[2], iy[2];
  if ((oy & 1) == 0) { nky = 1; ky[0] = 1; iy[0] = oy >> 1; ky[1] = 0; iy[1] = 0; }
  else { nky = 2; ky[0] = 2; iy[0] = (oy - 1) >> 1; ky[1] = 0; iy[1] = (oy + 1) >> 1; if (iy[1] >= p.Hi) nky = 1; }
  int nkz, kz[2], iz[2];
  if (KD == 1) { nkz = 1; kz[0] = 0; iz[0] = oz; kz[1] = 0; iz[1] = 0; }
  else if ((oz & 1) == 0) { nkz = 1; kz[0] = 1; iz[0] = oz >> 1; kz[1] = 0; iz[1] = 0; }
  else { nkz = 2; kz[0] = 2; iz[0] = (oz - 1) >> 1; kz[1] = 0; iz[1] = (oz + 1) >> 1; if (iz[1] >= p.Di) nkz = 1; }

  float acc[2][COUT_T];  // [0]: output 2i, [1]: output 2i+1
#pragma unroll
  for (int c = 0; c < COUT_T; ++c) acc[0][c] = acc[1][c] = 0.f;

  for (int ci0 = 0; ci0 < p.Cin; ci0 += CI_T) {
    __syncthreads();
    stage_weights<COUT_T>(sw, p, TAPS, ci0, co0);
    __syncthreads();
    if (!active) continue;
    const int nci = min(CI_T, p.Cin - ci0);
    for (int ci = 0; ci < nci; ++ci) {
      const float* xin = p.x + (long long)b * p.x_bs + (long long)(ci0 + ci) * p.Di * iplane;
      const float* wci = sw + ci * TAPS * COUT_T;
      for (int a = 0; a < nkz; ++a) {
        for (int c2 = 0; c2 < nky; ++c2) {
          const float* row = xin + (long long)iz[a] * iplane + (long long)iy[c2] * p.Wi;
          const float v0 = __ldg(row + i);
          const float v1 = (i + 1 < p.Wi) ? __ldg(row + i + 1) : 0.f;
          const float* wt = wci + ((kz[a] * 3 + ky[c2]) * 3) * COUT_T;
          const float4* w0 = reinterpret_cast<const float4*>(wt);               // kw = 0: odd output, input i+1
          const float4* w1 = reinterpret_cast<const float4*>(wt + COUT_T);      // kw = 1: even output, input i
          const float4* w2 = reinterpret_cast<const float4*>(wt + 2 * COUT_T);  // kw = 2: odd output, input i
#pragma unroll
          for (int q = 0; q < COUT_T / 4; ++q) {
            const float4 a0 = w0[q], a1 = w1[q], a2 = w2[q];
            acc[0][4 * q + 0] = fmaf(v0, a1.x, acc[0][4 * q + 0]);
            acc[0][4 * q + 1] = fmaf(v0, a1.y, acc[0][4 * q + 1]);
            acc[0][4 * q + 2] = fmaf(v0, a1.z, acc[0][4 * q + 2]);
            acc[0][4 * q + 3] = fmaf(v0, a1.w, acc[0][4 * q + 3]);
            acc[1][4 * q + 0] = fmaf(v0, a2.x, fmaf(v1, a0.x, acc[1][4 * q + 0]));
            acc[1][4 * q + 1] = fmaf(v0, a2.y, fmaf(v1, a0.y, acc[1][4 * q + 1]));
            acc[1][4 * q + 2] = fmaf(v0, a2.z, fmaf(v1, a0.z, acc[1][4 * q + 2]));
            acc[1][4 * q + 3] = fmaf(v0, a2.w, fmaf(v1, a0.w, acc[1][4 * q + 3]));
          }
        }
      }
    }
  }
  if (!active) return;
  const long long plane = (long long)p.Ho * p.Wo;
#pragma unroll
  for (int c = 0; c < COUT_T; ++c) {
    const int co = co0 + c;
    if (co >= p.Cout) break;
    float e = acc[0][c], o = acc[1][c];
    if (p.scale) {
      const float s = __ldg(p.scale + co), t = __ldg(p.shift + co);
      e = fmaf(e, s, t);
      o = fmaf(o, s, t);
    }
    if (p.relu) { e = fmaxf(e, 0.f); o = fmaxf(o, 0.f); }
    const long long off = ((long long)co * p.Do + oz) * plane + (long long)oy * p.Wo + 2 * i;
    if (p.skip) {
      const float2 sk = __ldg(reinterpret_cast<const float2*>(p.skip + (long long)b * p.skip_bs + off));
      e += sk.x; o += sk.y;
    }
    *reinterpret_cast<float2*>(p.y + (long long)b * p.y_bs + off) = make_float2(e, o);
  }
}

// ---------------------------------------------------------------------------------------------
template <int COUT_T, int KD, int STRIDE, int TD>
static int launch_fwd(ConvParams p, cudaStream_t st) {
  p.n_zblocks = ceil_div(p.Do, TD);
  p.n_cgroups = p.Cout_w / COUT_T;
  dim3 grid(ceil_div(p.Wo, 32), ceil_div(p.Ho, 4), p.B * p.n_zblocks * p.n_cgroups);
  DMVS_REQUIRE(grid.y <= 65535 && grid.z <= 65535, DMVS_ERR_BAD_SHAPE, "conv3d: grid too large");
  const size_t smem = sizeof(float) * CI_T * KD * 9 * COUT_T;
  conv_fwd_kernel<COUT_T, KD, STRIDE, TD><<<grid, dim3(32, 4), smem, st>>>(p);
  return check_launch("conv3d_fwd");
}

template <int COUT_T, int KD>
static int launch_tr(ConvParams p, cudaStream_t st) {
  p.n_zblocks = p.Do;
  p.n_cgroups = p.Cout_w / COUT_T;
  dim3 grid(ceil_div(p.Wi, 32), ceil_div(p.Ho, 4), p.B * p.n_zblocks * p.n_cgroups);
  DMVS_REQUIRE(grid.y <= 65535 && grid.z <= 65535, DMVS_ERR_BAD_SHAPE, "deconv3d: grid too large");
  const size_t smem = sizeof(float) * CI_T * KD * 9 * COUT_T;
  conv_tr_kernel<COUT_T, KD><<<grid, dim3(32, 4), smem, st>>>(p);
  return check_launch("conv3d_tr");
}

template <int COUT_T>
static int dispatch(const ConvParams& p, int kd, int stride, int transposed, cudaStream_t st) {
  if (transposed) return kd == 3 ? launch_tr<COUT_T, 3>(p, st) : launch_tr<COUT_T, 1>(p, st);
  if (kd == 3) {
    if (stride == 2) return launch_fwd<COUT_T, 3, 2, 1>(p, st);
    if (p.Do >= 2 && COUT_T <= 16) return launch_fwd<COUT_T, 3, 1, 2>(p, st);
    return launch_fwd<COUT_T, 3, 1, 1>(p, st);
  }
  return stride == 2 ? launch_fwd<COUT_T, 1, 2, 1>(p, st) : launch_fwd<COUT_T, 1, 1, 1>(p, st);
}

int conv_layer(const float* x, long long x_bs, const dmvs_conv_layer& L, const float* skip, long long skip_bs, float* y,
               long long y_bs, int B, int Cin, int Cout, int Di, int Hi, int Wi, int kd, int stride, int transposed, int relu,
               cudaStream_t st) {
  DMVS_REQUIRE(x && y && L.w, DMVS_ERR_BAD_POINTER, "conv3d: null pointer");
  DMVS_REQUIRE((L.scale == nullptr) == (L.shift == nullptr), DMVS_ERR_BAD_POINTER, "conv3d: scale and shift go together");
  DMVS_REQUIRE(kd == 1 || kd == 3, DMVS_ERR_BAD_SHAPE, "conv3d: kd=%d (1 or 3)", kd);
  DMVS_REQUIRE(stride == 1 || stride == 2, DMVS_ERR_BAD_SHAPE, "conv3d: stride=%d (1 or 2)", stride);
  DMVS_REQUIRE(!transposed || stride == 2, DMVS_ERR_BAD_SHAPE, "conv3d: transposed convs are stride 2");
  DMVS_REQUIRE(B >= 1 && Cin >= 1 && Di >= 1 && Hi >= 1 && Wi >= 1, DMVS_ERR_BAD_SHAPE, "conv3d: bad dims");
  ConvParams p;
  p.x = x; p.w = L.w; p.scale = L.scale; p.shift = L.shift; p.skip = skip; p.y = y;
  p.x_bs = x_bs; p.y_bs = y_bs; p.skip_bs = skip_bs;
  p.B = B; p.Cin = Cin; p.Cout = Cout; p.Di = Di; p.Hi = Hi; p.Wi = Wi; p.relu = relu;
  if (transposed) {
    p.Do = (kd == 3) ? 2 * Di : Di; p.Ho = 2 * Hi; p.Wo = 2 * Wi;
    DMVS_REQUIRE(aligned16(y) && (y_bs % 2 == 0) && (!skip || (aligned16(skip) && skip_bs % 2 == 0)), DMVS_ERR_BAD_POINTER,
                 "deconv3d: y/skip must be 16-byte aligned");
  } else if (stride == 2) {
    p.Do = (kd == 3) ? (Di - 1) / 2 + 1 : Di; p.Ho = (Hi - 1) / 2 + 1; p.Wo = (Wi - 1) / 2 + 1;
  } else {
    p.Do = Di; p.Ho = Hi; p.Wo = Wi;
  }
  p.n_zblocks = p.n_cgroups = 1;
  p.Cout_w = (Cout + 3) & ~3;
  if (p.Cout_w % 32 == 0) return dispatch<32>(p, kd, stride, transposed, st);
  if (p.Cout_w % 16 == 0) return dispatch<16>(p, kd, stride, transposed, st);
  if (p.Cout_w % 8 == 0) return dispatch<8>(p, kd, stride, transposed, st);
  return dispatch<4>(p, kd, stride, transposed, st);
}

}  // namespace dmvs
