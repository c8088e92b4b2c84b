// Pieces shared by the tensor-core convolution kernels (conv_tc2.cu, conv_kf.cu): formats, launch parameters, cell addressing,
// tensor maps over the cell layouts.
#pragma once
#include <cuda.h>
#include <string.h>

#include "tc_common.cuh"

namespace dmvs {

enum { FMT_F32 = 0, FMT_CH16 = 1, FMT_CH16P = 2, FMT_NHWC2 = 4, FMT_NHWC2H = 5 };
enum { M2_S1 = 0, M2_S2 = 1, M2_TR = 2, M2_C0 = 3, M2_PB = 4, M2_C0T = 5, M2_TRF = 6 };
__host__ __device__ constexpr bool is_c0(int mode) { return mode == M2_C0 || mode == M2_C0T; }
// TRF: transposed conv with the 27 taps folded by input shift: the taps that read the same shifted A view (shift in {0,1}^3,
// 8 of them) become ONE MMA whose N spans the 8 parity-class accumulators (zero weight columns where a class has no tap for
// that shift).  Every MMA re-reads its 4 KB A tile from shared memory at 64 B/clk whatever its N, so 8 MMAs instead of 27.
__host__ __device__ constexpr bool is_tr(int mode) { return mode == M2_TR || mode == M2_TRF; }
constexpr int T_H = 16, T_W = 8;
// issuing threads per tile for the prob / conv0 kernels.  2 paid while every tcgen05.mma cost the issuing thread ~11 instructions; with the
// elected-lane issue (one UTCHMMA per MMA) one thread per accumulator set is faster (R1 6.76 -> 6.66 ms per DTU view)
#ifndef SUBISSUE
#define SUBISSUE 1
#endif
// issuing warps per kernel (tiles alternate between them).  Same story: with the elected-lane issue ONE warp keeps up and the second only
// takes issue slots from the epilogue (R1 6.66 -> 6.52 ms per DTU view)
#ifndef TC2_MMA_WARPS
#define TC2_MMA_WARPS 1
#endif

struct Tc2Params {
  const float* x_f32;  // C0 only
  const uint4* wtc;
  const uint4* wtc_wide;  // `prob` only: image for the wide-tile kernel (dmvs_conv_layer.w_tc_kw), may be null
  const float* scale;
  const float* shift;
  const uint4* skip;  // CH16P, TR only
  void* y;
  long long x_bs;     // C0 only (elements)
  long long y_bs;     // fp32 output only: batch stride in elements (the logits tensor interleaves two branches)
  int B, Cin, Cout, Di, Hi, Wi, Do, Ho, Wo;
  int relu, out_fmt;
  int dbg;            // experiments (dmvs_debug_set("kf_dbg", bits)): 1 = epilogue releases without loading / storing, 2 = issuers commit without MMAs
  long long* trace;   // experiments: clock64() stamps of CTA 0's pipeline events (dmvs_debug_set_ptr("kf_trace", buffer), kf_dbg bit 3)
  int skip_prefetch;  // transposed layers: L2 prefetch of the next tile's skip cells
  int tiles_x, tiles_y, tiles_z, n_tiles;
};

__host__ __device__ constexpr int pad128(int v) { return (v + 127) / 128 * 128; }
__host__ __device__ constexpr int pow2c(int c) { return c <= 32 ? 32 : c <= 64 ? 64 : c <= 128 ? 128 : c <= 256 ? 256 : 512; }

// ------------------------------------------------------------------------------------------------ cell helpers
__device__ __forceinline__ void unpack_cell(const uint4& c, float (&v)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&c);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
// cell index of voxel (z, y, x) of plane `pl` in a CH16 / CH16P tensor with dims (D, H, W) and NP planes per batch entry
__device__ __forceinline__ long long cell_index(int fmt, int b, int np, int pl, int D, int H, int W, int z, int y, int x) {
  const long long row = ((long long)(b * np + pl) * D + z) * H + y;
  if (fmt == FMT_CH16P) {
    const int we = (W + 1) >> 1;
    return (row * 2 + (x & 1)) * we + (x >> 1);
  }
  return row * W + x;
}

// ------------------------------------------------------------------------------------------------ tensor maps (host)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  }
  return fn;
}

// tensor map over a CH16 (rank 4: {W*8 halfs, H, D, planes}) or CH16P (rank 5: {We*8, 2, H, D, planes}) tensor
inline int make_tmap(CUtensorMap* m, const void* base, int fmt, int planes, int D, int H, int W, int box_cells, int box_h, int box_d,
                     int box_planes = 1) {
  EncodeTiledFn enc = encode_fn();
  DMVS_REQUIRE(enc != nullptr, DMVS_ERR_CUDA, "conv_tc2: cuTensorMapEncodeTiled is not available from the driver");
  CUresult r;
  if (fmt == FMT_CH16) {
    const cuuint64_t dims[4] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)planes};
    const cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)D * H * W * 16};
    const cuuint32_t box[4] = {(cuuint32_t)box_cells * 8, (cuuint32_t)box_h, (cuuint32_t)box_d, (cuuint32_t)box_planes};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    const cuuint64_t we = (cuuint64_t)(W + 1) / 2;
    const cuuint64_t dims[5] = {we * 8, 2, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)planes};
    const cuuint64_t strides[4] = {we * 16, 2 * we * 16, (cuuint64_t)H * 2 * we * 16, (cuuint64_t)D * H * 2 * we * 16};
    const cuuint32_t box[5] = {(cuuint32_t)box_cells * 8, 1, (cuuint32_t)box_h, (cuuint32_t)box_d, 1};
    const cuuint32_t es[5] = {1, 1, 1, 1, 1};
    r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  DMVS_REQUIRE(r == CUDA_SUCCESS, DMVS_ERR_CUDA, "conv_tc2: cuTensorMapEncodeTiled failed (%d) for dims D=%d H=%d W=%d planes=%d", (int)r,
               D, H, W, planes);
  return DMVS_OK;
}


}  // namespace dmvs
