// N4: the geometric-consistency check of the depth-map fusion, the consumer of the path's output.
//
// Replaces reference filter/pcd.py:152-242 (reproject_with_depth_pytorch + check_geometric_consistency[_pytorch]) and the
// accumulation of filter_depth (pcd.py:283-304): for every reference pixel and source view
//   unproject with the reference depth -> source camera -> source pixel -> bilinear sample of the SOURCE depth map
//   (F.grid_sample, zeros padding, align_corners=True) -> unproject with the sampled depth -> back to the reference camera ->
//   reprojected depth and pixel -> mask = |p_reproj - p| < 1*alpha  and  |d_reproj - d| / d < 0.01*alpha,
// fused over all S source views in one pass: the reference depth map is read once, nothing but the requested outputs is
// written (the reference materialises ~12 [H*W]-sized temporaries per source view and copies them to the host).
// One thread = one reference pixel; the six small matrices per source (precomputed on the host with the reference's own
// torch calls) sit in shared memory.  fp32 throughout, dot products as FMA chains.
#include "common.cuh"

namespace dmvs {

constexpr int kGeoMats = 60;  // inv(K_ref) 9 | (E_src inv(E_ref))[:3,:4] 12 | K_src 9 | inv(K_src) 9 | (E_ref inv(E_src))[:3,:4] 12 | K_ref 9
constexpr int kGeoMaxSrc = 32;

__device__ __forceinline__ void mat3(const float* m, float a, float b, float c, float& x, float& y, float& z) {
  x = fmaf(m[2], c, fmaf(m[1], b, m[0] * a));
  y = fmaf(m[5], c, fmaf(m[4], b, m[3] * a));
  z = fmaf(m[8], c, fmaf(m[7], b, m[6] * a));
}
__device__ __forceinline__ void mat34(const float* m, float a, float b, float c, float& x, float& y, float& z) {
  x = fmaf(m[2], c, fmaf(m[1], b, m[0] * a)) + m[3];
  y = fmaf(m[6], c, fmaf(m[5], b, m[4] * a)) + m[7];
  z = fmaf(m[10], c, fmaf(m[9], b, m[8] * a)) + m[11];
}

__global__ void __launch_bounds__(256) geo_consistency_kernel(const float* __restrict__ depth_ref, const float* __restrict__ depth_src,
                                                              const float* __restrict__ mats, int S, int H, int W, float dist_th,
                                                              float rel_th, unsigned char* __restrict__ mask,
                                                              float* __restrict__ depth_reproj, float* __restrict__ xy_src,
                                                              int* __restrict__ mask_sum, float* __restrict__ depth_avg) {
  __shared__ float sm[kGeoMaxSrc * kGeoMats];
  for (int i = threadIdx.x; i < S * kGeoMats; i += 256) sm[i] = __ldg(mats + i);
  __syncthreads();
  const long long hw = (long long)H * W;
  const long long pix = (long long)blockIdx.x * 256 + threadIdx.x;
  if (pix >= hw) return;
  const int y = (int)(pix / W), x = (int)(pix - (long long)y * W);
  const float fx = (float)x, fy = (float)y;
  const float d0 = __ldg(depth_ref + pix);
  const float dref = (d0 == 0.0f) ? 1e-4f : d0;             // pcd.py:212: depth_ref[depth_ref == 0] = 1e-4 (before the ratio)
  const float half_w = (float)((double)(W - 1) / 2.0), half_h = (float)((double)(H - 1) / 2.0);
  int count = 0;
  float dsum = 0.0f;
  for (int s = 0; s < S; ++s) {
    const float* m = sm + s * kGeoMats;
    // step 1: reference pixel -> source view (pcd.py:164-172); the homogeneous pixel is scaled by the depth first
    float X, Y, Z;
    mat3(m, fx * d0, fy * d0, d0, X, Y, Z);
    float Xs, Ys, Zs;
    mat34(m + 9, X, Y, Z, Xs, Ys, Zs);
    float kx, ky, kz;
    mat3(m + 21, Xs, Ys, Zs, kx, ky, kz);
    const float u = kx / kz, v = ky / kz;
    // step 2: sample the source depth map (pcd.py:176-179; grid_sample bilinear / zeros / align_corners=True)
    const float un = u / half_w - 1.0f, vn = v / half_h - 1.0f;
    const float ix = (un + 1.0f) * half_w, iy = (vn + 1.0f) * half_h;
    const float x0f = floorf(ix), y0f = floorf(iy);
    const float wx1 = ix - x0f, wy1 = iy - y0f, wx0 = 1.0f - wx1, wy0 = 1.0f - wy1;
    float sd = 0.0f;
    const bool finite = fabsf(ix) < 1.0e9f && fabsf(iy) < 1.0e9f;
    if (finite) {
      const int x0 = (int)x0f, y0 = (int)y0f;
      const float* ds = depth_src + (long long)s * hw;
      const bool xa = x0 >= 0 && x0 < W, xb = x0 + 1 >= 0 && x0 + 1 < W, ya = y0 >= 0 && y0 < H, yb = y0 + 1 >= 0 && y0 + 1 < H;
      const float v00 = (xa && ya) ? __ldg(ds + (long long)y0 * W + x0) : 0.0f;
      const float v01 = (xb && ya) ? __ldg(ds + (long long)y0 * W + x0 + 1) : 0.0f;
      const float v10 = (xa && yb) ? __ldg(ds + (long long)(y0 + 1) * W + x0) : 0.0f;
      const float v11 = (xb && yb) ? __ldg(ds + (long long)(y0 + 1) * W + x0 + 1) : 0.0f;
      sd = v00 * (wx0 * wy0) + v01 * (wx1 * wy0) + v10 * (wx0 * wy1) + v11 * (wx1 * wy1);
    } else {
      sd = __int_as_float(0x7fc00000);  // the reference propagates NaN positions into the sample; the mask below rejects them
    }
    // back-projection with the sampled depth (pcd.py:186-198)
    mat3(m + 30, u * sd, v * sd, sd, X, Y, Z);
    float Xr, Yr, Zr;
    mat34(m + 39, X, Y, Z, Xr, Yr, Zr);
    float rx, ry, rz;
    mat3(m + 51, Xr, Yr, Zr, rx, ry, rz);
    if (rz == 0.0f) rz += 0.00001f;
    const float xr = rx / rz, yr = ry / rz;
    // consistency (pcd.py:209-221)
    const float ddx = xr - fx, ddy = yr - fy;
    const float dist = sqrtf(ddx * ddx + ddy * ddy);
    const float rel = fabsf(Zr - dref) / dref;
    const bool ok = (dist < dist_th) && (rel < rel_th);
    const float dr = ok ? Zr : 0.0f;
    if (mask) mask[(long long)s * hw + pix] = ok ? 1 : 0;
    if (depth_reproj) depth_reproj[(long long)s * hw + pix] = dr;
    if (xy_src) {
      xy_src[((long long)s * 2) * hw + pix] = un;
      xy_src[((long long)s * 2 + 1) * hw + pix] = vn;
    }
    count += ok ? 1 : 0;
    dsum += dr;  // python sum(): ((0 + d_1) + d_2) + ...
  }
  if (mask_sum) mask_sum[pix] = count;
  if (depth_avg) depth_avg[pix] = (dsum + dref) / (float)(count + 1);  // pcd.py:298 (the reference depth was patched in place)
}

// ---------------------------------------------------------------------------------------------------------------------------
// Dynamic-threshold variant: reference filter/dypcd_tanks.py:61-98 (reproject_with_depth, the numpy + cv2.remap form),
// :164-184 (check_geometric_consistency: nine threshold levels i = 2..10, dist < i*dist_base and rel < i*rel_diff_base) and the
// accumulation of filter_depth (:237-270).  What differs from the kernel above, because the reference's numpy code differs:
//   * the projective chain runs in float64 (numpy promotes int64 pixel grids x float32 depths to float64); the matrices stay the
//     float32 values numpy computed, promoted exactly
//   * the source depth is sampled like cv2.remap(INTER_LINEAR, BORDER_CONSTANT 0): coordinates rounded to 1/32 pixel
//     (cvRound(x * 32), ties to even), the four weights products of the 1-D float taps, float accumulation
//   * the pixel distance is float64 (float32 coordinates minus an int64 grid), the relative depth difference float32, zeros of the
//     reference depth are NOT patched (0/0 and x/0 fail every comparison)
//   * a (pixel, source) pair gets the smallest level it passes; masks[i-2] = level <= i; the reprojected depth is kept where level
//     10 passes; the fused mask is  OR_{i=2..S} (#sources passing level i) >= i  and the average is taken in float64.
__device__ __forceinline__ void dmat3(const float* m, double a, double b, double c, double& x, double& y, double& z) {
  x = fma((double)m[2], c, fma((double)m[1], b, (double)m[0] * a));
  y = fma((double)m[5], c, fma((double)m[4], b, (double)m[3] * a));
  z = fma((double)m[8], c, fma((double)m[7], b, (double)m[6] * a));
}
__device__ __forceinline__ void dmat34(const float* m, double a, double b, double c, double& x, double& y, double& z) {
  x = fma((double)m[2], c, fma((double)m[1], b, (double)m[0] * a)) + (double)m[3];
  y = fma((double)m[6], c, fma((double)m[5], b, (double)m[4] * a)) + (double)m[7];
  z = fma((double)m[10], c, fma((double)m[9], b, (double)m[8] * a)) + (double)m[11];
}

// cv2.remap, INTER_LINEAR, BORDER_CONSTANT(0) on a float32 image: fixed-point coordinates with 5 fractional bits
__device__ __forceinline__ float remap_linear_q5(const float* __restrict__ img, int H, int W, float xs, float ys) {
  if (!(fabsf(xs) < 3.0e7f) || !(fabsf(ys) < 3.0e7f)) return 0.0f;  // NaN / far outside (cvRound saturates): all four taps are border
  const int sx = __float2int_rn(xs * 32.0f), sy = __float2int_rn(ys * 32.0f);
  const int x0 = sx >> 5, y0 = sy >> 5;
  const float ax = (float)(sx & 31) * (1.0f / 32.0f), ay = (float)(sy & 31) * (1.0f / 32.0f);
  const float tx0 = 1.0f - ax, ty0 = 1.0f - ay;
  const bool xa = x0 >= 0 && x0 < W, xb = x0 + 1 >= 0 && x0 + 1 < W, ya = y0 >= 0 && y0 < H, yb = y0 + 1 >= 0 && y0 + 1 < H;
  const float v00 = (xa && ya) ? __ldg(img + (long long)y0 * W + x0) : 0.0f;
  const float v01 = (xb && ya) ? __ldg(img + (long long)y0 * W + x0 + 1) : 0.0f;
  const float v10 = (xa && yb) ? __ldg(img + (long long)(y0 + 1) * W + x0) : 0.0f;
  const float v11 = (xb && yb) ? __ldg(img + (long long)(y0 + 1) * W + x0 + 1) : 0.0f;
  return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(v00, ty0 * tx0), __fmul_rn(v01, ty0 * ax)), __fmul_rn(v10, ay * tx0)), __fmul_rn(v11, ay * ax));
}

__global__ void __launch_bounds__(256) geo_consistency_dynamic_kernel(const float* __restrict__ depth_ref, const float* __restrict__ depth_src,
                                                                      const float* __restrict__ mats, int S, int H, int W, double dist_base,
                                                                      double rel_base, unsigned char* __restrict__ level,
                                                                      float* __restrict__ depth_reproj, float* __restrict__ xy_src,
                                                                      int* __restrict__ mask_sum, unsigned char* __restrict__ geo_mask,
                                                                      float* __restrict__ depth_avg) {
  __shared__ float sm[kGeoMaxSrc * kGeoMats];
  for (int i = threadIdx.x; i < S * kGeoMats; i += 256) sm[i] = __ldg(mats + i);
  __syncthreads();
  const long long hw = (long long)H * W;
  const long long pix = (long long)blockIdx.x * 256 + threadIdx.x;
  if (pix >= hw) return;
  const int y = (int)(pix / W), x = (int)(pix - (long long)y * W);
  const float d0 = __ldg(depth_ref + pix);
  const double dx = (double)x, dy = (double)y, dd = (double)d0;
  int cnt[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) cnt[i] = 0;
  float dsum = 0.0f;
  for (int s = 0; s < S; ++s) {
    const float* m = sm + s * kGeoMats;
    double X, Y, Z, Xs, Ys, Zs, kx, ky, kz;
    dmat3(m, dx * dd, dy * dd, dd, X, Y, Z);            // :66-67
    dmat34(m + 9, X, Y, Z, Xs, Ys, Zs);                  // :69-70
    dmat3(m + 21, Xs, Ys, Zs, kx, ky, kz);               // :72
    const double u = kx / kz, v = ky / kz;               // :73
    const float xs = (float)u, ys = (float)v;            // :77-78
    const float sd = remap_linear_q5(depth_src + (long long)s * hw, H, W, xs, ys);  // :79
    const double sdd = (double)sd;
    dmat3(m + 30, u * sdd, v * sdd, sdd, X, Y, Z);       // :84-85
    double Xr, Yr, Zr, rx, ry, rz;
    dmat34(m + 39, X, Y, Z, Xr, Yr, Zr);                 // :87-88
    dmat3(m + 51, Xr, Yr, Zr, rx, ry, rz);               // :91
    if (rz == 0.0) rz += 0.00001;
    const float drep = (float)Zr, xr = (float)(rx / rz), yr = (float)(ry / rz);  // :90, :94-95
    // :170-175
    const double ex = (double)xr - dx, ey = (double)yr - dy;
    const double dist = sqrt(ex * ex + ey * ey);
    const float rel = __fdiv_rn(fabsf(__fsub_rn(drep, d0)), d0);
    int lv = 0;
#pragma unroll
    for (int i = 10; i >= 2; --i) {
      const bool ok = (dist < (double)i * dist_base) && (rel < (float)((double)i * rel_base));
      if (ok) lv = i;
      cnt[i - 2] += ok ? 1 : 0;
    }
    // thresholds grow with i, so the levels that pass form a suffix i >= lv and cnt[i-2] counts exactly masks[i-2]
    const bool ok10 = lv != 0;
    const float dr = ok10 ? drep : 0.0f;
    if (level) level[(long long)s * hw + pix] = (unsigned char)lv;
    if (depth_reproj) depth_reproj[(long long)s * hw + pix] = dr;
    if (xy_src) {
      xy_src[((long long)s * 2) * hw + pix] = xs;
      xy_src[((long long)s * 2 + 1) * hw + pix] = ys;
    }
    dsum = __fadd_rn(dsum, dr);
  }
  const int sum10 = cnt[8];
  if (mask_sum) mask_sum[pix] = sum10;
  if (geo_mask) {
    bool g = sum10 >= S + 1;  // :262 (dy_range = S + 1: never true, kept for fidelity)
#pragma unroll
    for (int i = 2; i <= 10; ++i)
      if (i <= S) g = g || (cnt[i - 2] >= i);
    geo_mask[pix] = g ? 1 : 0;
  }
  if (depth_avg) depth_avg[pix] = (float)((double)__fadd_rn(dsum, d0) / (double)(sum10 + 1));  // :256, float32 / int32 -> float64 in numpy
}

// Back-projection of a depth map to world points: reference filter/pcd.py:340-343 (the same lines in dypcd_tanks.py:310-313).
// numpy runs it in float64 (int64 pixel grid x depth), the float32 matrices promoted; the vertices are cast to float32 for the PLY.
// Every pixel is projected, [H,W,3]; the caller selects the valid ones.
__global__ void __launch_bounds__(256) backproject_kernel(const float* __restrict__ depth, const float* __restrict__ mats, int H, int W,
                                                          float* __restrict__ xyz) {
  __shared__ float sm[21];  // inv(K) 9 | inv(E)[:3,:4] 12
  if (threadIdx.x < 21) sm[threadIdx.x] = __ldg(mats + threadIdx.x);
  __syncthreads();
  const long long hw = (long long)H * W;
  const long long pix = (long long)blockIdx.x * 256 + threadIdx.x;
  if (pix >= hw) return;
  const int y = (int)(pix / W), x = (int)(pix - (long long)y * W);
  const double d = (double)__ldg(depth + pix);
  double X, Y, Z, wx, wy, wz;
  dmat3(sm, (double)x * d, (double)y * d, d, X, Y, Z);
  dmat34(sm + 9, X, Y, Z, wx, wy, wz);
  xyz[pix * 3 + 0] = (float)wx;
  xyz[pix * 3 + 1] = (float)wy;
  xyz[pix * 3 + 2] = (float)wz;
}

}  // namespace dmvs

extern "C" int dmvs_geo_consistency_f32(const float* depth_ref, const float* depth_src, const float* mats, int S, int H, int W,
                                        float dist_thresh, float rel_thresh, unsigned char* mask, float* depth_reproj, float* xy_src,
                                        int* mask_sum, float* depth_avg, void* stream) {
  using namespace dmvs;
  DMVS_REQUIRE(depth_ref && depth_src && mats, DMVS_ERR_BAD_POINTER, "geo_consistency: null pointer");
  DMVS_REQUIRE(mask || depth_reproj || xy_src || mask_sum || depth_avg, DMVS_ERR_BAD_POINTER, "geo_consistency: no output requested");
  DMVS_REQUIRE(S >= 1 && S <= kGeoMaxSrc && H >= 2 && W >= 2, DMVS_ERR_BAD_SHAPE, "geo_consistency: bad dims S=%d H=%d W=%d (S <= %d)", S, H, W,
               kGeoMaxSrc);
  const long long hw = (long long)H * W;
  geo_consistency_kernel<<<(unsigned)((hw + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      depth_ref, depth_src, mats, S, H, W, dist_thresh, rel_thresh, mask, depth_reproj, xy_src, mask_sum, depth_avg);
  return check_launch("geo_consistency");
}

extern "C" int dmvs_geo_consistency_dynamic_f32(const float* depth_ref, const float* depth_src, const float* mats, int S, int H, int W,
                                                double dist_base, double rel_diff_base, unsigned char* level, float* depth_reproj,
                                                float* xy_src, int* mask_sum, unsigned char* geo_mask, float* depth_avg, void* stream) {
  using namespace dmvs;
  DMVS_REQUIRE(depth_ref && depth_src && mats, DMVS_ERR_BAD_POINTER, "geo_consistency_dynamic: null pointer");
  DMVS_REQUIRE(level || depth_reproj || xy_src || mask_sum || geo_mask || depth_avg, DMVS_ERR_BAD_POINTER,
               "geo_consistency_dynamic: no output requested");
  DMVS_REQUIRE(S >= 1 && S <= kGeoMaxSrc && H >= 2 && W >= 2, DMVS_ERR_BAD_SHAPE, "geo_consistency_dynamic: bad dims S=%d H=%d W=%d (S <= %d)",
               S, H, W, kGeoMaxSrc);
  // the reference indexes its nine masks with i - 2 for i in [2, S]: more than 10 source views raise IndexError there
  DMVS_REQUIRE(!geo_mask || S <= 10, DMVS_ERR_BAD_SHAPE, "geo_consistency_dynamic: the fused mask is defined for at most 10 source views, got %d", S);
  const long long hw = (long long)H * W;
  geo_consistency_dynamic_kernel<<<(unsigned)((hw + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      depth_ref, depth_src, mats, S, H, W, dist_base, rel_diff_base, level, depth_reproj, xy_src, mask_sum, geo_mask, depth_avg);
  return check_launch("geo_consistency_dynamic");
}

extern "C" int dmvs_backproject_world_f32(const float* depth, const float* mats, int H, int W, float* xyz, void* stream) {
  using namespace dmvs;
  DMVS_REQUIRE(depth && mats && xyz, DMVS_ERR_BAD_POINTER, "backproject_world: null pointer");
  DMVS_REQUIRE(H >= 1 && W >= 1, DMVS_ERR_BAD_SHAPE, "backproject_world: bad dims H=%d W=%d", H, W);
  const long long hw = (long long)H * W;
  backproject_kernel<<<(unsigned)((hw + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(depth, mats, H, W, xyz);
  return check_launch("backproject_world");
}
