/*
 * Plain-C restatement of the reference's warp + group-correlation path.  TEST INFRASTRUCTURE ONLY:
 * it is the second, library-free checker for the W1 CUDA kernel and the multi-threaded CPU baseline
 * of that kernel.  Nothing under dmvsnet_b200/ links or loads it.
 *
 * Follows, in the reference's own order of operations:
 *   networks/module.py:227-241   pixel grid, rot @ (x,y,1), * depth, + trans, Z==0 patch, divide, normalise
 *   ATen grid_sampler_2d (CPU, bilinear / zeros / align_corners=True), as called at module.py:247-248:
 *       un-normalise ((g+1) * (size-1)/2), floor, weights (1-f, f), per-corner zero padding
 *   networks/mvsnet.py:139       interleaved 2-group product with the reference feature, mean over C/2
 *   networks/mvsnet.py:141-146   plain sum over the source views
 * Parity pinning: checked against the imported reference via tests/golden/warp_edge.npz and the
 * cascade fixtures (tests/test_oracle_golden.py).
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC (see oracle/build_oracle.py).  -ffast-math must NOT be used.
 */
#include <math.h>
#include <stddef.h>

static inline float corner(const float* img, int h, int w, float fx, float fy) {
  /* zeros padding: an out-of-image corner contributes 0 (even if its weight is non-zero) */
  if (!(fx > -1.0f && fx < (float)w && fy > -1.0f && fy < (float)h)) return 0.0f;
  return img[(size_t)(int)fy * w + (int)fx];
}

/* ref [B,C,h,w]; src: n_src pointers to [B,C,h,w]; rt [B,n_src,12]; hyp [B,D,h,w]; cost [B,2,D,h,w] */
int dmvs_oracle_warp_corr_f32(const float* ref, const float* const* src, int n_src, const float* rt, const float* hyp,
                              float* cost, int B, int C, int D, int h, int w) {
  const size_t hw = (size_t)h * w;
  const float half_w = (float)((double)(w - 1) / 2.0), half_h = (float)((double)(h - 1) / 2.0);
  const float inv_groups = 1.0f / (float)(C / 2);
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b) {
    for (int d = 0; d < D; ++d) {
      for (int y = 0; y < h; ++y) {
        for (int x = 0; x < w; ++x) {
          const float dep = hyp[((size_t)(b * D + d)) * hw + (size_t)y * w + x];
          float total0 = 0.0f, total1 = 0.0f;
          for (int s = 0; s < n_src; ++s) {
            const float* m = rt + (size_t)(b * n_src + s) * 12;
            const float rx = fmaf(m[1], (float)y, m[0] * (float)x) + m[2];
            const float ry = fmaf(m[4], (float)y, m[3] * (float)x) + m[5];
            const float rz = fmaf(m[7], (float)y, m[6] * (float)x) + m[8];
            const float X = rx * dep + m[9];
            const float Y = ry * dep + m[10];
            float Z = rz * dep + m[11];
            if (Z == 0.0f) Z += 1e-5f;
            const float un = (X / Z) / half_w - 1.0f;
            const float vn = (Y / Z) / half_h - 1.0f;
            const float ix = (un + 1.0f) * half_w;
            const float iy = (vn + 1.0f) * half_h;
            const float x0 = floorf(ix), y0 = floorf(iy);
            const float we = ix - x0, ww = 1.0f - we; /* east / west */
            const float ws = iy - y0, wn = 1.0f - ws; /* south / north */
            const float nw = wn * ww, ne = wn * we, sw = ws * ww, se = ws * we;
            const float* img = src[s] + (size_t)b * C * hw;
            const float* rp = ref + (size_t)b * C * hw + (size_t)y * w + x;
            float sum0 = 0.0f, sum1 = 0.0f;
            for (int c = 0; c < C; ++c) {
              const float* ch = img + (size_t)c * hw;
              const float val = corner(ch, h, w, x0, y0) * nw + corner(ch, h, w, x0 + 1.0f, y0) * ne +
                                corner(ch, h, w, x0, y0 + 1.0f) * sw + corner(ch, h, w, x0 + 1.0f, y0 + 1.0f) * se;
              const float prod = val * rp[(size_t)c * hw];
              if (c & 1) sum1 += prod; else sum0 += prod;
            }
            total0 += sum0 * inv_groups;
            total1 += sum1 * inv_groups;
          }
          cost[((size_t)(b * 2 + 0) * D + d) * hw + (size_t)y * w + x] = total0;
          cost[((size_t)(b * 2 + 1) * D + d) * hw + (size_t)y * w + x] = total1;
        }
      }
    }
  }
  return 0;
}

/*
 * Backward of the function above w.r.t. the feature maps (the W1 part): what autograd records for
 * networks/mvsnet.py:137-146 and F.grid_sample (module.py:247-249; ATen grid_sampler_2d_backward, bilinear / zeros /
 * align_corners=True).  The sampling grid is built under torch.no_grad() (module.py:222): no gradient to hyp / rt.
 *   grad_cost [B,2,D,h,w] -> grad_ref [B,C,h,w], grad_src[i] [B,C,h,w] (all overwritten).
 * Serial scatter-add in a fixed order (b, d, y, x, s, c): deterministic, the checker for the atomics of the CUDA kernel.
 */
int dmvs_oracle_warp_corr_backward_f32(const float* ref, const float* const* src, int n_src, const float* rt, const float* hyp,
                                       const float* grad_cost, float* grad_ref, float* const* grad_src, int B, int C, int D, int h,
                                       int w) {
  const size_t hw = (size_t)h * w;
  const float half_w = (float)((double)(w - 1) / 2.0), half_h = (float)((double)(h - 1) / 2.0);
  const float inv_groups = 1.0f / (float)(C / 2);
  for (size_t i = 0; i < (size_t)B * C * hw; ++i) grad_ref[i] = 0.0f;
  for (int s = 0; s < n_src; ++s)
    for (size_t i = 0; i < (size_t)B * C * hw; ++i) grad_src[s][i] = 0.0f;
  for (int b = 0; b < B; ++b) {
    for (int d = 0; d < D; ++d) {
      for (int y = 0; y < h; ++y) {
        for (int x = 0; x < w; ++x) {
          const size_t pix = (size_t)y * w + x;
          const float dep = hyp[((size_t)(b * D + d)) * hw + pix];
          const float g[2] = {grad_cost[((size_t)(b * 2 + 0) * D + d) * hw + pix] * inv_groups,
                              grad_cost[((size_t)(b * 2 + 1) * D + d) * hw + pix] * inv_groups};
          for (int s = 0; s < n_src; ++s) {
            const float* m = rt + (size_t)(b * n_src + s) * 12;
            const float rx = fmaf(m[1], (float)y, m[0] * (float)x) + m[2];
            const float ry = fmaf(m[4], (float)y, m[3] * (float)x) + m[5];
            const float rz = fmaf(m[7], (float)y, m[6] * (float)x) + m[8];
            const float X = rx * dep + m[9];
            const float Y = ry * dep + m[10];
            float Z = rz * dep + m[11];
            if (Z == 0.0f) Z += 1e-5f;
            const float ix = (((X / Z) / half_w - 1.0f) + 1.0f) * half_w;
            const float iy = (((Y / Z) / half_h - 1.0f) + 1.0f) * half_h;
            const float x0 = floorf(ix), y0 = floorf(iy);
            const float we = ix - x0, ww = 1.0f - we, ws = iy - y0, wn = 1.0f - ws;
            const float wgt[4] = {wn * ww, wn * we, ws * ww, ws * we};
            const float cx[4] = {x0, x0 + 1.0f, x0, x0 + 1.0f}, cy[4] = {y0, y0, y0 + 1.0f, y0 + 1.0f};
            for (int c = 0; c < C; ++c) {
              const size_t plane = ((size_t)b * C + c) * hw;
              const float r = ref[plane + pix];
              float warped = 0.0f;
              for (int k = 0; k < 4; ++k) {
                if (!(cx[k] > -1.0f && cx[k] < (float)w && cy[k] > -1.0f && cy[k] < (float)h)) continue; /* zeros padding */
                const size_t q = (size_t)(int)cy[k] * w + (int)cx[k];
                warped += src[s][plane + q] * wgt[k];
                grad_src[s][plane + q] += g[c & 1] * r * wgt[k];
              }
              grad_ref[plane + pix] += g[c & 1] * warped;
            }
          }
        }
      }
    }
  }
  return 0;
}
