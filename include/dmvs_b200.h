/*
 * dmvs_b200.h - C ABI of libdmvs_b200.so: the B200 (sm_100a) implementation of DMVSNet's
 * per-stage cost-volume hot path.
 *
 * The reference (DIVE128/DMVSNet) is pure Python/PyTorch and has no FFI; each entry point below
 * replaces a group of PyTorch library calls on the reference's forward path and cites them
 * (paths relative to the reference checkout).  The reference-side binding is a ctypes stub,
 * shown in INTEGRATION.md; dmvsnet_b200/_native.py is the one this repo ships.
 *
 * Conventions
 *   - plain pointers and sizes only; all tensor pointers are DEVICE pointers to contiguous fp32
 *     in the reference's layouts (NCHW features, [B,D,h,w] hypotheses, [B,C,D,h,w] volumes)
 *     unless a stride argument says otherwise;
 *   - `stream` is a cudaStream_t passed as void*; every call only enqueues work on it:
 *     no allocation, no synchronisation, no host<->device copy inside the library;
 *   - return 0 on success, a negative dmvs_status otherwise; dmvs_last_error() returns a
 *     thread-local message.  Nothing throws or exits across the ABI;
 *   - re-entrant per (device, stream); no global mutable state besides the error string.
 */
#ifndef DMVS_B200_H_
#define DMVS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DMVS_ABI_VERSION 20
#define DMVS_MAX_SRC 16 /* source views per call (reference configs use 2..10) */

typedef enum {
  DMVS_OK = 0,
  DMVS_ERR_BAD_SHAPE = -1,   /* unsupported C / D / odd sizes ... */
  DMVS_ERR_BAD_POINTER = -2, /* null or misaligned pointer */
  DMVS_ERR_WORKSPACE = -3,   /* workspace too small */
  DMVS_ERR_CUDA = -4         /* a CUDA runtime call or launch failed */
} dmvs_status;

int dmvs_abi_version(void);
const char* dmvs_last_error(void);
/* number of kernels this library has launched from the calling process (all threads) */
unsigned long long dmvs_launch_count(void);
/* tuning knobs for experiments: "tc2_max_ctas" (persistent CTAs per SM of the tensor-core convolutions, default 1; 2 measured no
 * gain), "head_px" (pixels per block of the depth head: 32 | 64 | 128, default 32), "pb_td8" (prob / conv0 on 8-plane tiles, default 1) */
int dmvs_debug_set(const char* key, int value);
/* experiments: a device buffer for an instrumented kernel ("kf_trace": 512 x 16 int64 clock stamps of CTA 0 of the folded convolutions) */
int dmvs_debug_set_ptr(const char* key, void* ptr);

/* ---------------------------------------------------------------------------------------------
 * W1  fused homography warp + 2-group correlation, summed over source views.
 * Replaces CostAgg.forward (networks/mvsnet.py:111-153) + homo_warping (networks/module.py:212-251)
 * i.e. torch meshgrid/matmul/div/stack, F.grid_sample(bilinear, zeros, align_corners=True),
 * the view*view product, .mean(1) and the += over views.  The [B,C,D,h,w] warped volume and the
 * sampling grid are never written to memory.
 *
 *   ref      [B,C,h,w], batch stride ref_bstride (elements)
 *   src      host array of n_src device pointers, each [B,C,h,w] with batch stride src_bstride
 *   rt       [B,n_src,12]: row-major 3x3 `rot` then 3 `trans` of P_src @ inv(P_ref) (module.py:223-225)
 *   hyp      [B,D,h,w] per-pixel depth hypotheses
 *   cost     [B,2,D,h,w]; only planes d_begin <= d < d_end are written (depth sharding).  Nullable if cost_cells is given.
 *   cost_cells  nullable: the same cost volume in the layout the tensor path's conv0 consumes (DMVS_FMT_COST2):
 *            [B][D][h][w+1] 16-byte cells, cell x = [hi g0, hi g1, lo g0, lo g1](voxel x-1) ++ the same of voxel x,
 *            fp16 hi/lo split (value = hi + lo); cells 0 and w carry the zero padding
 *   C in {8,16,32}; group g = channels {2j+g}; cost = (2/C) * sum_j ref[2j+g] * warped[2j+g], summed over views
 */
int dmvs_warp_corr_f32(const float* ref, long long ref_bstride, const float* const* src, long long src_bstride,
                       int n_src, const float* rt, const float* hyp, float* cost, void* cost_cells, int B, int C, int D, int h,
                       int w, int d_begin, int d_end, void* stream);

/* W1, channel-last sources.  Same result as dmvs_warp_corr_f32, but every SOURCE feature map is given channel-last,
 * [B,h,w,C] with `src_pixstride` floats between pixels (>= C, multiple of 4; FeatureNet's [B,2C,h,w] output in
 * channels_last memory format has stride 2C and its channel slices are consumed in place) and `src_bstride` between
 * batches; pointers 16-byte aligned.  The reference view is NCHW when `ref_pixstride` == 0, else channel-last too.  A bilinear footprint row is then one contiguous
 * run of 2*C floats that C/2 lanes fetch with one 16-byte load each (`src_cornerstride` = floats between the two x-corners of a
 * footprint; 0 = `src_pixstride`, the only value the Python layer passes), which is what makes the gather cheap when the
 * per-pixel hypotheses are rough (see csrc/warp_corr_nhwc.cu).  All other arguments as above. */
int dmvs_warp_corr_nhwc_f32(const float* ref, long long ref_bstride, int ref_pixstride, const float* const* src,
                            long long src_bstride, int src_pixstride, int src_cornerstride, int n_src, const float* rt, const float* hyp, float* cost, void* cost_cells,
                            int B, int C, int D, int h, int w, int d_begin, int d_end, void* stream);

/* W1, TMA-staged.  Same result and arguments as dmvs_warp_corr_nhwc_f32 plus a scratch byte array `flags` of
 * dmvs_warp_corr_flag_bytes(B, D, h, w) bytes (device memory, contents irrelevant on entry).  Two launches: (1) per 32x8
 * pixel tile x plane chunk, the source footprint is fetched into shared memory by one TMA box load per source (zero fill =
 * zeros padding) and gathered from there with conflict-free 16-byte loads; tiles whose footprint does not fit the box (rough
 * per-pixel hypotheses, depth discontinuities) are flagged instead; (2) the channel-last gather kernel computes exactly the
 * flagged (tile, plane) pairs.  Which pairs take which pass depends only on the inputs, not on [d_begin, d_end). */
int dmvs_warp_corr_staged_f32(const float* ref, long long ref_bstride, int ref_pixstride, const float* const* src,
                              long long src_bstride, int src_pixstride, int src_cornerstride, int n_src, const float* rt, const float* hyp, float* cost,
                              void* cost_cells, void* flags, int B, int C, int D, int h, int w, int d_begin, int d_end, void* stream);
size_t dmvs_warp_corr_flag_bytes(int B, int D, int h, int w);

/* W1, fp16-staged ("h16").  Same cost volume as dmvs_warp_corr_nhwc_f32 up to the rounding of the SOURCE feature maps to
 * fp16 (relative 2^-11 per element; the reference view, weights, products and sums stay fp32) and sample positions computed
 * with a refined reciprocal instead of IEEE divisions (<= 2e-4 px): cost within 1e-3 of max|cost|, regressed depth within the
 * 1e-3 contract (tests/test_gpu_h16.py).  Replaces the same reference lines (networks/mvsnet.py:111-153,
 * networks/module.py:212-251).
 *   src[i]   fp16 channel-last source maps [B,h,w,C], `src_pixstride` halfs between pixels (>= C, multiple of 8), 16-byte
 *            aligned (dmvs_features_nhwc_f16 makes them)
 * Per 16x16 pixel tile x plane chunk the footprint box of every source is derived from the tile's 8 corner projections,
 * fetched by TMA (zero fill = zeros padding) into a ring of shared-memory slots and gathered with conflict-free 16-byte loads;
 * a source whose box does not fit (rough hypotheses, wide baseline) is gathered from global memory by the same threads with
 * the same arithmetic - one launch, no scratch, result independent of the path taken and of [d_begin, d_end).
 * Row bands (single-view sharding): `ref`, `hyp`, `cost` / `cost_cells` may hold only rows [ref_row0, ref_row0 + h) of
 * the view while the source maps stay whole (`src_rows` rows; 0 = h): the absolute row enters the homography
 * (module.py:227-236), so a band's result carries the same bits as the same rows of the unsharded call. */
int dmvs_warp_corr_h16_f32(const float* ref, long long ref_bstride, int ref_pixstride, const void* const* src, long long src_bstride,
                           int src_pixstride, int n_src, const float* rt, const float* hyp, float* cost, void* cost_cells, int B, int C,
                           int D, int h, int w, int d_begin, int d_end, int ref_row0, int src_rows, void* stream);

/* fp32 feature map -> fp16 dense channel-last [B,h,w,C] (the source-map format of dmvs_warp_corr_h16_f32).  x is NCHW (dense
 * (c,h,w), `x_bstride` floats between batches) when `x_pixstride` == 0, else channel-last with that pixel stride. */
int dmvs_features_nhwc_f16(const float* x, long long x_bstride, int x_pixstride, void* y, int B, int C, int h, int w, void* stream);

/* W1 backward (the W1 part): gradients of dmvs_warp_corr_*'s cost volume w.r.t. the feature maps -
 * what autograd records for reference networks/mvsnet.py:137-146 (product, group mean, sum over views) and
 * networks/module.py:247-249 (F.grid_sample).  The sampling grid is built under torch.no_grad() in the reference
 * (module.py:222): no gradient w.r.t. hyp / rt.
 *   ref, src[i]   the forward's feature maps, channel-last ([B,h,w,C], `*_pixstride` floats between pixels, >= C, multiple of 4)
 *   grad_cost     [B,2,D,h,w]
 *   grad_ref, grad_src[i]   dense channel-last [B,h,w,C], OVERWRITTEN (zeroed, then accumulated with 16-byte vector
 *                 reductions; summation order across samples is not fixed -> equal to fp32 rounding, not bit for bit)
 * Samples whose position is not finite contribute nothing.  C in {8,16,32}. */
int dmvs_warp_corr_backward_f32(const float* ref, long long ref_bstride, int ref_pixstride, const float* const* src,
                                long long src_bstride, int src_pixstride, int n_src, const float* rt, const float* hyp,
                                const float* grad_cost, float* grad_ref, float* const* grad_src, int B, int C, int D, int h, int w,
                                void* stream);

/* NCHW -> channel-last repack of one feature map for the call above: x [B,C,h,w] (batch stride x_bstride) -> y [B,h,w,C]
 * dense.  Replaces nothing in the reference (it is `tensor.permute(0,2,3,1).contiguous()`); callers whose FeatureNet
 * already runs in channels_last skip it.  C in {8,16,32}. */
int dmvs_features_nhwc_f32(const float* x, long long x_bstride, float* y, int B, int C, int h, int w, void* stream);

/* ---------------------------------------------------------------------------------------------
 * N1  FeatureNet layers (the caller-side neighbour of the path).
 * Replaces the Conv2d wrapper (networks/module.py:28-69: nn.Conv2d + eval BatchNorm2d + ReLU) and the bare nn.Conv2d
 * heads / laterals of FeatureNet (module.py:274-340), fp32 direct convolution, padding = K/2.
 *   x       [B,Cin,Hi,Wi] NCHW          w  weights repacked [Cin][K][K][Cout]
 *   scale, shift  [Cout] or NULL (eval BatchNorm folded, or scale = NULL and shift = the conv bias)
 *   up_add  nullable [B,Cout,Ho/2,Wo/2]: nearest-x2-upsampled and added after the affine (FPN top-down path, module.py:329,334)
 *   y_nchw  nullable [B,Cout,Ho,Wo];  y_nhwc0 / y_nhwc1 nullable (both or none): channel-last [B,Ho,Wo,Cout/2] buffers that
 *           receive channels [0,Cout/2) and [Cout/2,Cout) - the `stageK` / `stageK_c` halves (module.py:326-336) in the
 *           layout dmvs_warp_corr_nhwc_f32 gathers from
 *   y_cells nullable: the output as DMVS_FMT_CH16 cells [B][2*Cout/8][1][Ho][Wo] (cells_s2d != 0: of the 2x2 pixel-unshuffled
 *           map, see dmvs_features_s2d_cells_f32) for the tensor-core
 *           3x3 layers (dmvs_conv3d_ch16 with kd = 1); at least one of the three output forms must be given
 *   (K, stride, Cin, Cout) must be one of FeatureNet's: (3,1,3,8) (3,1,8,8) (5,2,8,16) (3,1,16,16) (5,2,16,32)
 *   (3,1,32,32) (3,1,32,16) (1,1,32,64) (1,1,16,32) (1,1,8,32); Ho = (Hi + 2*(K/2) - K)/stride + 1. */
int dmvs_conv2d_f32(const float* x, const float* w, const float* scale, const float* shift, const float* up_add,
                    float* y_nchw, float* y_nhwc0, float* y_nhwc1, void* y_cells, int cells_s2d, int B, int Cin, int Cout, int Hi, int Wi,
                    int K, int stride, int relu, void* stream);

/* fp32 [B,C,H,W] -> DMVS_FMT_CH16 cells of the 2x2 pixel-unshuffled map [B][2*(4C)/8][1][H/2][W/2] (channel (dy*2+dx)*C + c at
 * block (y/2, x/2)): FeatureNet's 5x5 stride-2 layers (module.py:304,308) run as 3x3 stride-1 layers on 4C channels on the tensor
 * engine (kernel zero-padded to 6x6 and regrouped by the caller).  dmvs_conv2d_f32(cells_s2d = 1) emits the same layout directly. */
int dmvs_features_s2d_cells_f32(const float* x, void* y_cells, int B, int C, int H, int W, void* stream);

/* ---------------------------------------------------------------------------------------------
 * R1  3-D regularisation U-Nets.
 * Replaces CostRegNet / CostRegNet_refine (networks/module.py:342-436): nn.Conv3d / ConvTranspose3d /
 * Conv2d / ConvTranspose2d + eval-mode BatchNorm + ReLU + skip adds, both branches, concatenated.
 *
 * One layer = conv (no bias) -> y*scale[co] + shift[co] (eval BatchNorm) -> ReLU -> (+ skip).
 *   w      weights repacked as [tap][Cin][Cout] (tap = (kd*3+kh)*3+kw for 3x3x3, kh*3+kw for 1x3x3);
 *          for transposed convs the tap indexes the ConvTranspose kernel as stored by PyTorch
 *   scale, shift  [Cout] or NULL (then 1 / 0); relu != 0 applies max(.,0) before the skip add
 */
typedef struct {
  const float* w;
  const float* scale;
  const float* shift;
  /* optional: the same weights packed for the tcgen05 path (engine = DMVS_ENGINE_TENSOR), fp16 hi/lo split:
   * [chunk j = Cin/8][tap][kc = 2][n = 2*Cout_p][8 halfs], Cout_p = max(8, Cout rounded up to 8);
   * n < Cout_p: hi(W[tap][8j+k][n]) for kc = 0 and 1;  n >= Cout_p: lo(W[tap][8j+k][n-Cout_p]) for kc = 0, zero for kc = 1.
   * NULL: the layer always runs on the fp32 path. */
  const void* w_tc;
  /* optional folded images.  Cin = 16 / Cout = 16 (`conv2`) and Cin = 2 (`conv0`, `conv0_pair`): the w_tc image with the depth tap
   * folded into N, [chunk j][tap (kh,kw) | (kh)][kc = 2][n = 3 x 2*Cout_p, kd-major][8 halfs] (csrc/conv_kf.cu).
   * Cin = 8 / Cout = 2 (`prob`): the depth tap folded into the UMMA N dimension,
   * [1][9 taps (kh,kw)][kc = 2][n = 16][8 halfs] with n = 4*kd + co (hi, both kc) and n = 4*kd + 2 + co (lo, kc = 0).
   * Also, Cin = 16 / Cout = 8 transposed (`conv11`): the 27 taps folded by input shift,
   * [chunk j = 2][shift (sz,sy,sx) = 8][kc = 2][n = 8 parity classes x 16][8 halfs]: class block c = (pz,py,px) holds the w_tc
   * columns of tap k with k = 1 (p = 0, s = 0), 2 (p = 1, s = 0), 0 (p = 1, s = 1) per axis, zeros where p = 0 and s = 1. */
  const void* w_tc_kd;
  /* optional, Cin = 8 / Cout = 2 (`prob`) only: depth tap AND kw folded into N for the wide-tile kernel (csrc/conv_kf.cu),
   * [1][3 taps (kh)][kc = 2][n = 48][8 halfs] with n = 16*kd + 4*kw + co (hi, both kc) and n = 16*kd + 4*kw + 2 + co (lo, kc = 0);
   * columns 16*kd + 12 .. 16*kd + 15 stay zero.  NULL: `prob` runs on the w_tc_kd / w_tc kernels. */
  const void* w_tc_kw;
} dmvs_conv_layer;

/* which arithmetic a convolution call uses */
#define DMVS_ENGINE_FP32 0   /* CUDA-core fp32 FMA, bit-level faithful accumulation order aside */
#define DMVS_ENGINE_TENSOR 1 /* tcgen05 tensor cores on split fp16 operands (hi*hi + hi*lo + lo*hi, fp32 accumulate); layers
                                without a tensor specialisation silently use the fp32 kernels */

/* layers in order: conv0 conv1 conv2 conv3 conv4 conv5 conv6 conv7 conv9 conv11 prob */
#define DMVS_REGNET_LAYERS 11
typedef struct {
  dmvs_conv_layer layer[DMVS_REGNET_LAYERS];
  /* optional, read from branches[0] only (w_tc NULL = absent): conv0 of BOTH branches as one 2 -> 16 layer (channels 0..7 =
   * cosR_small, 8..15 = cosR_huge; w_tc in the K-packed conv0 layout with Cout = 16, scale / shift [16]).  Both branches
   * read the same cost volume and the tensor-core kernel's time is set by how often the 4 KB input tile is re-read, not
   * by N, so one launch with N = 32 replaces two with N = 16.  Used on the tensor engine with cost_cells and B == 1. */
  dmvs_conv_layer conv0_pair;
} dmvs_regnet_branch;

/* bytes of scratch dmvs_regnet_forward_f32 needs for these dimensions: conv0 of both branches (16 channels at full
 * resolution) plus, per branch, conv11's input (8 channels) and two buffers per coarser level (16 / 32 / 64 channels) */
size_t dmvs_regnet_workspace_bytes(int refine, int B, int D, int h, int w);

/*   branches   host array of 2 (cosR_small, cosR_huge) descriptors holding device pointers
 *   refine     0: CostRegNet_part (D % 8 == 0), 1: CostRegNet_part_refine (D == 4, 2-D bottleneck)
 *   cost       [B,2,D,h,w]   logits [B,4,D,h,w] (channels 0,1 = small branch, 2,3 = huge branch)
 *   h % 8 == 0 and w % 8 == 0 */
/*   cost_cells  nullable; with engine = DMVS_ENGINE_TENSOR the first layer then reads it by TMA and `cost` may be NULL
 *   Scheduling: everything is ordered after the work already enqueued on `stream`, and everything enqueued on `stream`
 *   after the call is ordered after it.  Internally (tensor engine, cost_cells, B == 1) the second branch runs on a
 *   per-device side stream, forked and joined by events: no host synchronisation. */
int dmvs_regnet_forward_f32(const dmvs_regnet_branch* branches, int refine, const float* cost, const void* cost_cells,
                            float* logits, void* workspace, size_t workspace_bytes, int B, int D, int h, int w, int engine,
                            void* stream);

/* The same with a branch selection: bit 0 = cosR_small (logit channels 0,1), bit 1 = cosR_huge (channels 2,3); only the selected
 * branches run and only their logit channels are written.  The two branches are independent (module.py:343-349), so two GPUs
 * can take one each and exchange their channel halves with one all-gather (dmvsnet_b200/parallel.py). */
int dmvs_regnet_forward_branches_f32(const dmvs_regnet_branch* branches, int refine, const float* cost, const void* cost_cells,
                                     float* logits, void* workspace, size_t workspace_bytes, int B, int D, int h, int w, int engine,
                                     int branch_mask, void* stream);

/* single layers (exposed for unit tests and for callers that want their own schedule).
 *   x [B,Cin,Di,Hi,Wi] -> y [B,Cout,Do,Ho,Wo];  kd in {1,3} (1 = the 2-D convs of the refine net)
 *   stride in {1,2}; transposed != 0: ConvTranspose(k=3, s=2, p=1, output_padding=1), Do = 2*Di ...
 *   skip (nullable) has the shape of y and is added after the ReLU (module.py:394-396) */
int dmvs_conv3d_f32(const float* x, const dmvs_conv_layer* layer, const float* skip, float* y, int B, int Cin, int Cout,
                    int Di, int Hi, int Wi, int kd, int stride, int transposed, int relu, int engine, void* stream);

/* Activation layouts of the tensor path.  CH16: 16-byte cells of 8 fp16, [B][plane][D][H][W], plane 2j = hi and plane
 * 2j+1 = lo of channels 8j..8j+7 (x = hi + lo, same bytes as fp32) - the layout the UMMA descriptors and TMA boxes consume.
 * CH16P: the same with every row stored column-parity split, [..][H][parity][ceil(W/2)] (inputs of stride-2 convs). */
#define DMVS_FMT_F32 0
#define DMVS_FMT_CH16 1
#define DMVS_FMT_CH16P 2
#define DMVS_FMT_COST2 3 /* conv0 input written by dmvs_warp_corr_f32(cost_cells), see there */
#define DMVS_FMT_NHWC2 4 /* output only, FeatureNet's 3x3 heads (kd = 1, Cin = 32, Cout = 16 / 32): two channel-last fp32 buffers
                            back to back, [2][B][D][H][W][Cout/2] = the `stageK` / `stageK_c` feature sets */
#define DMVS_FMT_NHWC2_F16 5 /* DMVS_FMT_NHWC2 followed, in the same buffer, by the same two sets rounded to fp16 ([2][B][D][H][W][Cout/2]
                                halfs): the source-map format of dmvs_warp_corr_h16_f32, written by the head's epilogue instead of a
                                separate dmvs_features_nhwc_f16 pass.  Buffer size: 2*B*D*H*W*(Cout/2) * (4 + 2) bytes */

/* fp32 NCDHW <-> CH16 / CH16P (C % 8 == 0; CH16P: W even).  to_ch16 != 0: x fp32 -> y cells; else x cells -> y fp32. */
int dmvs_convert_layout(const void* x, void* y, int B, int C, int D, int H, int W, int fmt, int to_ch16, void* stream);

/* One conv block of the tensor path on CH16 activations (TMA-fed persistent tcgen05 kernel; kd = 3: 3x3x3,
 * kd = 1: the 1x3x3 layers of the refine net's bottleneck and FeatureNet's 32-channel 3x3 heads out2 / out3
 * (networks/module.py:326-336), depth treated as a batch of planes).
 *   x     CH16 (stride 1, transposed), CH16P (stride 2); when Cin == 2 (conv0): fp32 [B,2,D,H,W] (in_cells == 0) or
 *         DMVS_FMT_COST2 cells (in_cells != 0)
 *   skip  CH16P with the output's shape, transposed convs only (nullable)
 *   y     out_fmt: DMVS_FMT_CH16 / DMVS_FMT_CH16P, or DMVS_FMT_F32 (needed when Cout < 8); transposed convs write CH16
 * Returns DMVS_ERR_BAD_SHAPE for (Cin, Cout, stride) combinations outside the U-Net's. layer->w_tc must be set. */
int dmvs_conv3d_ch16(const void* x, int in_cells, const dmvs_conv_layer* layer, const void* skip, void* y, int B, int Cin, int Cout,
                     int Di, int Hi, int Wi, int kd, int stride, int transposed, int relu, int out_fmt, void* stream);

/* ---------------------------------------------------------------------------------------------
 * E1  dual-depth head.  Replaces DepthNet.forward (networks/mvsnet.py:15-66) + depth_regression
 * (networks/module.py:454-460): softmax over D, expectation, min/max pairs, 6-value extrapolation
 * stacks, (row%4, col%2) selection, confidence.
 *   logits [B,4,D,h,w], hyp [B,D,h,w], interval: device scalar
 *   prob (nullable) [B,4,D,h,w]; d4 [B,4,h,w]; hyp_c [B,4,h,w]; conf [B,h,w]
 */
int dmvs_depth_head_f32(const float* logits, const float* hyp, const float* interval, float* prob, float* d4,
                        float* hyp_c, float* conf, int B, int D, int h, int w, void* stream);

/* E2  refine head.  Replaces DepthNet.refine (networks/mvsnet.py:67-100).
 *   logits_c [B,4,4,h,w], hyp_c [B,4,h,w] -> depth [B,h,w], conf [B,h,w], d4 [B,4,h,w] */
int dmvs_refine_head_f32(const float* logits_c, const float* hyp_c, const float* interval, float alpha, float* depth,
                         float* conf, float* d4, int B, int h, int w, void* stream);

/* ---------------------------------------------------------------------------------------------
 * S1  hypothesis sampler.  Replaces get_depth_range_samples (networks/module.py:556-649) and, for
 * stages > 0, the F.interpolate(bilinear, align_corners=False) that follows it (mvsnet.py:232-233).
 *
 * stage 0: depth_values [B,Nd] -> hyp [B,D,h,w], interval_out (device scalar).
 * stage>0: last_depth [B,h0,w0], interval_pixel (device scalar = ratio * depth_interval) ->
 *          hyp [B,D,h,w] (sampled at (h0,w0) with the checkerboard n/p ranges, then upsampled), interval_out.
 */
int dmvs_hypotheses_first_f32(const float* depth_values, int Nd, float* hyp, float* interval_out, int B, int D, int h,
                              int w, int inverse, void* stream);
int dmvs_hypotheses_next_f32(const float* last_depth, const float* interval_pixel, float* hyp, float* interval_out,
                             int B, int D, int h0, int w0, int h, int w, int inverse, void* stream);

/* ---------------------------------------------------------------------------------------------
 * N4  geometric-consistency check of the depth-map fusion (the consumer of the path's output).
 * Replaces reproject_with_depth_pytorch + check_geometric_consistency[_pytorch] (filter/pcd.py:152-242) and the accumulation
 * of filter_depth (pcd.py:283-304), fused over the S source views of one reference view.
 *   depth_ref [H,W], depth_src [S,H,W] (all views at the same resolution, as the reference requires)
 *   mats [S,60]: per source, row-major  inv(K_ref) 3x3 | (E_src @ inv(E_ref))[:3,:4] | K_src 3x3 | inv(K_src) 3x3 |
 *                (E_ref @ inv(E_src))[:3,:4] | K_ref 3x3  - computed by the caller with the reference's own calls
 *   dist_thresh, rel_thresh: 1*alpha and 0.01*alpha (pcd.py:219)
 *   outputs, each nullable: mask [S,H,W] (0/1), depth_reproj [S,H,W] (0 where the mask is 0), xy_src [S,2,H,W] (the NORMALISED
 *   source coordinates the reference returns), mask_sum [H,W] int32, depth_avg [H,W] = (sum_s depth_reproj + depth_ref) /
 *   (mask_sum + 1) with reference zeros replaced by 1e-4 like pcd.py:212.  S <= 32. */
int dmvs_geo_consistency_f32(const float* depth_ref, const float* depth_src, const float* mats, int S, int H, int W, float dist_thresh,
                             float rel_thresh, unsigned char* mask, float* depth_reproj, float* xy_src, int* mask_sum, float* depth_avg,
                             void* stream);

/* N4, dynamic-threshold variant.  Replaces reproject_with_depth + check_geometric_consistency of filter/dypcd_tanks.py:61-98,
 * 164-184 and the accumulation of its filter_depth (:237-270).  Same inputs; `mats` are the float32 matrices numpy computes
 * (np.linalg.inv / np.matmul on float32).  The projective chain runs in float64 and the source depth is sampled like
 * cv2.remap(INTER_LINEAR, BORDER_CONSTANT 0) - 1/32-pixel fixed-point coordinates - because that is what the reference's numpy
 * code does.  Nine threshold levels i = 2..10: dist < i * dist_base (float64) and rel < i * rel_diff_base (float32).
 *   outputs, each nullable: level [S,H,W] uint8 = the smallest level the pair passes, 0 = none (masks[i-2] = 0 < level <= i);
 *   depth_reproj [S,H,W] (0 where level 10 fails); xy_src [S,2,H,W] = the float32 PIXEL coordinates in the source view;
 *   mask_sum [H,W] int32 = number of sources passing level 10; geo_mask [H,W] 0/1 = OR_{i=2..S} (#sources passing level i) >= i
 *   (S <= 10 when requested: the reference raises IndexError beyond); depth_avg [H,W] = float32((sum_s depth_reproj + depth_ref)
 *   / float64(mask_sum + 1)); zeros of depth_ref are not patched in this variant. */
int dmvs_geo_consistency_dynamic_f32(const float* depth_ref, const float* depth_src, const float* mats, int S, int H, int W,
                                     double dist_base, double rel_diff_base, unsigned char* level, float* depth_reproj, float* xy_src,
                                     int* mask_sum, unsigned char* geo_mask, float* depth_avg, void* stream);

/* N4, point cloud: depth map -> world points for every pixel.  Replaces filter/pcd.py:340-343: xyz = inv(E)[:3,:4] @ [inv(K) @
 * ((x, y, 1) * depth); 1], evaluated in float64 like numpy does there, stored as float32 (the PLY's vertex type).
 *   depth [H,W]; mats [21] = inv(K) 3x3 | inv(E)[:3,:4] (float32, computed by the caller with np.linalg.inv); xyz [H,W,3]. */
int dmvs_backproject_world_f32(const float* depth, const float* mats, int H, int W, float* xyz, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DMVS_B200_H_ */
