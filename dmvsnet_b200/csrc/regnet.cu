// R1 driver: schedules the 2 x 11 layers of CostRegNet / CostRegNet_refine (reference
// networks/module.py:342-436) on one stream.  Stateless: weights arrive as device pointers in
// dmvs_regnet_branch, activations live in a caller-provided workspace (owned by torch's allocator).
#include "common.cuh"

namespace dmvs {

int conv_layer(const float* x, long long x_bs, const dmvs_conv_layer& L, const float* skip, long long skip_bs, float* y,
               long long y_bs, int B, int Cin, int Cout, int Di, int Hi, int Wi, int kd, int stride, int transposed, int relu,
               cudaStream_t st);
int conv_layer_tc2(const void* x, int in_cells, const dmvs_conv_layer& L, const void* skip, void* y, long long y_bs_f32, int B, int Cin,
                   int Cout, int Di, int Hi, int Wi, int kd, int stride, int transposed, int relu, int out_fmt, cudaStream_t st);
int convert_layout(const void* x, void* y, int B, int C, int D, int H, int W, int fmt, int to_ch16, cudaStream_t st);

// single layers on fp32 NCDHW activations always run on the exact fp32 kernels; the tensor engine works on the cell layouts
// (dmvs_conv3d_ch16 / the fused drivers below) and `engine` only selects it there
static int run_layer(int engine, const float* x, long long x_bs, const dmvs_conv_layer& L, const float* skip, long long skip_bs,
                     float* y, long long y_bs, int B, int Cin, int Cout, int Di, int Hi, int Wi, int kd, int stride, int transposed,
                     int relu, cudaStream_t st) {
  (void)engine;
  return conv_layer(x, x_bs, L, skip, skip_bs, y, y_bs, B, Cin, Cout, Di, Hi, Wi, kd, stride, transposed, relu, st);
}

int g_regnet_streams = 1;  // dmvs_debug_set("regnet_streams", 0 | 1): second branch on a side stream (tensor path, both branches)

namespace {

// One non-blocking side stream per (device, caller stream), created on first use and kept for the life of the process: callers
// that drive several cascades on different streams of one device do not meet in a shared side stream.
cudaStream_t side_stream(cudaStream_t main) {
  struct Entry { cudaStream_t main, side; bool used; };
  static Entry table[64][4] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  Entry* row = table[dev];
  for (int i = 0; i < 4; ++i)
    if (row[i].used && row[i].main == main) return row[i].side;
  for (int i = 0; i < 4; ++i) {
    if (!row[i].used) {
      if (cudaStreamCreateWithFlags(&row[i].side, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
      row[i].main = main;
      row[i].used = true;
      return row[i].side;
    }
  }
  return row[0].side;  // more than four caller streams on one device: share the first (ordering stays correct, overlap may not)
}

// `to` waits for everything enqueued on `from` so far (the event is released as soon as the wait has been enqueued)
bool stream_wait(cudaStream_t to, cudaStream_t from) {
  cudaEvent_t e;
  if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return false;
  const bool ok = cudaEventRecord(e, from) == cudaSuccess && cudaStreamWaitEvent(to, e, 0) == cudaSuccess;
  cudaEventDestroy(e);
  return ok;
}

// joins the side stream back into the caller's stream on EVERY exit path once the fork has happened: the caller frees (and
// torch recycles) the workspace as soon as the call returns, error or not
struct SideJoin {
  cudaStream_t main, side;
  ~SideJoin() {
    if (side) stream_wait(main, side);
  }
};

struct Level {
  int D, H, W;
  long long vox() const { return (long long)D * H * W; }
};

struct Plan {
  Level lv[4];
  // element offsets into the workspace
  long long c0, u11, c1, c2, c3, c4, c5, c6, total;
  long long second;  // element offset of the second branch's private copy of u11 .. c6 (two-stream schedule), 0 = none
};

Plan make_plan(int refine, int B, int D, int h, int w) {
  Plan p;
  p.lv[0] = {D, h, w};
  for (int k = 1; k < 4; ++k) {
    const Level& a = p.lv[k - 1];
    const bool flat = refine && k == 3;  // the refine net's bottleneck is 2-D (module.py:411-414)
    p.lv[k] = {flat ? a.D : (a.D - 1) / 2 + 1, (a.H - 1) / 2 + 1, (a.W - 1) / 2 + 1};
  }
  long long o = 0;
  auto take = [&](int ch, int lvl) { long long at = o; o += (long long)B * ch * p.lv[lvl].vox(); o = (o + 3) & ~3LL; return at; };
  p.c0 = take(16, 0); p.u11 = take(8, 0);  // c0: room for conv0 of both branches (conv0_pair), 8 channels each
  p.c1 = take(16, 1); p.c2 = take(16, 1);
  p.c3 = take(32, 2); p.c4 = take(32, 2);
  p.c5 = take(64, 3); p.c6 = take(64, 3);
  // the two branches are independent once conv0 has run: a second set of u11 .. c6 lets them run on two streams
  p.second = o - p.u11;
  o += p.second;
  p.total = o;
  return p;
}

}  // namespace
}  // namespace dmvs

using namespace dmvs;

extern "C" size_t dmvs_regnet_workspace_bytes(int refine, int B, int D, int h, int w) {
  if (B < 1 || D < 1 || h < 1 || w < 1) return 0;
  return (size_t)make_plan(refine, B, D, h, w).total * sizeof(float);
}

static int regnet_forward_impl(const dmvs_regnet_branch* branches, int refine, const float* cost, const void* cost_cells,
                               float* logits, void* workspace, size_t workspace_bytes, int B, int D, int h, int w, int engine,
                               int branch_mask, void* stream) {
  DMVS_REQUIRE(engine == DMVS_ENGINE_FP32 || engine == DMVS_ENGINE_TENSOR, DMVS_ERR_BAD_SHAPE, "regnet: unknown engine %d", engine);
  DMVS_REQUIRE(branches && (cost || cost_cells) && logits && workspace, DMVS_ERR_BAD_POINTER, "regnet: null pointer");
  DMVS_REQUIRE(B >= 1 && h >= 8 && w >= 8 && h % 8 == 0 && w % 8 == 0, DMVS_ERR_BAD_SHAPE,
               "regnet: h=%d w=%d must be positive multiples of 8", h, w);
  if (refine)
    DMVS_REQUIRE(D == 4, DMVS_ERR_BAD_SHAPE, "regnet(refine): D=%d, the refine net squeezes depth 4->2->1", D);
  else
    DMVS_REQUIRE(D >= 8 && D % 8 == 0, DMVS_ERR_BAD_SHAPE, "regnet: D=%d must be a positive multiple of 8", D);
  DMVS_REQUIRE(aligned16(workspace) && aligned16(logits), DMVS_ERR_BAD_POINTER, "regnet: workspace/logits must be 16-byte aligned");
  const Plan p = make_plan(refine, B, D, h, w);
  DMVS_REQUIRE(workspace_bytes >= (size_t)p.total * sizeof(float), DMVS_ERR_WORKSPACE, "regnet: workspace %zu < %zu bytes",
               workspace_bytes, (size_t)p.total * sizeof(float));
  cudaStream_t st = (cudaStream_t)stream;
  float* ws = (float*)workspace;
  float *c0 = ws + p.c0, *u11 = ws + p.u11, *c1 = ws + p.c1, *c2 = ws + p.c2, *c3 = ws + p.c3, *c4 = ws + p.c4, *c5 = ws + p.c5,
        *c6 = ws + p.c6;
  float *u9 = c1, *u7 = c3;  // c1 / c3 are dead once conv2 / conv4 have run
  const Level *L0 = &p.lv[0], *L1 = &p.lv[1], *L2 = &p.lv[2], *L3 = &p.lv[3];
  const long long V0 = L0->vox(), V1 = L1->vox(), V2 = L2->vox(), V3 = L3->vox();
  const int kd_mid = refine ? 1 : 3;
  bool tensor_ok = (engine == DMVS_ENGINE_TENSOR);
  for (int br = 0; br < 2 && tensor_ok; ++br)
    for (int i = 0; i < DMVS_REGNET_LAYERS; ++i)
      if (!branches[br].layer[i].w_tc) tensor_ok = false;
  DMVS_REQUIRE(cost || tensor_ok, DMVS_ERR_BAD_POINTER, "regnet: the fp32 engine needs the fp32 cost volume");
  if (tensor_ok) {
    // ---- tensor path: activations in CH16 / CH16P cells between layers, every 3x3x3 layer a TMA-fed tcgen05 kernel
    enum { F32 = DMVS_FMT_F32, CH = DMVS_FMT_CH16, CHP = DMVS_FMT_CH16P };
    // conv0 of both branches in one launch: a 2 -> 16 layer whose CH16P output holds branch 0 in planes 0,1 and branch 1
    // in planes 2,3 (contiguous per branch when B == 1)
    const bool pair = cost_cells && B == 1 && branch_mask == 3 && branches[0].conv0_pair.w_tc != nullptr;
    if (pair) {
      const int rc0 = conv_layer_tc2(cost_cells, 1, branches[0].conv0_pair, nullptr, c0, 0, B, 2, 16, L0->D, L0->H, L0->W, 3, 1, 0, 1, CHP, st);
      if (rc0 > 0) { set_error("regnet: conv0_pair has no tensor specialisation"); return DMVS_ERR_BAD_SHAPE; }
      if (rc0 != DMVS_OK) return rc0;
    }
    float* const c0_base = c0;
    // Two-stream schedule: after conv0 the branches share nothing, so branch 1 runs on a side stream with its own copy of
    // u11 .. c6.  The coarse levels' kernels (20-130 CTAs) then fill the SMs the other branch leaves idle; the full-resolution
    // layers, whose persistent CTAs do not fit an SM twice, queue behind each other as before.
    cudaStream_t const main_st = st;
    cudaStream_t side = (g_regnet_streams && pair && branch_mask == 3) ? side_stream(main_st) : nullptr;
    if (side && !stream_wait(side, main_st)) side = nullptr;
    SideJoin join{main_st, side};
    for (int br = 0; br < 2; ++br) {
      if (!((branch_mask >> br) & 1)) continue;
      const dmvs_conv_layer* L = branches[br].layer;
      int rc;
      c0 = pair ? c0_base + (long long)br * 8 * V0 : c0_base;
      if (br == 1 && side) {
        st = side;
        u11 = ws + p.u11 + p.second; c1 = ws + p.c1 + p.second; c2 = ws + p.c2 + p.second; c3 = ws + p.c3 + p.second;
        c4 = ws + p.c4 + p.second; c5 = ws + p.c5 + p.second; c6 = ws + p.c6 + p.second;
        u9 = c1; u7 = c3;
      }
#define TC2(...)                                                                                         \
  rc = conv_layer_tc2(__VA_ARGS__);                                                                      \
  if (rc > 0) { set_error("regnet: layer has no tensor specialisation"); return DMVS_ERR_BAD_SHAPE; }   \
  if (rc != DMVS_OK) return rc;
      //  x,   layer, skip,    y,  y_bs, B, Cin, Cout, Di,    Hi,    Wi,   stride, transposed, relu, out_fmt
      if (!pair) {
        TC2(cost_cells ? cost_cells : (const void*)cost, cost_cells ? 1 : 0, L[0], nullptr, c0, 0, B, 2, 8, L0->D, L0->H, L0->W, 3, 1, 0, 1, CHP, st);
      }
      TC2(c0, 0, L[1], nullptr, c1, 0, B, 8, 16, L0->D, L0->H, L0->W, 3, 2, 0, 1, CH, st);
      TC2(c1, 0, L[2], nullptr, c2, 0, B, 16, 16, L1->D, L1->H, L1->W, 3, 1, 0, 1, CHP, st);
      TC2(c2, 0, L[3], nullptr, c3, 0, B, 16, 32, L1->D, L1->H, L1->W, 3, 2, 0, 1, CH, st);
      TC2(c3, 0, L[4], nullptr, c4, 0, B, 32, 32, L2->D, L2->H, L2->W, 3, 1, 0, 1, CHP, st);
      if (!refine) {
        TC2(c4, 0, L[5], nullptr, c5, 0, B, 32, 64, L2->D, L2->H, L2->W, 3, 2, 0, 1, CH, st);
        TC2(c5, 0, L[6], nullptr, c6, 0, B, 64, 64, L3->D, L3->H, L3->W, 3, 1, 0, 1, CH, st);
        TC2(c6, 0, L[7], c4, u7, 0, B, 64, 32, L3->D, L3->H, L3->W, 3, 2, 1, 1, CH, st);
      } else {  // 2-D bottleneck (module.py:411-414): depth has been squeezed to one plane
        TC2(c4, 0, L[5], nullptr, c5, 0, B, 32, 64, L2->D, L2->H, L2->W, 1, 2, 0, 1, CH, st);
        TC2(c5, 0, L[6], nullptr, c6, 0, B, 64, 64, L3->D, L3->H, L3->W, 1, 1, 0, 1, CH, st);
        TC2(c6, 0, L[7], c4, u7, 0, B, 64, 32, L3->D, L3->H, L3->W, 1, 2, 1, 1, CH, st);
      }
      TC2(u7, 0, L[8], c2, u9, 0, B, 32, 16, L2->D, L2->H, L2->W, 3, 2, 1, 1, CH, st);
      TC2(u9, 0, L[9], c0, u11, 0, B, 16, 8, L1->D, L1->H, L1->W, 3, 2, 1, 1, CH, st);
      TC2(u11, 0, L[10], nullptr, logits + (long long)br * 2 * V0, 4 * V0, B, 8, 2, L0->D, L0->H, L0->W, 3, 1, 0, 0, F32, st);
#undef TC2
    }
    join.side = nullptr;  // normal exit: join here so that a failure is reported
    if (side && !stream_wait(main_st, side)) {
      set_error("regnet: joining the side stream failed: %s", cudaGetErrorString(cudaGetLastError()));
      return DMVS_ERR_CUDA;
    }
    return DMVS_OK;
  }
  for (int br = 0; br < 2; ++br) {
    if (!((branch_mask >> br) & 1)) continue;
    const dmvs_conv_layer* L = branches[br].layer;
    int rc;
#define RUN(...)                 \
  rc = run_layer(engine, __VA_ARGS__);  \
  if (rc != DMVS_OK) return rc;
    //   x,  x_bs,    layer, skip, skip_bs, y,   y_bs,   B, Cin, Cout, Di,    Hi,    Wi,    kd, stride, transposed, relu
    RUN(cost, 2 * V0, L[0], nullptr, 0, c0, 8 * V0, B, 2, 8, L0->D, L0->H, L0->W, 3, 1, 0, 1, st);
    RUN(c0, 8 * V0, L[1], nullptr, 0, c1, 16 * V1, B, 8, 16, L0->D, L0->H, L0->W, 3, 2, 0, 1, st);
    RUN(c1, 16 * V1, L[2], nullptr, 0, c2, 16 * V1, B, 16, 16, L1->D, L1->H, L1->W, 3, 1, 0, 1, st);
    RUN(c2, 16 * V1, L[3], nullptr, 0, c3, 32 * V2, B, 16, 32, L1->D, L1->H, L1->W, 3, 2, 0, 1, st);
    RUN(c3, 32 * V2, L[4], nullptr, 0, c4, 32 * V2, B, 32, 32, L2->D, L2->H, L2->W, 3, 1, 0, 1, st);
    RUN(c4, 32 * V2, L[5], nullptr, 0, c5, 64 * V3, B, 32, 64, L2->D, L2->H, L2->W, kd_mid, 2, 0, 1, st);
    RUN(c5, 64 * V3, L[6], nullptr, 0, c6, 64 * V3, B, 64, 64, L3->D, L3->H, L3->W, kd_mid, 1, 0, 1, st);
    RUN(c6, 64 * V3, L[7], c4, 32 * V2, u7, 32 * V2, B, 64, 32, L3->D, L3->H, L3->W, kd_mid, 2, 1, 1, st);
    RUN(u7, 32 * V2, L[8], c2, 16 * V1, u9, 16 * V1, B, 32, 16, L2->D, L2->H, L2->W, 3, 2, 1, 1, st);
    RUN(u9, 16 * V1, L[9], c0, 8 * V0, u11, 8 * V0, B, 16, 8, L1->D, L1->H, L1->W, 3, 2, 1, 1, st);
    RUN(u11, 8 * V0, L[10], nullptr, 0, logits + (long long)br * 2 * V0, 4 * V0, B, 8, 2, L0->D, L0->H, L0->W, 3, 1, 0, 0, st);
#undef RUN
  }
  return DMVS_OK;
}

extern "C" int dmvs_regnet_forward_f32(const dmvs_regnet_branch* branches, int refine, const float* cost, const void* cost_cells,
                                       float* logits, void* workspace, size_t workspace_bytes, int B, int D, int h, int w, int engine,
                                       void* stream) {
  return regnet_forward_impl(branches, refine, cost, cost_cells, logits, workspace, workspace_bytes, B, D, h, w, engine, 3, stream);
}

extern "C" int dmvs_regnet_forward_branches_f32(const dmvs_regnet_branch* branches, int refine, const float* cost, const void* cost_cells,
                                                float* logits, void* workspace, size_t workspace_bytes, int B, int D, int h, int w,
                                                int engine, int branch_mask, void* stream) {
  DMVS_REQUIRE(branch_mask >= 1 && branch_mask <= 3, DMVS_ERR_BAD_SHAPE, "regnet: branch_mask %d not in 1..3", branch_mask);
  return regnet_forward_impl(branches, refine, cost, cost_cells, logits, workspace, workspace_bytes, B, D, h, w, engine, branch_mask, stream);
}

extern "C" int dmvs_convert_layout(const void* x, void* y, int B, int C, int D, int H, int W, int fmt, int to_ch16, void* stream) {
  DMVS_REQUIRE(B >= 1 && C >= 8 && D >= 1 && H >= 1 && W >= 1, DMVS_ERR_BAD_SHAPE, "convert_layout: bad dims");
  return convert_layout(x, y, B, C, D, H, W, fmt, to_ch16, (cudaStream_t)stream);
}

extern "C" int dmvs_conv3d_ch16(const void* x, int in_cells, const dmvs_conv_layer* layer, const void* skip, void* y, int B, int Cin,
                                int Cout, int Di, int Hi, int Wi, int kd, int stride, int transposed, int relu, int out_fmt, void* stream) {
  DMVS_REQUIRE(layer && layer->w_tc, DMVS_ERR_BAD_POINTER, "conv3d_ch16: the layer needs packed tensor-core weights (w_tc)");
  DMVS_REQUIRE(B >= 1 && Di >= 1 && Hi >= 1 && Wi >= 1, DMVS_ERR_BAD_SHAPE, "conv3d_ch16: bad dims");
  const int rc = conv_layer_tc2(x, in_cells, *layer, skip, y, 0, B, Cin, Cout, Di, Hi, Wi, kd, stride, transposed, relu, out_fmt,
                                (cudaStream_t)stream);
  if (rc > 0) {
    set_error("conv3d_ch16: no tensor specialisation for Cin=%d Cout=%d stride=%d transposed=%d", Cin, Cout, stride, transposed);
    return DMVS_ERR_BAD_SHAPE;
  }
  return rc;
}

extern "C" int dmvs_conv3d_f32(const float* x, const dmvs_conv_layer* layer, const float* skip, float* y, int B, int Cin,
                               int Cout, int Di, int Hi, int Wi, int kd, int stride, int transposed, int relu, int engine,
                               void* stream) {
  DMVS_REQUIRE(layer, DMVS_ERR_BAD_POINTER, "conv3d: null layer");
  DMVS_REQUIRE(engine == DMVS_ENGINE_FP32 || engine == DMVS_ENGINE_TENSOR, DMVS_ERR_BAD_SHAPE, "conv3d: unknown engine %d", engine);
  DMVS_REQUIRE(Cout >= 1, DMVS_ERR_BAD_SHAPE, "conv3d: Cout=%d", Cout);
  int Do, Ho, Wo;
  if (transposed) { Do = (kd == 3) ? 2 * Di : Di; Ho = 2 * Hi; Wo = 2 * Wi; }
  else if (stride == 2) { Do = (kd == 3) ? (Di - 1) / 2 + 1 : Di; Ho = (Hi - 1) / 2 + 1; Wo = (Wi - 1) / 2 + 1; }
  else { Do = Di; Ho = Hi; Wo = Wi; }
  const long long xbs = (long long)Cin * Di * Hi * Wi, ybs = (long long)Cout * Do * Ho * Wo;
  return run_layer(engine, x, xbs, *layer, skip, ybs, y, ybs, B, Cin, Cout, Di, Hi, Wi, kd, stride, transposed, relu,
                   (cudaStream_t)stream);
}
