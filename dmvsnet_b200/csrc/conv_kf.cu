// R1 tensor-core path, stride-1 3x3x3 layers with the DEPTH TAP FOLDED INTO N and a march along z (sm_100a).
//
// conv_tc2.cu issues one tcgen05.mma per (output plane, tap, 8-channel chunk); each of them re-reads its 128-voxel A tile from
// shared memory (~64 cycles) for 8-32 cycles of math, so those layers are bound by the NUMBER of MMAs.  Here the three depth
// taps share one MMA: the weight columns are [kd = 0 | kd = 1 | kd = 2] (N = 3 x 2*Cout_p), and the accumulator belongs to the
// INPUT plane s: P[s][kd] = sum over (kh, kw, c) of in[s] * W[kd].  An output plane is then the sum of three column blocks of
// three neighbouring accumulators, out[t] = P[t-1][0] + P[t][1] + P[t+1][2], taken in the epilogue (TMEM loads are cheap).
// 9 * Cin/8 MMAs per plane instead of 27 * Cin/8.
//
//   * A CTA owns a contiguous range of the linearised (column, output plane) sequence, column = one 16 x 8 patch over all D planes;
//     ranges are equal to within one plane whatever the volume shape.  Inside a column fragment [t0, t1] it marches through the
//     input planes max(t0-1, 0) .. min(t1+1, D-1): every plane is staged once by TMA (one 18 x 10-cell box per hi/lo plane, OOB
//     zero fill = the padding in y / x; planes outside the volume are simply not visited = the padding in z).
//   * The accumulators form a ring of R TMEM slots.  Plane number g (counted over the CTA's whole range) lives in slot g % R;
//     `accfull[slot]` is committed by the issuing thread, `accempty[slot]` collects 3 reader arrivals per epilogue warp (the output planes s-1, s,
//     s+1; where one of them is outside the fragment its neighbour inside arrives in its place).
//   * Two issuing threads alternate planes (R is even, so a slot always belongs to the same thread and its barrier phases are
//     consumed in order); the smem stage ring is split between them the same way.
//
// Kinds: S1 = Cin, Cout multiples of 8 on CH16 cells (conv2); C0 = conv0 on the cost cells W1 emits (K packed along kw, 3 MMAs
// per plane); PB = `prob` (8 -> 2, fp32 logits out); PW = `prob` with kw folded into N as well, on wide tiles:
//
// SW = S1 for Cout_p = 32 (conv4), where 3 x 2*Cout_p columns per plane would leave room for two accumulators only: the lo(W) product
// moves from the N to the K dimension.  Columns are [kd][Cout_p] (N = 96); per tap and 8-channel chunk one MMA multiplies
// [A_hi | A_lo] with [hi(W); hi(W)], and per tap and PAIR of chunks one MMA multiplies [A_hi(chunk 2i) | A_hi(chunk 2i+1)] (the K
// halves are the two chunks' hi planes: LBO = 2 planes) with [lo(W, 2i); lo(W, 2i+1)].  1.5 MMAs per (tap, chunk) instead of 3.
//
//   * The full-resolution layers are bound by the TMA engine's rate for narrow rows (18 x 160-byte rows per box: ~235 cycles,
//     profiles/r2l_kf_probe.txt) and by the MMA count at about the same level.  PW attacks both: the tile is 4 rows x 32 voxels
//     (M row = y * 32 + x, so the 8-row core matrices are consecutive 128-byte pieces of 512-byte box rows: SBO = 128, no x halo in
//     shared memory), the box is 6 rows x 512 bytes, and A is never shifted in x: the accumulator of input voxel x carries all nine
//     (kd, kw) column blocks, out[t][x] = sum_kd sum_kw P[t+kd-1][x+kw-1][kd][kw].  The x shift is a lane shuffle in the epilogue
//     (an epilogue warp owns one 32-voxel row; lanes 0 and 31 are halo: tiles advance by 30 voxels).  3 MMAs of N = 48 per plane.
#include "conv_tc2.cuh"

namespace dmvs {

enum { KF_S1 = 0, KF_C0 = 1, KF_PB = 2, KF_PW = 3, KF_SW = 4 };

template <int KIND, int CIN, int COUT_P, int R, int STAGES, int NPART_, int CS_, int MW_ = 2, int NPROD_ = 1>
struct KF {
  static constexpr bool WIDE = KIND == KF_PW;
  static constexpr bool SPLITW = KIND == KF_SW;
  static constexpr int NB = WIDE ? 16 : SPLITW ? COUT_P : 2 * COUT_P;  // one kd block: [hi(W) | lo(W)] columns (PW: 3 kw x 4, padded to 16; SW: one column per output channel)
  static constexpr int NF = (3 * NB + 15) / 16 * 16;    // N of one tcgen05.mma (PB: 12 -> 16, PW: 36 -> 48)
  static constexpr int TW = WIDE ? 32 : T_W, TH = WIDE ? 4 : T_H;  // tile: TH rows x TW voxels = 128 accumulator lanes
  static constexpr int XSTEP = WIDE ? 30 : T_W;                    // valid output voxels per tile row
  static constexpr int SH = TH + 2, BW = WIDE ? TW : TW + 2;
  static constexpr int BLK_BYTES = SH * BW * 16;
  static constexpr int PLANE = BLK_BYTES;  // the hi / lo planes of all channel chunks arrive as ONE 4-D box: dense plane pitch
  static constexpr int CJ = (KIND == KF_C0) ? 1 : CIN / 8;
  static constexpr int NPLANE = (KIND == KF_C0) ? 1 : 2 * CJ;
  static constexpr int TAPS = (KIND == KF_C0 || WIDE) ? 3 : 9;
  static constexpr int A_LBO = (KIND == KF_C0) ? 32 : PLANE;
  static constexpr int A_SBO = WIDE ? 128 : BW * 16;
  static constexpr int B_TILE = 2 * NF * 16;
  static constexpr int B_BYTES = (SPLITW ? CJ + CJ / 2 : CJ) * TAPS * B_TILE;  // SW: the hi(W) image, then the lo(W) image of the chunk pairs
  static constexpr int STAGE_BYTES = pad128(NPLANE * PLANE);
  static constexpr int TX_BYTES = NPLANE * BLK_BYTES;
  static constexpr int OFF_B = STAGES * STAGE_BYTES;
  static constexpr int OFF_BAR = OFF_B + pad128(B_BYTES);
  static constexpr int SMEM = OFF_BAR + 8 * (2 * STAGES + 2 * R) + 16 + 128;
  static constexpr int TMEM_COLS = pow2c(R * NF);
  // epilogue: NPART groups of 4 warps interleave the output planes, CS groups share one plane by 8-channel chunk
  static constexpr int NPART = NPART_, CS = CS_;
  static constexpr int EPI_WARPS = 4 * NPART * CS;
  static constexpr int MMA_WARPS = MW_;
  static constexpr int HS = STAGES / MMA_WARPS;
  // ONE commit per plane: with as many smem stages as accumulator slots (stage = slot = g % R) "the MMAs of plane g are done" frees
  // the stage for the producer AND hands the accumulator to the epilogue - both wait on accfull[slot].  A tcgen05.commit costs
  // ~200 cycles and commits serialise per SM (tools/experiments/kf_trace.py): two per plane were the floor of the per-plane hand-off
  static constexpr bool ONEBAR = (STAGES == R);
  static constexpr int NPROD = NPROD_;  // TMA-issuing threads (one per warp): plane g is fetched by thread g % NPROD
  static constexpr int THREADS = (NPROD + MMA_WARPS + EPI_WARPS) * 32;
  static_assert(R % MMA_WARPS == 0 && R >= 4, "a slot must always belong to the same issuing thread; 3 live slots + 1");
  static_assert(MMA_WARPS <= 2, "more than two issuing threads make the readers' parity waits unsound (kf_protocol_sim.py)");
  static_assert(R * NF <= 512, "accumulator ring exceeds TMEM");
  static_assert((MMA_WARPS & (MMA_WARPS - 1)) == 0 && (NPROD & (NPROD - 1)) == 0, "role strides are powers of two");
  static_assert(STAGES % MMA_WARPS == 0, "stage ring is split between the issuing threads");
  static_assert(SMEM <= 227 * 1024, "pipeline does not fit shared memory");
  // a group waits only on the accumulators it reads; with more than R/2 plane-interleaved groups one of them can meet a slot a
  // full phase early and pass the parity test on the previous tenant (tools/experiments/kf_protocol_sim.py)
  static_assert((NPART & (NPART - 1)) == 0, "plane interleave is a power of two");
  static_assert(2 * NPART <= R, "too many plane-interleaved epilogue groups for the accumulator ring");
  static_assert(COUT_P % (8 * CS) == 0 || CS == 1, "channel chunks do not split evenly");
};

struct KfRange {  // the CTA's share of the (column, output plane) sequence
  long long o, o_end;
  int D;
  __device__ __forceinline__ KfRange(const Tc2Params& p) {
    const long long total = (long long)p.n_tiles * p.Do;
    o = total * blockIdx.x / gridDim.x;
    o_end = total * (blockIdx.x + 1) / gridDim.x;
    D = p.Do;
  }
  // next column fragment: outputs [t0, t1] of column `col`, input planes [sa, sb]
  __device__ __forceinline__ bool next(int& col, int& t0, int& t1, int& sa, int& sb) {
    if (o >= o_end) return false;
    col = (int)(o / D);
    t0 = (int)(o - (long long)col * D);
    const long long left = o_end - o;
    t1 = (left >= D - t0) ? D - 1 : t0 + (int)left - 1;
    sa = t0 > 0 ? t0 - 1 : 0;
    sb = t1 < D - 1 ? t1 + 1 : D - 1;
    o += t1 - t0 + 1;
    return true;
  }
};

// x0: first voxel of the accumulator rows (PW: one halo voxel left of the first output)
// experiment: time stamps of the first KF_TRACE_PLANES planes of CTA 0, 16 slots per plane
constexpr int KF_TRACE_PLANES = 512;
__device__ __forceinline__ void kf_stamp(const Tc2Params& p, int g, int ev) {
  if ((p.dbg & 8) && p.trace && blockIdx.x == 0 && g < KF_TRACE_PLANES) p.trace[g * 16 + ev] = clock64();
}

template <class Cfg>
__device__ __forceinline__ void kf_column(const Tc2Params& p, int col, int& x0, int& y0, int& b) {
  x0 = (col % p.tiles_x) * Cfg::XSTEP - (Cfg::WIDE ? 1 : 0);
  col /= p.tiles_x;
  y0 = (col % p.tiles_y) * Cfg::TH;
  b = col / p.tiles_y;
}

template <int KIND, int CIN, int COUT_P, int R, int STAGES, int NP, int CS, int MW, int NPR>
__global__ void __launch_bounds__(KF<KIND, CIN, COUT_P, R, STAGES, NP, CS, MW, NPR>::THREADS, 1)
    conv_kf_kernel(const __grid_constant__ Tc2Params p, const __grid_constant__ CUtensorMap tmap) {
  using Cfg = KF<KIND, CIN, COUT_P, R, STAGES, NP, CS, MW, NPR>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  uint8_t* sB = smem + Cfg::OFF_B;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* empty = full + STAGES;
  uint64_t* accfull = empty + STAGES;
  uint64_t* accempty = accfull + R;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accempty + R);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == NPR) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    for (int a = 0; a < R; ++a) {
      mbar_init(accfull + a, 1);
      mbar_init(accempty + a, 3 * 4 * CS);  // 3 readers x (4 warps x CS groups), one elected arrival per warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    uint4* wdst = reinterpret_cast<uint4*>(sB);
    for (int i = tid; i < Cfg::B_BYTES / 16; i += Cfg::THREADS) wdst[i] = __ldg(p.wtc + i);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // p.dbg bit 4: this launch takes part in programmatic dependent launch (see launch_kf)
  if (p.dbg & 16) pdl_launch();
  pdl_wait();

  KfRange range(p);
  int col, t0, t1, sa, sb;
  if (warp < NPR) {
    // ---------------------------------------------------------------- producer: one TMA box per (input plane, hi/lo plane)
    {  // the whole warp walks the loop, one elected lane issues (uniform registers, see the MMA issuers below)
      if (lane == 0) prefetch_tmap(&tmap);
      const int planes_per_b = 2 * CIN / 8;
      // a single thread runs this loop: every instruction and every taken branch of it is serial latency per plane, so each producer
      // steps straight to its own planes
      int g_frag = 0;
      while (range.next(col, t0, t1, sa, sb)) {
        int x0, y0, b;
        kf_column<Cfg>(p, col, x0, y0, b);
        const int first = (warp - g_frag) & (NPR - 1);
#pragma unroll 1
        for (int s = sa + first; s <= sb; s += NPR) {
          const int g = g_frag + (s - sa);
          const int m = g % MW, j = g / MW;
          const int st = Cfg::ONEBAR ? g % R : MW * (j % Cfg::HS) + m, u = Cfg::ONEBAR ? g / R : j / Cfg::HS;
          if (lane == 0) kf_stamp(p, g, 8);
          if (Cfg::ONEBAR) {
            if (g >= R) mbar_wait(accfull + st, (u - 1) & 1);  // plane g - R (the stage's previous tenant) has been multiplied
          } else {
            mbar_wait(empty + st, (u & 1) ^ 1);
          }
          if (lane == 0) kf_stamp(p, g, 0);
          uint8_t* dst = smem + st * Cfg::STAGE_BYTES;
          if (elect_one()) {
          mbar_expect_tx(full + st, Cfg::TX_BYTES);
          if (KIND == KF_C0) {  // cost cells: cell x = [voxel x-1 | voxel x], one plane per batch entry
            tma_load_4d(dst, &tmap, full + st, 8 * x0, y0 - 1, s, b);
          } else {
            // one box over the NPLANE consecutive planes of this batch entry (a TMA instruction costs ~235 cycles whatever its box)
            tma_load_4d(dst, &tmap, full + st, 8 * (Cfg::WIDE ? x0 : x0 - 1), y0 - 1, s, b * planes_per_b);
          }
          kf_stamp(p, g, 7);
          }
          __syncwarp();
        }
        g_frag += sb - sa + 1;
      }
    }
  } else if (warp < NPR + Cfg::MMA_WARPS) {
    // ---------------------------------------------------------------- MMA issuers: thread m takes the planes with g % MW == m
    // The WHOLE warp runs this loop and one elected lane issues: descriptors and addresses computed in warp-uniform code live in
    // uniform registers, whereas under `if (lane == 0)` the compiler keeps them in vector registers and wraps every tcgen05.mma in
    // a R2UR / ELECT waterfall (~11 instructions and a branch per MMA on the serial path of the issuing thread)
    const int me = warp - NPR;
    {
      constexpr uint32_t idesc = make_idesc(Cfg::NF);
      const uint64_t bdesc0 = make_desc(smem_u32(sB), Cfg::NF * 16, 128);
      int g_frag = 0;
      while (range.next(col, t0, t1, sa, sb)) {
        const int first = (me - g_frag) & (MW - 1);
#pragma unroll 1
        for (int s = sa + first; s <= sb; s += MW) {
          const int g = g_frag + (s - sa);
          const int slot = g % R, k = g / R;
          if (lane == 0) kf_stamp(p, g, 9);
          mbar_wait(accempty + slot, (k & 1) ^ 1);
          if (lane == 0) kf_stamp(p, g, 1);
          const int j = g / MW;
          const int st = Cfg::ONEBAR ? slot : MW * (j % Cfg::HS) + me, u = Cfg::ONEBAR ? k : j / Cfg::HS;
          mbar_wait(full + st, u & 1);
          if (lane == 0) kf_stamp(p, g, 2);
          tc_fence_after();
          const uint64_t adesc0 = make_desc(smem_u32(smem + st * Cfg::STAGE_BYTES), Cfg::A_LBO, Cfg::A_SBO);
          const uint32_t acc = tmem_base + slot * Cfg::NF;
          if (elect_one()) {
          if (!(p.dbg & 2))
#pragma unroll
          for (int tap = 0; tap < Cfg::TAPS; ++tap) {
            const int off = (KIND == KF_C0 || Cfg::WIDE) ? tap * Cfg::BW * 16 : ((tap / 3) * Cfg::BW + tap % 3) * 16;
#pragma unroll
            for (int cj = 0; cj < Cfg::CJ; ++cj) {
              const uint64_t ad = adesc0 + (uint64_t)(((KIND == KF_C0 ? 0 : 2 * cj * Cfg::PLANE) + off) >> 4);
              const uint64_t bd = bdesc0 + (uint64_t)(((cj * Cfg::TAPS + tap) * Cfg::B_TILE) >> 4);
              umma_f16(acc, ad, bd, idesc, (tap == 0 && cj == 0) ? 0u : 1u);
            }
            if (Cfg::SPLITW) {  // [A_hi(chunk 2i) | A_hi(chunk 2i+1)] x [lo(W, 2i); lo(W, 2i+1)]
              const uint64_t adesc_p = make_desc(smem_u32(smem + st * Cfg::STAGE_BYTES), 2 * Cfg::PLANE, Cfg::A_SBO);
#pragma unroll
              for (int ci = 0; ci < Cfg::CJ / 2; ++ci) {
                const uint64_t ad = adesc_p + (uint64_t)((4 * ci * Cfg::PLANE + off) >> 4);
                const uint64_t bd = bdesc0 + (uint64_t)((((Cfg::CJ + ci) * Cfg::TAPS + tap) * Cfg::B_TILE) >> 4);
                umma_f16(acc, ad, bd, idesc, 1u);
              }
            }
          }
          if (!Cfg::ONEBAR) umma_commit(empty + st);
          umma_commit(accfull + slot);
          kf_stamp(p, g, 3);
          }
          __syncwarp();
        }
        g_frag += sb - sa + 1;
      }
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------- epilogue: group (part, cs) takes every NPART-th output plane
    // and every CS-th channel chunk of it
    const int ew = warp - NPR - Cfg::MMA_WARPS;
    const int q = warp & 3, part = (ew >> 2) % Cfg::NPART, cs = (ew >> 2) / Cfg::NPART;
    const int hl = Cfg::WIDE ? q : q * 4 + (lane >> 3), wl = Cfg::WIDE ? lane : lane & 7;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int npo = p.Cout / 4;
    // BatchNorm scale / shift of this group's first channel chunk stay in registers (CS groups own one chunk each)
    float sc[8], sh[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int ch = 8 * cs + c;
      const bool on = p.scale && ch < p.Cout;
      sc[c] = on ? __ldg(p.scale + ch) : 1.f;
      sh[c] = on ? __ldg(p.shift + ch) : 0.f;
    }
    int g_base = 0, oc = 0;
    // output addressing: per fragment one base (cell or fp32 element of plane 0 at this thread's pixel), per plane one stride
    const long long zcells = (p.out_fmt == FMT_CH16P) ? (long long)p.Ho * (2 * ((p.Wo + 1) >> 1)) : (long long)p.Ho * p.Wo;
    const long long pstride = (long long)p.Do * zcells;  // hi plane -> lo plane of a CH16 / CH16P tensor
    while (range.next(col, t0, t1, sa, sb)) {
      int x0, y0, b;
      kf_column<Cfg>(p, col, x0, y0, b);
      const int oy = y0 + hl, ox = x0 + wl;
      const bool in_img = (oy < p.Ho) && (ox < p.Wo) && (!Cfg::WIDE || (lane >= 1 && lane <= 30));
      const long long base = (KIND == KF_PB || KIND == KF_PW) ? (long long)b * p.y_bs + (long long)oy * p.Wo + ox
                                             : cell_index(p.out_fmt, b, npo, 0, p.Do, p.Ho, p.Wo, 0, oy, ox);
      for (int t = t0 + ((part - oc) & (Cfg::NPART - 1)); t <= t1; t += Cfg::NPART) {
        const bool has_m = t > 0, has_p = t < p.Do - 1;
        const int g0 = g_base + (t - sa), gm = g0 - 1, gp = g0 + 1;
        if (has_m) mbar_wait(accfull + gm % R, (gm / R) & 1);
        mbar_wait(accfull + g0 % R, (g0 / R) & 1);
        if (has_p) mbar_wait(accfull + gp % R, (gp / R) & 1);
        if (q == 0 && lane == 0 && cs == 0) kf_stamp(p, g0, 4);
        tc_fence_after();
        // column blocks: kd = 0 of plane t-1, kd = 1 of plane t, kd = 2 of plane t+1 (a missing neighbour re-reads plane t and is
        // masked out, so the loads stay unconditional)
        const uint32_t am = lane_addr + (uint32_t)((has_m ? gm : g0) % R) * Cfg::NF;
        const uint32_t a0 = lane_addr + (uint32_t)(g0 % R) * Cfg::NF + Cfg::NB;
        const uint32_t ap = lane_addr + (uint32_t)((has_p ? gp : g0) % R) * Cfg::NF + 2 * Cfg::NB;
        const float fm = has_m ? 1.f : 0.f, fp = has_p ? 1.f : 0.f;
        // 3 reader arrivals per accumulator: this plane arrives for itself and for the readers outside the fragment it stands in
        // for.  Called as soon as the group's TMEM loads have landed in registers - the slot is free long before the stores.
        auto release = [&]() {
          if (q == 0 && lane == 0 && cs == 0) kf_stamp(p, g0, 5);
          tc_fence_before();
          __syncwarp();  // every lane's tcgen05.ld has completed (wait::ld) before lane 0 arrives for the warp
          if (q == 0 && lane == 0 && cs == 0) kf_stamp(p, g0, 6);
          if (lane == 0) {
            if (has_m) mbar_arrive_n(accempty + gm % R, (t == t0) ? 3 : 1);
            mbar_arrive_n(accempty + g0 % R, 1 + (t == t0 ? 1 : 0) + (t == t1 ? 1 : 0));
            if (has_p) mbar_arrive_n(accempty + gp % R, (t == t1) ? 3 : 1);
          }
        };
        if (p.dbg & 1) {
          release();
          continue;
        }
        if (KIND == KF_PW) {
          // NB = 16: columns [kw][hi co0, hi co1, lo co0, lo co1] (12 used) per kd.  s[kw][co] = sum over kd, then the x shift across lanes.
          uint32_t ra[8], rb[4], rc[8], rd[4], re[8], rf[4];
          tmem_ld8_issue(am, ra);
          tmem_ld4_issue(am + 8, rb);
          tmem_ld8_issue(a0, rc);
          tmem_ld4_issue(a0 + 8, rd);
          tmem_ld8_issue(ap, re);
          tmem_ld4_issue(ap + 8, rf);
          tmem_wait_ld();
          release();
          float sk[3][2];
#pragma unroll
          for (int kw = 0; kw < 3; ++kw)
#pragma unroll
            for (int co = 0; co < 2; ++co) {
              const int c = 4 * kw + co;  // hi column; lo column = c + 2
              const float vm = __uint_as_float(c < 8 ? ra[c] : rb[c - 8]) + __uint_as_float(c + 2 < 8 ? ra[c + 2] : rb[c + 2 - 8]);
              const float v0 = __uint_as_float(c < 8 ? rc[c] : rd[c - 8]) + __uint_as_float(c + 2 < 8 ? rc[c + 2] : rd[c + 2 - 8]);
              const float vp = __uint_as_float(c < 8 ? re[c] : rf[c - 8]) + __uint_as_float(c + 2 < 8 ? re[c + 2] : rf[c + 2 - 8]);
              sk[kw][co] = (fm * vm + v0) + fp * vp;
            }
          float* yf = reinterpret_cast<float*>(p.y) + base + (long long)t * zcells;
#pragma unroll
          for (int co = 0; co < 2; ++co) {
            // out[x] = P[x-1][kw = 0] + P[x][kw = 1] + P[x+1][kw = 2]
            const float left = __shfl_up_sync(0xffffffffu, sk[0][co], 1), right = __shfl_down_sync(0xffffffffu, sk[2][co], 1);
            float v = (left + sk[1][co]) + right;
            if (p.scale) v = fmaf(v, __ldg(p.scale + co), __ldg(p.shift + co));
            if (p.relu) v = fmaxf(v, 0.f);
            if (in_img && co < p.Cout) yf[(long long)co * pstride] = v;
          }
        } else if (KIND == KF_PB) {
          // NB = 4: columns [hi co0, hi co1, lo co0, lo co1] per kd; the x8 loads cover kd 0,1 (columns 0..7) and kd 2 (8..15)
          uint32_t r0[8], r1[8], r2[8];
          tmem_ld8_issue(am, r0);
          tmem_ld8_issue(a0 - Cfg::NB, r1);
          tmem_ld8_issue(ap, r2);
          tmem_wait_ld();
          release();
          if (in_img) {
            float* yf = reinterpret_cast<float*>(p.y) + base + (long long)t * zcells;
#pragma unroll
            for (int co = 0; co < 2; ++co) {
              const float vm = __uint_as_float(r0[co]) + __uint_as_float(r0[2 + co]);
              const float v0 = __uint_as_float(r1[4 + co]) + __uint_as_float(r1[6 + co]);
              const float vp = __uint_as_float(r2[co]) + __uint_as_float(r2[2 + co]);
              float v = (fm * vm + v0) + fp * vp;
              if (p.scale) v = fmaf(v, __ldg(p.scale + co), __ldg(p.shift + co));
              if (p.relu) v = fmaxf(v, 0.f);
              if (co < p.Cout) yf[(long long)co * pstride] = v;
            }
          }
        } else {
#pragma unroll 1
          for (int c0 = 8 * cs; c0 < COUT_P; c0 += 8 * Cfg::CS) {
            uint32_t rm[8], rml[8], r0[8], r0l[8], rp[8], rpl[8];
            tmem_ld8_issue(am + c0, rm);
            tmem_ld8_issue(a0 + c0, r0);
            tmem_ld8_issue(ap + c0, rp);
            if (!Cfg::SPLITW) {
              tmem_ld8_issue(am + COUT_P + c0, rml);
              tmem_ld8_issue(a0 + COUT_P + c0, r0l);
              tmem_ld8_issue(ap + COUT_P + c0, rpl);
            }
            tmem_wait_ld();
            if (Cfg::SPLITW) {
#pragma unroll
              for (int c = 0; c < 8; ++c) rml[c] = r0l[c] = rpl[c] = 0u;  // the lo(W) product sits in the same columns
            }
            if (c0 + 8 * Cfg::CS >= COUT_P) release();
            if (!in_img || c0 >= p.Cout) continue;
            float v[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const float vm = __uint_as_float(rm[c]) + __uint_as_float(rml[c]);
              const float v0 = __uint_as_float(r0[c]) + __uint_as_float(r0l[c]);
              const float vp = __uint_as_float(rp[c]) + __uint_as_float(rpl[c]);
              float a = (fm * vm + v0) + fp * vp;
              if (c0 == 8 * cs) a = fmaf(a, sc[c], sh[c]);
              else if (p.scale) a = fmaf(a, __ldg(p.scale + c0 + c), __ldg(p.shift + c0 + c));
              if (p.relu) a = fmaxf(a, 0.f);
              v[c] = a;
            }
            uint4 hi, lo;
            split_pack8(v, hi, lo);
            uint4* yc = reinterpret_cast<uint4*>(p.y) + base + (long long)t * zcells + (long long)(c0 >> 2) * pstride;
            yc[0] = hi;
            yc[pstride] = lo;
          }
        }
      }
      g_base += sb - sa + 1;
      oc += t1 - t0 + 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NPR) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

extern int g_tc2_max_ctas;
extern int g_tc2_pdl;
// dmvs_debug_set("kf", 0 | 1 | 2): 0 = the per-tap kernels of conv_tc2.cu everywhere; 1 = conv2 folded (x1.75 on B200);
// 2 = conv0 and prob folded as well (measured SLOWER: with 3 / 9 MMAs per plane they are bound by the per-plane hand-off
// between issuer and epilogue through a ring of only 4 - 8 accumulators, not by the MMA count)
int g_kf = 1;
// dmvs_debug_set("kf_wide", 0 | 1): `prob` on the wide-tile kernel (kd and kw folded).  Measured equal to the per-tap kernel (265 vs
// 262 us at 32 x 592 x 800): with 3 MMAs per plane the kernel is bound by its per-plane hand-offs, not by MMAs or TMA - off by default
int g_kf_wide = 0;
long long* g_kf_trace = nullptr;
int g_kf_pdl = 0;
int g_kf_dbg = 0;  // dmvs_debug_set("kf_dbg", bits): 1 = epilogue releases without loading / storing, 2 = issuers commit without MMAs
int g_kf_mw = 0;    // dmvs_debug_set("kf_mw", 0 | 1 | 2): MMA-issuing threads of the folded kernels, 0 = the default (2)

template <int KIND, int CIN, int COUT_P, int R, int STAGES, int NP, int CS, int MW = 2, int NPR = 1>
static int launch_kf(Tc2Params p, const void* x, cudaStream_t st) {
  using Cfg = KF<KIND, CIN, COUT_P, R, STAGES, NP, CS, MW, NPR>;
  // Programmatic dependent launch is OFF for the folded kernels by default (dmvs_debug_set("kf_pdl", 1) turns it on): with it, and
  // only with the cascade and FeatureNet running on concurrent streams (MVSNet.infer_many), one run in ~60 T&T view sets ended in an
  // `unspecified launch failure` (0 of 20 runs without it, tools/experiments/repro_tnt.py); the gain was ~0.03 ms per view
  const bool pdl = g_tc2_pdl && g_kf_pdl;
  p.dbg = g_kf_dbg | (pdl ? 16 : 0);
  p.trace = g_kf_trace;
  p.tiles_x = ceil_div(p.Wo, Cfg::XSTEP);
  p.tiles_y = ceil_div(p.Ho, Cfg::TH);
  p.tiles_z = 1;
  p.n_tiles = p.tiles_x * p.tiles_y * p.B;  // columns
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  int rc;
  if (KIND == KF_C0)  // [B][D][H][W+1] cost cells viewed as a CH16 tensor of width W+1 with one plane per batch entry
    rc = make_tmap(&tmap, x, FMT_CH16, p.B, p.Di, p.Hi, p.Wi + 1, Cfg::BW, Cfg::SH, 1);
  else
    rc = make_tmap(&tmap, x, FMT_CH16, p.B * 2 * CIN / 8, p.Di, p.Hi, p.Wi, Cfg::BW, Cfg::SH, 1, Cfg::NPLANE);
  if (rc != DMVS_OK) return rc;
  auto kern = conv_kf_kernel<KIND, CIN, COUT_P, R, STAGES, NP, CS, MW, NPR>;
  static PerDevice state;  // per template instance
  const int slot = current_device_slot();
  DMVS_REQUIRE(slot >= 0, DMVS_ERR_CUDA, "conv_kf: no current CUDA device");
  if (!state.configured[slot]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) {
      set_error("conv_kf: cudaFuncSetAttribute(%d bytes): %s", Cfg::SMEM, cudaGetErrorString(e));
      return DMVS_ERR_CUDA;
    }
    int occ = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, Cfg::THREADS, Cfg::SMEM) != cudaSuccess || occ < 1) occ = 1;
    const int by_tmem = 512 / Cfg::TMEM_COLS;
    state.value[slot] = occ < by_tmem ? occ : by_tmem;
    if (state.value[slot] < 1) state.value[slot] = 1;
    state.configured[slot] = true;
  }
  const int ctas_per_sm = state.value[slot];
  const long long want = (long long)kNumSMs * (ctas_per_sm < g_tc2_max_ctas ? ctas_per_sm : g_tc2_max_ctas);
  // every range pays up to two halo planes: no ranges shorter than 4 output planes
  const long long total = (long long)p.n_tiles * p.Do;
  long long grid = total / 4 < 1 ? 1 : total / 4;
  if (grid > want) grid = want;
  if (pdl) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(Cfg::THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, p, tmap);
  } else {
    kern<<<(unsigned)grid, Cfg::THREADS, Cfg::SMEM, st>>>(p, tmap);
  }
  char what[128];
  snprintf(what, sizeof(what), "conv_kf<kind %d, Cin %d, Cout_p %d, ring %d, stages %d> columns %d x %d planes", KIND, CIN, COUT_P, R, STAGES,
           p.n_tiles, p.Do);
  return check_launch(what);
}

// Stride-1 3x3x3 layers with a folded weight image (dmvs_conv_layer.w_tc_kd).  `p` arrives filled by conv_layer_tc2 (dims, pointers,
// out_fmt, y_bs) with p.wtc already pointing at the folded image.  Returns +1 if the shape has no folded specialisation.
int conv_layer_kf(const Tc2Params& p, const void* x, int in_cells, cudaStream_t st) {
  // Issuing threads: at most TWO.  A reader polls only the accumulators it reads, and its parity wait is sound only if the previous
  // tenant of the slot is known to be complete; with two issuers (planes alternate, each issuer completes its planes in order) every
  // output has just seen a plane of either issuer, with four it has not (tools/experiments/kf_protocol_sim.py with a starved issuer:
  // over-arrivals for (R, groups, issuers) = (8, 4, 4) and (4, 2, 4), none in 4000 adversarial schedules for the kinds below).
  if (!g_kf) return 1;
  if (p.Cin == 16 && p.Cout == 16 && (p.out_fmt == FMT_CH16 || p.out_fmt == FMT_CH16P)) {  // conv2
    if (g_kf_mw == 1) return launch_kf<KF_S1, 16, 16, 4, 4, 2, 2, 1, 1>(p, x, st);
    return launch_kf<KF_S1, 16, 16, 4, 4, 2, 2, 2>(p, x, st);
  }
  if (p.Cin == 32 && p.Cout == 32 && (p.out_fmt == FMT_CH16 || p.out_fmt == FMT_CH16P)) {  // conv4
    if (g_kf_mw == 1) return launch_kf<KF_SW, 32, 32, 4, 2, 2, 2, 1, 1>(p, x, st);
    return launch_kf<KF_SW, 32, 32, 4, 2, 2, 2, 2>(p, x, st);
  }
  if (p.Cin == 2 && in_cells) {
    if (g_kf < 2) return 1;
    if (p.Cout == 16) return launch_kf<KF_C0, 2, 16, 4, 4, 2, 2, 2>(p, x, st);  // conv0 of both branches
    if (p.Cout == 8) return launch_kf<KF_C0, 2, 8, 8, 8, 4, 1>(p, x, st);       // conv0
    return 1;
  }
  if (p.Cin == 8 && p.Cout == 2 && p.out_fmt == FMT_F32 && p.wtc_wide && g_kf_wide) {  // prob, kd and kw folded, wide tiles
    Tc2Params pw = p;
    pw.wtc = p.wtc_wide;
    return launch_kf<KF_PW, 8, 2, 10, 10, 4, 1, 2, 1>(pw, x, st);
  }
  if (p.Cin == 8 && p.Cout == 2 && p.out_fmt == FMT_F32) {  // prob
    if (g_kf_mw == 1) return launch_kf<KF_PB, 8, 2, 8, 8, 4, 1, 1, 1>(p, x, st);
    return launch_kf<KF_PB, 8, 2, 8, 8, 4, 1, 2, 1>(p, x, st);
  }
  return 1;
}

}  // namespace dmvs
