// R1 tensor-core path, second generation: TMA-fed, persistent, warp-specialised implicit-GEMM convolutions (sm_100a).
//
// Arithmetic: fp16 hi/lo split operands (x = hi + lo; hi*hi + hi*lo + lo*hi, only lo*lo is dropped), UMMA M = 128 voxels x N = 2*Cout, shifted-descriptor im2col,
// fp32 accumulators in TMEM) - what changes is how operands get to shared memory:
//
//   * Activations between tensor layers live in HBM in the layout the tensor core consumes ("CH16"): 16-byte cells of 8 fp16,
//     [B][plane][D][H][W] with plane 2j = hi(channels 8j..8j+7), plane 2j+1 = lo.  Same bytes as fp32.  Tensors that feed a
//     stride-2 conv are written column-parity split ("CH16P": [..][H][parity][W/2]) so that the even and odd input columns
//     of a stride-2 tap are contiguous cells again.
//   * One thread per CTA issues TMA box loads (cp.async.bulk.tensor, OOB zero fill = the conv padding) straight into the
//     UMMA canonical no-swizzle layout; weights of multi-pass layers ride along as 1-D bulk copies.  A ring of STAGES
//     smem stages is handed between producer and MMA issuer by full/empty mbarriers (expect_tx / tcgen05.commit).
//   * Two TMEM accumulator sets ping-pong between the MMA issuer and 8 epilogue warps (accfull / accempty mbarriers);
//     the epilogue applies eval-BatchNorm + ReLU (+ skip), splits to hi/lo fp16 and stores CH16 / CH16P cells (or fp32
//     NCDHW for the last layer).
//   * conv0 (Cin = 2): K packed along kw.  Mode C0T reads the cost volume in the cell layout the W1 kernel emits for it
//     (DMVS_FMT_COST2) by TMA like every other layer; mode C0 takes the plain fp32 cost volume through a thread-filled
//     producer (12 warps) in the same pipeline.
//   * prob (8 -> 2, mode PB): the depth tap kd is folded into N.  Cout = 2 uses only 4 of the 16 UMMA columns, so the
//     columns carry [kd][hi0 hi1 lo0 lo1]: one accumulator per INPUT plane holds the three kd partial sums, the
//     epilogue adds P[t+kd][kd] - (TD+2)*9 MMAs per tile instead of TD*27 at the same cost each.
#include "conv_tc2.cuh"

namespace dmvs {

template <int MODE, int CIN, int CIN_P, int NB, int TD, int STAGES, int KD = 3>
struct C2 {
  // KD = 1: the 2-D convolutions of the refine net's bottleneck (no taps, halo or stride along depth)
  static constexpr int SD = (KD == 1) ? TD : (MODE == M2_S2) ? 2 * TD + 1 : is_tr(MODE) ? TD + 1 : TD + 2;
  static constexpr int SH = (MODE == M2_S2) ? 2 * T_H + 1 : is_tr(MODE) ? T_H + 1 : T_H + 2;
  static constexpr int BW = (MODE == M2_S1 || is_c0(MODE) || MODE == M2_PB) ? T_W + 2 : T_W + 1;  // cells per staged row (S2: per parity block)
  static constexpr int ROWS = SD * SH;
  static constexpr int BLK_BYTES = ROWS * BW * 16;  // bytes one TMA box writes
  static constexpr int BLK_PITCH = pad128(BLK_BYTES);
  static constexpr int NBLK = (MODE == M2_S2) ? 2 : 1;
  static constexpr int PLANE = NBLK * BLK_PITCH;
  static constexpr int CJ = is_c0(MODE) ? 1 : CIN_P / 8;
  static constexpr int NPLANE = is_c0(MODE) ? 1 : 2 * CJ;
  static constexpr int NPASS = is_c0(MODE) ? 1 : CIN / CIN_P;
  static constexpr bool RESIDENT = NPASS == 1;
  static constexpr int TAPS = (MODE == M2_TRF) ? 8 : (is_c0(MODE) || MODE == M2_PB || KD == 1) ? 9 : 27;
  static constexpr int NMMA = (MODE == M2_TRF) ? 8 * NB : NB;  // N of one tcgen05.mma
  static constexpr int A_BYTES = NPLANE * PLANE;
  static constexpr int A_LBO = is_c0(MODE) ? 32 : PLANE;
  static constexpr int A_SBO = (MODE == M2_S2) ? 2 * BW * 16 : BW * 16;
  static constexpr int B_TILE = 2 * NMMA * 16;
  static constexpr int B_BYTES = CJ * TAPS * B_TILE;
  static constexpr int TX_BYTES = NPLANE * NBLK * BLK_BYTES + (RESIDENT ? 0 : B_BYTES);
  static constexpr int STAGE_BYTES = A_BYTES + (RESIDENT ? 0 : pad128(B_BYTES));
  static constexpr int OFF_B = STAGES * STAGE_BYTES;
  static constexpr int OFF_BAR = OFF_B + (RESIDENT ? pad128(B_BYTES) : 0);
  static constexpr int SMEM = OFF_BAR + 8 * (2 * STAGES + 4) + 16 + 128;  // + slack for manual 128-byte alignment
  static constexpr int NACC = is_tr(MODE) ? (KD == 1 ? 4 : 8) * TD : (MODE == M2_PB) ? TD + 2 : TD;
  static constexpr int COLS = NACC * NB;
  static constexpr int ACC_SETS = (2 * COLS <= 512) ? 2 : 1;
  static constexpr int TMEM_COLS = pow2c(ACC_SETS * COLS);
  static constexpr int PROD_WARPS = (MODE == M2_C0) ? 12 : 1;
  // epilogue warps: TMEM quadrant = warp % 4, so a multiple of 4; the full-resolution layers' epilogues are latency-bound
  // instruction streams (2 warps per scheduler issue 30-40 % of the cycles), so the cheap-in-registers modes get 16
  static constexpr int EPI_WARPS = (MODE == M2_PB || MODE == M2_C0T || MODE == M2_TRF) ? 16 : (MODE == M2_TR || MODE == M2_S2) ? 12 : 8;
  static constexpr int NPART = EPI_WARPS / 4;
  // one MMA-issuing warp per accumulator set: tile k is issued by warp k % MMA_WARPS into set k % 2, so the (serial,
  // single-thread) descriptor arithmetic of consecutive tiles overlaps
  // Each issuing warp owns a private slice of the stage ring (stage = MMA_WARPS*(j % HS) + warp, j = its own item count):
  // a full/empty mbarrier is then always consumed by one warp in order - parity waits cannot alias a phase two uses away.
  static constexpr int MMA_WARPS = (TC2_MMA_WARPS == 2 && ACC_SETS == 2 && STAGES % 2 == 0) ? 2 : 1;
  static constexpr int HS = STAGES / MMA_WARPS;
  // SUB issuing threads share one tile: thread `sub` issues the accumulator rows (planes) t with t % SUB == sub.  The MMAs of
  // different planes write different accumulators, so no order is needed between the threads; both commit to the stage's
  // empty barrier and to the accumulator-full barrier (arrival count SUB).  A single thread issues about one tcgen05.mma
  // per ~100 cycles; with two tile-alternating issuers the layers sit at 64 cycles per MMA, splitting prob / conv0 tiles over
  // two more threads brings them to 60 (4-way: no further gain - the smem -> tensor-core A read is the floor).
  static constexpr int SUB = (KD == 3 && TD >= 2 && (MODE == M2_PB || MODE == M2_C0T)) ? SUBISSUE : 1;  // S1 (conv2) measured slower with it
  static constexpr int ISSUE_WARPS = MMA_WARPS * SUB;
  static constexpr int THREADS = (PROD_WARPS + ISSUE_WARPS + EPI_WARPS) * 32;
  static_assert(COLS <= 512, "accumulators exceed TMEM");
  static_assert(SMEM <= 227 * 1024, "pipeline does not fit shared memory");
  static_assert(CIN % CIN_P == 0 || is_c0(MODE), "channel passes");
};

struct Tile2 {
  int x0, y0, z0, b;
};
__device__ __forceinline__ Tile2 decode2(const Tc2Params& p, int lt, int td) {
  Tile2 t;
  t.x0 = (lt % p.tiles_x) * T_W;
  lt /= p.tiles_x;
  t.y0 = (lt % p.tiles_y) * T_H;
  lt /= p.tiles_y;
  t.z0 = (lt % p.tiles_z) * td;
  t.b = lt / p.tiles_z;
  return t;
}

// ------------------------------------------------------------------------------------------------ MMA issue (one thread)
// The issuing thread is the serial resource of a persistent CTA, so the whole tap / plane / chunk nest is unrolled at
// compile time: every descriptor is "base descriptor + constant" (the 14-bit address field never carries: smem < 256 KB),
// i.e. one 64-bit add per operand and the tcgen05.mma itself.
template <int MODE, int CIN, int CIN_P, int NB, int TD, int STAGES, int KD>
__device__ __forceinline__ void issue2(uint64_t adesc0, uint64_t bdesc0, uint32_t acc_base, bool fresh, int sub) {
  using Cfg = C2<MODE, CIN, CIN_P, NB, TD, STAGES, KD>;
  constexpr uint32_t idesc = make_idesc(Cfg::NMMA);
  const uint32_t fresh_acc = fresh ? 0u : 1u;
#pragma unroll
  for (int t = 0; t < ((MODE == M2_PB) ? TD + 2 : TD); ++t) {
    if (Cfg::SUB > 1 && (t % Cfg::SUB) != sub) continue;
#pragma unroll
    for (int tap = 0; tap < Cfg::TAPS; ++tap) {
      int acc, off;
      bool first;
      if (MODE == M2_PB) {
        // t is the staged INPUT plane; its accumulator collects the (kh,kw) taps for all three kd at once
        const int kh = tap / 3, kw = tap % 3;
        acc = t;
        off = ((t * Cfg::SH + kh) * Cfg::BW + kw) * 16;
        first = tap == 0;
      } else if (MODE == M2_TRF) {
        // `tap` is the input shift (sz,sy,sx); its MMA covers the 8 class accumulators of plane t in one go
        const int sz = tap >> 2, sy = (tap >> 1) & 1, sx = tap & 1;
        acc = t * 8;
        off = (((t + sz) * Cfg::SH + sy) * Cfg::BW + sx) * 16;
        first = tap == 0;  // shift 0 has a tap for every class: it initialises all 8 accumulators
      } else if (MODE == M2_TR) {
        // tap (kz,ky,kx) of the transposed kernel feeds output parity class (kz!=1, ky!=1, kx!=1); k = 0 reads input +1
        const int kz = (KD == 3) ? tap / 9 : 1, ky = (tap % 9) / 3, kx = tap % 3;
        const int pz = kz != 1, py = ky != 1, px = kx != 1;
        acc = (KD == 3) ? t * 8 + pz * 4 + py * 2 + px : t * 4 + py * 2 + px;
        off = (((t + (kz == 0)) * Cfg::SH + (ky == 0)) * Cfg::BW + (kx == 0)) * 16;
        first = (kz == (pz ? 0 : 1)) && (ky == (py ? 0 : 1)) && (kx == (px ? 0 : 1));
      } else if (is_c0(MODE)) {
        const int kd = tap / 3, kh = tap % 3;
        acc = t;
        off = (((t + kd) * Cfg::SH + kh) * Cfg::BW) * 16;
        first = tap == 0;
      } else {
        const int kd = (KD == 3) ? tap / 9 : 0, kh = (tap % 9) / 3, kw = tap % 3;
        acc = t;
        first = tap == 0;
        if (MODE == M2_S2) {
          const int row = ((KD == 3 ? 2 * t : t) + kd) * Cfg::SH + kh;  // even block: kw = 1; odd block: kw = 0 at column 0, kw = 2 at column 1
          off = (kw == 1 ? 0 : Cfg::BLK_PITCH) + (row * Cfg::BW + (kw == 2 ? 1 : 0)) * 16;
        } else {
          off = (((t + kd) * Cfg::SH + kh) * Cfg::BW + kw) * 16;
        }
      }
#pragma unroll
      for (int j = 0; j < Cfg::CJ; ++j) {
        const uint64_t ad = adesc0 + (uint64_t)(((is_c0(MODE) ? 0 : (2 * j) * Cfg::PLANE) + off) >> 4);
        const uint64_t bd = bdesc0 + (uint64_t)(((j * Cfg::TAPS + tap) * Cfg::B_TILE) >> 4);
        umma_f16(acc_base + acc * NB, ad, bd, idesc, (first && j == 0) ? fresh_acc : 1u);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ epilogue
// One warp: TMEM lanes 32*(warp%4)..+31; `part` in [0, NPART) splits the planes (or plane x parity units for TR) between
// the two warps that share a quadrant.
template <int MODE, int NB, int TD, int KD, int NPART>
__device__ __forceinline__ void epilogue2(const Tc2Params& p, const Tile2& tc, uint32_t acc_base, int q, int lane, int part) {
  constexpr int COUT_P = NB / 2;
  const int hl = q * 4 + (lane >> 3), wl = lane & 7;
  const uint32_t lane_addr = acc_base + ((uint32_t)(q * 32) << 16);
  const int npo = p.Cout / 4;  // planes per batch entry of a CH16 output with Cout channels
  if (is_tr(MODE)) {
    const int iy = tc.y0 + hl, ix = tc.x0 + wl;
    const bool in_img = (iy < p.Hi) && (ix < p.Wi);
    uint4* yc = reinterpret_cast<uint4*>(p.y);
    constexpr int NZY = (KD == 3) ? 4 : 2;  // (pz, py) output parity classes per input plane
    for (int u = part; u < TD * NZY; u += NPART) {
      const int t = u / NZY, pzy = u % NZY;
      if (tc.z0 + t >= p.Di) break;  // warp-uniform
      const int oz = (KD == 3) ? 2 * (tc.z0 + t) + (pzy >> 1) : tc.z0 + t, oy = 2 * iy + (pzy & 1);
      const uint32_t te = lane_addr + (t * 2 * NZY + pzy * 2) * NB, to = te + NB;
#pragma unroll 1
      for (int c0 = 0; c0 < COUT_P; c0 += 8) {
        // the skip cells do not depend on the accumulators: their (HBM-latency) loads go out first, then the four TMEM
        // loads, then ONE wait - the latencies overlap instead of adding up
        const bool live = in_img && c0 < p.Cout;
        const int ph = (c0 >> 3) * 2;
        uint4 he = make_uint4(0, 0, 0, 0), hle = he, ho = he, hlo = he;
        if (p.skip && live) {  // CH16P tensor: even output column 2ix -> parity 0 cell ix, odd column -> parity 1 cell ix
          const long long ce = cell_index(FMT_CH16P, tc.b, npo, ph, p.Do, p.Ho, p.Wo, oz, oy, 2 * ix);
          const long long co_ = cell_index(FMT_CH16P, tc.b, npo, ph, p.Do, p.Ho, p.Wo, oz, oy, 2 * ix + 1);
          const long long pstride = (long long)p.Do * p.Ho * (2 * ((p.Wo + 1) >> 1));
          he = __ldg(p.skip + ce); hle = __ldg(p.skip + ce + pstride);
          ho = __ldg(p.skip + co_); hlo = __ldg(p.skip + co_ + pstride);
        }
        uint32_t re[8], rle[8], ro[8], rlo[8];
        tmem_ld8_issue(te + c0, re);
        tmem_ld8_issue(te + COUT_P + c0, rle);
        tmem_ld8_issue(to + c0, ro);
        tmem_ld8_issue(to + COUT_P + c0, rlo);
        tmem_wait_ld();
        if (!live) continue;
        float e[8], o[8], le[8], lo8[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          e[c] = __uint_as_float(re[c]); le[c] = __uint_as_float(rle[c]);
          o[c] = __uint_as_float(ro[c]); lo8[c] = __uint_as_float(rlo[c]);
        }
        float se[8], so[8];
        if (p.skip) {
          float a[8], b2[8];
          unpack_cell(he, a); unpack_cell(hle, b2);
#pragma unroll
          for (int c = 0; c < 8; ++c) se[c] = a[c] + b2[c];
          unpack_cell(ho, a); unpack_cell(hlo, b2);
#pragma unroll
          for (int c = 0; c < 8; ++c) so[c] = a[c] + b2[c];
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float ve = e[c] + le[c], vo = o[c] + lo8[c];
          if (p.scale) {
            const float sc = __ldg(p.scale + c0 + c), sh = __ldg(p.shift + c0 + c);
            ve = fmaf(ve, sc, sh);
            vo = fmaf(vo, sc, sh);
          }
          if (p.relu) { ve = fmaxf(ve, 0.f); vo = fmaxf(vo, 0.f); }
          if (p.skip) { ve += se[c]; vo += so[c]; }
          e[c] = ve; o[c] = vo;
        }
        uint4 ehi, elo, ohi, olo;
        split_pack8(e, ehi, elo);
        split_pack8(o, ohi, olo);
        const long long cell = cell_index(FMT_CH16, tc.b, npo, ph, p.Do, p.Ho, p.Wo, oz, oy, 2 * ix);
        const long long pstride = (long long)p.Do * p.Ho * p.Wo;
        yc[cell] = ehi; yc[cell + 1] = ohi;
        yc[cell + pstride] = elo; yc[cell + pstride + 1] = olo;
      }
    }
  } else if (MODE == M2_PB) {
    // columns of accumulator s (input plane z0 - 1 + s): [kd][hi co0, hi co1, lo co0, lo co1]; out[t] = sum_kd P[t + kd][kd]
    const int oy = tc.y0 + hl, ox = tc.x0 + wl;
    const bool in_img = (oy < p.Ho) && (ox < p.Wo);
    float* yf = reinterpret_cast<float*>(p.y);
    const long long oplane = (long long)p.Ho * p.Wo;
    for (int t = part; t < TD; t += NPART) {
      const int oz = tc.z0 + t;
      if (oz >= p.Do) break;  // warp-uniform
      uint32_t r0[8], r1[8], r2[8];
      tmem_ld8_issue(lane_addr + (t + 0) * NB, r0);
      tmem_ld8_issue(lane_addr + (t + 1) * NB, r1);
      tmem_ld8_issue(lane_addr + (t + 2) * NB + 8, r2);
      tmem_wait_ld();
      if (!in_img) continue;
      float a0[8], a1[8], b2[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) { a0[c] = __uint_as_float(r0[c]); a1[c] = __uint_as_float(r1[c]); b2[c] = __uint_as_float(r2[c]); }
      // kd = 0 -> columns 0..3 of plane t; kd = 1 -> columns 4..7 of plane t+1; kd = 2 -> columns 8..11 (= b2[0..3]) of plane t+2
#pragma unroll
      for (int co = 0; co < 2; ++co) {
        float v = (a0[co] + a0[2 + co]) + (a1[4 + co] + a1[6 + co]) + (b2[co] + b2[2 + co]);
        if (p.scale) v = fmaf(v, __ldg(p.scale + co), __ldg(p.shift + co));
        if (p.relu) v = fmaxf(v, 0.f);
        if (co < p.Cout) yf[(long long)tc.b * p.y_bs + ((long long)co * p.Do + oz) * oplane + (long long)oy * p.Wo + ox] = v;
      }
    }
  } else {
    const int oy = tc.y0 + hl, ox = tc.x0 + wl;
    const bool in_img = (oy < p.Ho) && (ox < p.Wo);
    for (int t = part; t < TD; t += NPART) {
      const int oz = tc.z0 + t;
      if (oz >= p.Do) break;  // warp-uniform
      const uint32_t taddr = lane_addr + t * NB;
#pragma unroll 1
      for (int c0 = 0; c0 < COUT_P; c0 += 8) {
        uint32_t rv[8], rl[8];
        tmem_ld8_issue(taddr + c0, rv);
        tmem_ld8_issue(taddr + COUT_P + c0, rl);
        tmem_wait_ld();
        if (!in_img || c0 >= p.Cout) continue;
        float v[8], l8[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) { v[c] = __uint_as_float(rv[c]); l8[c] = __uint_as_float(rl[c]); }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float a = v[c] + l8[c];
          if (p.scale && c0 + c < p.Cout) a = fmaf(a, __ldg(p.scale + c0 + c), __ldg(p.shift + c0 + c));
          if (p.relu) a = fmaxf(a, 0.f);
          v[c] = a;
        }
        if (p.out_fmt == FMT_F32) {
          float* yf = reinterpret_cast<float*>(p.y);
          const long long oplane = (long long)p.Ho * p.Wo;
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (c0 + c < p.Cout)
              yf[(long long)tc.b * p.y_bs + ((long long)(c0 + c) * p.Do + oz) * oplane + (long long)oy * p.Wo + ox] = v[c];
        } else if (p.out_fmt == FMT_NHWC2 || p.out_fmt == FMT_NHWC2H) {
          // two channel-last fp32 buffers back to back, [2][B][Do][Ho][Wo][Cout/2]: the lane's 8 channels are 32 contiguous bytes
          const int half = p.Cout >> 1;
          const int hsel = c0 >= half ? 1 : 0, cc = c0 - hsel * half;
          const long long set = (long long)p.B * p.Do * p.Ho * p.Wo * half;
          const long long at = (long long)hsel * set + ((((long long)tc.b * p.Do + oz) * p.Ho + oy) * p.Wo + ox) * half + cc;
          float* yb = reinterpret_cast<float*>(p.y) + at;
          *reinterpret_cast<float4*>(yb) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(yb + 4) = make_float4(v[4], v[5], v[6], v[7]);
          if (p.out_fmt == FMT_NHWC2H) {  // + the same two sets rounded to fp16 behind them (W1's source-map format): 16 bytes per lane
            __half* yh = reinterpret_cast<__half*>(reinterpret_cast<float*>(p.y) + 2 * set) + at;
            const __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
            const __half2 h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
            *reinterpret_cast<uint4*>(yh) = make_uint4(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1),
                                                       *reinterpret_cast<const uint32_t*>(&h2), *reinterpret_cast<const uint32_t*>(&h3));
          }
        } else {
          uint4 hi, lo;
          split_pack8(v, hi, lo);
          const int ph = (c0 >> 3) * 2;
          const long long cell = cell_index(p.out_fmt, tc.b, npo, ph, p.Do, p.Ho, p.Wo, oz, oy, ox);
          const long long pstride = (p.out_fmt == FMT_CH16P) ? (long long)p.Do * p.Ho * (2 * ((p.Wo + 1) >> 1))
                                                             : (long long)p.Do * p.Ho * p.Wo;
          uint4* yc = reinterpret_cast<uint4*>(p.y);
          yc[cell] = hi;
          yc[cell + pstride] = lo;
        }
      }
    }
  }
}

// Transposed layers: the skip cells of a tile do not depend on its accumulators, but their loads sit on the epilogue's critical
// path (HBM latency per unit, 2 units per warp and tile).  While the MMAs of tile k run, the epilogue warps pull the skip lines of
// tile k + 1 into L2; one lane per 8-pixel row segment touches each 128-byte line (hi / lo plane x even / odd columns).
template <int MODE, int NB, int TD, int KD, int NPART>
__device__ __forceinline__ void prefetch_skip2(const Tc2Params& p, const Tile2& tc, int q, int lane, int part) {
  if (!is_tr(MODE) || !p.skip || (lane & 7) != 0) return;
  constexpr int COUT_P = NB / 2;
  const int hl = q * 4 + (lane >> 3);
  const int iy = tc.y0 + hl, ix = tc.x0;
  if (iy >= p.Hi || ix >= p.Wi) return;
  const int npo = p.Cout / 4;
  constexpr int NZY = (KD == 3) ? 4 : 2;
  const long long pstride = (long long)p.Do * p.Ho * (2 * ((p.Wo + 1) >> 1));
  for (int u = part; u < TD * NZY; u += NPART) {
    const int t = u / NZY, pzy = u % NZY;
    if (tc.z0 + t >= p.Di) break;
    const int oz = (KD == 3) ? 2 * (tc.z0 + t) + (pzy >> 1) : tc.z0 + t, oy = 2 * iy + (pzy & 1);
    for (int c0 = 0; c0 < COUT_P && c0 < p.Cout; c0 += 8) {
      const int ph = (c0 >> 3) * 2;
      const uint4* ce = p.skip + cell_index(FMT_CH16P, tc.b, npo, ph, p.Do, p.Ho, p.Wo, oz, oy, 2 * ix);
      const uint4* co_ = p.skip + cell_index(FMT_CH16P, tc.b, npo, ph, p.Do, p.Ho, p.Wo, oz, oy, 2 * ix + 1);
      asm volatile("prefetch.global.L2 [%0];" ::"l"(ce));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(ce + pstride));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(co_));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(co_ + pstride));
    }
  }
}

// ------------------------------------------------------------------------------------------------ the kernel
template <int MODE, int CIN, int CIN_P, int NB, int TD, int STAGES, int KD>
__global__ void __launch_bounds__(C2<MODE, CIN, CIN_P, NB, TD, STAGES, KD>::THREADS, 1)
    conv_tc2_kernel(const __grid_constant__ Tc2Params p, const __grid_constant__ CUtensorMap tmap) {
  using Cfg = C2<MODE, CIN, CIN_P, NB, TD, STAGES, KD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  uint8_t* sB = smem + Cfg::OFF_B;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* empty = full + STAGES;
  uint64_t* accfull = empty + STAGES;
  uint64_t* accempty = accfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accempty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int MMA_WARP = Cfg::PROD_WARPS;

  if (warp == MMA_WARP) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full + s, (MODE == M2_C0) ? Cfg::PROD_WARPS * 32 : 1);
      mbar_init(empty + s, Cfg::SUB);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(accfull + a, Cfg::SUB);
      mbar_init(accempty + a, Cfg::EPI_WARPS * 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (Cfg::RESIDENT) {  // the layer's weights stay in smem for the life of the CTA
    uint4* wdst = reinterpret_cast<uint4*>(sB);
    for (int i = tid; i < Cfg::B_BYTES / 16; i += Cfg::THREADS) wdst[i] = __ldg(p.wtc + i);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // activations (and the skip tensor) are the predecessor's output: nothing of them is touched before this point
  pdl_launch();
  pdl_wait();

  if (warp < Cfg::PROD_WARPS) {
    // ---------------------------------------------------------------- producer
    if (MODE == M2_C0) {
      // fp32 cost volume (2 channels) -> cells [hi c0, hi c1, lo c0, lo c1](x) ++ the same of x+1, filled by 12 warps
      const long long iplane = (long long)p.Hi * p.Wi, cs = (long long)p.Di * iplane;
      int k = 0;
      for (int lt = blockIdx.x; lt < p.n_tiles; lt += gridDim.x, ++k) {
        const int m = k % Cfg::MMA_WARPS, j = k / Cfg::MMA_WARPS;
        const int s = Cfg::MMA_WARPS * (j % Cfg::HS) + m, u = j / Cfg::HS;
        const Tile2 tc = decode2(p, lt, TD);
        mbar_wait(empty + s, (u & 1) ^ 1);
        uint8_t* sA = smem + s * Cfg::STAGE_BYTES;
        // three cells per thread per trip, all 12 loads issued before any conversion (the fill is latency bound)
        constexpr int NCELL = Cfg::ROWS * Cfg::BW, PT = Cfg::PROD_WARPS * 32, UNR = 3;
        for (int base = tid; base < NCELL; base += UNR * PT) {
          float v[UNR][4];
#pragma unroll
          for (int r = 0; r < UNR; ++r) {
            const int sv = base + r * PT;
            v[r][0] = v[r][1] = v[r][2] = v[r][3] = 0.f;
            if (sv < NCELL) {
              const int sx = sv % Cfg::BW, sy = (sv / Cfg::BW) % Cfg::SH, sz = sv / (Cfg::BW * Cfg::SH);
              const int ix = tc.x0 + sx - 1, iy = tc.y0 + sy - 1, iz = tc.z0 + sz - 1;
              if (iy >= 0 && iy < p.Hi && iz >= 0 && iz < p.Di) {
                const float* src = p.x_f32 + (long long)tc.b * p.x_bs + (long long)iz * iplane + (long long)iy * p.Wi;
                if (ix >= 0 && ix < p.Wi) { v[r][0] = __ldg(src + ix); v[r][1] = __ldg(src + cs + ix); }
                if (ix + 1 >= 0 && ix + 1 < p.Wi) { v[r][2] = __ldg(src + ix + 1); v[r][3] = __ldg(src + cs + ix + 1); }
              }
            }
          }
#pragma unroll
          for (int r = 0; r < UNR; ++r) {
            const int sv = base + r * PT;
            if (sv >= NCELL) break;
            __half h[4], l[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              h[i] = __float2half_rn(v[r][i]);
              l[i] = __float2half_rn(v[r][i] - __half2float(h[i]));
            }
            const __half2 w0 = __halves2half2(h[0], h[1]), w1 = __halves2half2(l[0], l[1]);
            const __half2 w2 = __halves2half2(h[2], h[3]), w3 = __halves2half2(l[2], l[3]);
            *reinterpret_cast<uint4*>(sA + sv * 16) =
                make_uint4(*reinterpret_cast<const uint32_t*>(&w0), *reinterpret_cast<const uint32_t*>(&w1),
                           *reinterpret_cast<const uint32_t*>(&w2), *reinterpret_cast<const uint32_t*>(&w3));
          }
        }
        fence_proxy_async();
        mbar_arrive(full + s);
      }
    } else {
      // the whole warp walks the tile loop, one elected lane issues the TMA box loads into the UMMA layout (coordinates and
      // addresses stay in uniform registers); OOB zero fill is the convolution's padding
      if (lane == 0) prefetch_tmap(&tmap);
      const int planes_per_b = 2 * CIN / 8;
      int tile_k = 0;
      for (int lt = blockIdx.x; lt < p.n_tiles; lt += gridDim.x, ++tile_k) {
        const Tile2 tc = decode2(p, lt, TD);
        for (int pass = 0; pass < Cfg::NPASS; ++pass) {
          const int m = tile_k % Cfg::MMA_WARPS, j = (tile_k / Cfg::MMA_WARPS) * Cfg::NPASS + pass;
          const int s = Cfg::MMA_WARPS * (j % Cfg::HS) + m, u = j / Cfg::HS;
          mbar_wait(empty + s, (u & 1) ^ 1);
          uint8_t* st = smem + s * Cfg::STAGE_BYTES;
          if (elect_one()) {
          mbar_expect_tx(full + s, Cfg::TX_BYTES);
#pragma unroll 1
          for (int pl = 0; pl < Cfg::NPLANE; ++pl) {
            const int gpl = tc.b * planes_per_b + pass * Cfg::NPLANE + pl;
            uint8_t* dst = st + pl * Cfg::PLANE;
            if (MODE == M2_C0T) {  // cost cells: cell x = [voxel x-1 | voxel x], one plane per batch entry
              tma_load_4d(dst, &tmap, full + s, 8 * tc.x0, tc.y0 - 1, tc.z0 - 1, tc.b);
            } else if (MODE == M2_S1 || MODE == M2_PB) {
              tma_load_4d(dst, &tmap, full + s, 8 * (tc.x0 - 1), tc.y0 - 1, (KD == 3) ? tc.z0 - 1 : tc.z0, gpl);
            } else if (is_tr(MODE)) {
              tma_load_4d(dst, &tmap, full + s, 8 * tc.x0, tc.y0, tc.z0, gpl);
            } else {  // S2 on a CH16P tensor: even columns 2*(x0+i) = even cell x0+i; odd columns 2*(x0+i)-1 = odd cell x0+i-1
              const int zc = (KD == 3) ? 2 * tc.z0 - 1 : tc.z0;
              tma_load_5d(dst, &tmap, full + s, 8 * tc.x0, 0, 2 * tc.y0 - 1, zc, gpl);
              tma_load_5d(dst + Cfg::BLK_PITCH, &tmap, full + s, 8 * (tc.x0 - 1), 1, 2 * tc.y0 - 1, zc, gpl);
            }
          }
          if (!Cfg::RESIDENT)
            bulk_load(st + Cfg::A_BYTES, p.wtc + (size_t)pass * (Cfg::B_BYTES / 16), Cfg::B_BYTES, full + s);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp < MMA_WARP + Cfg::ISSUE_WARPS) {
    // ---------------------------------------------------------------- MMA issuers (one thread each)
    const int me = (warp - MMA_WARP) % Cfg::MMA_WARPS, sub = (warp - MMA_WARP) / Cfg::MMA_WARPS;
    int tile_k = 0;
    for (int lt = blockIdx.x; lt < p.n_tiles; lt += gridDim.x, ++tile_k) {
      // The WHOLE warp walks the tile loop and one elected lane issues: descriptors computed in warp-uniform code stay in uniform
      // registers (one UTCHMMA per MMA in SASS); under `if (lane == 0)` every tcgen05.mma was wrapped in an ~11-instruction R2UR / ELECT
      // waterfall on the serial path of the issuing thread
      if ((tile_k % Cfg::MMA_WARPS) == me) {
        const int a = (Cfg::ACC_SETS == 2) ? (tile_k & 1) : 0, v = (Cfg::ACC_SETS == 2) ? (tile_k >> 1) : tile_k;
        mbar_wait(accempty + a, (v & 1) ^ 1);
        for (int pass = 0; pass < Cfg::NPASS; ++pass) {
          const int j = (tile_k / Cfg::MMA_WARPS) * Cfg::NPASS + pass;
          const int s = Cfg::MMA_WARPS * (j % Cfg::HS) + me, u = j / Cfg::HS;
          mbar_wait(full + s, u & 1);
          tc_fence_after();
          uint8_t* st = smem + s * Cfg::STAGE_BYTES;
          const uint64_t adesc0 = make_desc(smem_u32(st), Cfg::A_LBO, Cfg::A_SBO);
          const uint64_t bdesc0 = make_desc(Cfg::RESIDENT ? smem_u32(sB) : smem_u32(st + Cfg::A_BYTES), Cfg::NMMA * 16, 128);
          if (elect_one()) {
            issue2<MODE, CIN, CIN_P, NB, TD, STAGES, KD>(adesc0, bdesc0, tmem_base + a * Cfg::COLS, pass == 0, sub);
            umma_commit(empty + s);
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(accfull + a);
      }
      __syncwarp();
    }
  } else {
    // ---------------------------------------------------------------- epilogue (8 warps)
    const int ew = warp - MMA_WARP - Cfg::ISSUE_WARPS;
    const int q = warp & 3, part = ew >> 2;
    int tile_k = 0;
    for (int lt = blockIdx.x; lt < p.n_tiles; lt += gridDim.x, ++tile_k) {
      const int a = (Cfg::ACC_SETS == 2) ? (tile_k & 1) : 0, v = (Cfg::ACC_SETS == 2) ? (tile_k >> 1) : tile_k;
      if (is_tr(MODE) && p.skip_prefetch && lt + (int)gridDim.x < p.n_tiles)
        prefetch_skip2<MODE, NB, TD, KD, Cfg::NPART>(p, decode2(p, lt + gridDim.x, TD), q, lane, part);
      mbar_wait(accfull + a, v & 1);
      tc_fence_after();
      epilogue2<MODE, NB, TD, KD, Cfg::NPART>(p, decode2(p, lt, TD), tmem_base + a * Cfg::COLS, q, lane, part);
      tc_fence_before();
      mbar_arrive(accempty + a);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------ layout converters
__global__ void __launch_bounds__(256) f32_to_ch16_kernel(const float* __restrict__ x, uint4* __restrict__ y, int C, int D, int H,
                                                          int W, int fmt, long long n_cells) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;  // one thread per (b, chunk j, z, y, x)
  if (i >= n_cells) return;
  const int xx = (int)(i % W);
  long long r = i / W;
  const int yy = (int)(r % H); r /= H;
  const int zz = (int)(r % D); r /= D;
  const int j = (int)(r % (C / 8));
  const int b = (int)(r / (C / 8));
  const long long vol = (long long)D * H * W;
  const float* src = x + ((long long)b * C + j * 8) * vol + ((long long)zz * H + yy) * W + xx;
  float v[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) v[c] = __ldg(src + c * vol);
  uint4 hi, lo;
  split_pack8(v, hi, lo);
  const long long cell = cell_index(fmt, b, C / 4, 2 * j, D, H, W, zz, yy, xx);
  const long long pstride = (fmt == FMT_CH16P) ? (long long)D * H * (2 * ((W + 1) >> 1)) : vol;
  y[cell] = hi;
  y[cell + pstride] = lo;
}

__global__ void __launch_bounds__(256) ch16_to_f32_kernel(const uint4* __restrict__ x, float* __restrict__ y, int C, int D, int H,
                                                          int W, int fmt, long long n_cells) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n_cells) return;
  const int xx = (int)(i % W);
  long long r = i / W;
  const int yy = (int)(r % H); r /= H;
  const int zz = (int)(r % D); r /= D;
  const int j = (int)(r % (C / 8));
  const int b = (int)(r / (C / 8));
  const long long vol = (long long)D * H * W;
  const long long cell = cell_index(fmt, b, C / 4, 2 * j, D, H, W, zz, yy, xx);
  const long long pstride = (fmt == FMT_CH16P) ? (long long)D * H * (2 * ((W + 1) >> 1)) : vol;
  float a[8], l[8];
  unpack_cell(__ldg(x + cell), a);
  unpack_cell(__ldg(x + cell + pstride), l);
  float* dst = y + ((long long)b * C + j * 8) * vol + ((long long)zz * H + yy) * W + xx;
#pragma unroll
  for (int c = 0; c < 8; ++c) dst[c * vol] = a[c] + l[c];
}

// ------------------------------------------------------------------------------------------------ host side
int conv_layer_kf(const Tc2Params& p, const void* x, int in_cells, cudaStream_t st);  // conv_kf.cu

int g_tc2_max_ctas = 1;
int g_tc2_skip_prefetch = 1;  // dmvs_debug_set("tc2_skip_prefetch", 0 | 1): transposed layers pull the next tile's skip lines into L2
int g_tc2_pdl = 1;      // programmatic dependent launch of the tensor convs (dmvs_debug_set("tc2_pdl", 0 | 1))
int g_pb_td8 = 1;       // debug knob (dmvs_debug_set("pb_td8", 0 | 1)): prob layer with 8-plane tiles  // debug knob (dmvs_debug_set("tc2_max_ctas", n))

template <int MODE, int CIN, int CIN_P, int NB, int TD, int STAGES, int KD = 3>
static int launch2(Tc2Params p, const void* x, cudaStream_t st) {
  using Cfg = C2<MODE, CIN, CIN_P, NB, TD, STAGES, KD>;
  const int gw = is_tr(MODE) ? p.Wi : p.Wo, gh = is_tr(MODE) ? p.Hi : p.Ho, gd = is_tr(MODE) ? p.Di : p.Do;
  p.tiles_x = ceil_div(gw, T_W);
  p.tiles_y = ceil_div(gh, T_H);
  p.tiles_z = ceil_div(gd, TD);
  p.n_tiles = p.tiles_x * p.tiles_y * p.tiles_z * p.B;
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  if (MODE == M2_C0) {
    p.x_f32 = reinterpret_cast<const float*>(x);
    p.x_bs = 2LL * p.Di * p.Hi * p.Wi;
  } else if (MODE == M2_C0T) {  // [B][D][H][W+1] cells viewed as a CH16 tensor of width W+1 with one plane per batch entry
    const int rc = make_tmap(&tmap, x, FMT_CH16, p.B, p.Di, p.Hi, p.Wi + 1, Cfg::BW, Cfg::SH, Cfg::SD);
    if (rc != DMVS_OK) return rc;
  } else {
    const int rc = make_tmap(&tmap, x, MODE == M2_S2 ? FMT_CH16P : FMT_CH16, p.B * 2 * CIN / 8, p.Di, p.Hi, p.Wi, Cfg::BW, Cfg::SH, Cfg::SD);
    if (rc != DMVS_OK) return rc;
  }
  auto kern = conv_tc2_kernel<MODE, CIN, CIN_P, NB, TD, STAGES, KD>;
  static PerDevice state;  // per template instance
  const int slot = current_device_slot();
  DMVS_REQUIRE(slot >= 0, DMVS_ERR_CUDA, "conv_tc2: no current CUDA device");
  if (!state.configured[slot]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    if (e != cudaSuccess) {
      set_error("conv_tc2: cudaFuncSetAttribute(%d bytes): %s", Cfg::SMEM, cudaGetErrorString(e));
      return DMVS_ERR_CUDA;
    }
    int occ = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, Cfg::THREADS, Cfg::SMEM) != cudaSuccess || occ < 1) occ = 1;
    const int by_tmem = 512 / Cfg::TMEM_COLS;
    state.value[slot] = occ < by_tmem ? occ : by_tmem;
    if (state.value[slot] < 1) state.value[slot] = 1;
    state.configured[slot] = true;
  }
  // persistent CTAs: as many per SM as shared memory, registers AND the 512 TMEM columns allow (the memory-bound
  // full-resolution layers need the second CTA's loads in flight to cover HBM latency), never more than g_tc2_max_ctas
  const int ctas_per_sm = state.value[slot];
  const int want = kNumSMs * (ctas_per_sm < g_tc2_max_ctas ? ctas_per_sm : g_tc2_max_ctas);
  const int grid = p.n_tiles < want ? p.n_tiles : want;
  if (g_tc2_pdl) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
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
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(p, tmap);
  }
  char what[128];
  snprintf(what, sizeof(what), "conv_tc2<mode %d, Cin %d/%d, N %d, TD %d, stages %d, kd %d> tiles %d", MODE, CIN, CIN_P, NB, TD, STAGES, KD, p.n_tiles);
  return check_launch(what);
}

// One conv block on CH16 activations.  x: CH16 (stride 1 / transposed), CH16P (stride 2) or fp32 NCDHW (Cin == 2);
// skip: CH16P (transposed only); y: out_fmt.  Returns +1 if the shape has no specialisation.
int conv_layer_tc2(const void* x, int in_cells, const dmvs_conv_layer& L, const void* skip, void* y, long long y_bs_f32, int B, int Cin,
                   int Cout, int Di, int Hi, int Wi, int kd, int stride, int transposed, int relu, int out_fmt, cudaStream_t st) {
  if (!L.w_tc || (kd != 1 && kd != 3)) return 1;
  DMVS_REQUIRE(x && y, DMVS_ERR_BAD_POINTER, "conv_tc2: null pointer");
  DMVS_REQUIRE(aligned16(L.w_tc) && (!L.w_tc_kw || aligned16(L.w_tc_kw)) && aligned16(x) && aligned16(y) && (!skip || aligned16(skip)), DMVS_ERR_BAD_POINTER,
               "conv_tc2: pointers must be 16-byte aligned");
  DMVS_REQUIRE(out_fmt == FMT_F32 || out_fmt == FMT_CH16 || out_fmt == FMT_CH16P || out_fmt == FMT_NHWC2 || out_fmt == FMT_NHWC2H,
               DMVS_ERR_BAD_SHAPE,
               "conv_tc2: bad out_fmt %d", out_fmt);
  DMVS_REQUIRE((out_fmt != FMT_NHWC2 && out_fmt != FMT_NHWC2H) || (kd == 1 && !transposed && stride == 1 && Cin == 32 && (Cout == 16 || Cout == 32)),
               DMVS_ERR_BAD_SHAPE,
               "conv_tc2: the split channel-last output exists for FeatureNet's 32-channel 3x3 heads only");
  Tc2Params p;
  memset(&p, 0, sizeof(p));
  p.wtc = reinterpret_cast<const uint4*>(L.w_tc); p.scale = L.scale; p.shift = L.shift;
  p.skip = reinterpret_cast<const uint4*>(skip); p.y = y;
  p.B = B; p.Cin = Cin; p.Cout = Cout; p.Di = Di; p.Hi = Hi; p.Wi = Wi; p.relu = relu; p.out_fmt = out_fmt;
  p.y_bs = y_bs_f32;
  p.skip_prefetch = g_tc2_skip_prefetch;
  if (kd == 1) {  // 2-D layers (the refine net's bottleneck, FeatureNet's 3x3 heads): depth is a batch of planes
    p.Do = Di;
    if (!transposed && stride == 1 && Cin == 64 && Cout == 32) {  // FeatureNet conv2.0 (5x5 stride 2 on 16 ch = 3x3 on 64 unshuffled ch)
      DMVS_REQUIRE(skip == nullptr && out_fmt == FMT_CH16, DMVS_ERR_BAD_SHAPE, "conv_tc2: the 64 -> 32 2-D layer writes CH16, no skip");
      p.Ho = Hi; p.Wo = Wi;
      return launch2<M2_S1, 64, 32, 64, 1, 2, 1>(p, x, st);
    }
    if (!transposed && stride == 1 && ((Cin == 32 && (Cout == 16 || Cout == 32)) || (Cin == 16 && Cout == 16) || (Cin == 8 && Cout == 8))) {
      // FeatureNet's 3x3 layers (out3 / out2 and conv2.1-2 / conv1.1-2): weights resident, all channel chunks in one pass
      DMVS_REQUIRE(skip == nullptr, DMVS_ERR_BAD_SHAPE, "conv_tc2: only transposed convs take a skip input");
      DMVS_REQUIRE(out_fmt == FMT_F32 || out_fmt == FMT_NHWC2 || out_fmt == FMT_NHWC2H || out_fmt == FMT_CH16, DMVS_ERR_BAD_SHAPE,
                   "conv_tc2: FeatureNet layers write fp32 (NCHW or split channel-last) or CH16");
      p.Ho = Hi; p.Wo = Wi;
      if (p.y_bs == 0) p.y_bs = (long long)Cout * Di * Hi * Wi;
      if (Cin == 8) return launch2<M2_S1, 8, 8, 16, 1, 4, 1>(p, x, st);
      if (Cin == 16) return launch2<M2_S1, 16, 16, 32, 1, 4, 1>(p, x, st);
      if (Cout == 16) return launch2<M2_S1, 32, 32, 32, 1, 4, 1>(p, x, st);
      return launch2<M2_S1, 32, 32, 64, 1, 3, 1>(p, x, st);
    }
    DMVS_REQUIRE(out_fmt == FMT_CH16, DMVS_ERR_BAD_SHAPE, "conv_tc2: the 2-D layers write CH16");
    if (transposed) {
      p.Ho = 2 * Hi; p.Wo = 2 * Wi;
      if (Cin == 64 && Cout == 32) return launch2<M2_TR, 64, 8, 64, 1, 4, 1>(p, x, st);  // conv7 (2-D)
      return 1;
    }
    DMVS_REQUIRE(skip == nullptr, DMVS_ERR_BAD_SHAPE, "conv_tc2: only transposed convs take a skip input");
    if (stride == 2) {
      p.Ho = (Hi - 1) / 2 + 1; p.Wo = (Wi - 1) / 2 + 1;
      DMVS_REQUIRE(Wi % 2 == 0, DMVS_ERR_BAD_SHAPE, "conv_tc2: stride-2 input width must be even");
      if (Cin == 32 && Cout == 64) return launch2<M2_S2, 32, 8, 128, 1, 2, 1>(p, x, st);  // conv5 (2-D)
      return 1;
    }
    p.Ho = Hi; p.Wo = Wi;
    if (Cin == 64 && Cout == 64) return launch2<M2_S1, 64, 8, 128, 1, 4, 1>(p, x, st);    // conv6 (2-D)
    return 1;
  }
  if (transposed) {
    p.Do = 2 * Di; p.Ho = 2 * Hi; p.Wo = 2 * Wi;
    DMVS_REQUIRE(out_fmt == FMT_CH16, DMVS_ERR_BAD_SHAPE, "conv_tc2: transposed convs write CH16");
    if (Cin == 16 && Cout == 8) {                                                     // conv11
      if (L.w_tc_kd) {  // taps folded by input shift (8 MMAs per plane and chunk instead of 27)
        p.wtc = reinterpret_cast<const uint4*>(L.w_tc_kd);
        return launch2<M2_TRF, 16, 16, 16, 2, 4>(p, x, st);
      }
      return launch2<M2_TR, 16, 16, 16, 2, 4>(p, x, st);
    }
    if (Cin == 32 && Cout == 16) return launch2<M2_TR, 32, 16, 32, 1, 2>(p, x, st);  // conv9, 2 passes, weights streamed
    if (Cin == 64 && Cout == 32) return launch2<M2_TR, 64, 8, 64, 1, 3>(p, x, st);   // conv7, 8 passes
    return 1;
  }
  DMVS_REQUIRE(skip == nullptr, DMVS_ERR_BAD_SHAPE, "conv_tc2: only transposed convs take a skip input");
  if (stride == 2) {
    p.Do = (Di - 1) / 2 + 1; p.Ho = (Hi - 1) / 2 + 1; p.Wo = (Wi - 1) / 2 + 1;
    DMVS_REQUIRE(Wi % 2 == 0, DMVS_ERR_BAD_SHAPE, "conv_tc2: stride-2 input width must be even");
    if (Cin == 8 && Cout == 16) return launch2<M2_S2, 8, 8, 32, 1, 3>(p, x, st);     // conv1
    if (Cin == 16 && Cout == 32) return launch2<M2_S2, 16, 8, 64, 1, 2>(p, x, st);   // conv3, 2 passes
    if (Cin == 32 && Cout == 64) return launch2<M2_S2, 32, 8, 128, 1, 1>(p, x, st);  // conv5, 4 passes
    return 1;
  }
  p.Do = Di; p.Ho = Hi; p.Wo = Wi;
  if (p.y_bs == 0) p.y_bs = (long long)Cout * Di * Hi * Wi;
  if (L.w_tc_kd) {  // depth tap folded into N, march along z (conv_kf.cu): conv0 on cost cells, conv2, prob
    Tc2Params pk = p;
    pk.wtc = reinterpret_cast<const uint4*>(L.w_tc_kd);
    pk.wtc_wide = reinterpret_cast<const uint4*>(L.w_tc_kw);
    const int rc = conv_layer_kf(pk, x, in_cells, st);
    if (rc <= 0) return rc;
  }
  if (Cin == 2 && Cout == 16 && in_cells) {                                          // conv0 of both branches (conv0_pair)
    if (Di >= 8 && g_pb_td8) return launch2<M2_C0T, 2, 2, 32, 8, 4>(p, x, st);
    return launch2<M2_C0T, 2, 2, 32, 4, 4>(p, x, st);
  }
  if (Cin == 2 && Cout == 8) {                                                       // conv0
    if (in_cells) return launch2<M2_C0T, 2, 2, 16, 4, 4>(p, x, st);
    return launch2<M2_C0, 2, 2, 16, 4, 4>(p, x, st);
  }
  if (Cin == 8 && Cout <= 8) {                                                       // prob (8 -> 2)
    DMVS_REQUIRE(out_fmt == FMT_F32 || Cout == 8, DMVS_ERR_BAD_SHAPE, "conv_tc2: Cout < 8 needs an fp32 output");
    if (Cout == 2 && L.w_tc_kd) {  // prob with kd folded into N
      p.wtc = reinterpret_cast<const uint4*>(L.w_tc_kd);
      // MMAs per output plane = 9 * (TD + 2) / TD: 8-plane tiles (10 accumulators, 2 stages) where the volume has them
      if (Di >= 8 && g_pb_td8) return launch2<M2_PB, 8, 8, 16, 8, 2>(p, x, st);
      return launch2<M2_PB, 8, 8, 16, 4, 4>(p, x, st);
    }
    return launch2<M2_S1, 8, 8, 16, 4, 4>(p, x, st);
  }
  if (Cin == 16 && Cout == 16) return launch2<M2_S1, 16, 16, 32, 2, 2>(p, x, st);    // conv2
  if (Cin == 32 && Cout == 32) return launch2<M2_S1, 32, 8, 64, 2, 2>(p, x, st);     // conv4, 4 passes
  if (Cin == 64 && Cout == 64) return launch2<M2_S1, 64, 8, 128, 1, 1>(p, x, st);    // conv6, 8 passes
  return 1;
}

int convert_layout(const void* x, void* y, int B, int C, int D, int H, int W, int fmt, int to_ch16, cudaStream_t st) {
  DMVS_REQUIRE(x && y && aligned16(x) && aligned16(y), DMVS_ERR_BAD_POINTER, "convert_layout: null or misaligned pointer");
  DMVS_REQUIRE(C % 8 == 0 && (fmt == FMT_CH16 || fmt == FMT_CH16P), DMVS_ERR_BAD_SHAPE, "convert_layout: C=%d fmt=%d", C, fmt);
  DMVS_REQUIRE(fmt != FMT_CH16P || W % 2 == 0, DMVS_ERR_BAD_SHAPE, "convert_layout: CH16P needs an even width");
  const long long n = (long long)B * (C / 8) * D * H * W;
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (to_ch16)
    f32_to_ch16_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(x), reinterpret_cast<uint4*>(y), C, D, H, W, fmt, n);
  else
    ch16_to_f32_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<float*>(y), C, D, H, W, fmt, n);
  return check_launch("convert_layout");
}

}  // namespace dmvs
