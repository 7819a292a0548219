// cspm_device.cuh — device-side building blocks of the B200 PatchMatch engine.
//
// Arithmetic contract: every f64 product / sum / quotient / sqrt on the cost
// and plane path is issued through the *_rn intrinsics, so the compiler can
// never contract a*b+c into an FMA and the operation ORDER is the one the
// reference executes (CSPM/plane_cost/pre_ss_pc.cc:84-116, MSVC SSE2 build:
// no FMA).  IEEE-754 add/mul/div/sqrt are correctly rounded on sm_100a, so the
// results are bit-identical to the reference's scalar code, not merely close.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// Build-time tuning knobs (kernel-variant experiments; the defaults are the measured best).
#ifndef CSPM_SP_RING
#define CSPM_SP_RING 16
#endif
// producer items in flight per thread in the ring sweep (1: one item per loop iteration; 2: two, branch-free).
// Measured at HD (spatial stage, ms): 640 threads ILP 1 / 2: 216.1 / 214.9; 512 threads ILP 2 (128 registers, no
// spills): 208.5; 768 threads ILP 2 (80 registers, spills): 232.6.  Cross-scale and 960x540 are unchanged.
#ifndef CSPM_RING_ILP
#define CSPM_RING_ILP 2
#endif
// (skipping switched-off candidates with a branch instead of computing and discarding them: 227 vs 208 ms; 512 threads
// with one item per iteration: 227 ms)
// (Round 2 also measured "column mode" producers — a thread owns one (pair, window column) and walks the rows 1-5
// at a time with the row-invariant set-up hoisted: 212-246 ms at HD against 208 with both candidates always
// computed, 337-442 ms with compile-time partner direction and skipped candidates (divergence at pair boundaries,
// spills at the 96 registers that 608-thread CTAs allow) — profiles/r02_variants_ring_col.txt; removed.)
// -DCSPM_PROFILE_LANES=1: diagnostic build that histograms, per executed window row of the bulk kernels, how many
// lanes of the warp were still evaluating (read back with cspm_debug_lane_hist; scratch/lane_hist.py).
#ifndef CSPM_PROFILE_LANES
#define CSPM_PROFILE_LANES 0
#endif
#ifndef CSPM_COOP_CS
#define CSPM_COOP_CS 1
#endif
// -DCSPM_SP_PROFILE=1: cycle accounting inside the ring sweep (clock64 + global counters on the critical path of
// CTA 0); compiled out of the default build.
#ifndef CSPM_SP_PROFILE
#define CSPM_SP_PROFILE 0
#endif
#ifndef CSPM_COOP_LANES
#define CSPM_COOP_LANES 8
#endif
#ifndef CSPM_COOP_LANES_CS
#define CSPM_COOP_LANES_CS 8
#endif
#ifndef CSPM_TAP_UNROLL
#define CSPM_TAP_UNROLL 5
#endif
#ifndef CSPM_BULK_BY
#define CSPM_BULK_BY 24
#endif
#ifndef CSPM_BULK_BY_CS
#define CSPM_BULK_BY_CS 20
#endif
#define CSPM_PRAGMA_(x) _Pragma(#x)
#define CSPM_UNROLL(n) CSPM_PRAGMA_(unroll n)

namespace cspm {

// -DCSPM_LUTC_REP=0: colour table not replicated in the bulk layout (round-1 layout)
#ifndef CSPM_LUTC_REP
#define CSPM_LUTC_REP 1
#endif
constexpr int kMaxScales = 8;
constexpr int kLutW = 1000;   // exp(-i/10) table, pre_ss_pc.cc:61-64
constexpr int kLutC = 32;     // colour term table: 0.1*min(k*0.3333333333,10), saturates at k>=31
constexpr int kLutWUsed = 766;      // reachable weight-table entries: L1 distance of two 8-bit BGR pixels <= 765
constexpr int kLutRepShift = 4;     // bulk layout: tables replicated 16x, one copy per lane mod 16
constexpr int kLutCRepShift = CSPM_LUTC_REP ? kLutRepShift : 0;
// Bulk layout of the dynamic shared memory (doubles): [weight table x16][colour table x16 | x1][mbarrier][strips]
constexpr int kBulkLutCOff = kLutWUsed << kLutRepShift;
constexpr int kBulkMbarOff = kBulkLutCOff + (kLutC << kLutCRepShift);
constexpr int kBulkCoopOff = kBulkMbarOff + 2;
constexpr double kDoubleEps = 0.00000001;  // commfunc.h:26

// Raw cost behind the window sum: GrdCC / CenCC volumes (PreSSPC, PreCSPC) or the image-interpolating
// cost of GrdPC / CSPC (plane_cost/grd_pc.cc, cspc.cc).
enum CcKind { CC_GRD = 0, CC_CEN = 1, CC_IMG = 2 };

// One pyramid level, both views.  Images are stored padded: `pad` columns on
// each side hold the reference's out-of-image partner constant
// (BORDER_THRES = 3 for every colour channel and for the gradient,
// cc/grd_cc.cpp:21-35), so the border variant of the raw cost is the general
// formula applied to a padded read.
struct ScaleDev {
  int w, h, max_disp;
  int pad, pitch;              // elements; row r, column x lives at r*pitch + pad + x
  const uint32_t* color[2];    // B | G<<8 | R<<16 (8-bit originals; byte 3 = 0)
  const double* grad[2];       // Sobel-x of float gray, f64 (cc/grd_cc.cpp:70-77)
  const unsigned long long* cg[2];  // GRD only: colour and gradient of a pixel in ONE 64-bit word (see cg_* below)
  const uint32_t* cen_hi[2];   // CEN only: bits 64..79 of the 80-bit census code; bits 0..63 live in grad[] (as raw u64)
  double max_cost[2];          // max over the whole raw cost volume, pre_ss_pc.cc:51-58
};

struct PcDev {
  int cc_kind;                 // CcKind
  int cross_scale;             // 0: PreSSPC, 1: PreCSPC
  int scale_num;               // 1 for PreSSPC
  int half_wnd;                // wnd_size / 2
  int monotone;                // 1: every scale weight is >= 0, so running sums never decrease (exact early termination
                               // and the certified comparisons of the wavefront sweep rely on it); 0 (reg_lambda < 0
                               // can give negative weights): every cost is summed in full
  int early_term;              // monotone, and not switched off for measurements (CSPM_NO_EARLY_TERM=1 when the plane
                               // cost is created): the bulk kernels stop a window sum once it has reached the bound
  ScaleDev sc[kMaxScales];
  double scale_wgt[kMaxScales];
  const double* lut_w;         // [kLutW]
  const double* lut_c;         // [kLutC]
  const double* lut_rep;       // [766*16 + kLutC]: weight table replicated 16x, then the colour table (bulk kernels)
  double k_grd;                // (1 - ALPHA), evaluated in double on the host
  double k_grd27;              // k_grd * 2^-27 (exact), for gradient differences carried as dG * 2^27 (packed GRD layout)
  uint32_t cg_exp_hi;          // 0x43300000 (high word of 2^52).  A kernel parameter rather than a literal so that
                               // the PRMT that rebuilds the gradient double takes it from the constant bank and
                               // keeps its selector as the immediate (as a literal, ptxas re-materialises the
                               // selector register before almost every PRMT: +3 moves per tap)
};

// ------------------------------------------------------------ strict f64 ops
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }

// ------------------------------------------------------------------ helpers
// commfunc.h:117-121 Round2Int: round-half-even of a double that fits int32.
__device__ __forceinline__ int round2int(double d) {
  d = dadd(d, 6755399441055744.0);
  return __double2loint(d);
}
// commfunc.h:129-145
__device__ __forceinline__ int handle_border(int loc, int size) {
  if (loc < 0) return loc + size;
  if (loc >= size) return loc - size;
  return loc;
}

struct PlaneD {
  double nx, ny, nz;  // unit normal as stored (sign not canonical)
  double a, b, c;     // d(x,y) = a*x + b*y + c
};

// plane.h:25-34 (norm_.dot(point_) accumulates s = 0; s += n[i]*p[i])
__device__ __forceinline__ void update_param(PlaneD& p, double px, double py, double pz) {
  double denom = fmax(fabs(p.nz), kDoubleEps);
  if (p.nz < 0.0) denom = -denom;
  p.a = ddiv(-p.nx, denom);
  p.b = ddiv(-p.ny, denom);
  double s = 0.0;
  s = dadd(s, dmul(p.nx, px));
  s = dadd(s, dmul(p.ny, py));
  s = dadd(s, dmul(p.nz, pz));
  p.c = ddiv(s, denom);
}
// cv::norm(Vec3d, NORM_L2)
__device__ __forceinline__ double norm_l2(double x, double y, double z) {
  double s = 0.0;
  s = dadd(s, dmul(x, x));
  s = dadd(s, dmul(y, y));
  s = dadd(s, dmul(z, z));
  return __dsqrt_rn(s);
}
// (a*x + b*y) + c  — cs_patchmatch.cc:249,318-320
__device__ __forceinline__ double plane_disp(const PlaneD& p, int x, int y) {
  return dadd(dadd(dmul(p.a, (double)x), dmul(p.b, (double)y)), p.c);
}

// ---------------------------------------------------------------------- RNG
// Philox4x32-10 (Salmon et al., SC'11).  A 64-bit draw is addressed by
// (seed, stream, slot, sub); see include/cspm_c.h "RNG contract".
__device__ __forceinline__ uint64_t rng_draw(uint64_t seed, uint64_t stream, uint32_t slot, uint32_t sub) {
  uint32_t c0 = slot, c1 = sub, c2 = (uint32_t)stream, c3 = (uint32_t)(stream >> 32);
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
    c0 = n0; c1 = l1; c2 = n2; c3 = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return (uint64_t)c0 | ((uint64_t)c1 << 32);
}
__device__ __forceinline__ double rng_u01(uint64_t seed, uint64_t stream, uint32_t slot, uint32_t sub) {
  return dmul((double)(rng_draw(seed, stream, slot, sub) >> 11), 1.1102230246251565e-16);
}
// RNG::uniform(a,b) = u*(b-a)+a
__device__ __forceinline__ double rng_uniform(uint64_t seed, uint64_t stream, uint32_t slot, uint32_t sub,
                                              double a, double b) {
  return dadd(dmul(rng_u01(seed, stream, slot, sub), dsub(b, a)), a);
}
// fill(Vec3d, NORMAL, mean, std): uniform direction on the sphere (Marsaglia
// 1972) — only + - * sqrt, hence bit-reproducible against the CPU oracle.
__device__ __forceinline__ void rng_normal3(uint64_t seed, uint64_t stream, uint32_t slot, double mean,
                                            double std, double v[3]) {
  double x = 0.0, y = 0.0, z = 1.0;
  for (uint32_t t = 0; t < 32; ++t) {
    double u1 = dsub(dmul(2.0, rng_u01(seed, stream, slot, 2 * t)), 1.0);
    double u2 = dsub(dmul(2.0, rng_u01(seed, stream, slot, 2 * t + 1)), 1.0);
    double s = dadd(dmul(u1, u1), dmul(u2, u2));
    if (s < 1.0) {
      double r = __dsqrt_rn(dsub(1.0, s));
      x = dmul(dmul(2.0, u1), r);
      y = dmul(dmul(2.0, u2), r);
      z = dsub(1.0, dmul(2.0, s));
      break;
    }
  }
  v[0] = dadd(mean, dmul(std, x));
  v[1] = dadd(mean, dmul(std, y));
  v[2] = dadd(mean, dmul(std, z));
}
__host__ __device__ __forceinline__ uint64_t stream_init(int h, int v, int y) {
  return 1ull + (uint64_t)v * h + y;
}
__host__ __device__ __forceinline__ uint64_t stream_refine(int h, int n_ref, int iter, int step, int v, int y) {
  return 1ull + 2ull * h + (((uint64_t)iter * n_ref + step) * 2ull + v) * h + y;
}

// ----------------------------------------------------------------- raw cost
// GrdCC raw cost of pixel q (this view v) against the partner at disparity d,
// computed on the fly instead of read from a (max_disp+1)-slice volume
// (cc/grd_cc.cpp:4-19 and :85-108 / :131-153).  The colour term only depends
// on the integer L1 distance of two 8-bit pixels, so it is a 32-entry table.
struct RawCtx {
  const uint32_t* col_o;  // other view, padded
  const double* grd_o;    // GRD: x-gradient; CEN: census bits 0..63 (raw u64 stored in the f64 slot)
  const uint32_t* hi_o;   // CEN: census bits 64..79
  const unsigned long long* cg_o;  // GRD: packed colour + gradient of the other view
  const double* lut_c;
  double k_grd;
  double k_grd27;
  uint32_t cg_exp_hi;
  int dir;  // partner x = q_x + dir*d : -1 for the left view, +1 for the right view
  int w;    // CEN: image width (partner outside the image costs CENCUS_BIT)
};
// min(|x|, TAU_GRD) as the reference's `grdDiff > TAU_GRD ? TAU_GRD : grdDiff` (no NaN handling needed:
// inputs are finite), cheaper than fmin(), which must honour NaN semantics.
__device__ __forceinline__ double abs_min2(double x) {
  const double ax = fabs(x);
  return ax > 2.0 ? 2.0 : ax;
}
// Packed GRD pixel.  The float-gray x-gradient G = g[x+1] - g[x-1] of 8-bit images is always an integer
// multiple of 2^-27 with |G| <= 510 (each f32 gray R*0.299f + G*0.587f + B*0.114f is a multiple of 2^-27 no
// larger than 255; k_pack verifies it for every pixel), so n = G * 2^27 + 2^37 is a non-negative 38-bit
// integer.  One 64-bit word per pixel:
//     bits  0..37  n            bits 40..47 B    48..55 G    56..63 R
// One 64-bit load per pixel replaces a 32-bit colour load plus a 64-bit gradient load (the bulk kernels are
// bound by LSU issue: 6 -> 3 loads per tap), and decoding needs no conversion instruction:
//   * gradient: byte 0 of hi under 0x433000.. (one PRMT) over the untouched low word is the bit pattern of the double
//     2^52 + n (n sits in the low mantissa bits).  The difference of two such doubles is the exact integer
//     (G_q - G_p) * 2^27; the clamp min(|dG|, TAU_GRD = 2) becomes min(|.|, 2^28), and the product with
//     k_grd * 2^-27 rounds exactly like k_grd * min(|dG|, 2) (power-of-two scalings, same real product).
//   * colour: hi & 0xFFFFFF00 keeps B,G,R in bytes 1..3 with byte 0 clear on both operands, so the 4-byte
//     SAD instruction returns the 3-channel L1 distance unchanged.
constexpr uint32_t kCgColorMask = 0xFFFFFF00u;
__device__ __forceinline__ uint32_t cg_color_key(uint32_t bgr) { return bgr << 8; }  // unpacked colour -> SAD operand
struct CgPix {
  uint32_t key;  // colour in bytes 1..3 (SAD operand)
  double g;      // 2^52 + G * 2^27 + 2^37
};
__device__ __forceinline__ CgPix cg_load(const unsigned long long* cg, int idx, uint32_t exp_hi) {
  const uint2 w = *reinterpret_cast<const uint2*>(cg + idx);  // (ld.global.nc via __ldg: measured, no gain)
  CgPix p;
  p.key = w.y & kCgColorMask;
  p.g = __hiloint2double((int)__byte_perm(w.y, exp_hi, 0x7650), (int)w.x);  // byte 0 of hi under the exponent of 2^52: one PRMT
  return p;
}
// min(|x|, 2^28).  (An integer-pipe version on the high word was measured slower: 358 vs 349 ms per HD refine.)
__device__ __forceinline__ double abs_min_2p28(double x) {
  const double ax = fabs(x);
  return ax > 268435456.0 ? 268435456.0 : ax;
}
// raw_cost<CC_GRD> on the packed layout (cc/grd_cc.cpp:4-19).  The gradient term k_grd * min(|dG|, TAU_GRD) is
// formed as min(|k * dG|, k * TAU_GRD): multiplication by k > 0 is monotone under round-to-nearest and
// commutes with |.|, so both sides of the min are the correctly rounded products the reference computes in
// either case — and the absolute value rides as an operand modifier on the final add (one FP64 op less).
// LCS: log2 of the replication factor of the colour table (bulk layout: one copy per lane mod 16, like the weight
// table, so that the data-dependent 64-bit lookups of a half-warp never share a bank pair)
template <int LCS = 0>
__device__ __forceinline__ double raw_cost_cg(const unsigned long long* cg_o, const double* lut_c, double k_grd27,
                                              uint32_t exp_hi, int idx, uint32_t kq, double gq) {
  const CgPix p = cg_load(cg_o, idx, exp_hi);
  const int sad = min((int)__vsadu4(kq, p.key), kLutC - 1);
  const double clr = lut_c[sad << LCS];                   // ALPHA * min(sum*0.3333333333, TAU_CLR)
  const double kd = dmul(k_grd27, dsub(gq, p.g));         // (1-ALPHA) * (lG-rG), signed
  const double k_tau = dmul(k_grd27, 268435456.0);        // (1-ALPHA) * TAU_GRD (exact power-of-two scaling)
  const double grd = fabs(kd) > k_tau ? k_tau : kd;
  return dadd(clr, fabs(grd));
}
// CC = CC_GRD: cq = packed colour of q, gq = gradient of q (qhi unused).
// CC = CC_CEN: CenCC raw cost (cc/cen_cc.cc:48-67 / :116-135): Hamming distance of the 80-bit census
// codes, CENCUS_BIT = 80 when the partner lies outside the image; gq carries code bits 0..63, qhi bits 64..79.
template <int CC>
__device__ __forceinline__ double raw_cost(const RawCtx& r, int row_off, int qx, uint32_t cq, double gq, uint32_t qhi, int d) {
  const int px = qx + r.dir * d;
  const int idx = row_off + px;
  if (CC == CC_CEN) {
    if (px < 0 || px >= r.w) return 80.0;
    const unsigned long long lo = (unsigned long long)__double_as_longlong(gq) ^ (unsigned long long)__double_as_longlong(r.grd_o[idx]);
    const uint32_t hi = (qhi ^ r.hi_o[idx]) & 0xFFFFu;
    return (double)(__popcll(lo) + __popc(hi));
  }
  const uint32_t cp = r.col_o[idx];
  const double gp = r.grd_o[idx];
  const int sad = min((int)__vsadu4(cq, cp), kLutC - 1);
  const double clr = r.lut_c[sad];               // ALPHA * min(sum*0.3333333333, TAU_CLR)
  const double grd = abs_min2(dsub(gq, gp));     // min(|lG-rG|, TAU_GRD)
  return dadd(clr, dmul(r.k_grd, grd));
}

// One window tap's addend (pre_ss_pc.cc:93-112).  q is inside the image.
struct TapCtx {
  const uint32_t* col_v;  // this view, padded
  const double* grd_v;
  const uint32_t* hi_v;   // CEN only
  const unsigned long long* cg_v;  // GRD only: packed colour + gradient of this view
  const double* lut_w;
  RawCtx raw;
  double max_cost;
  int max_disp;
};
// the plane-dependent part: wgt * (max_cost | lerp of the two raw costs)
// tap_value_qd takes the plane's disparity at q, q_disp = a*q_x + (b*q_y + c) (pre_ss_pc.cc:88,99), already formed.
// Iq / gqp: colour and gradient of q as returned by load_guide (GRD: SAD key and packed-layout double).
// DIR: the view's partner direction when known at compile time (-1 left view, +1 right view), 0 = read t.raw.dir.
// ALLVALID: the caller has proved that trunc(q_disp) lies in [1, max_disp) for every tap of the window
// (GRD only): the validity test and the max_cost select are compiled out.
template <int CC, int DIR = 0, bool ALLVALID = false, int LCS = 0>
__device__ __forceinline__ double tap_value_qd(const TapCtx& t, int row_off, int qx, uint32_t Iq, double gqp, double wgt,
                                               double q_disp) {
  const int fl = __double2int_rz(q_disp);  // static_cast<int>: truncation toward zero
  if (CC == CC_GRD) {
    // Branch-free: an out-of-range disparity is clamped for ADDRESSING only (the padded rows make any
    // d in [1, max_disp] readable) and the reference's max_cost is selected at the end.  Straight-line code
    // lets the scheduler overlap the loads of the unrolled taps instead of serialising one divergent region
    // per tap (measured 423 -> 388 ms per HD run); the selected values are the same, so results stay bit-identical.
    const bool valid = ALLVALID || (fl > 0 && fl < t.max_disp);
    const int flc = ALLVALID ? fl : (valid ? fl : 1);
    const double fw = dsub((double)(flc + 1), q_disp);
    const int dir = DIR != 0 ? DIR : t.raw.dir;  // compile-time: the second partner load is the first +- 8 bytes
    const int i0 = row_off + qx + dir * flc;
    const double c0 = raw_cost_cg<LCS>(t.raw.cg_o, t.raw.lut_c, t.raw.k_grd27, t.raw.cg_exp_hi, i0, Iq, gqp);
    const double c1 = raw_cost_cg<LCS>(t.raw.cg_o, t.raw.lut_c, t.raw.k_grd27, t.raw.cg_exp_hi, i0 + dir, Iq, gqp);
    const double tmp = dadd(dmul(fw, c0), dmul(dsub(1.0, fw), c1));
    return dmul(wgt, ALLVALID ? tmp : (valid ? tmp : t.max_cost));
  }
  if (fl <= 0 || fl >= t.max_disp) return dmul(wgt, t.max_cost);
  if (CC == CC_IMG) {
    // GrdPC / CSPC tap (grd_pc.cc:147-174, cspc.cc:150-175): interpolate the partner IMAGE linearly at the
    // sub-pixel position q_x -+ d (truncation toward zero, single wrap-around), then TAD colour + gradient.
    const double other_x = dadd((double)qx, t.raw.dir > 0 ? q_disp : -q_disp);  // q_x + (2*view-1)*q_disp
    int floor_x = __double2int_rz(other_x);
    int ceil_x = floor_x + 1;
    const double floor_wgt = dsub((double)ceil_x, other_x);
    floor_x = handle_border(floor_x, t.raw.w);
    ceil_x = handle_border(ceil_x, t.raw.w);
    const uint32_t If = t.raw.col_o[row_off + floor_x], Ic = t.raw.col_o[row_off + ceil_x];
    double clr = 0.0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int q8 = (Iq >> (8 * c)) & 255, c8 = (Ic >> (8 * c)) & 255, f8 = (If >> (8 * c)) & 255;
      const double term = fabs(dadd((double)(q8 - c8), dmul(floor_wgt, (double)(c8 - f8))));
      clr = c == 0 ? term : dadd(clr, term);
    }
    clr = dmul(clr, 0.33333333333333);
    const double Gf = t.raw.grd_o[row_off + floor_x], Gc = t.raw.grd_o[row_off + ceil_x];
    const double grd = fabs(dadd(dsub(t.grd_v[row_off + qx], Gc), dmul(floor_wgt, dsub(Gc, Gf))));
    const double val = dadd(dmul(0.1, clr > 10.0 ? 10.0 : clr), dmul(t.raw.k_grd, grd > 2.0 ? 2.0 : grd));
    return dmul(wgt, val);
  }
  const double fw = dsub((double)(fl + 1), q_disp);
  const double gq = t.grd_v[row_off + qx];
  const uint32_t qhi = CC == CC_CEN ? t.hi_v[row_off + qx] : 0u;
  const double c0 = raw_cost<CC>(t.raw, row_off, qx, Iq, gq, qhi, fl);
  const double c1 = raw_cost<CC>(t.raw, row_off, qx, Iq, gq, qhi, fl + 1);
  const double tmp = dadd(dmul(fw, c0), dmul(dsub(1.0, fw), c1));
  return dmul(wgt, tmp);
}
template <int CC, int DIR = 0, bool ALLVALID = false, int LCS = 0>
__device__ __forceinline__ double tap_value(const TapCtx& t, int row_off, int qx, uint32_t Iq, double gqp, double wgt,
                                            double a, double q_disp_y) {
  return tap_value_qd<CC, DIR, ALLVALID, LCS>(t, row_off, qx, Iq, gqp, wgt, dadd(dmul(a, (double)qx), q_disp_y));
}
// Guide pixel q of this view.  GRD: the SAD key and the gradient double of the packed word (centre colours
// compared against it must go through guide_key); other kinds: the plain packed colour.
template <int CC>
__device__ __forceinline__ void load_guide(const TapCtx& t, int idx, uint32_t* Iq, double* gqp) {
  if (CC == CC_GRD) {
    const CgPix p = cg_load(t.cg_v, idx, t.raw.cg_exp_hi);
    *Iq = p.key;
    *gqp = p.g;
  } else {
    *Iq = t.col_v[idx];
    *gqp = 0.0;
  }
}
template <int CC>
__device__ __forceinline__ uint32_t guide_key(uint32_t bgr) { return CC == CC_GRD ? cg_color_key(bgr) : bgr; }
// LWS: log2 of the replication factor of the weight table.  With LWS = 4 every lane reads its own copy
// (t.lut_w already points at copy lane&15, entries are 16 doubles apart): the data-dependent lookup is
// then free of shared-memory bank conflicts (measured 3.8 conflict wavefronts per warp-tap otherwise).
template <int CC, int LWS = 0, int DIR = 0, bool ALLVALID = false>
__device__ __forceinline__ double tap_term(const TapCtx& t, int row_off, int qx, uint32_t Ip, double a,
                                           double q_disp_y) {
  uint32_t Iq;
  double gqp;
  load_guide<CC>(t, row_off + qx, &Iq, &gqp);
  const double wgt = t.lut_w[__vsadu4(Ip, Iq) << LWS];
  return tap_value<CC, DIR, ALLVALID, (LWS > 0 ? kLutCRepShift : 0)>(t, row_off, qx, Iq, gqp, wgt, a, q_disp_y);
}

__device__ __forceinline__ TapCtx make_tap_ctx(const PcDev& pc, int s, int v, const double* lut_w,
                                               const double* lut_c) {
  const ScaleDev& sc = pc.sc[s];
  TapCtx t;
  t.col_v = sc.color[v];
  t.grd_v = sc.grad[v];
  t.hi_v = sc.cen_hi[v];
  t.cg_v = sc.cg[v];
  t.lut_w = lut_w;
  t.raw.cg_o = sc.cg[1 - v];
  t.raw.k_grd27 = pc.k_grd27;
  t.raw.cg_exp_hi = pc.cg_exp_hi;
  t.raw.col_o = sc.color[1 - v];
  t.raw.grd_o = sc.grad[1 - v];
  t.raw.hi_o = sc.cen_hi[1 - v];
  t.raw.w = sc.w;
  t.raw.lut_c = lut_c;
  t.raw.k_grd = pc.k_grd;
  t.raw.dir = v == 0 ? -1 : 1;
  t.max_cost = sc.max_cost[v];
  t.max_disp = sc.max_disp;
  return t;
}

__device__ unsigned long long g_lane_hist[33];  // CSPM_PROFILE_LANES: window rows executed with k active lanes
// -DCSPM_COUNT_TAPS=1: diagnostic build that counts the window taps the bulk kernels actually execute (per-lane
// early termination skips the rest); read back with cspm_debug_taps_executed.  Never the benchmarked build.
#ifndef CSPM_COUNT_TAPS
#define CSPM_COUNT_TAPS 0
#endif
__device__ unsigned long long g_taps_exec;
#if CSPM_COUNT_TAPS
#define CSPM_TAPCOUNT(nn)                                                                     \
  do {                                                                                        \
    const unsigned int m__ = __activemask();                                                  \
    const unsigned int t__ = __reduce_add_sync(m__, (unsigned int)(nn));                      \
    if ((int)(threadIdx.x & 31) == __ffs(m__) - 1) atomicAdd(&g_taps_exec, (unsigned long long)t__); \
  } while (0)
#else
#define CSPM_TAPCOUNT(nn) do { } while (0)
#endif

// PreSSPC::GetPlaneCost window at one scale, literal row-major accumulation
// order (pre_ss_pc.cc:84-116 / pre_cs_pc.cc:151-181).
//
// Exact early termination: every addend is >= 0 and IEEE addition is monotone,
// so the running sum never decreases.  The caller only asks "is the cost <
// bound?" (strict, cs_patchmatch.cc:182,192,201,209,270,335); once the running
// total — for PreCSPC: cost_prev + sum*w_s, evaluated with the same roundings
// as the final expression — has reached `bound`, the answer is "no" whatever
// the remaining taps add, and the (partial) value returned is still >= bound.
// Pass bound = +inf to get the full sum.  Checked once per window row.
// CSM: 1 / 0 = the plane cost is / is not cross-scale, known at compile time (bulk kernels); -1 = read pc.cross_scale.
// DIR as in tap_value_qd.
// Cooperative tail (single-scale bulk kernels, coop != nullptr).  Early termination is per lane, so a warp often ends up
// with a handful of lanes still evaluating while the others idle (HD refinement: 15 % of all executed window
// rows run with <= 4 of 32 lanes, profiles/r01_lane_hist.txt).  Once at most kCoopLanes lanes of a FULL warp are
// left, the warp changes shape: all 32 lanes compute the <= 35 addends of the survivors' current window rows
// (four survivors at a time, their taps laid end to end and dealt to the lanes), park them in a per-warp
// shared-memory strip, and each surviving lane adds its own row to its running sum in tap order.  The sum is formed by the
// same operations in the same order as in the per-lane loop (bit-identical) and the per-row termination test is
// unchanged.  Measured at HD: refinement 337 -> 316 ms with the switch at <= 8 lanes (6 / 12 / 16: 320 / 316 / 323;
// one survivor at a time instead of groups of four: 328).
constexpr int kCoopLanes = CSPM_COOP_LANES;  // switch to the cooperative tail at <= this many live lanes; 0: never
constexpr int kCoopGroup = 4;                 // survivors handled together
constexpr int kCoopSlot = 36;                 // doubles per survivor (windows up to 35 wide; larger ones skip the tail)
constexpr int kCoopStride = kCoopGroup * kCoopSlot;  // doubles per warp strip
template <int CC, int LWS, int CSM, int DIR, bool ALLVALID = false>
__device__ __forceinline__ double window_rows(const PcDev& pc, const ScaleDev& sc, const TapCtx& t, uint32_t Ip, int x0, int x1,
                                              int y0, int y1, double a, double b, double c, double bound, double cost_prev,
                                              double w_s, double* coop = nullptr, bool active = true) {
  // active = false: this lane has nothing to evaluate (its candidate was already rejected at a finer scale) and is
  // only here to HELP the warp's survivors in the cooperative tail; it returns 0.
  const bool cross = CSM < 0 ? pc.cross_scale != 0 : CSM != 0;
  double cost = 0.0;
  // (cross-scale: the tail is a small loss there — C3 refinement 200 vs 196 ms with the grouped version, HD S=3 1 % with the
  // ungrouped one: the deeper scales run with partial warps, and the extra live state spills; CSPM_COOP_CS=1 re-enables it)
  if (kCoopLanes == 0 || CSM < 0 || (CSM != 0 && !CSPM_COOP_CS) || CC != CC_GRD || coop == nullptr) {
    if (!active) return 0.0;
    for (int qy = y0; qy <= y1; ++qy) {
      const double q_disp_y = dadd(dmul(b, (double)qy), c);
      const int row_off = qy * sc.pitch + sc.pad;
#if CSPM_PROFILE_LANES
      {
        const unsigned int m = __activemask();
        if ((int)(threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(&g_lane_hist[__popc(m)], 1ull);
      }
#endif
      CSPM_UNROLL(CSPM_TAP_UNROLL)  // independent taps in flight per thread (35 = 7 x 5); 24 warps per SM x 5 taps measured best
      for (int qx = x0; qx <= x1; ++qx) cost = dadd(cost, tap_term<CC, LWS, DIR, ALLVALID>(t, row_off, qx, Ip, a, q_disp_y));
      CSPM_TAPCOUNT(x1 - x0 + 1);
      const double total = cross ? dadd(cost_prev, dmul(cost, w_s)) : cost;
      if (total >= bound) break;
    }
    return cost;
  }
  // ---- bulk kernels: all lanes of the warp stay in the row loop together (same y0 / y1: one image row per warp)
  const unsigned int mask = __activemask();
  const int lane = threadIdx.x & 31;
  bool done = !active;
  int qy = y0;
  for (; qy <= y1; ++qy) {
#if CSPM_PROFILE_LANES
    {
      const unsigned int act = __ballot_sync(mask, !done);
      if (lane == __ffs(mask) - 1) atomicAdd(&g_lane_hist[__popc(act)], 1ull);
    }
#endif
    if (!done) {
      const double q_disp_y = dadd(dmul(b, (double)qy), c);
      const int row_off = qy * sc.pitch + sc.pad;
      CSPM_UNROLL(CSPM_TAP_UNROLL)
      for (int qx = x0; qx <= x1; ++qx) cost = dadd(cost, tap_term<CC, LWS, DIR, ALLVALID>(t, row_off, qx, Ip, a, q_disp_y));
      CSPM_TAPCOUNT(x1 - x0 + 1);
      const double total = cross ? dadd(cost_prev, dmul(cost, w_s)) : cost;
      if (total >= bound) done = true;
    }
    const unsigned int alive = __ballot_sync(mask, !done);
    if (alive == 0u) return cost;
    if (mask == 0xffffffffu && __popc(alive) <= (CSM == 1 ? CSPM_COOP_LANES_CS : kCoopLanes) && 2 * pc.half_wnd + 1 <= 35) { ++qy; break; }
  }
  // ---- cooperative tail: rows qy .. y1 of the surviving lanes, kCoopGroup survivors at a time.  The taps of the
  // group's rows are laid end to end (survivor k, tap j <-> index k * wnd + j) and dealt to the 32 lanes pass by
  // pass, so a group of 4 needs 5 passes for its 140 taps; then the owners add their own rows side by side.
  const int wnd = 2 * pc.half_wnd + 1;
  const int my_n = x1 - x0 + 1;
  for (; qy <= y1; ++qy) {
    unsigned int todo = __ballot_sync(mask, !done);
    if (todo == 0u) break;
    const int row_off = qy * sc.pitch + sc.pad;
    const double my_qdy = dadd(dmul(b, (double)qy), c);
    while (todo) {
      int src_lane[kCoopGroup];
      unsigned int grp = 0u;
      int gsz = 0;
#pragma unroll
      for (int g = 0; g < kCoopGroup; ++g) {
        src_lane[g] = todo ? __ffs(todo) - 1 : 0;
        if (todo) { grp |= todo & (0u - todo); todo &= todo - 1u; ++gsz; }
      }
      const int ntap = gsz * wnd;
      for (int base = 0; base < ntap; base += 32) {
        const int idx = base + lane;
        const bool act = idx < ntap;
        const int k = act ? (wnd == 35 ? idx / 35 : idx / wnd) : 0;
        const int j = idx - k * wnd;
        int src = src_lane[0];
#pragma unroll
        for (int g = 1; g < kCoopGroup; ++g) src = k == g ? src_lane[g] : src;
        const double aL = __shfl_sync(mask, a, src), qdyL = __shfl_sync(mask, my_qdy, src);
        const uint32_t IpL = __shfl_sync(mask, Ip, src);
        const int x0L = __shfl_sync(mask, x0, src), nL = __shfl_sync(mask, my_n, src);
        if (act && j < nL) coop[k * kCoopSlot + j] = tap_term<CC, LWS, DIR, ALLVALID>(t, row_off, x0L + j, IpL, aL, qdyL);
      }
      __syncwarp(mask);
      if ((grp >> lane) & 1u) {
        const volatile double* cp = coop + __popc(grp & ((1u << lane) - 1u)) * kCoopSlot;
        for (int jj = 0; jj < my_n; ++jj) cost = dadd(cost, cp[jj]);
        CSPM_TAPCOUNT(my_n);
        const double total = cross ? dadd(cost_prev, dmul(cost, w_s)) : cost;
        if (total >= bound) done = true;
      }
      __syncwarp(mask);
    }
  }
  return cost;
}
// COOP = false: never take the cooperative tail (callers whose lanes do not share one image row).
template <int CC, int LWS = 0, int CSM = -1, bool COOP = true>
__device__ __forceinline__ double window_cost(const PcDev& pc, int s, int v, int rx, int ry, double a, double b,
                                              double c, const double* lut_w, const double* lut_c, double bound,
                                              double cost_prev, double w_s, bool active = true) {
  const ScaleDev& sc = pc.sc[s];
  const TapCtx t = make_tap_ctx(pc, s, v, lut_w, lut_c);
  const int hw = pc.half_wnd;
  const uint32_t Ip = guide_key<CC>(t.col_v[ry * sc.pitch + sc.pad + rx]);
  const int y0 = max(ry - hw, 0), y1 = min(ry + hw, sc.h - 1);
  const int x0 = max(rx - hw, 0), x1 = min(rx + hw, sc.w - 1);
  if (CC == CC_GRD && CSM >= 0) {  // bulk kernels: the view is CTA-uniform, specialise the partner direction
    // The plane's disparity q_disp = a*q_x + (b*q_y + c) is, as computed (rounded products and sums are
    // monotone), monotone in q_x for fixed q_y and in q_y for fixed q_x, so its extremes over the window
    // are at the four corners.  If every corner truncates into [1, max_disp) — for every lane of the warp,
    // so that the warp stays on one code path — no tap of this window can be invalid and the loop without
    // the validity test / max_cost select is used (5 of ~63 instructions per tap).
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int qx = (k & 1) ? x1 : x0, qy = (k & 2) ? y1 : y0;
      const int fl = __double2int_rz(dadd(dmul(a, (double)qx), dadd(dmul(b, (double)qy), c)));
      ok = ok && fl > 0 && fl < t.max_disp;
    }
    ok = ok || !active;  // a helper lane must not spoil the warp's vote
    const bool all_ok = __all_sync(__activemask(), ok);
    // per-warp strip for the cooperative tail: behind the colour table and the mbarrier of the bulk layout
    double* coop = COOP ? const_cast<double*>(lut_w) - (threadIdx.x & 15) + kBulkCoopOff + (size_t)threadIdx.y * kCoopStride : nullptr;
    if (all_ok)
      return v == 0 ? window_rows<CC, LWS, CSM, -1, true>(pc, sc, t, Ip, x0, x1, y0, y1, a, b, c, bound, cost_prev, w_s, coop, active)
                    : window_rows<CC, LWS, CSM, 1, true>(pc, sc, t, Ip, x0, x1, y0, y1, a, b, c, bound, cost_prev, w_s, coop, active);
    return v == 0 ? window_rows<CC, LWS, CSM, -1>(pc, sc, t, Ip, x0, x1, y0, y1, a, b, c, bound, cost_prev, w_s, coop, active)
                  : window_rows<CC, LWS, CSM, 1>(pc, sc, t, Ip, x0, x1, y0, y1, a, b, c, bound, cost_prev, w_s, coop, active);
  }
  return window_rows<CC, LWS, CSM, 0>(pc, sc, t, Ip, x0, x1, y0, y1, a, b, c, bound, cost_prev, w_s, nullptr, active);
}

// IPlaneCost::GetPlaneCost: PreSSPC (pre_ss_pc.cc:74-118) or PreCSPC
// (pre_cs_pc.cc:133-188).  Returns the exact cost if it is < bound, otherwise
// some value >= bound (see window_cost).
template <int CC, int LWS = 0, int CSM = -1, bool COOP = true>
__device__ __forceinline__ double plane_cost(const PcDev& pc, int v, int rx, int ry, const PlaneD& p,
                                             const double* lut_w, const double* lut_c, double bound) {
  // (bound is a by-value copy: callers keep comparing against their own)
  const bool cross = CSM < 0 ? pc.cross_scale != 0 : CSM != 0;
  if (!pc.early_term) bound = __longlong_as_double(0x7ff0000000000000ll);  // sums not monotone (negative scale weight), or switched off
  if (!cross) return window_cost<CC, LWS, CSM, COOP>(pc, 0, v, rx, ry, p.a, p.b, p.c, lut_w, lut_c, bound, 0.0, 1.0);
  double cost = 0.0;
  double cur_disp = plane_disp(p, rx, ry);
  int cx = rx, cy = ry;
  if (CSM == 1 && COOP && CSPM_COOP_CS && CC == CC_GRD) {
    // Bulk kernels: a lane whose candidate is rejected at scale s (cost >= bound) does not leave the loop; it walks
    // the remaining scales as a HELPER, so that the warp is complete when the few survivors reach the coarser
    // scales and the cooperative tail (window_rows) can spread their window rows over all 32 lanes.  Without this
    // the coarse scales ran with a handful of live lanes per warp (18 of 32 thread-instructions per instruction in
    // the cross-scale refinement, profiles/r02_ncu_k_refine_hd_cs3.txt).  The loop ends when every lane is done.
    const unsigned int m = __activemask();
    bool fin = false;
    for (int s = 0; s < pc.scale_num; ++s) {
      PlaneD q = p;
      update_param(q, (double)cx, (double)cy, cur_disp);
      const double sc_cost = window_cost<CC, LWS, CSM, COOP>(pc, s, v, cx, cy, q.a, q.b, q.c, lut_w, lut_c, bound, cost, pc.scale_wgt[s], !fin);
      if (!fin) {
        cost = dadd(cost, dmul(sc_cost, pc.scale_wgt[s]));
        if (cost >= bound) fin = true;
      }
      if (__all_sync(m, fin)) break;
      cy /= 2;
      cx /= 2;
      cur_disp = ddiv(cur_disp, 2.0);
    }
    return cost;
  }
  for (int s = 0; s < pc.scale_num; ++s) {
    PlaneD q = p;
    update_param(q, (double)cx, (double)cy, cur_disp);
    const double sc_cost = window_cost<CC, LWS, CSM, COOP>(pc, s, v, cx, cy, q.a, q.b, q.c, lut_w, lut_c, bound, cost, pc.scale_wgt[s]);
    cost = dadd(cost, dmul(sc_cost, pc.scale_wgt[s]));
    if (cost >= bound) break;
    cy /= 2;
    cx /= 2;
    cur_disp = ddiv(cur_disp, 2.0);
  }
  return cost;
}

// Stage the two lookup tables in shared memory (8.25 KB).
__device__ __forceinline__ void stage_luts(const PcDev& pc, double* s_lut_w, double* s_lut_c) {
  for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < kLutW; i += blockDim.x * blockDim.y) s_lut_w[i] = pc.lut_w[i];
  for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < kLutC; i += blockDim.x * blockDim.y) s_lut_c[i] = pc.lut_c[i];
  __syncthreads();
}

// Bulk (thread-per-pixel) kernels: weight table replicated 16x, one copy per lane-mod-16, so that the
// 64-bit data-dependent lookups of a half-warp always hit 16 different bank pairs.  Only the 766
// reachable entries (L1 distance of two 8-bit BGR pixels <= 765) are staged: 98 KB.
constexpr size_t kBulkLutBytes = sizeof(double) * (size_t)kBulkMbarOff;  // 102 144 B (98 304 un-replicated), multiple of 16
__device__ __forceinline__ const double* bulk_lut_w(const double* s_dyn) { return s_dyn + (threadIdx.x & 15); }
__device__ __forceinline__ const double* bulk_lut_c(const double* s_dyn) {
  return s_dyn + kBulkLutCOff + (kLutCRepShift ? (threadIdx.x & 15) : 0);
}
__device__ __forceinline__ unsigned long long* bulk_mbar(double* s_dyn) { return reinterpret_cast<unsigned long long*>(s_dyn + kBulkMbarOff); }
constexpr int kBulkWarpsMax = CSPM_BULK_BY > CSPM_BULK_BY_CS ? CSPM_BULK_BY : CSPM_BULK_BY_CS;
// + the mbarrier + one cooperative-tail strip per warp (126 KB with 24 warps: still the 132 KB carve-out)
constexpr size_t kBulkSmemBytes = kBulkLutBytes + 16 + (size_t)kBulkWarpsMax * kCoopStride * sizeof(double);
// The replicated table (+ the colour table right behind it) is prepared once per plane cost in global
// memory (PcDev::lut_rep) and brought into shared memory by ONE TMA bulk copy (cp.async.bulk, SASS
// UBLKCP) issued by thread 0 and tracked by an mbarrier; the other threads meanwhile draw their
// random candidate (Philox) and only wait on the barrier right before the first table lookup.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_stage_luts_begin(const PcDev& pc, double* s_lut, unsigned long long* s_mbar) {
  const bool leader = threadIdx.x == 0 && threadIdx.y == 0;
  if (leader) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(s_mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (leader) {
    const uint32_t bytes = (uint32_t)kBulkLutBytes;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(s_mbar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(s_lut)),
                 "l"(pc.lut_rep), "r"(bytes), "r"(smem_u32(s_mbar))
                 : "memory");
  }
}
__device__ __forceinline__ void tma_stage_luts_wait(unsigned long long* s_mbar) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(s_mbar))
        : "memory");
  }
}

// ===========================================================================
// Streaming evaluator ("engine"): GetPlaneCost for up to 32 planes at once with
// the lanes of the PRODUCER warps laid along consecutive window taps (so the
// guide row, the gradient row and the partner row segment of one plane are read
// coalesced whatever the plane is), and ONE CONSUMER warp — the last warp of
// the CTA, i.e. the highest scheduling priority on its sub-partition — whose
// lane L adds the addends of chain L in the reference's row-major order
// (pre_ss_pc.cc:84-116).  Producers fill chunk i+1 (whole window rows) of a
// double-buffered shared-memory array while the consumer drains chunk i, so 32
// serial f64 add chains run side by side, overlapped with addend production.
// The sums are therefore rounded exactly like the reference's scalar loop.
//
//   WND   window size if known at compile time (35 in the reference), 0 = runtime
//   ROWS  window rows per chunk for WND = 35 (buffer holds ROWS*35 addends per chain)
//   NT    threads per CTA (NT/32 - 1 producer warps + 1 consumer warp)
//   PAIRS chains 2p and 2p+1 share centre pixel and view (the two candidates of
//         one pixel in spatial propagation): the adaptive weight is computed once
// ===========================================================================
struct ChainDesc {  // per-scale description of one chain, written by its consumer lane
  double a, b, c;
  int v, cx, cy, on;
  uint32_t Ip;
  int pad_;
};

template <int ROWS>
struct EngineCfg {
  static constexpr int kChunkMax = ROWS * 35;
  static constexpr int kStride = kChunkMax + 1 + (kChunkMax & 1);  // odd number of doubles: conflict-free chain reads
  static constexpr size_t kTermsDoubles = 2 * 32 * (size_t)kStride;
  static constexpr size_t kSmemBytes = sizeof(double) * (kLutW + kLutC + kTermsDoubles) + 32 * sizeof(ChainDesc);
};

struct EngineSmem {
  double* lut_w;
  double* lut_c;
  double* terms;
  ChainDesc* ch;
};
template <int ROWS>
__device__ __forceinline__ EngineSmem engine_smem(double* dyn) {
  EngineSmem e;
  e.lut_w = dyn;
  e.lut_c = e.lut_w + kLutW;
  e.terms = e.lut_c + kLutC;
  e.ch = reinterpret_cast<ChainDesc*>(e.terms + EngineCfg<ROWS>::kTermsDoubles);
  return e;
}

// Every thread of the CTA calls this with CTA-uniform control flow.  Consumer
// lanes pass their chain (ch_on, view, reference pixel, plane, bound); the
// aggregated cost comes back in the consumer lanes (other threads get 0).
//
// Chains come in PAIRS (2p, 2p+1) that share centre pixel and view — the two
// candidates of one pixel in spatial propagation — so the adaptive weight of a
// tap is computed once per pair.  A production item is (pair, tap); items are
// spread evenly over all producer threads, lanes along consecutive taps.
//
// Exact early termination (see window_cost): after each chunk the consumer
// lane compares its running total with `bound`; once reached, the chain is
// switched off (its remaining addends are neither produced nor added) and the
// partial total (>= bound) is returned.
__device__ unsigned long long g_eng_prof[4];
  // diagnostic: {consumer busy, producer warp 0 busy, chunk loop total, -} of CTA gridDim/2
struct PairDesc {  // 64 B, written by the consumer lanes, read (broadcast) by the producers
  double a0, b0, c0, a1, b1, c1;
  int cx, cy;
  uint32_t Ip;
  int flags;  // bit0: view, bit1: chain 2p on, bit2: chain 2p+1 on
};

template <int CC, int WND, int ROWS, int NT>
__device__ __forceinline__ double engine_eval_pairs(const PcDev& pc, const EngineSmem& sm, int nch, bool ch_on, int ch_v,
                                                    int ch_x, int ch_y, const PlaneD& ch_plane, double ch_bound) {
  constexpr int kStride = EngineCfg<ROWS>::kStride;
  constexpr int kProducers = NT - 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool consumer = warp == NT / 32 - 1;
  const int wnd = WND > 0 ? WND : 2 * pc.half_wnd + 1;
  const int chunk_rows = WND > 0 ? ROWS : max(1, EngineCfg<ROWS>::kChunkMax / wnd);
  const int nchunks = (wnd + chunk_rows - 1) / chunk_rows;
  const int npairs = (nch + 1) >> 1;
  PairDesc* pd = reinterpret_cast<PairDesc*>(sm.ch);  // 16 x 64 B <= 32 x sizeof(ChainDesc)
  double cost = 0.0, disp = 0.0, sum = 0.0;
  int cx = ch_x, cy = ch_y;
  bool live = ch_on;
  if (consumer && ch_on && pc.cross_scale) disp = plane_disp(ch_plane, ch_x, ch_y);
  for (int s = 0; s < pc.scale_num; ++s) {
    const ScaleDev& sc = pc.sc[s];
    if (consumer) {
      PlaneD q = ch_plane;
      if (live && pc.cross_scale) update_param(q, (double)cx, (double)cy, disp);  // pre_cs_pc.cc:144
      const int on_other = __shfl_xor_sync(0xffffffffu, (int)live, 1);
      if (lane < 2 * npairs) {
        PairDesc& d = pd[lane >> 1];
        if (lane & 1) {
          d.a1 = q.a; d.b1 = q.b; d.c1 = q.c;
        } else {
          d.a0 = q.a; d.b0 = q.b; d.c0 = q.c;
          d.cx = cx; d.cy = cy;
          d.Ip = guide_key<CC>(sc.color[ch_v][cy * sc.pitch + sc.pad + cx]);
          d.flags = ch_v | (live ? 2 : 0) | (on_other ? 4 : 0);
        }
      }
      sum = 0.0;
    }
    __syncthreads();
    // image pointers of this scale for both views (uniform), selected per item by the pair's view
    const uint32_t* col0 = sc.color[0];
    const uint32_t* col1 = sc.color[1];
    const double* grd0 = sc.grad[0];
    const double* grd1 = sc.grad[1];
    const uint32_t* hi0 = sc.cen_hi[0];
    const uint32_t* hi1 = sc.cen_hi[1];
#if CSPM_SP_PROFILE
    const bool prof = blockIdx.x == gridDim.x / 2 && lane == 0;
    const long long loop_t0 = clock64();
#endif
    for (int ci = 0; ci <= nchunks; ++ci) {
#if CSPM_SP_PROFILE
      const long long it_t0 = clock64();
#endif
      if (!consumer) {
        if (ci < nchunks) {  // produce chunk ci (whole window rows)
          const int dy0 = ci * chunk_rows, len = min(chunk_rows, wnd - dy0) * wnd;
          double* buf = sm.terms + (size_t)(ci & 1) * 32 * kStride;
          const unsigned int magic = 0xFFFFFFFFu / (unsigned int)len + 1u;  // it / len for it < 2^16
          const int nitems = npairs * len;
          for (int it = threadIdx.x; it < nitems; it += kProducers) {
            const int gi = (int)__umulhi((unsigned int)it, magic), kk = it - gi * len;
            const PairDesc& d = pd[gi];
            const int flags = d.flags;
            if (!(flags & 6)) continue;
            const int dy = kk / wnd, dx = kk - dy * wnd;
            const int qy = d.cy + dy0 + dy - pc.half_wnd, qx = d.cx + dx - pc.half_wnd;
            double t0 = 0.0, t1 = 0.0;  // +0.0: additive identity of these non-negative sums
            if (qy >= 0 && qy < sc.h && qx >= 0 && qx < sc.w) {
              const int v = flags & 1;
              TapCtx tc;
              tc.col_v = v ? col1 : col0;
              tc.grd_v = v ? grd1 : grd0;
              tc.hi_v = v ? hi1 : hi0;
              tc.cg_v = sc.cg[v];
              tc.raw.cg_o = sc.cg[1 - v];
              tc.raw.k_grd27 = pc.k_grd27;
            tc.raw.cg_exp_hi = pc.cg_exp_hi;
              tc.lut_w = sm.lut_w;
              tc.raw.col_o = v ? col0 : col1;
              tc.raw.grd_o = v ? grd0 : grd1;
              tc.raw.hi_o = v ? hi0 : hi1;
              tc.raw.w = sc.w;
              tc.raw.lut_c = sm.lut_c;
              tc.raw.k_grd = pc.k_grd;
              tc.raw.dir = v ? 1 : -1;
              tc.max_cost = sc.max_cost[v];
              tc.max_disp = sc.max_disp;
              const int row_off = qy * sc.pitch + sc.pad;
              uint32_t Iq;
              double gqp;
              load_guide<CC>(tc, row_off + qx, &Iq, &gqp);
              const double wgt = sm.lut_w[__vsadu4(d.Ip, Iq)];
              const double qyd = (double)qy;
              if (flags & 2) t0 = tap_value<CC>(tc, row_off, qx, Iq, gqp, wgt, d.a0, dadd(dmul(d.b0, qyd), d.c0));
              if (flags & 4) t1 = tap_value<CC>(tc, row_off, qx, Iq, gqp, wgt, d.a1, dadd(dmul(d.b1, qyd), d.c1));
            }
            buf[(2 * gi) * kStride + kk] = t0;
            buf[(2 * gi + 1) * kStride + kk] = t1;
          }
        }
      } else {
        if (ci > 0 && live) {  // consume chunk ci-1 in tap order
          const int dy0 = (ci - 1) * chunk_rows, rows = min(chunk_rows, wnd - dy0);
          const double* tp = sm.terms + (size_t)((ci - 1) & 1) * 32 * kStride + lane * kStride;
          if (WND > 0) {
            for (int r = 0; r < rows; ++r) {
#pragma unroll
              for (int kk = 0; kk < (WND > 0 ? WND : 1); ++kk) sum = dadd(sum, tp[r * (WND > 0 ? WND : 1) + kk]);
            }
          } else {
            const int len = rows * wnd;
#pragma unroll 8
            for (int kk = 0; kk < len; ++kk) sum = dadd(sum, tp[kk]);
          }
          const double total = pc.cross_scale ? dadd(cost, dmul(sum, pc.scale_wgt[s])) : sum;
          if (total >= ch_bound) live = false;  // rejected for certain: stop this chain
        }
        // publish switched-off chains so producers skip them from the next chunk on
        const int live_other = __shfl_xor_sync(0xffffffffu, (int)live, 1);
        if (ci > 0 && lane < 2 * npairs && !(lane & 1)) {
          const int f = ch_v | (live ? 2 : 0) | (live_other ? 4 : 0);
          if (f != pd[lane >> 1].flags) pd[lane >> 1].flags = f;
        }
      }
#if CSPM_SP_PROFILE
      if (prof && consumer) g_eng_prof[0] += (unsigned long long)(clock64() - it_t0);
      if (prof && warp == 0) g_eng_prof[1] += (unsigned long long)(clock64() - it_t0);
#endif
      __syncthreads();
    }
#if CSPM_SP_PROFILE
    if (prof && warp == 0) g_eng_prof[2] += (unsigned long long)(clock64() - loop_t0);
#endif
    if (consumer && ch_on) {
      if (pc.cross_scale) cost = dadd(cost, dmul(sum, pc.scale_wgt[s]));  // pre_cs_pc.cc:182
      else cost = sum;
      disp = ddiv(disp, 2.0);
    }
    cx /= 2;
    cy /= 2;
  }
  return cost;
}

// ---------------------------------------------------------------------------
// Ring-buffer form of engine_eval_pairs for the reference's 35x35 window: no
// CTA-wide barrier inside an evaluation.  Producers stream through ALL window
// rows of all pairs (items ordered row-major, so rows complete in order) and
// publish per-row completion counts in shared memory; the consumer warp waits
// for row r, adds its 35 addends per chain in tap order, and publishes its
// progress so producers never overwrite a row that is still unread (ring of
// kRing rows).  This removes the per-chunk barriers and the load-quantisation
// loss of the chunked pipeline (1.3 items per thread per chunk -> 2 rounds).
// ---------------------------------------------------------------------------
constexpr int kRing = CSPM_SP_RING;   // window rows in flight
constexpr int kRingChain = 37;       // doubles per (row, chain): 35 addends, odd stride -> conflict-free reads
constexpr int kRingRowDoubles = 32 * kRingChain;
struct ViewCtx {     // everything a producer needs about (scale, view), copied out of the kernel parameters once
  const uint32_t* col_v;  // per scale so that items read it with fixed-offset shared loads instead of
  const double* grd_v;    // dynamically indexed constant-bank loads
  const uint32_t* col_o;
  const double* grd_o;
  const uint32_t* hi_v;   // CEN only
  const uint32_t* hi_o;
  const unsigned long long* cg_v;  // GRD only
  const unsigned long long* cg_o;
  double max_cost;
  int dir, pad_;
};
struct RingCtl {
  ViewCtx view[2];
  int w, h, pitch, pad, max_disp, pad_[3];
  int done[36];      // addends produced per window row (items, not chains)
  int consumed;      // rows fully read by the consumer
  int pad2_[3];
};
constexpr size_t kRingSmemBytes = sizeof(double) * (kLutW + kLutC + (size_t)kRing * kRingRowDoubles) + 16 * sizeof(PairDesc) + sizeof(RingCtl);

struct RingSmem {
  double* lut_w;
  double* lut_c;
  double* ring;
  PairDesc* pd;
  RingCtl* ctl;
};
__device__ __forceinline__ RingSmem ring_smem(double* dyn) {
  RingSmem e;
  e.lut_w = dyn;
  e.lut_c = e.lut_w + kLutW;
  e.ring = e.lut_c + kLutC;
  e.pd = reinterpret_cast<PairDesc*>(e.ring + (size_t)kRing * kRingRowDoubles);
  e.ctl = reinterpret_cast<RingCtl*>(e.pd + 16);
  return e;
}

template <int CC, int NT>
__device__ __forceinline__ double engine_eval_pairs_ring(const PcDev& pc, const RingSmem& sm, int nch, bool ch_on, int ch_v,
                                                         int ch_x, int ch_y, const PlaneD& ch_plane, double ch_bound) {
  constexpr int WND = 35, HW = 17;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool consumer = warp == NT / 32 - 1;
  constexpr int kProducers = NT - 32;
  const int npairs = (nch + 1) >> 1;
  const int row_items = npairs * WND;
  volatile int* v_done = sm.ctl->done;
  volatile int* v_consumed = &sm.ctl->consumed;
  double cost = 0.0, disp = 0.0, sum = 0.0;
  int cx = ch_x, cy = ch_y;
  bool live = ch_on;
  if (consumer && ch_on && pc.cross_scale) disp = plane_disp(ch_plane, ch_x, ch_y);
  for (int s = 0; s < pc.scale_num; ++s) {
    const ScaleDev& sc = pc.sc[s];
    if (consumer) {
      PlaneD q = ch_plane;
      if (live && pc.cross_scale) update_param(q, (double)cx, (double)cy, disp);  // pre_cs_pc.cc:144
      const int on_other = __shfl_xor_sync(0xffffffffu, (int)live, 1);
      if (lane < 2 * npairs) {
        PairDesc& d = sm.pd[lane >> 1];
        if (lane & 1) {
          d.a1 = q.a; d.b1 = q.b; d.c1 = q.c;
        } else {
          d.a0 = q.a; d.b0 = q.b; d.c0 = q.c;
          d.cx = cx; d.cy = cy;
          d.Ip = guide_key<CC>(sc.color[ch_v][cy * sc.pitch + sc.pad + cx]);
          d.flags = ch_v | (live ? 2 : 0) | (on_other ? 4 : 0);
        }
      }
      for (int r = lane; r < 36; r += 32) sm.ctl->done[r] = 0;
      if (lane == 0) sm.ctl->consumed = 0;
      if (lane < 2) {
        ViewCtx& vc = sm.ctl->view[lane];
        vc.col_v = sc.color[lane]; vc.grd_v = sc.grad[lane];
        vc.col_o = sc.color[1 - lane]; vc.grd_o = sc.grad[1 - lane];
        vc.hi_v = sc.cen_hi[lane]; vc.hi_o = sc.cen_hi[1 - lane];
        vc.cg_v = sc.cg[lane]; vc.cg_o = sc.cg[1 - lane];
        vc.max_cost = sc.max_cost[lane];
        vc.dir = lane ? 1 : -1;
      }
      if (lane == 2) {
        sm.ctl->w = sc.w; sm.ctl->h = sc.h; sm.ctl->pitch = sc.pitch; sm.ctl->pad = sc.pad; sm.ctl->max_disp = sc.max_disp;
      }
      sum = 0.0;
    }
    __syncthreads();
    if (!consumer) {
      const int sw = sm.ctl->w, sh = sm.ctl->h, spitch = sm.ctl->pitch, spad = sm.ctl->pad, smaxd = sm.ctl->max_disp;
      const double k_grd = pc.k_grd;
      const unsigned int magic_row = 0xFFFFFFFFu / (unsigned int)row_items + 1u;
      const int total = WND * row_items;
#if CSPM_RING_ILP == 2
      // Two items per thread and iteration (rows r and ~r + 1 of the window), written as straight-line, branch-free
      // code — coordinates clamped for addressing, results selected at the end — so that the loads of the second
      // item are in flight while the first is computed: a producer thread is otherwise one ~190-instruction
      // dependent chain per window row, and the consumer waits for it.
      for (int it = threadIdx.x; it < total; it += 2 * kProducers) {
        int rr[2], gg[2], xx[2], fl[2];
        bool okk[2];
        double tv[2][2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int itk = it + k * kProducers;
          okk[k] = itk < total;
          const int itc = okk[k] ? itk : it;
          rr[k] = (int)__umulhi((unsigned int)itc, magic_row);
          const int rem = itc - rr[k] * row_items;
          gg[k] = rem / WND;
          xx[k] = rem - gg[k] * WND;
        }
        if (rr[1] >= kRing) {  // ring slots still unread?  (rr[1] >= rr[0])
          while (*v_consumed <= rr[1] - kRing) __nanosleep(40);
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const PairDesc& d = sm.pd[gg[k]];
          fl[k] = *reinterpret_cast<const volatile int*>(&d.flags);
          const int qy = d.cy + rr[k] - HW, qx = d.cx + xx[k] - HW;
          const bool inb = qy >= 0 && qy < sh && qx >= 0 && qx < sw;
          const int qyc = min(max(qy, 0), sh - 1), qxc = min(max(qx, 0), sw - 1);
          const ViewCtx& vc = sm.ctl->view[fl[k] & 1];
          TapCtx tc;
          tc.col_v = vc.col_v;
          tc.grd_v = vc.grd_v;
          tc.hi_v = vc.hi_v;
          tc.cg_v = vc.cg_v;
          tc.raw.cg_o = vc.cg_o;
          tc.raw.k_grd27 = pc.k_grd27;
          tc.raw.cg_exp_hi = pc.cg_exp_hi;
          tc.lut_w = sm.lut_w;
          tc.raw.col_o = vc.col_o;
          tc.raw.grd_o = vc.grd_o;
          tc.raw.hi_o = vc.hi_o;
          tc.raw.w = sw;
          tc.raw.lut_c = sm.lut_c;
          tc.raw.k_grd = k_grd;
          tc.raw.dir = vc.dir;
          tc.max_cost = vc.max_cost;
          tc.max_disp = smaxd;
          const int row_off = qyc * spitch + spad;
          uint32_t Iq;
          double gqp;
          load_guide<CC>(tc, row_off + qxc, &Iq, &gqp);
          const double wgt = sm.lut_w[__vsadu4(d.Ip, Iq)];
          const double qyd = (double)qyc;
          const double v0 = tap_value<CC>(tc, row_off, qxc, Iq, gqp, wgt, d.a0, dadd(dmul(d.b0, qyd), d.c0));
          const double v1 = tap_value<CC>(tc, row_off, qxc, Iq, gqp, wgt, d.a1, dadd(dmul(d.b1, qyd), d.c1));
          tv[k][0] = (inb && (fl[k] & 2)) ? v0 : 0.0;  // +0.0: additive identity of these non-negative sums
          tv[k][1] = (inb && (fl[k] & 4)) ? v1 : 0.0;
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          if (okk[k] && (fl[k] & 6)) {
            double* slot = sm.ring + (size_t)(rr[k] % kRing) * kRingRowDoubles + (2 * gg[k]) * kRingChain + xx[k];
            slot[0] = tv[k][0];
            slot[kRingChain] = tv[k][1];
          }
        }
        __threadfence_block();
        atomicAdd(&sm.ctl->done[rr[0]], 1);
        if (okk[1]) atomicAdd(&sm.ctl->done[rr[1]], 1);
      }
#else
      for (int it = threadIdx.x; it < total; it += kProducers) {
        const int r = (int)__umulhi((unsigned int)it, magic_row);
        const int rem = it - r * row_items;
        const int gi = rem / WND, dx = rem - gi * WND;
        if (r >= kRing) {  // ring slot still unread?
          while (*v_consumed <= r - kRing) __nanosleep(40);
        }
        const PairDesc& d = sm.pd[gi];
        const int flags = *reinterpret_cast<const volatile int*>(&d.flags);
        double t0 = 0.0, t1 = 0.0;  // +0.0: additive identity of these non-negative sums
        if (flags & 6) {
          const int qy = d.cy + r - HW, qx = d.cx + dx - HW;
          if (qy >= 0 && qy < sh && qx >= 0 && qx < sw) {
            const ViewCtx& vc = sm.ctl->view[flags & 1];
            TapCtx tc;
            tc.col_v = vc.col_v;
            tc.grd_v = vc.grd_v;
            tc.hi_v = vc.hi_v;
            tc.cg_v = vc.cg_v;
            tc.raw.cg_o = vc.cg_o;
            tc.raw.k_grd27 = pc.k_grd27;
            tc.raw.cg_exp_hi = pc.cg_exp_hi;
            tc.lut_w = sm.lut_w;
            tc.raw.col_o = vc.col_o;
            tc.raw.grd_o = vc.grd_o;
            tc.raw.hi_o = vc.hi_o;
            tc.raw.w = sw;
            tc.raw.lut_c = sm.lut_c;
            tc.raw.k_grd = k_grd;
            tc.raw.dir = vc.dir;
            tc.max_cost = vc.max_cost;
            tc.max_disp = smaxd;
            const int row_off = qy * spitch + spad;
            uint32_t Iq;
            double gqp;
            load_guide<CC>(tc, row_off + qx, &Iq, &gqp);
            const double wgt = sm.lut_w[__vsadu4(d.Ip, Iq)];
            const double qyd = (double)qy;
            if (flags & 2) t0 = tap_value<CC>(tc, row_off, qx, Iq, gqp, wgt, d.a0, dadd(dmul(d.b0, qyd), d.c0));
            if (flags & 4) t1 = tap_value<CC>(tc, row_off, qx, Iq, gqp, wgt, d.a1, dadd(dmul(d.b1, qyd), d.c1));
          }
          double* slot = sm.ring + (size_t)(r % kRing) * kRingRowDoubles + (2 * gi) * kRingChain + dx;
          slot[0] = t0;
          slot[kRingChain] = t1;
        }
        __threadfence_block();  // red.release.cta (ATOMS.ADD instead of the aggregated ATOMS.POPC.INC) was measured slower: 265 vs 221 ms per HD run
        atomicAdd(&sm.ctl->done[r], 1);
      }
#endif
    } else {
      for (int r = 0; r < WND; ++r) {
        while (v_done[r] < row_items) __nanosleep(20);  // (a pure spin was measured slower: 225 vs 221 ms per HD run)
        __threadfence_block();
        const bool was_live = live;
        if (live) {
          const volatile double* tp = sm.ring + (size_t)(r % kRing) * kRingRowDoubles + lane * kRingChain;
#pragma unroll
          for (int k0 = 0; k0 < WND; k0 += 12) {  // 12 loads in flight, then 12 dependent adds
            double t[12];
#pragma unroll
            for (int k = 0; k < 12; ++k)
              if (k0 + k < WND) t[k] = tp[k0 + k];
#pragma unroll
            for (int k = 0; k < 12; ++k)
              if (k0 + k < WND) sum = dadd(sum, t[k]);
          }  // (requesting the next 12 before adding the current 12 was measured slower: 224 vs 221 ms)
          const double total = pc.cross_scale ? dadd(cost, dmul(sum, pc.scale_wgt[s])) : sum;
          if (total >= ch_bound) live = false;  // rejected for certain: stop this chain
        }
        if (__any_sync(0xffffffffu, was_live != live)) {  // publish only when a chain was switched off in this row
          const int live_other = __shfl_xor_sync(0xffffffffu, (int)live, 1);
          if (lane < 2 * npairs && !(lane & 1)) {
            const int f = ch_v | (live ? 2 : 0) | (live_other ? 4 : 0);
            if (f != sm.pd[lane >> 1].flags) *reinterpret_cast<volatile int*>(&sm.pd[lane >> 1].flags) = f;
          }
        }
        __syncwarp();
        if (lane == 0) *v_consumed = r + 1;
      }
    }
    __syncthreads();
    if (consumer && ch_on) {
      if (pc.cross_scale) cost = dadd(cost, dmul(sum, pc.scale_wgt[s]));  // pre_cs_pc.cc:182
      else cost = sum;
      disp = ddiv(disp, 2.0);
    }
    cx /= 2;
    cy /= 2;
  }
  return cost;
}

}  // namespace cspm
