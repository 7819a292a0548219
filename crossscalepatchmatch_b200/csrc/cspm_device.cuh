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

namespace cspm {

constexpr int kMaxScales = 8;
constexpr int kLutW = 1000;   // exp(-i/10) table, pre_ss_pc.cc:61-64
constexpr int kLutC = 32;     // colour term table: 0.1*min(k*0.3333333333,10), saturates at k>=31
constexpr double kDoubleEps = 0.00000001;  // commfunc.h:26

enum CcKind { CC_GRD = 0, CC_CEN = 1 };

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
  const uint4* census[2];      // 80-bit census codes in the low 80 bits (CEN only)
  double max_cost[2];          // max over the whole raw cost volume, pre_ss_pc.cc:51-58
};

struct PcDev {
  int cc_kind;                 // CcKind
  int cross_scale;             // 0: PreSSPC, 1: PreCSPC
  int scale_num;               // 1 for PreSSPC
  int half_wnd;                // wnd_size / 2
  ScaleDev sc[kMaxScales];
  double scale_wgt[kMaxScales];
  const double* lut_w;         // [kLutW]
  const double* lut_c;         // [kLutC]
  double k_grd;                // (1 - ALPHA), evaluated in double on the host
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
  const double* grd_o;
  const double* lut_c;
  double k_grd;
  int dir;  // partner x = q_x + dir*d : -1 for the left view, +1 for the right view
};
__device__ __forceinline__ double raw_grd(const RawCtx& r, int row_off, int qx, uint32_t cq, double gq, int d) {
  const int idx = row_off + qx + r.dir * d;
  const uint32_t cp = r.col_o[idx];
  const double gp = r.grd_o[idx];
  const int sad = min((int)__vsadu4(cq, cp), kLutC - 1);
  const double clr = r.lut_c[sad];                 // ALPHA * min(sum*0.3333333333, TAU_CLR)
  const double grd = fmin(fabs(dsub(gq, gp)), 2.0);  // min(|lG-rG|, TAU_GRD)
  return dadd(clr, dmul(r.k_grd, grd));
}

// One window tap's addend (pre_ss_pc.cc:93-112).  q is inside the image.
struct TapCtx {
  const uint32_t* col_v;  // this view, padded
  const double* grd_v;
  const double* lut_w;
  RawCtx raw;
  double max_cost;
  int max_disp;
};
__device__ __forceinline__ double tap_term(const TapCtx& t, int row_off, int qx, uint32_t Ip, double a,
                                           double q_disp_y) {
  const uint32_t Iq = t.col_v[row_off + qx];
  const double wgt = t.lut_w[__vsadu4(Ip, Iq)];
  const double q_disp = dadd(dmul(a, (double)qx), q_disp_y);
  const int fl = __double2int_rz(q_disp);  // static_cast<int>: truncation toward zero
  if (fl <= 0 || fl >= t.max_disp) return dmul(wgt, t.max_cost);
  const double fw = dsub((double)(fl + 1), q_disp);
  const double gq = t.grd_v[row_off + qx];
  const double c0 = raw_grd(t.raw, row_off, qx, Iq, gq, fl);
  const double c1 = raw_grd(t.raw, row_off, qx, Iq, gq, fl + 1);
  const double tmp = dadd(dmul(fw, c0), dmul(dsub(1.0, fw), c1));
  return dmul(wgt, tmp);
}

__device__ __forceinline__ TapCtx make_tap_ctx(const PcDev& pc, int s, int v, const double* lut_w,
                                               const double* lut_c) {
  const ScaleDev& sc = pc.sc[s];
  TapCtx t;
  t.col_v = sc.color[v];
  t.grd_v = sc.grad[v];
  t.lut_w = lut_w;
  t.raw.col_o = sc.color[1 - v];
  t.raw.grd_o = sc.grad[1 - v];
  t.raw.lut_c = lut_c;
  t.raw.k_grd = pc.k_grd;
  t.raw.dir = v == 0 ? -1 : 1;
  t.max_cost = sc.max_cost[v];
  t.max_disp = sc.max_disp;
  return t;
}

// PreSSPC::GetPlaneCost window at one scale, literal row-major accumulation
// order (pre_ss_pc.cc:84-116 / pre_cs_pc.cc:151-181).
__device__ __forceinline__ double window_cost(const PcDev& pc, int s, int v, int rx, int ry, double a, double b,
                                              double c, const double* lut_w, const double* lut_c) {
  const ScaleDev& sc = pc.sc[s];
  const TapCtx t = make_tap_ctx(pc, s, v, lut_w, lut_c);
  const int hw = pc.half_wnd;
  const uint32_t Ip = t.col_v[ry * sc.pitch + sc.pad + rx];
  const int y0 = max(ry - hw, 0), y1 = min(ry + hw, sc.h - 1);
  const int x0 = max(rx - hw, 0), x1 = min(rx + hw, sc.w - 1);
  double cost = 0.0;
  for (int qy = y0; qy <= y1; ++qy) {
    const double q_disp_y = dadd(dmul(b, (double)qy), c);
    const int row_off = qy * sc.pitch + sc.pad;
    for (int qx = x0; qx <= x1; ++qx) cost = dadd(cost, tap_term(t, row_off, qx, Ip, a, q_disp_y));
  }
  return cost;
}

// IPlaneCost::GetPlaneCost: PreSSPC (pre_ss_pc.cc:74-118) or PreCSPC
// (pre_cs_pc.cc:133-188).
__device__ __forceinline__ double plane_cost(const PcDev& pc, int v, int rx, int ry, const PlaneD& p,
                                             const double* lut_w, const double* lut_c) {
  if (!pc.cross_scale) return window_cost(pc, 0, v, rx, ry, p.a, p.b, p.c, lut_w, lut_c);
  double cost = 0.0;
  double cur_disp = plane_disp(p, rx, ry);
  int cx = rx, cy = ry;
  for (int s = 0; s < pc.scale_num; ++s) {
    PlaneD q = p;
    update_param(q, (double)cx, (double)cy, cur_disp);
    const double sc_cost = window_cost(pc, s, v, cx, cy, q.a, q.b, q.c, lut_w, lut_c);
    cost = dadd(cost, dmul(sc_cost, pc.scale_wgt[s]));
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

}  // namespace cspm
