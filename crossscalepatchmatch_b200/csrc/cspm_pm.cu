// cspm_pm.cu — CSPatchMatch side of libcspm_b200.so: the PatchMatch driver
// loop (cs_patchmatch.cc:51-109) as CUDA kernels.
//
//   k_init            InitRandomPlane     (:115-148)   thread per pixel
//   k_spatial_diag    SpatialPropagation  (:163-216)   anti-diagonal wavefront,
//                                                      one CTA per pixel
//   k_view_*          ViewPropagation     (:229-277)   scatter + ordered resolve
//   k_refine          PlaneRefinement     (:292-345)   thread per pixel, all
//                                                      n_ref halvings fused
//   k_plane2disp      PlaneToDisp         (:590-602)
//   k_pp_*            PostProcessing      (:347-588)
//
// Every stage reproduces the reference's serial semantics exactly (see
// DESIGN.md "Schedules"), so planes, costs and disparities are bit-identical
// to the reference code run with the same RNG streams.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>

#include <math_constants.h>

#include "cspm_internal.h"

namespace cspm {

struct PmDev {
  int w, h, max_dis, n_ref, dis_scale;
  uint64_t seed;
  double* st[2];  // 7 planes of w*h: nx ny nz a b c cost
  int cost_consistent;  // every stored cost is GetPlaneCost(stored plane) under the plane cost of this call
};

__device__ __forceinline__ PlaneD load_plane(const double* st, size_t n, size_t i) {
  PlaneD p;
  p.nx = st[i]; p.ny = st[n + i]; p.nz = st[2 * n + i];
  p.a = st[3 * n + i]; p.b = st[4 * n + i]; p.c = st[5 * n + i];
  return p;
}
__device__ __forceinline__ void store_plane(double* st, size_t n, size_t i, const PlaneD& p, double cost) {
  st[i] = p.nx; st[n + i] = p.ny; st[2 * n + i] = p.nz;
  st[3 * n + i] = p.a; st[4 * n + i] = p.b; st[5 * n + i] = p.c;
  st[6 * n + i] = cost;
}

// ------------------------------------------------------------------- init
#ifndef CSPM_BULK_MINB
#define CSPM_BULK_MINB 1
#endif
constexpr int kBulkBySs = CSPM_BULK_BY;     // bulk kernels: 32 x kBulkBy pixels per CTA (98 KB replicated table per CTA) ...
constexpr int kBulkByCs = CSPM_BULK_BY_CS;  // ... chosen separately for the cross-scale instantiations (more live state)
template <bool CS> constexpr int bulk_by() { return CS ? kBulkByCs : kBulkBySs; }
// PlaneRefinement, cross-scale: 16 warps (128 registers, no spills) measured best once the rejected lanes stay on as
// helpers (HD S=3 refinement 805 ms against 819 at 20 warps, 838 at 24; init / view propagation prefer 20)
#ifndef CSPM_REFINE_BY_CS
#define CSPM_REFINE_BY_CS 16
#endif
template <bool CS> constexpr int refine_by() { return CS ? CSPM_REFINE_BY_CS : kBulkBySs; }
constexpr int kBulkMinB = CSPM_BULK_MINB;  // CTAs per SM the register allocation must allow
template <int CC, bool CS>
__global__ void __launch_bounds__(32 * bulk_by<CS>(), kBulkMinB) k_init(PcDev pc, PmDev pm) {
  extern __shared__ __align__(16) double s_dyn[];
  const double* s_lut_c = bulk_lut_c(s_dyn);  // this lane's copy of the colour table
  unsigned long long* s_mbar = bulk_mbar(s_dyn);
  tma_stage_luts_begin(pc, s_dyn, s_mbar);  // TMA bulk copy in flight while the candidate is drawn
  const double* s_lut_w = bulk_lut_w(s_dyn);
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int v = blockIdx.z;
  if (x >= pm.w || y >= pm.h) return;
  const uint64_t stream = stream_init(pm.h, v, y);
  // rng.uniform(kDoubleEps, max_dis)  (:134-135)
  const double z = rng_uniform(pm.seed, stream, 2u * x, 0, kDoubleEps, (double)pm.max_dis);
  double rn[3];
  rng_normal3(pm.seed, stream, 2u * x + 1, 0.0, 1.0, rn);  // (:138)
  const double denom = fmax(norm_l2(rn[0], rn[1], rn[2]), kDoubleEps);
  const double inv = ddiv(1.0, denom);  // Vec / double multiplies by the reciprocal
  PlaneD p;
  p.nx = dmul(rn[0], inv); p.ny = dmul(rn[1], inv); p.nz = dmul(rn[2], inv);
  update_param(p, (double)x, (double)y, z);
  tma_stage_luts_wait(s_mbar);
  const double cost = plane_cost<CC, kLutRepShift, CS ? 1 : 0>(pc, v, x, y, p, s_lut_w, s_lut_c, CUDART_INF);
  store_plane(pm.st[v], (size_t)pm.w * pm.h, (size_t)y * pm.w + x, p, cost);
}

// ----------------------------------------------------------------- refine
template <int CC, bool CS>
__global__ void __launch_bounds__(32 * refine_by<CS>(), kBulkMinB) k_refine(PcDev pc, PmDev pm, int iter) {
  extern __shared__ __align__(16) double s_dyn[];
  const double* s_lut_c = bulk_lut_c(s_dyn);  // this lane's copy of the colour table
  unsigned long long* s_mbar = bulk_mbar(s_dyn);
  tma_stage_luts_begin(pc, s_dyn, s_mbar);  // TMA bulk copy in flight while the candidate is drawn
  const double* s_lut_w = bulk_lut_w(s_dyn);
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int v = blockIdx.z;
  if (x >= pm.w || y >= pm.h) return;
  const size_t n = (size_t)pm.w * pm.h, i = (size_t)y * pm.w + x;
  PlaneD cur = load_plane(pm.st[v], n, i);
  double cur_cost = pm.st[v][6 * n + i];
  double z_iter = ddiv((double)pm.max_dis, 2.0), n_iter = 1.0;  // kMaxNorm_
  tma_stage_luts_wait(s_mbar);  // tables must have landed before the first lookup
  for (int step = 0; step < pm.n_ref; ++step) {
    const uint64_t stream = stream_refine(pm.h, pm.n_ref, iter, step, v, y);
    const double dz = rng_uniform(pm.seed, stream, 2u * x, 0, -z_iter, z_iter);
    const double z = dadd(plane_disp(cur, x, y), dz);
    PlaneD q;
    q.nx = dadd(cur.nx, rng_uniform(pm.seed, stream, 2u * x + 1, 0, -n_iter, n_iter));
    q.ny = dadd(cur.ny, rng_uniform(pm.seed, stream, 2u * x + 1, 1, -n_iter, n_iter));
    q.nz = dadd(cur.nz, rng_uniform(pm.seed, stream, 2u * x + 1, 2, -n_iter, n_iter));
    const double denom = fmax(norm_l2(q.nx, q.ny, q.nz), kDoubleEps);
    const double inv = ddiv(1.0, denom);
    q.nx = dmul(q.nx, inv); q.ny = dmul(q.ny, inv); q.nz = dmul(q.nz, inv);
    update_param(q, (double)x, (double)y, z);
    const double cost = plane_cost<CC, kLutRepShift, CS ? 1 : 0>(pc, v, x, y, q, s_lut_w, s_lut_c, cur_cost);
    if (cost < cur_cost) {
      cur = q;
      cur_cost = cost;
    }
    z_iter = ddiv(z_iter, 2.0);
    n_iter = ddiv(n_iter, 2.0);
  }
  store_plane(pm.st[v], n, i, cur, cur_cost);
}

// --------------------------------------------------------------- red-black
// CSPM_SCHEDULE_REDBLACK / _REDBLACK8: SpatialPropagation as two checkerboard half-sweeps per iteration (the
// throughput schedule BASELINE.json's north_star names; NOT the reference's raster order — bit-identical to the
// oracle's schedule 1 / 2, oracle.cc SpatialPropagationRedBlack).  One half-sweep = one launch: thread per pixel
// of the active colour, candidates = the planes of (x-1,y), (x+1,y), (x,y-1), (x,y+1) [+ the same at distance 5],
// in that order, each tested against the running minimum with the strict `<`.  All offsets are odd, so the
// neighbours have the other colour and nobody writes them during the launch: no ordering between threads.
// Same bulk evaluator as refinement (exact early termination against the running minimum, cooperative tail).
__device__ __forceinline__ bool plane_bits_eq(const PlaneD& p, const PlaneD& q) {
  auto eq = [](double a, double b) { return __double_as_longlong(a) == __double_as_longlong(b); };
  return eq(p.a, q.a) && eq(p.b, q.b) && eq(p.c, q.c) && eq(p.nx, q.nx) && eq(p.ny, q.ny) && eq(p.nz, q.nz);
}
template <int CC, bool CS>
__global__ void __launch_bounds__(32 * bulk_by<CS>(), kBulkMinB) k_redblack(PcDev pc, PmDev pm, int colour, int noff) {
  extern __shared__ __align__(16) double s_dyn[];
  const double* s_lut_c = bulk_lut_c(s_dyn);  // this lane's copy of the colour table
  unsigned long long* s_mbar = bulk_mbar(s_dyn);
  tma_stage_luts_begin(pc, s_dyn, s_mbar);
  const double* s_lut_w = bulk_lut_w(s_dyn);
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int x = 2 * (blockIdx.x * blockDim.x + threadIdx.x) + ((y + colour) & 1);
  const int v = blockIdx.z;
  if (x >= pm.w || y >= pm.h) return;
  const size_t n = (size_t)pm.w * pm.h, i = (size_t)y * pm.w + x;
  PlaneD cur = load_plane(pm.st[v], n, i);
  double cur_cost = pm.st[v][6 * n + i];
  bool adopted = false;
  tma_stage_luts_wait(s_mbar);
  for (int k = 0; k < noff; ++k) {
    const int o = k < 4 ? 1 : 5;
    const int nx = x + ((k & 2) ? 0 : ((k & 1) ? o : -o)), ny = y + ((k & 2) ? ((k & 1) ? o : -o) : 0);
    if (nx < 0 || nx >= pm.w || ny < 0 || ny >= pm.h) continue;
    const PlaneD cand = load_plane(pm.st[v], n, (size_t)ny * pm.w + nx);
    // exact shortcut: a candidate that is bitwise the plane the pixel currently holds costs exactly cur_cost
    // (if that plane was adopted in this launch, or — while every stored cost is known to be GetPlaneCost of its
    // plane — if it is the pixel's own) and cannot pass the strict `<`
    if ((adopted || pm.cost_consistent) && plane_bits_eq(cand, cur)) continue;
    const double cost = plane_cost<CC, kLutRepShift, CS ? 1 : 0>(pc, v, x, y, cand, s_lut_w, s_lut_c, cur_cost);
    if (cost < cur_cost) {
      cur = cand;
      cur_cost = cost;
      adopted = true;
    }
  }
  if (adopted) store_plane(pm.st[v], n, i, cur, cur_cost);
}

// ---------------------------------------------------------------- spatial
// One anti-diagonal of the serial raster scan.  Pixel (u,t) in scan
// coordinates (u+t == diag) depends only on (u-1,t) and (u,t-1), both final
// on the previous diagonal, so all pixels of a diagonal run concurrently and
// the result equals the sequential Gauss-Seidel sweep.  One CTA per pixel:
// phase 1 computes every window tap's addend in parallel, phase 2 adds them
// in the reference's row-major order (one thread per candidate), so the sum
// is rounded exactly like the scalar loop.
template <int CC>
__global__ void __launch_bounds__(128) k_spatial_diag(PcDev pc, PmDev pm, int diag, int u_lo, int rev) {
  extern __shared__ double s_dyn[];
  double* s_lut_w = s_dyn;
  double* s_lut_c = s_lut_w + kLutW;
  double* s_terms = s_lut_c + kLutC;  // [2][wnd*wnd]
  __shared__ double s_cost[2];
  stage_luts(pc, s_lut_w, s_lut_c);

  const int v = blockIdx.y;
  const int u = u_lo + blockIdx.x, t = diag - u;
  const int x = rev ? pm.w - 1 - u : u, y = rev ? pm.h - 1 - t : t;
  const int inc = rev ? -1 : 1;
  const size_t n = (size_t)pm.w * pm.h, i = (size_t)y * pm.w + x;
  const bool has[2] = {u >= 1, t >= 1};
  PlaneD cand[2];
  if (has[0]) cand[0] = load_plane(pm.st[v], n, i - inc);                      // (x - x_inc, y)
  if (has[1]) cand[1] = load_plane(pm.st[v], n, i - (ptrdiff_t)inc * pm.w);    // (x, y - y_inc)

  const int wnd = 2 * pc.half_wnd + 1, wtaps = wnd * wnd;
  double cost[2] = {0.0, 0.0};
  double cur_disp[2];
  for (int c = 0; c < 2; ++c) cur_disp[c] = has[c] ? plane_disp(cand[c], x, y) : 0.0;
  int cx = x, cy = y;
  for (int s = 0; s < pc.scale_num; ++s) {
    const ScaleDev& sc = pc.sc[s];
    const TapCtx tc = make_tap_ctx(pc, s, v, s_lut_w, s_lut_c);
    PlaneD q[2];
    for (int c = 0; c < 2; ++c) {
      q[c] = cand[c];
      if (pc.cross_scale && has[c]) update_param(q[c], (double)cx, (double)cy, cur_disp[c]);
    }
    const uint32_t Ip = guide_key<CC>(tc.col_v[cy * sc.pitch + sc.pad + cx]);
    for (int k = threadIdx.x; k < 2 * wtaps; k += blockDim.x) {
      const int c = k / wtaps, tap = k - c * wtaps;
      if (!has[c]) continue;
      const int qy = cy + tap / wnd - pc.half_wnd, qx = cx + tap % wnd - pc.half_wnd;
      double term = 0.0;  // +0.0 is the additive identity for these non-negative sums
      if (qy >= 0 && qy < sc.h && qx >= 0 && qx < sc.w) {
        const double q_disp_y = dadd(dmul(q[c].b, (double)qy), q[c].c);
        term = tap_term<CC>(tc, qy * sc.pitch + sc.pad, qx, Ip, q[c].a, q_disp_y);
      }
      s_terms[k] = term;
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && threadIdx.x < 64) {
      const int c = threadIdx.x >> 5;
      if (has[c]) {
        double sum = 0.0;
        const double* tp = s_terms + c * wtaps;
        for (int k = 0; k < wtaps; ++k) sum = dadd(sum, tp[k]);
        s_cost[c] = sum;
      }
    }
    __syncthreads();
    for (int c = 0; c < 2; ++c) {
      if (!has[c]) continue;
      if (pc.cross_scale) cost[c] = dadd(cost[c], dmul(s_cost[c], pc.scale_wgt[s]));
      else cost[c] = s_cost[c];
      cur_disp[c] = ddiv(cur_disp[c], 2.0);
    }
    cx /= 2;
    cy /= 2;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double min_cost = pm.st[v][6 * n + i];
    int best = -1;
    if (has[0] && cost[0] < min_cost) { min_cost = cost[0]; best = 0; }   // x-axis neighbour first (:198-204)
    if (has[1] && cost[1] < min_cost) { min_cost = cost[1]; best = 1; }   // then y-axis neighbour (:206-212)
    if (best >= 0) store_plane(pm.st[v], n, i, cand[best], min_cost);
  }
}

// ------------------------------------------------------------------- view
// ViewPropagation for reference view v: every pixel (x,y) of the OTHER view
// proposes its plane to (cor_x, y).  Sequentially the target keeps the lowest
// cost and, among equal costs, the proposal met first in scan order (strict
// `<`, :270); the original plane wins ties.  Parallel form: evaluate all
// proposals, atomicMin the cost bits per target, then atomicMin the scan
// position among the proposals that hit the minimum, then the winner writes.
// f64 -> u64, order-preserving for every finite value and the infinities (costs can be negative when reg_lambda < 0
// makes a scale weight negative): atomicMin on the keys is the minimum of the costs.
__device__ __forceinline__ unsigned long long cost_key(double c) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(c);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__global__ void k_view_begin(PmDev pm, int v, unsigned long long* tmin, unsigned int* win) {
  const size_t n = (size_t)pm.w * pm.h;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  tmin[i] = cost_key(pm.st[v][6 * n + i]);
  win[i] = 0xFFFFFFFFu;
}
__device__ __forceinline__ void view_candidate(const PmDev& pm, int v, int x, int y, PlaneD* q, int* cor_x) {
  const size_t n = (size_t)pm.w * pm.h;
  const PlaneD o = load_plane(pm.st[1 - v], n, (size_t)y * pm.w + x);
  double disp = plane_disp(o, x, y);
  if (disp < 0.0) disp = 0.0;
  if (disp >= pm.max_dis) disp = dsub((double)pm.max_dis, 1.0);
  const int r = round2int(disp);
  *cor_x = handle_border(v == 0 ? x + r : x - r, pm.w);
  *q = o;
  update_param(*q, (double)*cor_x, (double)y, disp);
}
template <int CC, bool CS>
__global__ void __launch_bounds__(32 * bulk_by<CS>(), kBulkMinB) k_view_eval(PcDev pc, PmDev pm, int v, double* cand_cost, int* cand_tx,
                                                   unsigned long long* tmin, unsigned long long* tap_counter) {
  extern __shared__ __align__(16) double s_dyn[];
  const double* s_lut_c = bulk_lut_c(s_dyn);  // this lane's copy of the colour table
  unsigned long long* s_mbar = bulk_mbar(s_dyn);
  tma_stage_luts_begin(pc, s_dyn, s_mbar);  // TMA bulk copy in flight while the candidate is drawn
  const double* s_lut_w = bulk_lut_w(s_dyn);
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  unsigned long long taps = 0;
  if (x < pm.w && y < pm.h) {
    PlaneD q;
    int cor_x;
    view_candidate(pm, v, x, y, &q, &cor_x);
    // a proposal can only win if it beats the target's cost at the start of this pass
    const double bound = pm.st[v][6 * (size_t)pm.w * pm.h + (size_t)y * pm.w + cor_x];
    tma_stage_luts_wait(s_mbar);
    const double cost = plane_cost<CC, kLutRepShift, CS ? 1 : 0>(pc, v, cor_x, y, q, s_lut_w, s_lut_c, bound);
    const size_t i = (size_t)y * pm.w + x;
    cand_cost[i] = cost;
    cand_tx[i] = cor_x;
    atomicMin(&tmin[(size_t)y * pm.w + cor_x], cost_key(cost));
    int cx = cor_x, cy = y;
    for (int s = 0; s < pc.scale_num; ++s) {
      const ScaleDev& sc = pc.sc[s];
      taps += (unsigned long long)(min(cx + pc.half_wnd, sc.w - 1) - max(cx - pc.half_wnd, 0) + 1) *
              (min(cy + pc.half_wnd, sc.h - 1) - max(cy - pc.half_wnd, 0) + 1);
      cx /= 2; cy /= 2;
    }
  }
  for (int o = 16; o > 0; o >>= 1) taps += __shfl_xor_sync(0xffffffffu, taps, o);
  if (((threadIdx.y * blockDim.x + threadIdx.x) & 31) == 0 && taps) atomicAdd(tap_counter, taps);
}
__global__ void k_view_resolve(PmDev pm, int v, int rev, const double* cand_cost, const int* cand_tx,
                               const unsigned long long* tmin, unsigned int* win) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= pm.w) return;
  const size_t n = (size_t)pm.w * pm.h, i = (size_t)y * pm.w + x;
  const size_t t = (size_t)y * pm.w + cand_tx[i];
  const double cost = cand_cost[i];
  if (cost_key(cost) == tmin[t] && cost < pm.st[v][6 * n + t])
    atomicMin(&win[t], (unsigned int)(rev ? pm.w - 1 - x : x));
}
__global__ void k_view_commit(PmDev pm, int v, int rev, const double* cand_cost, const int* cand_tx,
                              const unsigned int* win) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= pm.w) return;
  const size_t n = (size_t)pm.w * pm.h, i = (size_t)y * pm.w + x;
  const size_t t = (size_t)y * pm.w + cand_tx[i];
  if (win[t] != (unsigned int)(rev ? pm.w - 1 - x : x)) return;
  PlaneD q;
  int cor_x;
  view_candidate(pm, v, x, y, &q, &cor_x);
  store_plane(pm.st[v], n, t, q, cand_cost[i]);
}

#include "cspm_pm_engine.cuh"
#include "cspm_pm_wave.cuh"

// -------------------------------------------------------------- plane2disp
__global__ void k_plane2disp(PmDev pm, uint8_t* dis0, uint8_t* dis1) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, v = blockIdx.z;
  if (x >= pm.w) return;
  const size_t n = (size_t)pm.w * pm.h, i = (size_t)y * pm.w + x;
  const double* st = pm.st[v];
  // param().dot(Vec3d(x, y, 1.0)): s = 0; s += a*x; s += b*y; s += c*1.0
  double d = 0.0;
  d = dadd(d, dmul(st[3 * n + i], (double)x));
  d = dadd(d, dmul(st[4 * n + i], (double)y));
  d = dadd(d, dmul(st[5 * n + i], 1.0));
  const int r = round2int(dmul(d, (double)pm.dis_scale));
  (v == 0 ? dis0 : dis1)[i] = (uint8_t)min(max(r, 0), 255);  // saturate_cast<uchar>
}

// ---------------------------------------------------------- post-processing
// LeftRightCheck (:347-369)
__global__ void k_pp_lrcheck(PmDev pm, const uint8_t* dis0, const uint8_t* dis1, int* valid0, int* valid1) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, v = blockIdx.z;
  if (x >= pm.w) return;
  const uint8_t* cur = v == 0 ? dis0 : dis1;
  const uint8_t* oth = v == 0 ? dis1 : dis0;
  const size_t row = (size_t)y * pm.w;
  int ok = 0;
  const double cur_dis = ddiv(dmul((double)cur[row + x], 1.0), (double)pm.dis_scale);
  const int other_x = x + (2 * v - 1) * round2int(cur_dis);
  if (other_x >= 0 && other_x < pm.w) {
    const double other_dis = ddiv(dmul((double)oth[row + other_x], 1.0), (double)pm.dis_scale);
    if (fabs(dsub(cur_dis, other_dis)) <= 0.5 && cur_dis > 0.0) ok = 1;
  }
  (v == 0 ? valid0 : valid1)[row + x] = ok;
}
// FillInvalid (:370-428): thread per pixel, scans its row for the nearest
// valid pixel on each side (valid flags are not modified by this stage).
__global__ void k_pp_fill(PmDev pm, uint8_t* dis0, uint8_t* dis1, const int* valid0, const int* valid1) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, v = blockIdx.z;
  if (x >= pm.w) return;
  const int* yv = (v == 0 ? valid0 : valid1) + (size_t)y * pm.w;
  if (yv[x]) return;
  uint8_t* out = (v == 0 ? dis0 : dis1) + (size_t)y * pm.w;
  int l = x, r = x;
  while (l >= 0 && !yv[l]) --l;
  while (r < pm.w && !yv[r]) ++r;
  const size_t n = (size_t)pm.w * pm.h;
  const double* st = pm.st[v];
  auto disp_of = [&](int px) {
    const size_t i = (size_t)y * pm.w + px;
    double d = 0.0;
    d = dadd(d, dmul(st[3 * n + i], (double)x));
    d = dadd(d, dmul(st[4 * n + i], (double)y));
    d = dadd(d, dmul(st[5 * n + i], 1.0));
    return d;
  };
  const bool lf = l >= 0, rf = r < pm.w;
  double d;
  if (lf && rf) {
    const double ld = disp_of(l), rd = disp_of(r);
    d = ld <= rd ? ld : rd;
  } else if (lf) d = disp_of(l);
  else if (rf) d = disp_of(r);
  else return;
  out[x] = (uint8_t)min(max(pm.dis_scale * round2int(d), 0), 255);
}
// WeightedMedian (:430-506): thread per invalid pixel, 256-bin f64 histogram
// in shared memory.  Only VALID neighbours contribute and valid pixels are
// never rewritten, so reading `dis` while other threads update invalid pixels
// is order-free; we still read from a snapshot to keep it race-free.
// The weighted median only touches pixels that failed the left-right check (typically ~10 %): their indices are
// compacted first (order is irrelevant, every pixel is independent) so that the 16 histogram slots of a CTA are
// all in use (HD, both views: 37 -> 19 ms).
__global__ void k_pp_compact(const int* __restrict__ valid, int n, int* __restrict__ list, unsigned int* __restrict__ count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool take = i < n && !valid[i];
  const unsigned int m = __ballot_sync(0xffffffffu, take);
  if (!m) return;
  const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
  unsigned int base = 0;
  if (lane == leader) base = atomicAdd(count, (unsigned int)__popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (take) list[base + __popc(m & ((1u << lane) - 1u))] = i;
}
__global__ void __launch_bounds__(16) k_pp_wmedian(PmDev pm, int v, const uint8_t* __restrict__ bgr,
                                                   const uint8_t* __restrict__ dis_in, uint8_t* dis_out,
                                                   const int* __restrict__ valid, const double* __restrict__ lut_w,
                                                   int half_wnd, const int* __restrict__ list,
                                                   const unsigned int* __restrict__ count) {
  __shared__ double hist[16][257];
  const unsigned int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= *count) return;
  const int pix = list[slot];
  const int y = pix / pm.w, x = pix - y * pm.w;
  const size_t row = (size_t)y * pm.w;
  double* hs = hist[threadIdx.x];
  for (int d = 0; d < 256; ++d) hs[d] = 0.0;
  const uint8_t* p = bgr + (row + x) * 3;
  double sum_wgt = 0.0;
  for (int wy = -half_wnd; wy <= half_wnd; ++wy) {
    const int qy = y + wy;
    if (qy < 0 || qy >= pm.h) continue;
    for (int wx = -half_wnd; wx <= half_wnd; ++wx) {
      const int qx = x + wx;
      if (qx < 0 || qx >= pm.w) continue;
      const size_t qi = (size_t)qy * pm.w + qx;
      if (!valid[qi]) continue;
      const uint8_t* q = bgr + qi * 3;
      const int diff = abs((int)p[0] - q[0]) + abs((int)p[1] - q[1]) + abs((int)p[2] - q[2]);
      const double wgt = lut_w[diff];
      hs[dis_in[qi]] = dadd(hs[dis_in[qi]], wgt);
      sum_wgt = dadd(sum_wgt, wgt);
    }
  }
  const double median_wgt = ddiv(sum_wgt, 2.0);
  double acc = 0.0;
  int median = 0;
  for (int d = 0; d < 256; ++d) {
    acc = dadd(acc, hs[d]);
    if (acc >= median_wgt) { median = d; break; }
  }
  if (median_wgt > 0.0) dis_out[row + x] = (uint8_t)median;
}

__global__ void k_fill_f64(double* p, size_t n, double v) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// -------------------------------------------------------------------- host
static PmDev pm_dev(const cspm_pm* pm) {
  PmDev d;
  d.w = pm->w; d.h = pm->h; d.max_dis = pm->max_dis; d.n_ref = pm->n_ref; d.dis_scale = pm->dis_scale;
  d.seed = pm->seed;
  d.st[0] = pm->st[0]; d.st[1] = pm->st[1];
  d.cost_consistent = 0;
  return d;
}

static void check_pair(const cspm_pm* pm, const cspm_pc* pc) {
  CSPM_REQUIRE(pm, "null patch match handle");
  CSPM_REQUIRE(pc, "null plane cost");
  CSPM_REQUIRE(pc->device == pm->device, "plane cost and patch match live on different devices");
  CSPM_REQUIRE(pc->dims[0][0] == pm->w && pc->dims[0][1] == pm->h, "plane cost and patch match image sizes differ");
}

template <typename K>
static void bulk_smem(K kern) {
  CSPM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBulkSmemBytes));
}

// Stored costs stay attributable to ONE plane cost only while every cost-writing stage uses that one.
static void note_cost_writer(cspm_pm* pm, const cspm_pc* pc, bool rewrites_all) {
  if (rewrites_all) pm->cost_owner = pc->uid;
  else if (pm->cost_owner != pc->uid) pm->cost_owner = 0;
}

static void run_init(cspm_pm* pm, const cspm_pc* pc) {
  note_cost_writer(pm, pc, true);
  CSPM_DISPATCH_CC_CS(pc->dev.cc_kind, pc->dev.cross_scale, constexpr int by = bulk_by<CS>(); dim3 b(32, by), g((pm->w + 31) / 32, (pm->h + by - 1) / by, 2);
                      bulk_smem(k_init<CC, CS>); k_init<CC, CS><<<g, b, kBulkSmemBytes, pm->stream>>>(pc->dev, pm_dev(pm)));
  pm->stats.kernel_launches++;
  pm->stats.plane_evals += 2ull * pm->w * pm->h;
  pm->stats.window_taps += 2ull * pc->taps_per_image_eval;
}

static void run_refine(cspm_pm* pm, const cspm_pc* pc, int iter) {
  note_cost_writer(pm, pc, false);
  CSPM_DISPATCH_CC_CS(pc->dev.cc_kind, pc->dev.cross_scale, constexpr int by = refine_by<CS>(); dim3 b(32, by), g((pm->w + 31) / 32, (pm->h + by - 1) / by, 2);
                      bulk_smem(k_refine<CC, CS>); k_refine<CC, CS><<<g, b, kBulkSmemBytes, pm->stream>>>(pc->dev, pm_dev(pm), iter));
  pm->stats.kernel_launches++;
  pm->stats.plane_evals += 2ull * pm->n_ref * pm->w * pm->h;
  pm->stats.window_taps += 2ull * pm->n_ref * pc->taps_per_image_eval;
}

// Which sweep implementation: "wave" (default: parallel sums + certified comparisons + exact cost fix-up,
// cspm_pm_wave.cuh), "ring" (round 1: streaming evaluator with one serial add chain per candidate,
// cspm_pm_engine.cuh) or "diag" (one launch per anti-diagonal).  All three are bit-identical; the environment
// variable CSPM_SPATIAL exists for A/B measurements and for the tests that cross-check them.
enum SpatialMode { SP_AUTO = -1, SP_WAVE = 0, SP_RING = 1, SP_DIAG = 2 };
static SpatialMode spatial_mode() {
  const char* e = getenv("CSPM_SPATIAL");
  if (!e) return SP_AUTO;
  const std::string m(e);
  return m == "diag" ? SP_DIAG : m == "ring" ? SP_RING : m == "wave" ? SP_WAVE : SP_AUTO;
}
// CSPM_SP_CTAS=n caps the persistent grid (tests: several pixels per CTA and multi-pass diagonals on small images)
static int spatial_cta_cap() {
  const char* e = getenv("CSPM_SP_CTAS");
  return e ? std::max(1, atoi(e)) : 0;
}

constexpr int kMaxBatch = 16;  // pairs per batched sweep

// One wavefront sweep (cspm_pm_wave.cuh) over n >= 1 pairs of identical geometry in a single cooperative launch
// on pm[0]'s stream, followed per pair by the compaction and the exact cost fix-up of the pixels that adopted a
// plane.  Returns false if the launch is refused (the caller falls back to per-diagonal launches).
static bool launch_wave(cspm_pm* const* pm, const cspm_pc* const* pc, int npairs, int rev, int sms, int cap) {
  cspm_pm* pm0 = pm[0];
  const PcDev& pcd0 = pc[0]->dev;
  const int wnd = 2 * pcd0.half_wnd + 1;
  const size_t n = (size_t)pm0->w * pm0->h;
  cudaStream_t st = pm0->stream;
  const bool batch = npairs > 1;
  void (*kern)(PcDev, PmDev, const WvPair*, int, int, unsigned int*, uint8_t*, unsigned long long*, double, int) = nullptr;
  if (batch) CSPM_DISPATCH_CC(pcd0.cc_kind, kern = k_spatial_wave<CC, true>);
  else CSPM_DISPATCH_CC(pcd0.cc_kind, kern = k_spatial_wave<CC, false>);
  int occ = 0;
  CSPM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWvSmemBytes));
  CSPM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kWvThreads, kWvSmemBytes));
  if (occ < 1) return false;
  std::vector<WvPair> desc(npairs);
  bool monotone = true;
  for (int p = 0; p < npairs; ++p) {
    desc[p].pc = pc[p]->dev;
    desc[p].pm = pm_dev(pm[p]);
    desc[p].pm.cost_consistent = pm[p]->cost_owner == pc[p]->uid;
    desc[p].changed = pm[p]->wave_changed;
    monotone = monotone && pc[p]->dev.monotone;
    CSPM_CUDA(cudaMemsetAsync(pm[p]->wave_changed, 0, 2 * n, st));
    CSPM_CUDA(cudaMemsetAsync(pm[p]->wave_count, 0, 2 * sizeof(unsigned int), st));
  }
  if (batch) CSPM_CUDA(cudaMemcpyAsync(pm0->wave_pairs, desc.data(), sizeof(WvPair) * npairs, cudaMemcpyHostToDevice, st));
  // a CTA only helps while it has a pixel: never more CTAs than the longest diagonal has pixels
  int grid = std::min(sms, std::min(pm0->w, pm0->h));
  if (cap) grid = std::min(grid, cap);
  CSPM_CUDA(cudaMemsetAsync(pm0->grid_barrier, 0, sizeof(unsigned int) * 32 * 1024, st));
  // |row-major chain - tree sum| <= eps * tree sum for n non-negative addends (header of cspm_pm_wave.cuh)
  double eps = 4.0 * (double)wnd * wnd * pcd0.scale_num * 1.1102230246251565e-16;
  static const bool env_exact = [] { const char* e = getenv("CSPM_WAVE_EXACT"); return e && atoi(e) != 0; }();
  int force_exact = (env_exact || !monotone) ? 1 : 0;
  PcDev pc_arg = desc[0].pc;
  PmDev pm_arg = desc[0].pm;
  const WvPair* pairs_arg = batch ? (const WvPair*)pm0->wave_pairs : nullptr;
  int np_arg = npairs, rev_arg = rev;
  unsigned int* prog = pm0->grid_barrier;
  uint8_t* changed = pm0->wave_changed;
  unsigned long long* wst = pm0->wave_stats;
  void* args[] = {&pc_arg, &pm_arg, &pairs_arg, &np_arg, &rev_arg, &prog, &changed, &wst, &eps, &force_exact};
  const cudaError_t ce = cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(kWvThreads), args, kWvSmemBytes, st);
  if (ce != cudaSuccess) {  // e.g. SMs partly taken by another context
    cudaGetLastError();
    return false;
  }
  pm0->stats.kernel_launches++;
  for (int p = 0; p < npairs; ++p) {
    const PcDev& pcd = pc[p]->dev;
    const PmDev d = desc[p].pm;
    k_wave_compact<<<dim3((unsigned)((n + 255) / 256), 2), 256, 0, st>>>(pm[p]->wave_changed, (int)n, pm[p]->wave_list, pm[p]->wave_count);
    CSPM_DISPATCH_CC_CS(pcd.cc_kind, pcd.cross_scale, constexpr int by = bulk_by<CS>(); bulk_smem(k_cost_fix<CC, CS>);
                        k_cost_fix<CC, CS><<<dim3((unsigned)((n + 32 * by - 1) / (32 * by)), 2), dim3(32, by), kBulkSmemBytes, st>>>(
                            pcd, d, pm[p]->wave_list, pm[p]->wave_count));
    pm[p]->stats.kernel_launches += 2;
  }
  return true;
}

static void run_redblack(cspm_pm* pm, const cspm_pc* pc, int iter) {
  note_cost_writer(pm, pc, false);
  PmDev d = pm_dev(pm);
  d.cost_consistent = pm->cost_owner == pc->uid;
  const int noff = pm->schedule == CSPM_SCHEDULE_REDBLACK8 ? 8 : 4;
  for (int half = 0; half < 2; ++half) {
    const int colour = (iter + half) & 1;
    CSPM_DISPATCH_CC_CS(pc->dev.cc_kind, pc->dev.cross_scale, constexpr int by = bulk_by<CS>(); dim3 b(32, by), g(((pm->w + 1) / 2 + 31) / 32, (pm->h + by - 1) / by, 2);
                        bulk_smem(k_redblack<CC, CS>); k_redblack<CC, CS><<<g, b, kBulkSmemBytes, pm->stream>>>(pc->dev, d, colour, noff));
    pm->stats.kernel_launches++;
  }
  // every pixel is active in exactly one half-sweep; one evaluation per in-image neighbour (duplicates included)
  uint64_t ev = 0;
  for (int k = 0; k < noff; ++k) {
    const int o = k < 4 ? 1 : 5;
    const int ox = (k & 2) ? 0 : o, oy = (k & 2) ? o : 0;
    ev += 2ull * (uint64_t)std::max(0, pm->w - ox) * (uint64_t)std::max(0, pm->h - oy);
  }
  pm->stats.plane_evals += ev;
  pm->stats.window_taps += (uint64_t)((double)pc->taps_per_image_eval * ((double)ev / ((double)pm->w * pm->h)));
}

static void spatial_account(cspm_pm* pm, const cspm_pc* pc) {
  // evaluations: every pixel with an x-neighbour + every pixel with a y-neighbour, both views
  const uint64_t ev = 2ull * ((uint64_t)(pm->w - 1) * pm->h + (uint64_t)pm->w * (pm->h - 1));
  pm->stats.plane_evals += ev;
  // tap count: whole-image totals minus the first scan column / row (mirror-symmetric)
  pm->stats.window_taps += (uint64_t)((double)pc->taps_per_image_eval * ((double)ev / ((double)pm->w * pm->h)));
}

static void run_spatial(cspm_pm* pm, const cspm_pc* pc, int iter) {
  if (pm->schedule != CSPM_SCHEDULE_RASTER) return run_redblack(pm, pc, iter);
  const int rev = (iter % 2 == 0) ? 0 : 1;
  const int wnd = 2 * pc->dev.half_wnd + 1;
  note_cost_writer(pm, pc, false);
  PmDev d = pm_dev(pm);
  d.cost_consistent = pm->cost_owner == pc->uid;
  PcDev pcd = pc->dev;
  SpatialMode mode = spatial_mode();
  const size_t n = (size_t)pm->w * pm->h;
  int sms = 0, coop = 0, occ = 0;
  if (mode != SP_DIAG) {
    CSPM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, pm->device));
    CSPM_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, pm->device));
    if (!coop) mode = SP_DIAG;
  }
  // Measured on B200 (profiles/r02_sweep_modes.txt, spatial stage of a 3-iteration run, wave / ring):
  //   450x375 ss 35.5 / 43.8 ms, 450x375 S=3 73 / 115, 1242x375 S=3 181 / 232   (<= 2.5 pixels per CTA and diagonal)
  //   960x540 ss 86 / 83 (3.6), 1920x1080 ss 368 / 216, S=3 781 / 564 (7.3), 3840x2160 ss 1200 / 736 (14.6)
  // With few pixels per CTA a diagonal is bound by latency, which the wavefront kernel's parallel sums cut; with
  // many it is bound by instructions per tap and by the serial resolution of near-ties, where the ring wins.
  if (mode == SP_AUTO) mode = std::min(pm->w, pm->h) <= 3 * sms ? SP_WAVE : SP_RING;
  const int cap = spatial_cta_cap();
  if (mode == SP_WAVE && !launch_wave(&pm, &pc, 1, rev, sms, cap)) mode = SP_DIAG;
  if (mode == SP_RING) {
    const size_t smem_p = wnd == 35 ? kRingSmemBytes : EngineCfg<kSpRows>::kSmemBytes;
    void (*kern)(PcDev, PmDev, int, unsigned int*) = nullptr;
    CSPM_DISPATCH_CC(pcd.cc_kind, kern = wnd == 35 ? k_spatial_persist<CC, 35> : k_spatial_persist<CC, 0>);
    CSPM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p));
    CSPM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kSpThreads, smem_p));
    if (occ < kSpCtasPerSm) mode = SP_DIAG;
    else {
      CSPM_CUDA(cudaMemsetAsync(pm->grid_barrier, 0, sizeof(unsigned int) * 32 * 1024, pm->stream));
      int rev_arg = rev;
      unsigned int* bar = pm->grid_barrier;
      void* args[] = {&pcd, &d, &rev_arg, &bar};
      const int grid = cap ? std::min(cap, sms * kSpCtasPerSm) : sms * kSpCtasPerSm;
      const cudaError_t ce = cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(kSpThreads), args, smem_p, pm->stream);
      if (ce == cudaSuccess) pm->stats.kernel_launches++;
      else {
        cudaGetLastError();
        mode = SP_DIAG;
      }
    }
  }
  if (mode == SP_DIAG) {
    const size_t smem = sizeof(double) * (kLutW + kLutC + 2 * (size_t)wnd * wnd);
    CSPM_DISPATCH_CC(pcd.cc_kind, CSPM_CUDA(cudaFuncSetAttribute(k_spatial_diag<CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
    for (int diag = 1; diag <= pm->w + pm->h - 2; ++diag) {
      const int u_lo = std::max(0, diag - (pm->h - 1)), u_hi = std::min(diag, pm->w - 1);
      dim3 g(u_hi - u_lo + 1, 2);
      CSPM_DISPATCH_CC(pcd.cc_kind, k_spatial_diag<CC><<<g, 128, smem, pm->stream>>>(pcd, d, diag, u_lo, rev));
      pm->stats.kernel_launches++;
    }
  }
  spatial_account(pm, pc);
}

// SpatialPropagation of n pairs of identical geometry (cspm_pm_patchmatch_batch): one batched wavefront launch
// where the wavefront kernel is the sweep of choice, otherwise pair after pair.
static void run_spatial_batch(cspm_pm* const* pm, const cspm_pc* const* pc, int n, int iter) {
  SpatialMode mode = spatial_mode();
  int sms = 0, coop = 0;
  CSPM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, pm[0]->device));
  CSPM_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, pm[0]->device));
  // batching pays where the sweep is latency-bound; the crossover moves up with the batch (measured)
  if (mode == SP_AUTO) mode = std::min(pm[0]->w, pm[0]->h) <= 4 * sms ? SP_WAVE : SP_RING;
  if (n == 1 || mode != SP_WAVE || !coop || pm[0]->schedule != CSPM_SCHEDULE_RASTER) {
    for (int p = 0; p < n; ++p) run_spatial(pm[p], pc[p], iter);
    return;
  }
  const int rev = (iter % 2 == 0) ? 0 : 1;
  for (int p0 = 0; p0 < n; p0 += kMaxBatch) {
    const int np = std::min(kMaxBatch, n - p0);
    for (int p = p0; p < p0 + np; ++p) note_cost_writer(pm[p], pc[p], false);
    if (launch_wave(pm + p0, pc + p0, np, rev, sms, spatial_cta_cap())) {
      for (int p = p0; p < p0 + np; ++p) spatial_account(pm[p], pc[p]);
    } else {
      for (int p = p0; p < p0 + np; ++p) run_spatial(pm[p], pc[p], iter);
    }
  }
}

static void run_view(cspm_pm* pm, const cspm_pc* pc, int iter) {
  const int rev = (iter % 2 == 0) ? 0 : 1;
  note_cost_writer(pm, pc, false);
  const PmDev d = pm_dev(pm);
  const size_t n = (size_t)pm->w * pm->h;
  for (int v = 0; v < 2; ++v) {
    k_view_begin<<<(unsigned)((n + 255) / 256), 256, 0, pm->stream>>>(d, v, pm->tmin, pm->win);
    CSPM_DISPATCH_CC_CS(pc->dev.cc_kind, pc->dev.cross_scale, constexpr int by = bulk_by<CS>(); dim3 b(32, by), g((pm->w + 31) / 32, (pm->h + by - 1) / by);
                     bulk_smem(k_view_eval<CC, CS>);
                     k_view_eval<CC, CS><<<g, b, kBulkSmemBytes, pm->stream>>>(pc->dev, d, v, pm->cand_cost, pm->cand_tx, pm->tmin, pm->tap_counter));
    dim3 g1((pm->w + 127) / 128, pm->h);
    k_view_resolve<<<g1, 128, 0, pm->stream>>>(d, v, rev, pm->cand_cost, pm->cand_tx, pm->tmin, pm->win);
    k_view_commit<<<g1, 128, 0, pm->stream>>>(d, v, rev, pm->cand_cost, pm->cand_tx, pm->win);
    pm->stats.kernel_launches += 4;
    pm->stats.plane_evals += n;
  }
}

static void run_plane2disp(cspm_pm* pm) {
  dim3 g((pm->w + 127) / 128, pm->h, 2);
  k_plane2disp<<<g, 128, 0, pm->stream>>>(pm_dev(pm), pm->dis[0], pm->dis[1]);
  pm->stats.kernel_launches++;
}

static void run_postproc(cspm_pm* pm, const cspm_pc* pc) {
  const PmDev d = pm_dev(pm);
  dim3 g((pm->w + 127) / 128, pm->h, 2);
  k_pp_lrcheck<<<g, 128, 0, pm->stream>>>(d, pm->dis[0], pm->dis[1], pm->valid[0], pm->valid[1]);
  k_pp_fill<<<g, 128, 0, pm->stream>>>(d, pm->dis[0], pm->dis[1], pm->valid[0], pm->valid[1]);
  pm->stats.kernel_launches += 2;
  const size_t n = (size_t)pm->w * pm->h;
  for (int v = 0; v < 2; ++v) {
    CSPM_CUDA(cudaMemcpyAsync(pm->dis_tmp, pm->dis[v], n, cudaMemcpyDeviceToDevice, pm->stream));
    // invalid pixels -> list (view-propagation scratch is free here: cand_tx holds the list, win[0] its length)
    CSPM_CUDA(cudaMemsetAsync(pm->win, 0, sizeof(unsigned int), pm->stream));
    k_pp_compact<<<(unsigned)((n + 255) / 256), 256, 0, pm->stream>>>(pm->valid[v], (int)n, pm->cand_tx, pm->win);
    // WeightedMedian(valid, 35, WMF_GAMMA) — WMF_GAMMA == WGT_GAMMA == 10, same exp table.  The grid covers the
    // worst case (every pixel invalid); CTAs beyond the list length exit at once.
    k_pp_wmedian<<<(unsigned)((n + 15) / 16), 16, 0, pm->stream>>>(d, v, pm->d_bgr[v], pm->dis_tmp, pm->dis[v], pm->valid[v],
                                                                    pc->dev.lut_w, 35 / 2, pm->cand_tx, pm->win);
    pm->stats.kernel_launches += 2;
  }
}

static void sync_taps(cspm_pm* pm) {
  unsigned long long t = 0;
  CSPM_CUDA(cudaMemcpyAsync(&t, pm->tap_counter, sizeof(t), cudaMemcpyDeviceToHost, pm->stream));
  CSPM_CUDA(cudaStreamSynchronize(pm->stream));
  unsigned long long ws[kWvStats] = {};
  CSPM_CUDA(cudaMemcpyAsync(ws, pm->wave_stats, sizeof(ws), cudaMemcpyDeviceToHost, pm->stream));
  CSPM_CUDA(cudaStreamSynchronize(pm->stream));
  pm->stats.window_taps += t;
  pm->stats.spatial_tasks += ws[0];
  pm->stats.spatial_candidates += ws[1];
  pm->stats.spatial_adopted += ws[2];
  pm->stats.spatial_exact_tasks += ws[3];
  CSPM_CUDA(cudaMemsetAsync(pm->tap_counter, 0, sizeof(t), pm->stream));
  CSPM_CUDA(cudaMemsetAsync(pm->wave_stats, 0, sizeof(ws), pm->stream));
}

}  // namespace cspm

using namespace cspm;

extern "C" {

static cspm_pm_t* pm_create(const uint8_t* l, const uint8_t* r, bool on_device, int w, int h, int stride, int max_dis,
                            int dis_scale, uint64_t seed, int schedule, int device) {
  return guard_ptr<cspm_pm>([&]() -> cspm_pm* {
    CSPM_REQUIRE(l && r, "null image pointer");
    CSPM_REQUIRE(w > 0 && h > 0 && stride >= 3 * w, "bad image geometry");
    CSPM_REQUIRE(max_dis >= 1 && dis_scale >= 1, "max_dis and dis_scale must be >= 1");
    // ViewPropagation wraps x +- round(d) ONCE (HandleBorder, commfunc.h:129-145); with max_dis > width the
    // reference indexes past its row arrays (undefined behaviour), so that domain is rejected here.
    CSPM_REQUIRE(max_dis <= w, "max_dis must not exceed the image width (single wrap-around in ViewPropagation)");
    CSPM_REQUIRE(schedule == CSPM_SCHEDULE_RASTER || schedule == CSPM_SCHEDULE_REDBLACK || schedule == CSPM_SCHEDULE_REDBLACK8,
                 "unknown schedule");
    use_device(device);
    cspm_pm* pm = new cspm_pm();
    try {
      pm->device = device;
      pm->w = w; pm->h = h; pm->max_dis = max_dis; pm->dis_scale = dis_scale; pm->schedule = schedule;
      pm->seed = seed;
      int nr = 0;
      for (double z = max_dis / 2.0; z >= 0.1; z /= 2.0) ++nr;  // kZStopThres_ = 0.1
      pm->n_ref = nr;
      pm->stream = acquire_stream(&pm->own_stream);
      pm->arena.stream = pm->stream;
      const size_t n = (size_t)w * h;
      for (int v = 0; v < 2; ++v) {
        pm->st[v] = pm->arena.alloc<double>(7 * n);
        pm->dis[v] = pm->arena.alloc<uint8_t>(n);
        pm->d_bgr[v] = pm->arena.alloc<uint8_t>(3 * n);
        pm->valid[v] = pm->arena.alloc<int>(n);
        CSPM_CUDA(cudaMemsetAsync(pm->dis[v], 0, n, pm->stream));                  // Mat::zeros (:15)
        CSPM_CUDA(cudaMemsetAsync(pm->st[v], 0, sizeof(double) * 6 * n, pm->stream));  // Plane() = zeros
        CSPM_CUDA(cudaMemcpy2DAsync(pm->d_bgr[v], (size_t)3 * w, v == 0 ? l : r, (size_t)stride, (size_t)3 * w, h,
                                    on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, pm->stream));
        // min_cost_ = kDoubleMax (:25)
        k_fill_f64<<<(unsigned)((n + 255) / 256), 256, 0, pm->stream>>>(pm->st[v] + 6 * n, n, 1.7976931348623157e308);
        pm->stats.kernel_launches++;
      }
      pm->cand_cost = pm->arena.alloc<double>(n);
      pm->cand_tx = pm->arena.alloc<int>(n);
      pm->tmin = pm->arena.alloc<unsigned long long>(n);
      pm->win = pm->arena.alloc<unsigned int>(n);
      pm->dis_tmp = pm->arena.alloc<uint8_t>(n);
      pm->tap_counter = pm->arena.alloc<unsigned long long>(1);
      pm->grid_barrier = pm->arena.alloc<unsigned int>(32 * 1024);  // one 128-B line per CTA
      pm->wave_changed = pm->arena.alloc<uint8_t>(2 * n);
      pm->wave_list = pm->arena.alloc<int>(2 * n);
      pm->wave_count = pm->arena.alloc<unsigned int>(2);
      pm->wave_stats = pm->arena.alloc<unsigned long long>(kWvStats);
      pm->wave_pairs = pm->arena.alloc<WvPair>(kMaxBatch);
      CSPM_CUDA(cudaMemsetAsync(pm->wave_stats, 0, sizeof(unsigned long long) * kWvStats, pm->stream));
      CSPM_CUDA(cudaMemsetAsync(pm->tap_counter, 0, sizeof(unsigned long long), pm->stream));
      CSPM_CUDA(cudaStreamSynchronize(pm->stream));
      return pm;
    } catch (...) {
      pm->arena.release();
      if (pm->stream && pm->own_stream) cudaStreamDestroy(pm->stream);
      delete pm;
      throw;
    }
  });
}
cspm_pm_t* cspm_pm_create(const uint8_t* l, const uint8_t* r, int w, int h, int stride, int max_dis, int dis_scale,
                          uint64_t seed, int schedule, int device) {
  return pm_create(l, r, false, w, h, stride, max_dis, dis_scale, seed, schedule, device);
}
cspm_pm_t* cspm_pm_create_dev(const uint8_t* l, const uint8_t* r, int w, int h, int stride, int max_dis, int dis_scale,
                              uint64_t seed, int schedule, int device) {
  return pm_create(l, r, true, w, h, stride, max_dis, dis_scale, seed, schedule, device);
}

void cspm_pm_destroy(cspm_pm_t* pm) {
  if (!pm) return;
  cudaSetDevice(pm->device);
  pm->arena.release();
  cudaStreamSynchronize(pm->stream);
  if (pm->own_stream) cudaStreamDestroy(pm->stream);
  delete pm;
}

int cspm_pm_n_ref(const cspm_pm_t* pm) { return pm ? pm->n_ref : CSPM_EINVAL; }

int cspm_pm_stage(cspm_pm_t* pm, int stage, int iter, const cspm_pc_t* pc) {
  return guard([&] {
    CSPM_REQUIRE(pm, "null patch match handle");
    CSPM_REQUIRE(stage >= 0 && stage <= 5 && iter >= 0, "bad stage / iteration");
    if (stage != CSPM_STAGE_PLANE2DISP) check_pair(pm, pc);
    use_device(pm->device);
    switch (stage) {
      case CSPM_STAGE_INIT: run_init(pm, pc); break;
      case CSPM_STAGE_SPATIAL: run_spatial(pm, pc, iter); break;
      case CSPM_STAGE_VIEW: run_view(pm, pc, iter); break;
      case CSPM_STAGE_REFINE: run_refine(pm, pc, iter); break;
      case CSPM_STAGE_PLANE2DISP: run_plane2disp(pm); break;
      case CSPM_STAGE_POSTPROC: run_postproc(pm, pc); break;
    }
    CSPM_CUDA(cudaGetLastError());
    sync_taps(pm);
  });
}

int cspm_pm_patchmatch(cspm_pm_t* pm, int iter_num, const cspm_pc_t* pc, int use_pp) {
  return guard([&] {
    check_pair(pm, pc);
    CSPM_REQUIRE(iter_num >= 0, "negative iteration count");
    use_device(pm->device);
    // stage timing: events on the handle's stream
    struct Events {  // destroyed on every exit path
      std::vector<cudaEvent_t> v;
      ~Events() { for (cudaEvent_t e : v) cudaEventDestroy(e); }
    } evs;
    std::vector<cudaEvent_t>& ev = evs.v;
    std::vector<int> ev_stage;
    auto mark = [&](int stage) {
      cudaEvent_t e;
      CSPM_CUDA(cudaEventCreate(&e));
      ev.push_back(e);
      CSPM_CUDA(cudaEventRecord(e, pm->stream));
      ev_stage.push_back(stage);
    };
    mark(CSPM_STAGE_INIT);
    run_init(pm, pc);
    for (int i = 0; i < iter_num; ++i) {
      mark(CSPM_STAGE_SPATIAL);
      run_spatial(pm, pc, i);
      mark(CSPM_STAGE_VIEW);
      run_view(pm, pc, i);
      mark(CSPM_STAGE_REFINE);
      run_refine(pm, pc, i);
    }
    mark(CSPM_STAGE_PLANE2DISP);
    run_plane2disp(pm);
    if (use_pp) {
      mark(CSPM_STAGE_POSTPROC);
      run_postproc(pm, pc);
    }
    mark(-1);
    CSPM_CUDA(cudaGetLastError());
    sync_taps(pm);
    for (int s = 0; s < 6; ++s) pm->stage_ms[s] = 0.f;
    for (size_t k = 0; k + 1 < ev.size(); ++k) {
      float ms = 0.f;
      CSPM_CUDA(cudaEventElapsedTime(&ms, ev[k], ev[k + 1]));
      pm->stage_ms[ev_stage[k]] += ms;
    }
  });
}

int cspm_pm_patchmatch_batch(cspm_pm_t* const* pm, const cspm_pc_t* const* pc, int n, int iter_num, int use_pp) {
  if (n == 1 && pm && pc) return cspm_pm_patchmatch(pm[0], iter_num, pc[0], use_pp);
  return guard([&] {
    CSPM_REQUIRE(pm && pc && n >= 1, "null batch");
    CSPM_REQUIRE(iter_num >= 0, "negative iteration count");
    for (int p = 0; p < n; ++p) {
      check_pair(pm[p], pc[p]);
      CSPM_REQUIRE(pm[p]->device == pm[0]->device, "all pairs of a batch must live on one device");
      CSPM_REQUIRE(pm[p]->w == pm[0]->w && pm[p]->h == pm[0]->h && pm[p]->schedule == pm[0]->schedule,
                   "all pairs of a batch must have the same image size and schedule");
      const cspm_pc *a = pc[p], *b = pc[0];
      CSPM_REQUIRE(a->wnd_size == b->wnd_size && a->scale_num == b->scale_num && a->dev.cc_kind == b->dev.cc_kind &&
                       a->dev.cross_scale == b->dev.cross_scale && a->scale_wgt == b->scale_wgt &&
                       std::memcmp(a->dims, b->dims, sizeof a->dims) == 0,
                   "all plane costs of a batch must be of the same kind and geometry (cost, window, scales, reg_lambda, max_disp)");
      for (int q = 0; q < p; ++q) CSPM_REQUIRE(pm[q] != pm[p], "a pair appears twice in the batch");
    }
    use_device(pm[0]->device);
    // everything runs on the first pair's stream (handles are idle between calls: every entry point synchronises)
    struct Streams {
      cspm_pm_t* const* pm;
      int n;
      std::vector<cudaStream_t> saved;
      ~Streams() { for (int p = 0; p < n; ++p) pm[p]->stream = saved[p]; }
    } streams{pm, n, {}};
    for (int p = 0; p < n; ++p) {
      streams.saved.push_back(pm[p]->stream);
      pm[p]->stream = pm[0]->stream;
    }
    cudaStream_t st = pm[0]->stream;
    struct Events {
      std::vector<cudaEvent_t> v;
      ~Events() { for (cudaEvent_t e : v) cudaEventDestroy(e); }
    } evs;
    std::vector<int> ev_stage;
    auto mark = [&](int stage) {
      cudaEvent_t e;
      CSPM_CUDA(cudaEventCreate(&e));
      evs.v.push_back(e);
      CSPM_CUDA(cudaEventRecord(e, st));
      ev_stage.push_back(stage);
    };
    mark(CSPM_STAGE_INIT);
    for (int p = 0; p < n; ++p) run_init(pm[p], pc[p]);
    for (int i = 0; i < iter_num; ++i) {
      mark(CSPM_STAGE_SPATIAL);
      run_spatial_batch(pm, pc, n, i);
      mark(CSPM_STAGE_VIEW);
      for (int p = 0; p < n; ++p) run_view(pm[p], pc[p], i);
      mark(CSPM_STAGE_REFINE);
      for (int p = 0; p < n; ++p) run_refine(pm[p], pc[p], i);
    }
    mark(CSPM_STAGE_PLANE2DISP);
    for (int p = 0; p < n; ++p) run_plane2disp(pm[p]);
    if (use_pp) {
      mark(CSPM_STAGE_POSTPROC);
      for (int p = 0; p < n; ++p) run_postproc(pm[p], pc[p]);
    }
    mark(-1);
    CSPM_CUDA(cudaGetLastError());
    for (int p = 0; p < n; ++p) sync_taps(pm[p]);
    for (int p = 0; p < n; ++p)
      for (int s = 0; s < 6; ++s) pm[p]->stage_ms[s] = 0.f;
    for (size_t k = 0; k + 1 < evs.v.size(); ++k) {  // stage times of the whole batch, reported on the first pair
      float ms = 0.f;
      CSPM_CUDA(cudaEventElapsedTime(&ms, evs.v[k], evs.v[k + 1]));
      pm[0]->stage_ms[ev_stage[k]] += ms;
    }
  });
}

int cspm_pm_get_dis(const cspm_pm_t* pm, int view, uint8_t* out, int out_stride) {
  return guard([&] {
    CSPM_REQUIRE(pm && out && (view == 0 || view == 1) && out_stride >= pm->w, "bad argument");
    use_device(pm->device);
    CSPM_CUDA(cudaMemcpy2DAsync(out, (size_t)out_stride, pm->dis[view], (size_t)pm->w, (size_t)pm->w, pm->h,
                                cudaMemcpyDeviceToHost, pm->stream));
    CSPM_CUDA(cudaStreamSynchronize(pm->stream));
  });
}
int cspm_pm_set_dis(cspm_pm_t* pm, int view, const uint8_t* in, int in_stride) {
  return guard([&] {
    CSPM_REQUIRE(pm && in && (view == 0 || view == 1) && in_stride >= pm->w, "bad argument");
    use_device(pm->device);
    CSPM_CUDA(cudaMemcpy2DAsync(pm->dis[view], (size_t)pm->w, in, (size_t)in_stride, (size_t)pm->w, pm->h,
                                cudaMemcpyHostToDevice, pm->stream));
    CSPM_CUDA(cudaStreamSynchronize(pm->stream));
  });
}

int cspm_pm_get_planes(const cspm_pm_t* pm, int view, double* normal, double* param, double* min_cost) {
  return guard([&] {
    CSPM_REQUIRE(pm && (view == 0 || view == 1), "bad argument");
    use_device(pm->device);
    const size_t n = (size_t)pm->w * pm->h;
    std::vector<double> h(7 * n);
    CSPM_CUDA(cudaMemcpyAsync(h.data(), pm->st[view], sizeof(double) * 7 * n, cudaMemcpyDeviceToHost, pm->stream));
    CSPM_CUDA(cudaStreamSynchronize(pm->stream));
    for (size_t i = 0; i < n; ++i) {
      for (int k = 0; k < 3; ++k) {
        if (normal) normal[3 * i + k] = h[k * n + i];
        if (param) param[3 * i + k] = h[(3 + k) * n + i];
      }
      if (min_cost) min_cost[i] = h[6 * n + i];
    }
  });
}
int cspm_pm_set_planes(cspm_pm_t* pm, int view, const double* normal, const double* param, const double* min_cost) {
  return guard([&] {
    CSPM_REQUIRE(pm && normal && param && min_cost && (view == 0 || view == 1), "bad argument");
    use_device(pm->device);
    const size_t n = (size_t)pm->w * pm->h;
    std::vector<double> h(7 * n);
    for (size_t i = 0; i < n; ++i) {
      for (int k = 0; k < 3; ++k) {
        h[k * n + i] = normal[3 * i + k];
        h[(3 + k) * n + i] = param[3 * i + k];
      }
      h[6 * n + i] = min_cost[i];
    }
    CSPM_CUDA(cudaMemcpyAsync(pm->st[view], h.data(), sizeof(double) * 7 * n, cudaMemcpyHostToDevice, pm->stream));
    CSPM_CUDA(cudaStreamSynchronize(pm->stream));
    pm->cost_owner = 0;  // caller-supplied costs: no longer known to be GetPlaneCost of the stored planes
  });
}

int cspm_debug_taps_executed(unsigned long long* out, int reset) {
  return guard([&] {
    CSPM_REQUIRE(out, "null argument");
    CSPM_CUDA(cudaDeviceSynchronize());
    CSPM_CUDA(cudaMemcpyFromSymbol(out, g_taps_exec, sizeof(unsigned long long)));
    if (reset) {
      unsigned long long z = 0;
      CSPM_CUDA(cudaMemcpyToSymbol(g_taps_exec, &z, sizeof z));
    }
  });
}

int cspm_debug_wave_samples(double* out256, unsigned int* n) {
  return guard([&] {
    CSPM_REQUIRE(out256 && n, "null argument");
    *n = 0;
#if CSPM_WAVE_PROFILE
    CSPM_CUDA(cudaDeviceSynchronize());
    CSPM_CUDA(cudaMemcpyFromSymbol(out256, g_wv_samples, sizeof(double) * 256));
    CSPM_CUDA(cudaMemcpyFromSymbol(n, g_wv_nsamples, sizeof(unsigned int)));
    unsigned int z = 0;
    CSPM_CUDA(cudaMemcpyToSymbol(g_wv_nsamples, &z, sizeof z));
#endif
  });
}

int cspm_debug_lane_hist(unsigned long long* out33, int reset) {
  return guard([&] {
    CSPM_REQUIRE(out33, "null argument");
    CSPM_CUDA(cudaDeviceSynchronize());
    CSPM_CUDA(cudaMemcpyFromSymbol(out33, g_lane_hist, sizeof(unsigned long long) * 33));
    if (reset) {
      unsigned long long z[33] = {};
      CSPM_CUDA(cudaMemcpyToSymbol(g_lane_hist, z, sizeof z));
    }
  });
}

int cspm_debug_counters(unsigned long long* out8, int reset) {
  return guard([&] {
    CSPM_REQUIRE(out8, "null argument");
    CSPM_CUDA(cudaDeviceSynchronize());
#if CSPM_WAVE_PROFILE
    CSPM_CUDA(cudaMemcpyFromSymbol(out8, g_sp_prof, sizeof(unsigned long long) * 8));
#else
    CSPM_CUDA(cudaMemcpyFromSymbol(out8, g_sp_prof, sizeof(unsigned long long) * 4));
    CSPM_CUDA(cudaMemcpyFromSymbol(out8 + 4, g_eng_prof, sizeof(unsigned long long) * 4));
#endif
    if (reset) {
      unsigned long long z[8] = {};
      CSPM_CUDA(cudaMemcpyToSymbol(g_sp_prof, z, sizeof z));
      CSPM_CUDA(cudaMemcpyToSymbol(g_eng_prof, z, sizeof(unsigned long long) * 4));
    }
  });
}

void* cspm_pm_stream(const cspm_pm_t* pm) { return pm ? (void*)pm->stream : nullptr; }
int cspm_pm_stats(const cspm_pm_t* pm, cspm_stats_t* out) {
  if (!pm || !out) return CSPM_EINVAL;
  *out = pm->stats;
  return CSPM_OK;
}
int cspm_pm_stage_ms(const cspm_pm_t* pm, float* ms) {
  if (!pm || !ms) return CSPM_EINVAL;
  for (int s = 0; s < 6; ++s) ms[s] = pm->stage_ms[s];
  return CSPM_OK;
}

}  // extern "C"
