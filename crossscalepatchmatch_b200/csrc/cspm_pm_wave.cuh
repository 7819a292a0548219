// cspm_pm_wave.cuh — SpatialPropagation (cs_patchmatch.cc:163-216) as a persistent anti-diagonal wavefront
// whose per-diagonal step is a PARALLEL reduction instead of a serial add chain.  Included by cspm_pm.cu inside
// namespace cspm.
//
// What must be exact, and what need not be.  Pixel (x,y) tests the planes of (x-1,y) and (x,y-1) against its own
// cost with a strict `<` (:198-212).  The pixels of the next diagonal need this pixel's PLANE — i.e. the outcome
// of the two comparisons — not its cost; the cost is first read again by the next stage.  So the sweep only has to
// get the comparisons right, and it does so with certified bounds:
//   * every window addend is >= 0 (weights, raw costs and — checked on the host — scale weights are), so for ANY
//     summation order of the n addends the computed sum differs from the real sum s by at most gamma_(n-1) * s
//     (Higham, Accuracy and Stability of Numerical Algorithms, eq. 4.4).  The reference's row-major chain S and the
//     tree sum T formed here (lane partials, warp shuffles, a short serial tail) therefore satisfy
//     |S - T| <= eps * T with eps = 4 n u (u = 2^-53; > 2 gamma_n / (1 - gamma_n), and it absorbs the 2 extra roundings
//     per scale of the cross-scale combination);
//   * a comparison whose two intervals do not overlap has the same outcome for S as for T: taken at once;
//   * if they overlap — in practice real ties: near the image border whole families of planes have all their partners
//     outside the image and cost the same, tap by tap up to the rounding of the interpolation (~1 % of the tasks at
//     1920x1080 D=256) — the task is UNDECIDED and goes, CTA-wide, through
//       step A, a term-wise certificate: RN addition is monotone, so if every addend of the candidate is >= the
//         corresponding addend of the pixel's own plane, every partial sum of its chain is too and `S_cand < m` is
//         false without adding anything (m is GetPlaneCost of the own plane while cost_consistent); likewise
//         y-candidate >= x-candidate term-wise gives S_1 >= S_0;
//       step B, the exact path, for what is still open: all threads write the window's addends to shared memory,
//         lane 0 / 1 of warp 0 add the x- / y-candidate's in the reference's row-major order, and the reference's
//         comparison is applied to those sums.
// A pixel that adopts a plane on the fast path gets the tree sum as a PLACEHOLDER cost and a mark in `changed`;
// right after the sweep k_wave_compact + k_cost_fix recompute the cost of every marked pixel with the bit-exact
// bulk evaluator (thread per pixel, row-major order), before anything reads it.  Planes, costs and disparities
// therefore stay bit-identical to the reference (tests: stage parity after every sweep; the same with
// CSPM_WAVE_EXACT=1, which sends every task down the exact path, and with CSPM_SP_CTAS=n, which shrinks the grid;
// the inequality and the decision logic on the CPU in tests/test_wave_bound_math.py).
//
// Work layout per CTA and diagonal: a set-up phase (one warp per task = pixel x view) loads the neighbours' planes,
// applies the two exact shortcuts and files, per pyramid level, the planes as GetPlaneCost sees them; the
// candidates that survived are numbered, their windows are cut into chunks of 32 taps (chunk q < wnd = window row
// q with lane <-> column, the remaining columns follow as linear chunks), and the chunk list is dealt to the warps in
// equal contiguous runs, so every warp has the same load whatever the shortcuts removed.  Lanes keep partial sums, a
// warp reduces once per candidate it touched, warp 0 forms the tree sums and decides.  No thread waits on another
// inside the evaluation; the per-diagonal critical path is ~candidates * 39 / warps chunk iterations instead of a
// 1225-long dependent add chain.  BATCH = true runs the sweeps of several independent pairs in one launch.
#pragma once

// CTA size and row unrolling, measured on 450x375 / 450x375 S=3 / 1242x375 S=3 (spatial stage, ms;
// profiles/r02_variants_wave.txt): 768 threads x2: 36.9 / 70.3 / 171.1; 768 x4: 38.5 / 71.5 / 173.9 (80 registers, spills);
// 640 x5: 36.1 / 71.8 / 175.9; 640 x3: 34.8 / 69.3 / 170.5.
#ifndef CSPM_WAVE_THREADS
#define CSPM_WAVE_THREADS 640
#endif
#ifndef CSPM_WAVE_UNROLL
#define CSPM_WAVE_UNROLL 3
#endif

// -DCSPM_WAVE_PROFILE=1: cycle accounting of one CTA (grid / 2), read back with cspm_debug_counters:
// {neighbour wait, barrier, evaluation (warp 0), wait for the slowest warp, decision, barrier, publish, diagonals}
#ifndef CSPM_WAVE_PROFILE
#define CSPM_WAVE_PROFILE 0
#endif
#if CSPM_WAVE_PROFILE
#define WV_PROF(slot)                                                   \
  do {                                                                  \
    if (blockIdx.x == gridDim.x / 2 && threadIdx.x == 0) {              \
      const long long now__ = clock64();                                \
      g_sp_prof[slot] += (unsigned long long)(now__ - prof_t);          \
      prof_t = now__;                                                   \
    }                                                                   \
  } while (0)
#else
#define WV_PROF(slot) do { } while (0)
#endif
constexpr int kWvThreads = CSPM_WAVE_THREADS;
constexpr int kWvWarps = kWvThreads / 32;
constexpr int kWvTasks = 32;  // tasks per pass (a CTA owns <= 8 pixels x 2 views per diagonal at HD)
// candidate evaluations one warp's contiguous run of chunks can touch: a pass has E <= 2 * kWvTasks evaluations of cpt
// chunks each, a run is ceil(E * cpt / kWvWarps) chunks long and may start anywhere inside an evaluation
constexpr int kWvSlots = (2 * kWvTasks + kWvWarps - 1) / kWvWarps + 1;
constexpr int kWvMaxTaps = 63 * 63;  // wnd_size <= 63 (cspm_pc_create)

struct WvTask {
  double cand[3][6];       // nx ny nz a b c of the x-axis neighbour, the y-axis neighbour, the pixel itself
  double m;                // the pixel's cost at the start of the step (:201,209 compare against it)
  unsigned long long idx;  // y * w + x
  int flags;               // bit0 view, bit1 x-candidate to evaluate, bit2 y-candidate to evaluate
  int x, y, pair;          // pair: which stereo pair of a batched sweep (0 otherwise)
};
// One stereo pair of a batched sweep (cspm_pm_patchmatch_batch): the sweeps of several independent pairs of the
// same geometry run in ONE cooperative launch, diagonal by diagonal in lockstep, so that the per-diagonal latency
// (neighbour wait, barriers, decision, publish: more than half of a small image's sweep) is paid once per diagonal
// instead of once per diagonal and pair.
struct WvPair {
  PcDev pc;
  PmDev pm;
  uint8_t* changed;
};
struct WvScale {  // one task at one pyramid level, as GetPlaneCost sees it (pre_cs_pc.cc:139-147,183-185)
  double abc[3][3];               // disparity plane (a, b, c) of x-candidate, y-candidate, own plane at this scale
  int cx, cy;                     // centre pixel (x >> s, y >> s)
  uint32_t Ip;                    // its colour as SAD key
  int pad_;
};
struct WvCtl {
  WvTask task[kWvTasks];
  WvScale sd[kWvTasks][kMaxScales];
  double partial[kWvWarps][kWvSlots];  // per warp: tree sum of its share of each candidate evaluation it touched
  double amb_T[kWvTasks][2];           // tree sums of the undecided tasks
  double xbuf[2][kWvMaxTaps + 1];  // exact path: every addend of one window, per candidate, row-major
  int amb[kWvTasks];
  int n_amb, pad_[3];
};
constexpr size_t kWvSmemBytes = kBulkLutBytes + 16 + sizeof(WvCtl);

// counters of one sweep: {tasks, candidates evaluated, planes adopted on the fast path, tasks sent down the exact path}
constexpr int kWvStats = 4;
#if CSPM_WAVE_PROFILE
__device__ double g_wv_samples[64][4];  // first undecided tasks of the profiled CTA: T0, T1, m, flags
__device__ unsigned int g_wv_nsamples;
#endif

// Plane of candidate `p` (six numbers nx ny nz a b c) at scale s, centre (cx, cy).
__device__ __forceinline__ void wave_cand_at_scale(const PcDev& pc, const double p[6], int x, int y, int s, int cx, int cy,
                                                   double* a, double* b, double* c) {
  if (!pc.cross_scale) {
    *a = p[3]; *b = p[4]; *c = p[5];
    return;
  }
  PlaneD pl;
  pl.nx = p[0]; pl.ny = p[1]; pl.nz = p[2]; pl.a = p[3]; pl.b = p[4]; pl.c = p[5];
  double disp = plane_disp(pl, x, y);
  for (int k = 0; k < s; ++k) disp = ddiv(disp, 2.0);
  update_param(pl, (double)cx, (double)cy, disp);
  *a = pl.a; *b = pl.b; *c = pl.c;
}

// addend of window tap k (row-major) of one (task, scale) for the plane (a, b, c); +0.0 outside the image
template <int CC, int DIR>
__device__ __forceinline__ double wave_tap(const TapCtx& tc, const ScaleDev& sc, int cx, int cy, uint32_t Ip, double a, double b,
                                           double c, int hw, int wnd, int wtaps, unsigned int magic, int k) {
  const int dy = (int)__umulhi((unsigned int)k, magic), dx = k - dy * wnd;
  const int qx = cx + dx - hw, qy = cy + dy - hw;
  if (k < wtaps && qx >= 0 && qx < sc.w && qy >= 0 && qy < sc.h)
    return tap_term<CC, kLutRepShift, DIR>(tc, qy * sc.pitch + sc.pad, qx, Ip, a, dadd(dmul(b, (double)qy), c));
  return 0.0;
}

// Chunks [ch, ch + nch) of one (candidate, scale), accumulated into this lane's partial sum.  For windows at least
// a warp wide (the reference's 35), chunk q < wnd is window ROW q with lane <-> column (columns 0..31): the column
// test, a * q_x and the pointers are per-lane constants of the whole segment and the row part of the disparity is
// warp-uniform, which saves a third of the instructions of the linear mapping; the remaining wnd - 32 columns of all
// rows follow as (wnd - 32) * wnd / 32 linear chunks.  The number of chunks, ceil(wnd^2 / 32), is the same either
// way.  Any order is fine here: these are the tree sums (the exact paths index taps linearly, wave_tap).
template <int CC, int DIR>
__device__ __forceinline__ double wave_eval_segment(const TapCtx& tc, const ScaleDev& sc, int cx, int cy, uint32_t Ip, double a,
                                                    double b, double c, int hw, int wnd, int wtaps, unsigned int magic, int ch,
                                                    int nch, double ws, bool cross, double acc) {
  const int lane = threadIdx.x & 31;
  if (wnd < 32) {
    CSPM_UNROLL(CSPM_WAVE_UNROLL)
    for (int q = ch; q < ch + nch; ++q) {
      const double tv = wave_tap<CC, DIR>(tc, sc, cx, cy, Ip, a, b, c, hw, wnd, wtaps, magic, q * 32 + lane);
      acc = dadd(acc, cross ? dmul(tv, ws) : tv);
    }
    return acc;
  }
  const int sw = sc.w, sh = sc.h, pitch = sc.pitch, pad = sc.pad;
  const int qx = cx - hw + lane;
  const bool in_x = qx >= 0 && qx < sw;
  const int qxc = min(max(qx, 0), sw - 1);
  const double ax = dmul(a, (double)qxc);
  const int rows_end = min(ch + nch, wnd);
  // branch-free rows (coordinates clamped for addressing, the addend selected at the end): the loads of the
  // unrolled rows overlap instead of each row being its own divergent region
  CSPM_UNROLL(CSPM_WAVE_UNROLL)
  for (int q = ch; q < rows_end; ++q) {
    const int qy = cy - hw + q;
    const bool inb = in_x && qy >= 0 && qy < sh;
    const int qyc = min(max(qy, 0), sh - 1);
    const int row_off = qyc * pitch + pad;
    uint32_t Iq;
    double gqp;
    load_guide<CC>(tc, row_off + qxc, &Iq, &gqp);
    const double wgt = tc.lut_w[__vsadu4(Ip, Iq) << kLutRepShift];
    const double tv = tap_value_qd<CC, DIR, false, kLutCRepShift>(tc, row_off, qxc, Iq, gqp, wgt, dadd(ax, dadd(dmul(b, (double)qyc), c)));
    const double tw = cross ? dmul(tv, ws) : tv;
    acc = dadd(acc, inb ? tw : 0.0);
  }
  const int remc = wnd - 32, nrem = remc * wnd;
  for (int q = max(ch, wnd); q < ch + nch; ++q) {
    const int r = (q - wnd) * 32 + lane;
    if (r < nrem) {
      const int dy = r / remc, dx = 32 + r - dy * remc;
      const int rx = cx - hw + dx, ry = cy - hw + dy;
      if (rx >= 0 && rx < sw && ry >= 0 && ry < sh) {
        const double tv = tap_term<CC, kLutRepShift, DIR>(tc, ry * pitch + pad, rx, Ip, a, dadd(dmul(b, (double)ry), c));
        acc = dadd(acc, cross ? dmul(tv, ws) : tv);
      }
    }
  }
  return acc;
}

__device__ __forceinline__ double wave_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = dadd(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// BATCH = false: one pair, described by the kernel parameters pc / pm / changed (constant bank).  BATCH = true:
// npairs pairs described by pairs[] in global memory; pc / pm (pair 0) still provide everything the pairs share
// (image size, window, scales, weights, tables).
template <int CC, bool BATCH>
__global__ void __launch_bounds__(kWvThreads, 1) k_spatial_wave(PcDev pc, PmDev pm, const WvPair* __restrict__ pairs, int npairs, int rev,
                                                                unsigned int* progress, uint8_t* changed, unsigned long long* wstats,
                                                                double eps, int force_exact) {
  auto PC = [&](int p) -> const PcDev& { return BATCH ? pairs[p].pc : pc; };
  auto PM = [&](int p) -> const PmDev& { return BATCH ? pairs[p].pm : pm; };
  auto CHG = [&](int p) -> uint8_t* { return BATCH ? pairs[p].changed : changed; };
  if (!BATCH) npairs = 1;
  extern __shared__ __align__(16) double s_dyn[];
  const double* s_lut_c = bulk_lut_c(s_dyn);  // this lane's copy of the colour table
  unsigned long long* s_mbar = bulk_mbar(s_dyn);
  WvCtl* ctl = reinterpret_cast<WvCtl*>(s_mbar + 2);
  tma_stage_luts_begin(pc, s_dyn, s_mbar);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double* lut_w = bulk_lut_w(s_dyn);  // this lane's copy of the 16x replicated weight table
  const size_t n = (size_t)pm.w * pm.h;
  const int inc = rev ? -1 : 1;
  const int hw = pc.half_wnd, wnd = 2 * hw + 1, wtaps = wnd * wnd;
  const unsigned int magic = 0xFFFFFFFFu / (unsigned int)wnd + 1u;  // k / wnd for k < 2^16
  const int S = pc.scale_num, cps = (wtaps + 31) >> 5, cpt = S * cps;
  const bool cross = pc.cross_scale != 0;
  if (threadIdx.x == 0) ctl->n_amb = 0;
  unsigned int st_tasks = 0, st_cands = 0, st_fast = 0, st_exact = 0;
  tma_stage_luts_wait(s_mbar);
#if CSPM_WAVE_PROFILE
  long long prof_t = clock64();
#endif

  for (int diag = 1; diag <= pm.w + pm.h - 2; ++diag) {
    const int u_lo = max(0, diag - (pm.h - 1)), u_hi = min(diag, pm.w - 1);
    // balanced contiguous partition of the diagonal's pixels over the CTAs (a CTA's range barely moves from one
    // diagonal to the next, so its window data stays in this SM's L1)
    const int npix = u_hi - u_lo + 1;
    const int pq = npix / (int)gridDim.x, pr = npix % (int)gridDim.x;
    const int pix0 = (int)blockIdx.x * pq + min((int)blockIdx.x, pr);
    const int cnt = pq + ((int)blockIdx.x < pr ? 1 : 0);
    const int ntasks_pp = 2 * cnt;                // tasks (pixel x view) of one pair on this diagonal
    const int ntasks_mine = npairs * ntasks_pp;   // all pairs of the batch, pair-major
    // dataflow synchronisation: wait until the CTAs that own my pixels' x / y neighbours have published diag - 1
    if (diag >= 2 && cnt > 0 && warp == 0) {
      const int pu_lo = max(0, diag - 1 - (pm.h - 1)), pu_hi = min(diag - 1, pm.w - 1);
      const int pn = pu_hi - pu_lo + 1, pq2 = pn / (int)gridDim.x, pr2 = pn % (int)gridDim.x;
      const int delta = u_lo - pu_lo;
      const int j_lo = max(0, pix0 + delta - 1), j_hi = min(pn - 1, pix0 + cnt - 1 + delta);
      auto owner = [&](int j) {
        const int big = pr2 * (pq2 + 1);
        return j < big ? j / (pq2 + 1) : pr2 + (pq2 > 0 ? (j - big) / pq2 : 0);
      };
      const int o_lo = owner(j_lo), o_hi = owner(j_hi);
      for (int o = o_lo + lane; o <= o_hi; o += 32) {
        unsigned int seen;
        do {
          CSPM_SP_POLL(seen, progress + 32 * o);
        } while ((int)seen < diag - 1);
      }
      __syncwarp();
    }
    WV_PROF(0);
    __syncthreads();
    WV_PROF(1);
    for (int base = 0; base < ntasks_mine; base += kWvTasks) {
      const int nt = min(kWvTasks, ntasks_mine - base);
      // ---------------------------------------------------------------- set-up: one warp per task
      for (int ti = warp; ti < nt; ti += kWvWarps) {
        const int pair = BATCH ? (base + ti) / ntasks_pp : 0;
        const PmDev& pmp = PM(pair);
        const int tid = 2 * pix0 + (base + ti) - pair * ntasks_pp;
        const int u = u_lo + (tid >> 1), t = diag - u;
        const int t_v = tid & 1;
        const int t_x = rev ? pm.w - 1 - u : u, t_y = rev ? pm.h - 1 - t : t;
        const size_t idx = (size_t)t_y * pm.w + t_x;
        bool on0 = u >= 1, on1 = t >= 1;
        // lanes 0..5: x-neighbour's plane, 8..13: y-neighbour's, 16..22: own plane and cost.  The neighbours were
        // written by (possibly) another SM on the previous diagonal: read through L2.
        const double* st = pmp.st[t_v];
        double val = 0.0;
        if (lane < 6) { if (on0) val = __ldcg(st + (size_t)lane * n + (idx - inc)); }
        else if (lane >= 8 && lane < 14) { if (on1) val = __ldcg(st + (size_t)(lane - 8) * n + (idx - (ptrdiff_t)inc * pm.w)); }
        else if (lane >= 16 && lane < 23) val = __ldcg(st + (size_t)(lane - 16) * n + idx);
        double c0[6], c1[6], ow[6];
        bool own0 = true, own1 = true, same = true;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          c0[k] = __shfl_sync(0xffffffffu, val, k);
          c1[k] = __shfl_sync(0xffffffffu, val, 8 + k);
          ow[k] = __shfl_sync(0xffffffffu, val, 16 + k);
          own0 &= bits_eq(c0[k], ow[k]);
          own1 &= bits_eq(c1[k], ow[k]);
          same &= bits_eq(c0[k], c1[k]);
        }
        // exact shortcut 1: a candidate that is bitwise the pixel's own plane re-evaluates to exactly the stored
        // cost (only while every stored cost is known to be GetPlaneCost of its plane) and cannot pass `<`
        if (pmp.cost_consistent) { on0 = on0 && !own0; on1 = on1 && !own1; }
        // exact shortcut 2: the y-neighbour carries the x-neighbour's plane: same cost, so after the x-candidate
        // was tested it can no longer be strictly lower than the minimum
        if (on0 && on1 && same) on1 = false;
        WvTask& T = ctl->task[ti];
        if (lane < 6) T.cand[0][lane] = val;
        else if (lane >= 8 && lane < 14) T.cand[1][lane - 8] = val;
        else if (lane >= 16 && lane < 22) T.cand[2][lane - 16] = val;
        else if (lane == 22) T.m = val;
        if (lane == 0) { T.idx = idx; T.flags = t_v | (on0 ? 2 : 0) | (on1 ? 4 : 0); T.x = t_x; T.y = t_y; T.pair = pair; }
        if (lane < S && (on0 || on1)) {  // lane s prepares pyramid level s
          const int s = lane, cx = t_x >> s, cy = t_y >> s;
          const ScaleDev& sc = PC(pair).sc[s];
          WvScale& D = ctl->sd[ti][s];
          wave_cand_at_scale(pc, c0, t_x, t_y, s, cx, cy, &D.abc[0][0], &D.abc[0][1], &D.abc[0][2]);
          wave_cand_at_scale(pc, c1, t_x, t_y, s, cx, cy, &D.abc[1][0], &D.abc[1][1], &D.abc[1][2]);
          wave_cand_at_scale(pc, ow, t_x, t_y, s, cx, cy, &D.abc[2][0], &D.abc[2][1], &D.abc[2][2]);
          D.cx = cx; D.cy = cy;
          D.Ip = guide_key<CC>(sc.color[t_v][cy * sc.pitch + sc.pad + cx]);
        }
      }
      __syncthreads();
      // ---------------------------------------------------------------- evaluate: tree sums of all candidates.
      // The candidates that survived the shortcuts are numbered 0 .. E-1 (task order, x before y); their
      // E * cpt chunks of 32 taps are dealt to the warps in contiguous runs, so every warp has the same load
      // whatever the shortcuts removed.  Every warp derives the numbering from the task flags (lane <-> task).
      const int my_flags = lane < nt ? ctl->task[lane].flags : 0;
      const int my_non = ((my_flags >> 1) & 1) + ((my_flags >> 2) & 1);
      int incl = my_non;  // inclusive prefix over the tasks
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
      }
      const int E = __shfl_sync(0xffffffffu, incl, 31);
      const int total = E * cpt, per = max(1, (total + kWvWarps - 1) / kWvWarps);
      const int g0 = warp * per, g1 = min(total, g0 + per);
      const int first_eval = g0 / cpt;
      for (int g = g0; g < g1;) {
        const int e = g / cpt;
        const int task = __ffs(__ballot_sync(0xffffffffu, incl > e)) - 1;
        const int excl_t = __shfl_sync(0xffffffffu, incl - my_non, task);
        const int flags = __shfl_sync(0xffffffffu, my_flags, task);
        const int cand = (e - excl_t == 0 && (flags & 2)) ? 0 : 1;
        const int t_v = flags & 1;
        const PcDev& pcp = PC(BATCH ? ctl->task[task].pair : 0);
        double acc = 0.0;
        const int g_end = min(g1, (e + 1) * cpt);
        while (g < g_end) {
          const int rem = g - e * cpt, s = rem / cps, ch = rem - s * cps;
          const int nch = min(cps - ch, g_end - g);
          const ScaleDev& sc = pcp.sc[s];
          const TapCtx tc = make_tap_ctx(pcp, s, t_v, lut_w, s_lut_c);
          const WvScale& D = ctl->sd[task][s];
          const double a = D.abc[cand][0], b = D.abc[cand][1], c = D.abc[cand][2];
          const int cx = D.cx, cy = D.cy;
          const uint32_t Ip = D.Ip;
          const double ws = pc.scale_wgt[s];
          if (t_v == 0) acc = wave_eval_segment<CC, -1>(tc, sc, cx, cy, Ip, a, b, c, hw, wnd, wtaps, magic, ch, nch, ws, cross, acc);
          else acc = wave_eval_segment<CC, 1>(tc, sc, cx, cy, Ip, a, b, c, hw, wnd, wtaps, magic, ch, nch, ws, cross, acc);
          g += nch;
        }
        const double r = wave_warp_sum(acc);
        if (lane == 0) ctl->partial[warp][e - first_eval] = r;
      }
      WV_PROF(2);
      __syncthreads();
      WV_PROF(3);
      // ---------------------------------------------------------------- decide: certified comparisons
      if (warp == 0) {
        double Tc[2] = {0.0, 0.0};
        if (lane < nt) {
          int e = incl - my_non;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            if (!(my_flags & (2 << c))) continue;
            const int w_first = (e * cpt) / per, w_last = ((e + 1) * cpt - 1) / per;
            double tsum = 0.0;
            for (int w = w_first; w <= w_last; ++w) tsum = dadd(tsum, ctl->partial[w][e - (w * per) / cpt]);
            Tc[c] = tsum;
            ++e;
          }
          const WvTask& T = ctl->task[lane];
          const double T0 = Tc[0], T1 = Tc[1];
          const bool e0 = (my_flags & 2) != 0, e1 = (my_flags & 4) != 0;
          const double m = T.m;
          const double lo0 = dmul(T0, dsub(1.0, eps)), hi0 = dmul(T0, dadd(1.0, eps));
          const double lo1 = dmul(T1, dsub(1.0, eps)), hi1 = dmul(T1, dadd(1.0, eps));
          int best = -1;
          bool amb = force_exact && (e0 || e1);
          double min_lo = m, min_hi = m;
          if (!amb && e0) {  // x-axis neighbour first (:198-204)
            if (hi0 < m) { best = 0; min_lo = lo0; min_hi = hi0; }
            else if (!(lo0 >= m)) amb = true;
          }
          if (!amb && e1) {  // then the y-axis neighbour against the updated minimum (:206-212)
            if (hi1 < min_lo) best = 1;
            else if (!(lo1 >= min_hi)) amb = true;
          }
          st_tasks += 1;
          st_cands += (e0 ? 1 : 0) + (e1 ? 1 : 0);
          if (amb) {
            const int slot = atomicAdd(&ctl->n_amb, 1);
            ctl->amb[slot] = lane;
            ctl->amb_T[slot][0] = T0;
            ctl->amb_T[slot][1] = T1;
#if CSPM_WAVE_PROFILE
            if (blockIdx.x == gridDim.x / 2 && !force_exact) {
              const unsigned int k = atomicAdd(&g_wv_nsamples, 1u);
              if (k < 64) { g_wv_samples[k][0] = T0; g_wv_samples[k][1] = T1; g_wv_samples[k][2] = m; g_wv_samples[k][3] = (double)T.flags; }
            }
#endif
          } else if (best >= 0) {
            const int v = T.flags & 1;
            double* st = PM(T.pair).st[v];
            const double* p = T.cand[best];
            st[T.idx] = p[0]; st[n + T.idx] = p[1]; st[2 * n + T.idx] = p[2];
            st[3 * n + T.idx] = p[3]; st[4 * n + T.idx] = p[4]; st[5 * n + T.idx] = p[5];
            st[6 * n + T.idx] = best == 0 ? T0 : T1;  // placeholder until k_cost_fix
            CHG(T.pair)[(size_t)v * n + T.idx] = 1;
            st_fast += 1;
          }
        }
      }
      WV_PROF(4);
      __syncthreads();
      WV_PROF(5);
      // ---------------------------------------------------------------- the undecided tasks (near-ties; ~1 %)
      // Step A — term-wise certificate, all threads: the addends of both candidates AND of the pixel's own plane.
      // RN addition is monotone, so if candidate addend >= own addend at EVERY tap (every scale), every partial
      // sum of the candidate's row-major chain is >= the own plane's, hence S_cand >= S_own = m (the stored cost IS
      // GetPlaneCost of the own plane while cost_consistent): `S_cand < m` is false without summing anything.
      // Likewise y-candidate >= x-candidate term-wise gives S_1 >= S_0.  This settles the exact ties that make up
      // most undecided tasks (planes whose partners all lie outside the image cost the same, tap by tap).
      // Step B — what is still open: lane 0 / 1 of warp 0 add the x- / y-candidate's addends in the reference's
      // row-major order (pre_ss_pc.cc:84-116, pre_cs_pc.cc:151-182) and the reference's comparison is applied.
      const int n_amb = ctl->n_amb;
      for (int ai = 0; ai < n_amb; ++ai) {
        const int ti = ctl->amb[ai];
        const WvTask& T = ctl->task[ti];
        const int flags = T.flags, t_v = flags & 1;
        const PcDev& pcp = PC(T.pair);
        const PmDev& pmp = PM(T.pair);
        const bool on0 = (flags & 2) != 0, on1 = (flags & 4) != 0;
        bool ge0 = true, ge1o = true, ge10 = true;  // this thread's taps: t0 >= own, t1 >= own, t1 >= t0
        if (!force_exact) {
          for (int s = 0; s < S; ++s) {
            const ScaleDev& sc = pcp.sc[s];
            const TapCtx tc = make_tap_ctx(pcp, s, t_v, lut_w, s_lut_c);
            const WvScale& D = ctl->sd[ti][s];
            for (int k = threadIdx.x; k < wtaps; k += kWvThreads) {
              const double to = wave_tap<CC, 0>(tc, sc, D.cx, D.cy, D.Ip, D.abc[2][0], D.abc[2][1], D.abc[2][2], hw, wnd, wtaps, magic, k);
              const double t0 = on0 ? wave_tap<CC, 0>(tc, sc, D.cx, D.cy, D.Ip, D.abc[0][0], D.abc[0][1], D.abc[0][2], hw, wnd, wtaps, magic, k) : to;
              const double t1 = on1 ? wave_tap<CC, 0>(tc, sc, D.cx, D.cy, D.Ip, D.abc[1][0], D.abc[1][1], D.abc[1][2], hw, wnd, wtaps, magic, k) : to;
              ge0 = ge0 && t0 >= to;
              ge1o = ge1o && t1 >= to;
              ge10 = ge10 && t1 >= t0;
            }
          }
        }
        const int votes = __syncthreads_and((ge0 ? 1 : 0)) | (__syncthreads_and((ge1o ? 1 : 0)) << 1) | (__syncthreads_and((ge10 ? 1 : 0)) << 2);
        // the decision with the certificates as extra knowledge (CTA-uniform)
        const double m = T.m, T0 = ctl->amb_T[ai][0], T1 = ctl->amb_T[ai][1];
        const bool cert_ok = !force_exact && pmp.cost_consistent;
        int best = -1;
        bool open_q = force_exact != 0;
        if (!open_q) {
          const double lo0 = dmul(T0, dsub(1.0, eps)), hi0 = dmul(T0, dadd(1.0, eps));
          const double lo1 = dmul(T1, dsub(1.0, eps)), hi1 = dmul(T1, dadd(1.0, eps));
          bool acc0 = false;
          if (on0) {
            if (hi0 < m) acc0 = true;
            else if (lo0 >= m || (cert_ok && (votes & 1))) acc0 = false;
            else open_q = true;
          }
          if (!open_q && on1) {
            if (acc0) {  // against S_0
              if (hi1 < lo0) best = 1;
              else if (lo1 >= hi0 || (votes & 4)) best = 0;
              else open_q = true;
            } else {     // against m
              if (hi1 < m) best = 1;
              else if (lo1 >= m || (cert_ok && (votes & 2))) best = -1;
              else open_q = true;
            }
          } else if (!open_q && acc0) best = 0;
        }
        if (!open_q) {
          if (best >= 0 && threadIdx.x == 0) {
            double* st = pmp.st[t_v];
            const double* p = T.cand[best];
            st[T.idx] = p[0]; st[n + T.idx] = p[1]; st[2 * n + T.idx] = p[2];
            st[3 * n + T.idx] = p[3]; st[4 * n + T.idx] = p[4]; st[5 * n + T.idx] = p[5];
            st[6 * n + T.idx] = best == 0 ? T0 : T1;  // placeholder until k_cost_fix
            CHG(T.pair)[(size_t)t_v * n + T.idx] = 1;
          }
          continue;
        }
        // Step B
        double cost = 0.0;  // lane 0 / 1 of warp 0: the candidate's running cost over the scales
        for (int s = 0; s < S; ++s) {
          const ScaleDev& sc = pcp.sc[s];
          const TapCtx tc = make_tap_ctx(pcp, s, t_v, lut_w, s_lut_c);
          const WvScale& D = ctl->sd[ti][s];
          for (int k = threadIdx.x; k < wtaps; k += kWvThreads) {
            ctl->xbuf[0][k] = on0 ? wave_tap<CC, 0>(tc, sc, D.cx, D.cy, D.Ip, D.abc[0][0], D.abc[0][1], D.abc[0][2], hw, wnd, wtaps, magic, k) : 0.0;
            ctl->xbuf[1][k] = on1 ? wave_tap<CC, 0>(tc, sc, D.cx, D.cy, D.Ip, D.abc[1][0], D.abc[1][1], D.abc[1][2], hw, wnd, wtaps, magic, k) : 0.0;
          }
          __syncthreads();
          if (warp == 0 && lane < 2 && (lane == 0 ? on0 : on1)) {
            const int y0 = max(D.cy - hw, 0), y1 = min(D.cy + hw, sc.h - 1), x0 = max(D.cx - hw, 0), x1 = min(D.cx + hw, sc.w - 1);
            const volatile double* xb = ctl->xbuf[lane];
            double sum = 0.0;
            for (int qy = y0; qy <= y1; ++qy) {
              const volatile double* row = xb + (qy - D.cy + hw) * wnd + (x0 - D.cx + hw);
              const int nc = x1 - x0 + 1;
              int j = 0;
              for (; j + 8 <= nc; j += 8) {  // 8 loads in flight, then 8 dependent adds
                double t[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) t[q] = row[j + q];
#pragma unroll
                for (int q = 0; q < 8; ++q) sum = dadd(sum, t[q]);
              }
              for (; j < nc; ++j) sum = dadd(sum, row[j]);
            }
            cost = cross ? dadd(cost, dmul(sum, pc.scale_wgt[s])) : sum;  // pre_cs_pc.cc:182
          }
          __syncthreads();
        }
        if (warp == 0) {
          const double s0 = __shfl_sync(0xffffffffu, cost, 0), s1 = __shfl_sync(0xffffffffu, cost, 1);
          if (lane == 0) {
            double min_cost = m;
            int bb = -1;
            if (on0 && s0 < min_cost) { min_cost = s0; bb = 0; }
            if (on1 && s1 < min_cost) { min_cost = s1; bb = 1; }
            if (bb >= 0) {
              PlaneD p;
              const double* c = T.cand[bb];
              p.nx = c[0]; p.ny = c[1]; p.nz = c[2]; p.a = c[3]; p.b = c[4]; p.c = c[5];
              store_plane(pmp.st[t_v], n, (size_t)T.idx, p, min_cost);
            }
            st_exact += 1;
          }
        }
      }
      if (n_amb > 0) {
        __syncthreads();
        if (threadIdx.x == 0) ctl->n_amb = 0;
      }
    }
    // publish: the deciding threads' plane stores happen-before this release through the CTA barrier above
    // (release is cumulative); one store per CTA and diagonal
    if (threadIdx.x == 0)
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(progress + 32 * blockIdx.x), "r"((unsigned int)diag) : "memory");
    WV_PROF(6);
#if CSPM_WAVE_PROFILE
    if (blockIdx.x == gridDim.x / 2 && threadIdx.x == 0) g_sp_prof[7] += 1;
#endif
  }
  if (warp == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      st_tasks += __shfl_xor_sync(0xffffffffu, st_tasks, o);
      st_cands += __shfl_xor_sync(0xffffffffu, st_cands, o);
      st_fast += __shfl_xor_sync(0xffffffffu, st_fast, o);
      st_exact += __shfl_xor_sync(0xffffffffu, st_exact, o);
    }
    if (lane == 0) {
      atomicAdd(wstats + 0, (unsigned long long)st_tasks);
      atomicAdd(wstats + 1, (unsigned long long)st_cands);
      atomicAdd(wstats + 2, (unsigned long long)st_fast);
      atomicAdd(wstats + 3, (unsigned long long)st_exact);
    }
  }
}

// Pixels marked by the sweep, as per-view lists in (nearly) row-major order, so that the lanes of a warp of
// k_cost_fix still work on neighbouring pixels of one image row.
__global__ void k_wave_compact(const uint8_t* __restrict__ changed, int n, int* __restrict__ list, unsigned int* __restrict__ count) {
  const int v = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool take = i < n && changed[(size_t)v * n + i];
  const unsigned int m = __ballot_sync(0xffffffffu, take);
  if (!m) return;
  const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
  unsigned int base = 0;
  if (lane == leader) base = atomicAdd(count + v, (unsigned int)__popc(m));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (take) list[(size_t)v * n + base + __popc(m & ((1u << lane) - 1u))] = i;
}

// Exact cost of every pixel that adopted a plane on the sweep's fast path: GetPlaneCost of its final plane, by the
// same bit-exact thread-per-pixel evaluator as InitRandomPlane (no bound: the full sum).
template <int CC, bool CS>
__global__ void __launch_bounds__(32 * bulk_by<CS>(), kBulkMinB) k_cost_fix(PcDev pc, PmDev pm, const int* __restrict__ list,
                                                                            const unsigned int* __restrict__ count) {
  extern __shared__ __align__(16) double s_dyn[];
  constexpr unsigned int kCta = 32u * (CS ? kBulkByCs : kBulkBySs);
  const int v = blockIdx.y;
  const unsigned int cnt = count[v];
  if (blockIdx.x * kCta >= cnt) return;  // CTA-uniform, before the bulk copy is issued
  const double* s_lut_c = bulk_lut_c(s_dyn);  // this lane's copy of the colour table
  unsigned long long* s_mbar = bulk_mbar(s_dyn);
  tma_stage_luts_begin(pc, s_dyn, s_mbar);
  const double* s_lut_w = bulk_lut_w(s_dyn);
  const unsigned int slot = blockIdx.x * kCta + threadIdx.y * 32 + threadIdx.x;
  tma_stage_luts_wait(s_mbar);
  if (slot >= cnt) return;
  const size_t n = (size_t)pm.w * pm.h;
  const int i = list[(size_t)v * n + slot];
  const int y = i / pm.w, x = i - y * pm.w;
  const PlaneD p = load_plane(pm.st[v], n, (size_t)i);
  // lanes of a warp sit on different image rows here: the per-lane loop, not the cooperative tail
  pm.st[v][6 * n + i] = plane_cost<CC, kLutRepShift, CS ? 1 : 0, false>(pc, v, x, y, p, s_lut_w, s_lut_c, CUDART_INF);
}
