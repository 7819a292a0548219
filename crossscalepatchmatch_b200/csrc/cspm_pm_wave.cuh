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
//   * if they overlap (in practice: exact ties, e.g. two different planes whose windows are invalid everywhere)
//     the task takes the EXACT path: one warp computes each window row's addends, two lanes add them in the
//     reference's order, and the reference's comparison is applied to those sums.
// A pixel that adopts a plane on the fast path gets the tree sum as a PLACEHOLDER cost and a mark in `changed`;
// right after the sweep k_cost_fix recomputes the cost of every marked pixel with the bit-exact bulk evaluator
// (thread per pixel, row-major order), before anything reads it.  Planes, costs and disparities therefore stay
// bit-identical to the reference (tests: stage parity, and the same with CSPM_WAVE_EXACT=1, which sends every
// task down the exact path).
//
// Work layout per CTA and diagonal: the CTA's tasks (pixel x view, two candidates sharing the adaptive weight) are
// cut into chunks of 32 consecutive window taps (x scales); the chunk list is dealt to the warps in contiguous
// runs, lanes along taps (coalesced guide / partner rows whatever the plane), each lane keeps running partial sums
// and a warp reduces them once per task it touched.  No thread ever waits on another inside the evaluation; the
// per-diagonal critical path is ~cnt * 39 / warps chunk iterations instead of a 1225-long dependent add chain.
#pragma once

#ifndef CSPM_WAVE_THREADS
#define CSPM_WAVE_THREADS 768
#endif
#ifndef CSPM_WAVE_UNROLL
#define CSPM_WAVE_UNROLL 2
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
constexpr int kWvSlots = 4;   // tasks one warp's contiguous run can touch: <= ceil(kWvTasks / kWvWarps) + 2

struct WvTask {
  double cand[2][6];       // nx ny nz a b c of the x-axis and the y-axis neighbour
  double m;                // the pixel's cost at the start of the step (:201,209 compare against it)
  unsigned long long idx;  // y * w + x
  int flags;               // bit0 view, bit1 x-candidate to evaluate, bit2 y-candidate to evaluate
  int x, y, pad_;
};
struct WvCtl {
  WvTask task[kWvTasks];
  double partial[kWvWarps][kWvSlots][2];
  double strip[kWvWarps][2][64];  // exact path: one window row of addends per candidate
  int amb[kWvTasks];
  int n_amb, pad_[3];
};
constexpr size_t kWvSmemBytes = kBulkLutBytes + 16 + sizeof(WvCtl);

// counters of one sweep: {tasks, candidates evaluated, planes adopted on the fast path, tasks sent down the exact path}
constexpr int kWvStats = 4;

struct WvCand {
  double a, b, c;
};

// Plane of candidate `p` (six numbers nx ny nz a b c) as GetPlaneCost sees it at scale s (pre_cs_pc.cc:139-147,183-185).
__device__ __forceinline__ WvCand wave_cand_at_scale(const PcDev& pc, const double p[6], int x, int y, int s, int cx, int cy) {
  WvCand q;
  if (!pc.cross_scale) {
    q.a = p[3]; q.b = p[4]; q.c = p[5];
    return q;
  }
  PlaneD pl;
  pl.nx = p[0]; pl.ny = p[1]; pl.nz = p[2]; pl.a = p[3]; pl.b = p[4]; pl.c = p[5];
  double disp = plane_disp(pl, x, y);
  for (int k = 0; k < s; ++k) disp = ddiv(disp, 2.0);
  update_param(pl, (double)cx, (double)cy, disp);
  q.a = pl.a; q.b = pl.b; q.c = pl.c;
  return q;
}

// chunks [ch0, ch0 + nch) of one (task, scale): lane <-> tap ch * 32 + lane in row-major window order
template <int CC, int DIR>
__device__ __forceinline__ void wave_eval_chunks(const TapCtx& tc, const ScaleDev& sc, int cx, int cy, int hw, int wnd, int wtaps,
                                                 unsigned int magic, int ch0, int nch, uint32_t Ip, bool on0, bool on1,
                                                 const WvCand& q0, const WvCand& q1, double ws, bool cross, double& acc0,
                                                 double& acc1) {
  const int lane = threadIdx.x & 31;
  CSPM_UNROLL(CSPM_WAVE_UNROLL)
  for (int ch = ch0; ch < ch0 + nch; ++ch) {
    const int k = ch * 32 + lane;
    const int dy = (int)__umulhi((unsigned int)k, magic), dx = k - dy * wnd;
    const int qx = cx + dx - hw, qy = cy + dy - hw;
    if (k < wtaps && qx >= 0 && qx < sc.w && qy >= 0 && qy < sc.h) {
      const int row_off = qy * sc.pitch + sc.pad;
      uint32_t Iq;
      double gqp;
      load_guide<CC>(tc, row_off + qx, &Iq, &gqp);
      const double wgt = tc.lut_w[__vsadu4(Ip, Iq) << kLutRepShift];
      const double qyd = (double)qy;
      if (on0) {
        const double t = tap_value<CC, DIR>(tc, row_off, qx, Iq, gqp, wgt, q0.a, dadd(dmul(q0.b, qyd), q0.c));
        acc0 = dadd(acc0, cross ? dmul(t, ws) : t);
      }
      if (on1) {
        const double t = tap_value<CC, DIR>(tc, row_off, qx, Iq, gqp, wgt, q1.a, dadd(dmul(q1.b, qyd), q1.c));
        acc1 = dadd(acc1, cross ? dmul(t, ws) : t);
      }
    }
  }
}

__device__ __forceinline__ double wave_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = dadd(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Exact GetPlaneCost of both candidates of one task by one warp: lanes compute a window row's addends, lane 0 / 1
// add the x- / y-candidate's in the reference's order (pre_ss_pc.cc:84-116, pre_cs_pc.cc:151-182).  Returns the
// candidate's cost in lanes 0 and 1.
template <int CC>
__device__ __forceinline__ double wave_exact_pair(const PcDev& pc, const WvTask& T, const double* lut_w, const double* lut_c,
                                                  double (*strip)[64]) {
  const int lane = threadIdx.x & 31;
  const int v = T.flags & 1, hw = pc.half_wnd;
  const bool on0 = (T.flags & 2) != 0, on1 = (T.flags & 4) != 0;
  const bool mine = lane == 0 ? on0 : (lane == 1 ? on1 : false);
  double cost = 0.0;
  int cx = T.x, cy = T.y;
  for (int s = 0; s < pc.scale_num; ++s) {
    const ScaleDev& sc = pc.sc[s];
    const TapCtx tc = make_tap_ctx(pc, s, v, lut_w, lut_c);
    const uint32_t Ip = guide_key<CC>(tc.col_v[cy * sc.pitch + sc.pad + cx]);
    const WvCand q0 = wave_cand_at_scale(pc, T.cand[0], T.x, T.y, s, cx, cy);
    const WvCand q1 = wave_cand_at_scale(pc, T.cand[1], T.x, T.y, s, cx, cy);
    const int y0 = max(cy - hw, 0), y1 = min(cy + hw, sc.h - 1), x0 = max(cx - hw, 0), x1 = min(cx + hw, sc.w - 1);
    const int ncols = x1 - x0 + 1;
    double sum = 0.0;
    for (int qy = y0; qy <= y1; ++qy) {
      const int row_off = qy * sc.pitch + sc.pad;
      const double qyd = (double)qy;
      for (int j = lane; j < ncols; j += 32) {
        const int qx = x0 + j;
        uint32_t Iq;
        double gqp;
        load_guide<CC>(tc, row_off + qx, &Iq, &gqp);
        const double wgt = tc.lut_w[__vsadu4(Ip, Iq) << kLutRepShift];
        if (on0) strip[0][j] = tap_value<CC>(tc, row_off, qx, Iq, gqp, wgt, q0.a, dadd(dmul(q0.b, qyd), q0.c));
        if (on1) strip[1][j] = tap_value<CC>(tc, row_off, qx, Iq, gqp, wgt, q1.a, dadd(dmul(q1.b, qyd), q1.c));
      }
      __syncwarp();
      if (mine) {
        const volatile double* sp = strip[lane];
        for (int j = 0; j < ncols; ++j) sum = dadd(sum, sp[j]);
      }
      __syncwarp();
    }
    if (pc.cross_scale) cost = dadd(cost, dmul(sum, pc.scale_wgt[s]));  // pre_cs_pc.cc:182
    else cost = sum;
    cx /= 2;
    cy /= 2;
  }
  return cost;
}

template <int CC>
__global__ void __launch_bounds__(kWvThreads, 1) k_spatial_wave(PcDev pc, PmDev pm, int rev, unsigned int* progress, uint8_t* changed,
                                                                unsigned long long* wstats, double eps, int force_exact) {
  extern __shared__ __align__(16) double s_dyn[];
  double* s_lut_c = s_dyn + (size_t)kLutWUsed * (1 << kLutRepShift);
  unsigned long long* s_mbar = reinterpret_cast<unsigned long long*>(s_lut_c + kLutC);
  WvCtl* ctl = reinterpret_cast<WvCtl*>(s_mbar + 2);
  tma_stage_luts_begin(pc, s_dyn, s_mbar);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const double* lut_w = s_dyn + (lane & 15);  // this lane's copy of the 16x replicated weight table
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
    const int ntasks_mine = 2 * cnt;
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
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(progress + 32 * o) : "memory");
        } while ((int)seen < diag - 1);
      }
    }
    WV_PROF(0);
    __syncthreads();
    WV_PROF(1);
    for (int base = 0; base < ntasks_mine; base += kWvTasks) {
      const int nt = min(kWvTasks, ntasks_mine - base);
      // ---------------------------------------------------------------- evaluate: tree sums of all candidates
      const int total = nt * cpt, per = (total + kWvWarps - 1) / kWvWarps;
      const int g0 = warp * per, g1 = min(total, g0 + per);
      const int first_task = g0 < total ? g0 / cpt : 0;
      int cur_task = -1;
      double acc0 = 0.0, acc1 = 0.0;
      // the current task, broadcast to every lane of the warp
      double c0[6], c1[6];
      int t_v = 0, t_x = 0, t_y = 0;
      bool on0 = false, on1 = false;
      for (int g = g0; g < g1;) {
        const int task = g / cpt, rem = g - task * cpt, s = rem / cps, ch = rem - s * cps;
        if (task != cur_task) {
          if (cur_task >= 0) {
            const double r0 = wave_warp_sum(acc0), r1 = wave_warp_sum(acc1);
            if (lane == 0) { ctl->partial[warp][cur_task - first_task][0] = r0; ctl->partial[warp][cur_task - first_task][1] = r1; }
            acc0 = 0.0; acc1 = 0.0;
          }
          cur_task = task;
          const int tid = 2 * pix0 + base + task;
          const int u = u_lo + (tid >> 1), t = diag - u;
          t_v = tid & 1;
          t_x = rev ? pm.w - 1 - u : u;
          t_y = rev ? pm.h - 1 - t : t;
          const size_t idx = (size_t)t_y * pm.w + t_x;
          on0 = u >= 1;
          on1 = t >= 1;
          // lanes 0..5: x-neighbour's plane, 8..13: y-neighbour's, 16..22: own plane and cost.  The neighbours were
          // written by (possibly) another SM on the previous diagonal: read through L2.
          const double* st = pm.st[t_v];
          double val = 0.0;
          if (lane < 6) { if (on0) val = __ldcg(st + (size_t)lane * n + (idx - inc)); }
          else if (lane >= 8 && lane < 14) { if (on1) val = __ldcg(st + (size_t)(lane - 8) * n + (idx - (ptrdiff_t)inc * pm.w)); }
          else if (lane >= 16 && lane < 23) val = __ldcg(st + (size_t)(lane - 16) * n + idx);
          bool own0 = true, own1 = true, same = true;
#pragma unroll
          for (int k = 0; k < 6; ++k) {
            c0[k] = __shfl_sync(0xffffffffu, val, k);
            c1[k] = __shfl_sync(0xffffffffu, val, 8 + k);
            const double o = __shfl_sync(0xffffffffu, val, 16 + k);
            own0 &= bits_eq(c0[k], o);
            own1 &= bits_eq(c1[k], o);
            same &= bits_eq(c0[k], c1[k]);
          }
          // exact shortcut 1: a candidate that is bitwise the pixel's own plane re-evaluates to exactly the stored
          // cost (only while every stored cost is known to be GetPlaneCost of its plane) and cannot pass `<`
          if (pm.cost_consistent) { on0 = on0 && !own0; on1 = on1 && !own1; }
          // exact shortcut 2: the y-neighbour carries the x-neighbour's plane: same cost, so after the x-candidate
          // was tested it can no longer be strictly lower than the minimum
          if (on0 && on1 && same) on1 = false;
          if (rem == 0) {  // the warp that owns the task's first chunk files it for the decision phase
            WvTask& T = ctl->task[task];
            if (lane < 6) T.cand[0][lane] = val;
            else if (lane >= 8 && lane < 14) T.cand[1][lane - 8] = val;
            else if (lane == 22) T.m = val;
            if (lane == 0) { T.idx = idx; T.flags = t_v | (on0 ? 2 : 0) | (on1 ? 4 : 0); T.x = t_x; T.y = t_y; }
          }
        }
        const int nch = min(cps - ch, g1 - g);
        if (on0 || on1) {
          const int cx = t_x >> s, cy = t_y >> s;
          const ScaleDev& sc = pc.sc[s];
          const TapCtx tc = make_tap_ctx(pc, s, t_v, lut_w, s_lut_c);
          const uint32_t Ip = guide_key<CC>(tc.col_v[cy * sc.pitch + sc.pad + cx]);
          const WvCand q0 = wave_cand_at_scale(pc, c0, t_x, t_y, s, cx, cy);
          const WvCand q1 = wave_cand_at_scale(pc, c1, t_x, t_y, s, cx, cy);
          const double ws = pc.scale_wgt[s];
          if (t_v == 0) wave_eval_chunks<CC, -1>(tc, sc, cx, cy, hw, wnd, wtaps, magic, ch, nch, Ip, on0, on1, q0, q1, ws, cross, acc0, acc1);
          else wave_eval_chunks<CC, 1>(tc, sc, cx, cy, hw, wnd, wtaps, magic, ch, nch, Ip, on0, on1, q0, q1, ws, cross, acc0, acc1);
        }
        g += nch;
      }
      if (cur_task >= 0) {
        const double r0 = wave_warp_sum(acc0), r1 = wave_warp_sum(acc1);
        if (lane == 0) { ctl->partial[warp][cur_task - first_task][0] = r0; ctl->partial[warp][cur_task - first_task][1] = r1; }
      }
      WV_PROF(2);
      __syncthreads();
      WV_PROF(3);
      // ---------------------------------------------------------------- decide: certified comparisons
      if (warp == 0 && lane < nt) {
        const WvTask& T = ctl->task[lane];
        const int w_first = (lane * cpt) / per, w_last = ((lane + 1) * cpt - 1) / per;
        double T0 = 0.0, T1 = 0.0;
        for (int w = w_first; w <= w_last; ++w) {
          const int slot = lane - (w * per) / cpt;
          T0 = dadd(T0, ctl->partial[w][slot][0]);
          T1 = dadd(T1, ctl->partial[w][slot][1]);
        }
        const bool e0 = (T.flags & 2) != 0, e1 = (T.flags & 4) != 0;
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
          ctl->amb[atomicAdd(&ctl->n_amb, 1)] = lane;
          st_exact += 1;
        } else if (best >= 0) {
          const int v = T.flags & 1;
          double* st = pm.st[v];
          const double* p = T.cand[best];
          st[T.idx] = p[0]; st[n + T.idx] = p[1]; st[2 * n + T.idx] = p[2];
          st[3 * n + T.idx] = p[3]; st[4 * n + T.idx] = p[4]; st[5 * n + T.idx] = p[5];
          st[6 * n + T.idx] = best == 0 ? T0 : T1;  // placeholder until k_cost_fix
          changed[(size_t)v * n + T.idx] = 1;
          st_fast += 1;
        }
      }
      WV_PROF(4);
      __syncthreads();
      WV_PROF(5);
      // ---------------------------------------------------------------- exact path for the undecided tasks (rare)
      const int n_amb = ctl->n_amb;
      if (n_amb > 0) {
        for (int ai = warp; ai < n_amb; ai += kWvWarps) {
          const WvTask& T = ctl->task[ctl->amb[ai]];
          const double cst = wave_exact_pair<CC>(pc, T, lut_w, s_lut_c, ctl->strip[warp]);
          const double s0 = __shfl_sync(0xffffffffu, cst, 0), s1 = __shfl_sync(0xffffffffu, cst, 1);
          if (lane == 0) {
            double min_cost = T.m;
            int best = -1;
            if ((T.flags & 2) && s0 < min_cost) { min_cost = s0; best = 0; }
            if ((T.flags & 4) && s1 < min_cost) { min_cost = s1; best = 1; }
            if (best >= 0) {
              PlaneD p;
              const double* c = T.cand[best];
              p.nx = c[0]; p.ny = c[1]; p.nz = c[2]; p.a = c[3]; p.b = c[4]; p.c = c[5];
              store_plane(pm.st[T.flags & 1], n, (size_t)T.idx, p, min_cost);
            }
          }
        }
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
  double* s_lut_c = s_dyn + (size_t)kLutWUsed * (1 << kLutRepShift);
  unsigned long long* s_mbar = reinterpret_cast<unsigned long long*>(s_lut_c + kLutC);
  tma_stage_luts_begin(pc, s_dyn, s_mbar);
  const double* s_lut_w = s_dyn + (threadIdx.x & 15);
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
