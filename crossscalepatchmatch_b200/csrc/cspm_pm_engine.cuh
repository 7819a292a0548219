// cspm_pm_engine.cuh — PatchMatch stages on the streaming evaluator
// (cspm_device.cuh: engine_eval).  Included by cspm_pm.cu inside namespace cspm.
//
//   k_spatial_persist   SpatialPropagation: one cooperative launch per sweep;
//                       every CTA walks the anti-diagonals, owns a contiguous
//                       run of each diagonal's pixels (16 pixels x 2 candidates
//                       per pass) and synchronises with its partition
//                       neighbours through per-CTA progress flags (dataflow,
//                       no grid-wide barrier)
// (The bulk stages — init, refinement, view propagation — stay thread-per-pixel:
// an engine version of them was measured 1.8x slower, 950 vs 535 ms per HD run,
// because every pixel has only one live candidate at a time.)
#pragma once

#ifndef CSPM_SP_THREADS
#define CSPM_SP_THREADS 512
#endif
#ifndef CSPM_SP_LEANFENCE
#define CSPM_SP_LEANFENCE 1
#endif
#ifndef CSPM_SPATIAL_DEDUP
#define CSPM_SPATIAL_DEDUP 1
#endif
// Neighbour-progress poll of the persistent sweeps.  ld.acquire.gpu compiles to LDG.STRONG.GPU + CCTL.IVALL, i.e.
// every poll iteration invalidates the SM's L1.  A relaxed poll (-DCSPM_SP_ACQUIRE_POLL=0) would be enough in
// practice — the only data another CTA writes during the kernel is the plane state, which is read exclusively
// with ld.global.cg (LDG.STRONG.GPU, served by L2), and L1 only holds the immutable images — but it was measured
// to buy 1 % of the sweep at most (HD ring 212.5 vs 215.4 ms, 450x375 wavefront 34.6 vs 35.3 ms), so the
// formally synchronising acquire stays the default.
#ifndef CSPM_SP_ACQUIRE_POLL
#define CSPM_SP_ACQUIRE_POLL 1
#endif
#if CSPM_SP_ACQUIRE_POLL
#define CSPM_SP_POLL(dst, ptr) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(dst) : "l"(ptr) : "memory")
#else
#define CSPM_SP_POLL(dst, ptr) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(dst) : "l"(ptr) : "memory")
#endif
__device__ __forceinline__ bool bits_eq(double a, double b) { return __double_as_longlong(a) == __double_as_longlong(b); }

constexpr int kSpThreads = CSPM_SP_THREADS;
constexpr int kSpCtasPerSm = 1;
constexpr int kSpTasks = 16;  // pixels per CTA per pass: one producer warp per pixel at least
constexpr int kSpRows = 5;      // window rows per pipeline chunk (WND = 35)

// Cycle accounting of the persistent sweep (CTA 0 only; read back through
// cspm_debug_counters): {neighbour wait + task setup, evaluation, accept, publish}.
__device__ unsigned long long g_sp_prof[8];
#if CSPM_SP_PROFILE
#define SP_PROF(slot)                                              \
  do {                                                             \
    if (blockIdx.x == 0 && threadIdx.x == 0) {                     \
      const long long now__ = clock64();                           \
      g_sp_prof[slot] += (unsigned long long)(now__ - prof_t);     \
      prof_t = now__;                                              \
    }                                                              \
  } while (0)
#else
#define SP_PROF(slot) do { } while (0)
#endif

template <int CC, int WND>
__global__ void __launch_bounds__(kSpThreads, kSpCtasPerSm) k_spatial_persist(PcDev pc, PmDev pm, int rev, unsigned int* barrier_ctr) {
  extern __shared__ double s_dyn[];
  const EngineSmem sm = engine_smem<kSpRows>(s_dyn);   // chunked pipeline (generic window size)
  const RingSmem rs = ring_smem(s_dyn);                // ring pipeline (WND == 35); same LUT placement
  stage_luts(pc, sm.lut_w, sm.lut_c);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool consumer = warp == kSpThreads / 32 - 1;
  const size_t n = (size_t)pm.w * pm.h;
  const int inc = rev ? -1 : 1;
#if CSPM_SP_PROFILE
  long long prof_t = clock64();
#endif

  for (int diag = 1; diag <= pm.w + pm.h - 2; ++diag) {
    const int u_lo = max(0, diag - (pm.h - 1)), u_hi = min(diag, pm.w - 1);
    // Balanced CONTIGUOUS partition of the diagonal's pixels over the CTAs: neighbouring pixels share
    // 94 % of their windows and a CTA's range barely moves from one diagonal to the next, so the
    // window data stays in this SM's L1 (a round-robin assignment re-fetched everything from L2).
    const int npix = u_hi - u_lo + 1;
    const int pq = npix / (int)gridDim.x, pr = npix % (int)gridDim.x;
    const int pix0 = (int)blockIdx.x * pq + min((int)blockIdx.x, pr);
    const int cnt = pq + ((int)blockIdx.x < pr ? 1 : 0);
    const int ntasks_mine = 2 * cnt;  // task = pixel of one view
    // Dataflow synchronisation instead of a grid-wide barrier: my pixels only read the final planes of
    // their x/y neighbours on the previous diagonal, which belong to this CTA or its partition
    // neighbours.  Wait until exactly those CTAs have published that diagonal.
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
          CSPM_SP_POLL(seen, barrier_ctr + 32 * o);
        } while ((int)seen < diag - 1);
      }
#if !CSPM_SP_LEANFENCE
      __threadfence();
#endif
    }
    __syncthreads();
    for (int base = 0; base < ntasks_mine; base += kSpTasks) {
      const int nt = min(kSpTasks, ntasks_mine - base);
      // consumer lane L <-> (task L>>1, candidate L&1): x-axis neighbour, y-axis neighbour
      PlaneD ch_plane = {};
      double ch_min = 0.0;
      size_t ch_idx = 0;
      bool ch_on = false;
      int ch_v = 0, ch_x = 0, ch_y = 0;
      const int ti_c = lane >> 1, cc = lane & 1;
      if (consumer && ti_c < nt) {
        const int tid = 2 * pix0 + base + ti_c;
        const int u = u_lo + (tid >> 1), t = diag - u;
        ch_v = tid & 1;
        ch_x = rev ? pm.w - 1 - u : u;
        ch_y = rev ? pm.h - 1 - t : t;
        ch_idx = (size_t)ch_y * pm.w + ch_x;
        ch_on = cc == 0 ? u >= 1 : t >= 1;
        const double* st = pm.st[ch_v];
        if (ch_on) {
          // neighbour planes were written by another SM on the previous diagonal: bypass L1
          const size_t j = cc == 0 ? ch_idx - inc : ch_idx - (ptrdiff_t)inc * pm.w;
          ch_plane.nx = __ldcg(st + j); ch_plane.ny = __ldcg(st + n + j); ch_plane.nz = __ldcg(st + 2 * n + j);
          ch_plane.a = __ldcg(st + 3 * n + j); ch_plane.b = __ldcg(st + 4 * n + j); ch_plane.c = __ldcg(st + 5 * n + j);
        }
        if (cc == 0) ch_min = __ldcg(st + 6 * n + ch_idx);
#if CSPM_SPATIAL_DEDUP
        // Exact shortcut 1: a candidate whose six plane numbers are bitwise the pixel's own re-evaluates to
        // exactly min_cost (the stored cost is always GetPlaneCost of the stored plane), and `cost < min_cost`
        // (cs_patchmatch.cc:201,209) is false: not evaluated at all.
        if (ch_on && pm.cost_consistent) {
          const bool own = bits_eq(ch_plane.a, __ldcg(st + 3 * n + ch_idx)) && bits_eq(ch_plane.b, __ldcg(st + 4 * n + ch_idx)) &&
                           bits_eq(ch_plane.c, __ldcg(st + 5 * n + ch_idx)) && bits_eq(ch_plane.nx, __ldcg(st + ch_idx)) &&
                           bits_eq(ch_plane.ny, __ldcg(st + n + ch_idx)) && bits_eq(ch_plane.nz, __ldcg(st + 2 * n + ch_idx));
          if (own) ch_on = false;
        }
#endif
      }
#if CSPM_SPATIAL_DEDUP
      // Exact shortcut 2: the y-neighbour carries the same plane as the x-neighbour (common once a plane has
      // swept a region).  Its cost equals the x-candidate's: if that one was accepted the minimum IS that
      // cost, otherwise it was >= the minimum — either way `cost < min_cost` fails, so it is not evaluated.
      if (consumer) {
        // (every lane executes every shuffle: no short-circuit evaluation here)
        bool same = ch_on & (__shfl_xor_sync(0xffffffffu, (int)ch_on, 1) != 0);
        same &= bits_eq(ch_plane.a, __shfl_xor_sync(0xffffffffu, ch_plane.a, 1));
        same &= bits_eq(ch_plane.b, __shfl_xor_sync(0xffffffffu, ch_plane.b, 1));
        same &= bits_eq(ch_plane.c, __shfl_xor_sync(0xffffffffu, ch_plane.c, 1));
        same &= bits_eq(ch_plane.nx, __shfl_xor_sync(0xffffffffu, ch_plane.nx, 1));
        same &= bits_eq(ch_plane.ny, __shfl_xor_sync(0xffffffffu, ch_plane.ny, 1));
        same &= bits_eq(ch_plane.nz, __shfl_xor_sync(0xffffffffu, ch_plane.nz, 1));
        if (cc == 1 && same) ch_on = false;
      }
#endif
      // both candidates must beat the pixel's cost at the start of the step (cs_patchmatch.cc:201,209)
      // (negative scale weights, reg_lambda < 0: sums are not monotone, no early termination)
      const double ch_bound = pc.early_term ? __shfl_sync(0xffffffffu, ch_min, lane & ~1) : CUDART_INF;
      SP_PROF(0);
      const double ch_cost = WND == 35
          ? engine_eval_pairs_ring<CC, kSpThreads>(pc, rs, 2 * nt, ch_on, ch_v, ch_x, ch_y, ch_plane, ch_bound)
          : engine_eval_pairs<CC, WND, kSpRows, kSpThreads>(pc, sm, 2 * nt, ch_on, ch_v, ch_x, ch_y, ch_plane, ch_bound);
      SP_PROF(1);
      // acceptance in the reference's order: x-axis neighbour, then y-axis neighbour (:198-212)
      if (consumer) {
        const double c1 = __shfl_down_sync(0xffffffffu, ch_cost, 1);
        const bool on1 = __shfl_down_sync(0xffffffffu, (int)ch_on, 1) != 0;
        int best = -1;
        double min_cost = ch_min;
        if (cc == 0 && ti_c < nt) {
          if (ch_on && ch_cost < min_cost) { min_cost = ch_cost; best = 0; }
          if (on1 && c1 < min_cost) { min_cost = c1; best = 1; }
        }
        const int best_up = __shfl_up_sync(0xffffffffu, best, 1);
        const int my_best = cc == 0 ? best : best_up;
        if (ti_c < nt && my_best == cc) {  // the lane that owns the winning candidate writes it
          store_plane(pm.st[ch_v], n, ch_idx, ch_plane, ch_cost);
#if !CSPM_SP_LEANFENCE
          __threadfence();
#endif
        }
      }
      SP_PROF(2);
    }
    // publish: every writer fenced its stores; one release store per CTA and diagonal
    __syncthreads();
    if (threadIdx.x == 0) {
#if !CSPM_SP_LEANFENCE
      __threadfence();
#endif
      // release is cumulative: the winners' plane stores happen-before it through the CTA barrier above
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(barrier_ctr + 32 * blockIdx.x), "r"((unsigned int)diag) : "memory");
    }
    SP_PROF(3);
  }
}

