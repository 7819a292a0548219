/* cspm_c.h — C ABI of libcspm_b200.so, the B200-native drop-in for the
 * PatchMatch-stereo + cross-scale plane-cost hot path of
 * rookiepig/CrossScalePatchMatch (CSPM).
 *
 * The reference has no FFI: its plugin surface is four C++ interfaces linked
 * into one executable.  Each entry point below names the reference interface
 * it replaces (file:line relative to the reference repo).  INTEGRATION.md
 * shows the few-line adapter classes a CSPM maintainer would add
 * (include/cspm.hpp already provides them under the reference's own names).
 *
 * Conventions: plain pointers and sizes only; every handle owns one CUDA
 * device + one stream (chosen at create time) and all of its device memory;
 * images are 8-bit BGR interleaved (CV_8UC3), `stride` in bytes; functions
 * return 0 on success or a negative CSPM_E* code, creators return NULL, and
 * cspm_last_error() gives the message for the calling thread.  There is no
 * CPU fallback: without a CUDA device every compute call fails with
 * CSPM_ENODEV.
 *
 * RNG contract (replaces cv::RNG(time(NULL)), cs_patchmatch.cc:32,130,309,
 * which is unreproducible by construction): a 64-bit draw is
 * Philox4x32-10(key = seed, counter = (slot, sub, stream_lo, stream_hi)),
 * low two words.  `stream` stands for the reference's per-row RNG object:
 *   InitRandomPlane, view v, row y            : 1 + v*H + y
 *   PlaneRefinement, iter i, step k, view v, y: 1 + 2H + ((i*n_ref+k)*2+v)*H + y
 * (i.e. the running count of time(NULL) calls of a single-threaded reference
 * run); `slot` is the index of the rng.uniform()/rng.fill() call on that
 * object (2x and 2x+1 for pixel x), `sub` indexes draws inside one call.
 * uniform(a,b) = u*(b-a)+a with u = (draw>>11)*2^-53; fill(NORMAL) returns a
 * point uniform on the unit sphere (Marsaglia 1972) — its only consumer
 * normalises the vector (cs_patchmatch.cc:138-140).
 */
#ifndef CSPM_C_H_
#define CSPM_C_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSPM_OK 0
#define CSPM_EINVAL (-1)  /* bad argument (NULL, size, unknown name) */
#define CSPM_ENODEV (-2)  /* no usable CUDA device (there is no CPU fallback) */
#define CSPM_ENOMEM (-3)
#define CSPM_EUNSUPPORTED (-4)
#define CSPM_ECUDA (-5)   /* any other CUDA runtime failure (launch, copy, ...); the message names the call */

#define CSPM_VIEW_LEFT 0  /* RefView::kLeft,  commfunc.h:29 */
#define CSPM_VIEW_RIGHT 1 /* RefView::kRight */

/* Propagation schedules of SpatialPropagation (the other stages are the same in all of them).
 * RASTER reproduces the reference's serial Gauss-Seidel scan exactly (anti-diagonal wavefront): results are
 * bit-identical to the reference code.  REDBLACK / REDBLACK8 are the throughput schedules (not in the reference;
 * BASELINE.json's north_star names them): iteration i makes two checkerboard half-sweeps, colour (i + half) & 1
 * first; a pixel of the active colour tests, in this order, the planes of (x-1,y), (x+1,y), (x,y-1), (x,y+1) —
 * REDBLACK8 then also (x-5,y), (x+5,y), (x,y-5), (x,y+5) — against its running minimum with the strict `<`.
 * Bit-identical to the oracle's schedule 1 / 2 (oracle/oracle.cc SpatialPropagationRedBlack), not to the raster scan. */
#define CSPM_SCHEDULE_RASTER 0
#define CSPM_SCHEDULE_REDBLACK 1
#define CSPM_SCHEDULE_REDBLACK8 2

/* PatchMatch stages for cspm_pm_stage (parity tests drive them one by one). */
#define CSPM_STAGE_INIT 0       /* InitRandomPlane     cs_patchmatch.cc:115-148 */
#define CSPM_STAGE_SPATIAL 1    /* SpatialPropagation  cs_patchmatch.cc:163-216 */
#define CSPM_STAGE_VIEW 2       /* ViewPropagation     cs_patchmatch.cc:229-277 */
#define CSPM_STAGE_REFINE 3     /* PlaneRefinement     cs_patchmatch.cc:292-345 */
#define CSPM_STAGE_PLANE2DISP 4 /* PlaneToDisp         cs_patchmatch.cc:590-602 */
#define CSPM_STAGE_POSTPROC 5   /* PostProcessing      cs_patchmatch.cc:508-588 */

typedef struct cspm_cc cspm_cc_t; /* CCMethod      cc_method.h:15-33   */
typedef struct cspm_pc cspm_pc_t; /* IPlaneCost    plane_cost/i_plane_cost.h:13-34 */
typedef struct cspm_pm cspm_pm_t; /* CSPatchMatch  cs_patchmatch.h:16-147 */

const char* cspm_last_error(void);
/* CSPM_E* code of the calling thread's last failed call (creators return NULL: this is how to tell EINVAL from
 * ENODEV / ENOMEM / EUNSUPPORTED). */
int cspm_last_error_code(void);
/* Number of visible CUDA devices (0 when there is none); never fails. */
int cspm_device_count(void);
const char* cspm_version(void);
/* Stream injection: handles created by the calling thread AFTER this call
 * launch on `cuda_stream` (a cudaStream_t; e.g. the host application's or
 * PyTorch's current stream) instead of creating a private stream; pass NULL to
 * go back to private streams.  The stream must outlive those handles. */
void cspm_set_thread_stream(void* cuda_stream);

/* ---- CCMethod ---------------------------------------------------------- */
/* GetCCType(name), main.cc:39-55: "GRD" -> GrdCC, "CEN" -> CenCC, else NULL. */
cspm_cc_t* cspm_cc_create(const char* name);
void cspm_cc_destroy(cspm_cc_t* cc);
const char* cspm_cc_name(const cspm_cc_t* cc);
/* CCMethod::buildCV / buildRightCV, cc_method.h:31-32 (GrdCC: cc/grd_cc.cpp:60-154).
 * l_rgb/r_rgb: host, H x W x 3 f64 RGB 0..255 (CV_64FC3); cost_vol: host,
 * max_dis planar slices [d][y][x] f64.  Parity / debug export: the fused
 * PatchMatch path never materialises the volume. */
int cspm_cc_build_cv(const cspm_cc_t* cc, const double* l_rgb, const double* r_rgb, int w, int h,
                     int max_dis, int right_view, double* cost_vol, int device);

/* ---- IPlaneCost -------------------------------------------------------- */
/* PreSSPC(l_img, r_img, max_disp, wnd_size, cc_method), plane_cost/pre_ss_pc.cc:12-65 */
cspm_pc_t* cspm_pc_create_ss(const uint8_t* l, const uint8_t* r, int w, int h, int stride,
                             int max_disp, int wnd_size, const cspm_cc_t* cc, int device);
/* PreCSPC(l_img, r_img, max_disp, wnd_size, scale_num, cc_method, reg_lambda),
 * plane_cost/pre_cs_pc.cc:12-115.  scale_num in [2, 8]. */
cspm_pc_t* cspm_pc_create_cs(const uint8_t* l, const uint8_t* r, int w, int h, int stride,
                             int max_disp, int wnd_size, int scale_num, const cspm_cc_t* cc,
                             double reg_lambda, int device);
/* Same, but l/r are DEVICE pointers on `device` (inputs already resident in HBM). */
cspm_pc_t* cspm_pc_create_ss_dev(const uint8_t* d_l, const uint8_t* d_r, int w, int h, int stride,
                                 int max_disp, int wnd_size, const cspm_cc_t* cc, int device);
cspm_pc_t* cspm_pc_create_cs_dev(const uint8_t* d_l, const uint8_t* d_r, int w, int h, int stride,
                                 int max_disp, int wnd_size, int scale_num, const cspm_cc_t* cc,
                                 double reg_lambda, int device);
/* GrdPC(l_img, r_img, max_disp, wnd_size), plane_cost/grd_pc.cc:12-66, and
 * CSPC(l_img, r_img, max_disp, wnd_size, scale_num, reg_lambda), plane_cost/cspc.cc:3-92: the
 * alternative IPlaneCost plugins that interpolate the partner IMAGE at the sub-pixel position instead of
 * reading a cost volume (no CCMethod involved).  Require max_disp <= w (single wrap-around). */
cspm_pc_t* cspm_pc_create_grdpc(const uint8_t* l, const uint8_t* r, int w, int h, int stride,
                                int max_disp, int wnd_size, int device);
cspm_pc_t* cspm_pc_create_cspc(const uint8_t* l, const uint8_t* r, int w, int h, int stride,
                               int max_disp, int wnd_size, int scale_num, double reg_lambda, int device);
void cspm_pc_destroy(cspm_pc_t* pc);
/* Batched IPlaneCost::GetPlaneCost(ref_x, ref_y, plane, view)
 * (pre_ss_pc.cc:74-118 / pre_cs_pc.cc:133-188): query i is the plane with
 * unit normal normal_xyz[3i..] through point_xyz[3i..], evaluated at pixel
 * (x[i], y[i]) of view[i].  All pointers are host pointers. */
int cspm_pc_get_plane_cost(const cspm_pc_t* pc, int n, const int* x, const int* y,
                           const double* normal_xyz, const double* point_xyz, const int* view,
                           double* out_cost);
/* scale_num; scale_wgt[S] (pre_cs_pc.cc:105-109); max_cost[2][S]
 * (pre_ss_pc.cc:51-58); dims[S][3] = (w, h, max_disp) per scale.  Any output
 * may be NULL.  Returns S. */
int cspm_pc_info(const cspm_pc_t* pc, double* scale_wgt, double* max_cost, int* dims);
/* Pyramid image of scale s (pre_cs_pc.cc:45), dense BGR, host pointer. */
int cspm_pc_scale_image(const cspm_pc_t* pc, int s, int view, uint8_t* out);

/* ---- CSPatchMatch ------------------------------------------------------ */
/* CSPatchMatch(l_img, r_img, max_dis, dis_scale), cs_patchmatch.cc:3-34.  Requires 1 <= max_dis <= w:
 * ViewPropagation wraps x +- round(d) once (HandleBorder, commfunc.h:129-145) and the reference runs
 * past its row arrays beyond that. */
cspm_pm_t* cspm_pm_create(const uint8_t* l, const uint8_t* r, int w, int h, int stride, int max_dis,
                          int dis_scale, uint64_t seed, int schedule, int device);
/* Same, but l/r are DEVICE pointers on `device`. */
cspm_pm_t* cspm_pm_create_dev(const uint8_t* d_l, const uint8_t* d_r, int w, int h, int stride,
                              int max_dis, int dis_scale, uint64_t seed, int schedule, int device);
void cspm_pm_destroy(cspm_pm_t* pm);
/* PatchMatch(iter_num, plane_cost, use_pp), cs_patchmatch.cc:51-109.  pc must
 * live on the same device.  Synchronous: returns when the disparity maps are
 * final in device memory. */
int cspm_pm_patchmatch(cspm_pm_t* pm, int iter_num, const cspm_pc_t* pc, int use_pp);
/* The same for n independent stereo pairs of identical geometry (image size, schedule; plane costs of the same kind,
 * window, scales, reg_lambda and max_disp) on one device: results are exactly those of n separate
 * cspm_pm_patchmatch calls.  The pairs' SpatialPropagation sweeps share one cooperative launch and advance
 * diagonal by diagonal in lockstep, so the per-diagonal latency that dominates small images is paid once per
 * batch (throughput mode; the reference processes one pair per process, main.cc:57-139).  The call runs on the
 * first pair's stream; counters of the shared sweep and the stage times of the whole batch are reported on pm[0]. */
int cspm_pm_patchmatch_batch(cspm_pm_t* const* pm, const cspm_pc_t* const* pc, int n, int iter_num, int use_pp);
/* One stage of the loop above (CSPM_STAGE_*); `iter` selects scan direction
 * and RNG stream exactly as the loop index i does in cs_patchmatch.cc:65-102. */
int cspm_pm_stage(cspm_pm_t* pm, int stage, int iter, const cspm_pc_t* pc);
/* dis(view), cs_patchmatch.cc:111-113: H x W uint8 to a host buffer. */
int cspm_pm_get_dis(const cspm_pm_t* pm, int view, uint8_t* out, int out_stride);
int cspm_pm_set_dis(cspm_pm_t* pm, int view, const uint8_t* in, int in_stride);
/* Parity export / import of the per-pixel state: normal[h*w*3], param[h*w*3]
 * (a,b,c), min_cost[h*w]; host pointers, any may be NULL on get.
 * The imported costs need not belong to the imported planes: results stay those
 * of the reference loops on that state.  (The spatial sweep's "candidate equals
 * the pixel's own plane, so it cannot win" shortcut is only taken while every
 * stored cost is known to be GetPlaneCost of its plane under the plane cost of
 * the current call — i.e. from InitRandomPlane on, with one plane cost — and
 * set_planes / a different plane cost switch it off.) */
int cspm_pm_get_planes(const cspm_pm_t* pm, int view, double* normal, double* param, double* min_cost);
int cspm_pm_set_planes(cspm_pm_t* pm, int view, const double* normal, const double* param,
                       const double* min_cost);
/* Number of refinement halvings (cs_patchmatch.cc:299-301,342-343). */
int cspm_pm_n_ref(const cspm_pm_t* pm);

/* ---- measurement hooks -------------------------------------------------- */
/* The cudaStream_t (as void*) the handle launches on, for CUDA-event timing. */
void* cspm_pc_stream(const cspm_pc_t* pc);
void* cspm_pm_stream(const cspm_pm_t* pm);
/* Counters since handle creation: kernels launched by this library on the
 * handle's stream, plane evaluations and in-image window taps issued. */
typedef struct cspm_stats {
  uint64_t kernel_launches;
  uint64_t plane_evals;
  uint64_t window_taps; /* nominal: every in-image tap of every evaluation, whether or not early termination cut it */
  /* SpatialPropagation sweeps (wavefront kernel): pixel tasks, candidate planes evaluated (after the exact
   * duplicate shortcuts), planes adopted from certified comparisons, tasks that needed the exact serial path */
  uint64_t spatial_tasks;
  uint64_t spatial_candidates;
  uint64_t spatial_adopted;
  uint64_t spatial_exact_tasks;
} cspm_stats_t;
int cspm_pm_stats(const cspm_pm_t* pm, cspm_stats_t* out);
int cspm_pc_stats(const cspm_pc_t* pc, cspm_stats_t* out);
/* In-image (border-clipped) window taps of ONE GetPlaneCost evaluation at every
 * pixel of one view, summed over scales: the unit count behind the roofline. */
uint64_t cspm_pc_taps_per_image_eval(const cspm_pc_t* pc);
/* Device milliseconds spent in each CSPM_STAGE_* of the last cspm_pm_patchmatch
 * call (CUDA events on the handle's stream); ms[6]. */
int cspm_pm_stage_ms(const cspm_pm_t* pm, float* ms);
/* Diagnostic: SM-cycle totals spent by CTA 0 of the persistent spatial sweep in
 * {task setup, per-scale prep, chunk pipeline, accept, grid barrier, -, -, -}
 * on the current device since the last reset; out8[8]. */
int cspm_debug_counters(unsigned long long* out8, int reset);
/* Diagnostic builds (-DCSPM_COUNT_TAPS=1) only, zero otherwise: window taps the bulk kernels (init, view,
 * refinement, cost fix-up) actually EXECUTED on the current device since the last reset — the nominal count
 * minus what exact early termination skipped. */
int cspm_debug_taps_executed(unsigned long long* out, int reset);
/* Diagnostic builds (-DCSPM_WAVE_PROFILE=1) only (n = 0 otherwise): {T0, T1, m, flags} of the first <= 64 tasks one
 * CTA of the wavefront sweep could not decide from its tree sums (n = how many it met); out256[64 * 4]. */
int cspm_debug_wave_samples(double* out256, unsigned int* n);
/* Diagnostic builds (-DCSPM_PROFILE_LANES=1) only, zeros otherwise: out33[k] = window rows the bulk kernels
 * executed with k of a warp's 32 lanes still evaluating (per-lane early termination leaves the others idle). */
int cspm_debug_lane_hist(unsigned long long* out33, int reset);

#ifdef __cplusplus
}
#endif
#endif /* CSPM_C_H_ */
