// oracle/ref_wrap.cc — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// C-ABI veneer over the reference's own classes (CSPatchMatch, PreSSPC,
// PreCSPC, GrdCC, CenCC) so that Python tests can drive the UNMODIFIED
// reference sources, compiled where they lie under /root/reference against
// oracle/cvshim.  Linked into oracle/_ref/libcspm_ref.so (oracle/Makefile).
// No reference source is copied: this file only #includes the headers.
#include <fcntl.h>
#include <unistd.h>

#include <opencv2/opencv.hpp>  // the shim; pulls every std header first

#define private public  // read plane_/min_cost_/dis_ and call the stage functions
#include "cs_patchmatch.h"
#include "plane_cost/pre_ss_pc.h"
#include "plane_cost/pre_cs_pc.h"
#include "plane_cost/grd_pc.h"
#include "plane_cost/cspc.h"
#include "cc/grd_cc.h"
#include "cc/cen_cc.h"
#undef private

namespace {
// The reference prints progress on stdout (cs_patchmatch.cc:53,66,...;
// grd_cc.cpp:86); silence it so harness output stays parseable.
struct Quiet {
  int saved;
  Quiet() {
    fflush(stdout); std::cout.flush();
    saved = dup(1);
    int nul = open("/dev/null", O_WRONLY);
    dup2(nul, 1);
    close(nul);
  }
  ~Quiet() {
    fflush(stdout); std::cout.flush();
    dup2(saved, 1);
    close(saved);
  }
};
CCMethod* get_cc(const char* name) {  // main.cc:39-55
  Quiet q;
  std::string n(name);
  if (n == "GRD") return new GrdCC();
  if (n == "CEN") return new CenCC();
  return NULL;
}
int n_ref_of(int max_dis) {
  int n = 0;
  for (double z = max_dis / 2.0; z >= 0.1; z /= 2.0) ++n;
  return n;
}
}  // namespace

extern "C" {

void ref_set_seed(uint64_t seed) { g_cspm_shim_seed = seed; }
void ref_set_time_mode(int mode) { g_cspm_shim_time_mode = mode; }

int ref_cc_build_cv(const char* name, const double* l_rgb, const double* r_rgb, int w, int h,
                    int max_dis, int right_view, double* vol) {
  CCMethod* cc = get_cc(name);
  if (!cc) return -1;
  Quiet q;
  Mat l(h, w, CV_64FC3, (void*)l_rgb, (size_t)w * 24), r(h, w, CV_64FC3, (void*)r_rgb, (size_t)w * 24);
  std::vector<Mat> cv(max_dis);
  for (int d = 0; d < max_dis; ++d) cv[d] = Mat::zeros(h, w, CV_64FC1);
  if (right_view) cc->buildRightCV(l, r, max_dis, cv.data());
  else cc->buildCV(l, r, max_dis, cv.data());
  for (int d = 0; d < max_dis; ++d)
    for (int y = 0; y < h; ++y)
      memcpy(vol + ((size_t)d * h + y) * w, cv[d].ptr<double>(y), sizeof(double) * w);
  delete cc;
  return 0;
}

struct RefPC {
  IPlaneCost* pc;
  PreSSPC* ss;
  PreCSPC* cs;
  GrdPC* gp;
  CSPC* cp;
};

void* ref_pc_create(const uint8_t* l, const uint8_t* r, int w, int h, int stride, int max_disp,
                    int wnd_size, int scale_num, double reg_lambda, const char* cc_name) {
  if (std::string(cc_name) == "IMG") {  // GrdPC / CSPC take no CCMethod (grd_pc.h:20-22, cspc.h:21-23)
    Quiet q;
    Mat lm(h, w, CV_8UC3, (void*)l, (size_t)stride), rm(h, w, CV_8UC3, (void*)r, (size_t)stride);
    RefPC* p = new RefPC{NULL, NULL, NULL, NULL, NULL};
    if (scale_num == 0) p->pc = p->gp = new GrdPC(lm, rm, max_disp, wnd_size);
    else p->pc = p->cp = new CSPC(lm, rm, max_disp, wnd_size, scale_num, reg_lambda);
    return p;
  }
  CCMethod* cc = get_cc(cc_name);
  if (!cc) return NULL;
  Quiet q;
  Mat lm(h, w, CV_8UC3, (void*)l, (size_t)stride), rm(h, w, CV_8UC3, (void*)r, (size_t)stride);
  RefPC* p = new RefPC{NULL, NULL, NULL, NULL, NULL};
  if (scale_num == 0) p->pc = p->ss = new PreSSPC(lm, rm, max_disp, wnd_size, cc);
  else p->pc = p->cs = new PreCSPC(lm, rm, max_disp, wnd_size, scale_num, cc, reg_lambda);
  delete cc;
  return p;
}
void ref_pc_destroy(void* p_) {
  RefPC* p = (RefPC*)p_;
  delete p->pc;
  delete p;
}
void ref_pc_get_plane_cost(const void* p_, int n, const int* x, const int* y, const double* normal,
                           const double* point, const int* view, double* out, int nthreads) {
  const RefPC* p = (const RefPC*)p_;
  if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 64)
  for (int i = 0; i < n; ++i) {
    Plane pl(Vec3d(normal[3 * i], normal[3 * i + 1], normal[3 * i + 2]),
             Point3d(point[3 * i], point[3 * i + 1], point[3 * i + 2]));
    out[i] = p->pc->GetPlaneCost(x[i], y[i], pl, (RefView)view[i]);
  }
}
int ref_pc_info(const void* p_, double* scale_wgt, double* max_cost, int* dims) {
  const RefPC* p = (const RefPC*)p_;
  if (p->ss) {
    if (scale_wgt) scale_wgt[0] = 1.0;
    if (max_cost) { max_cost[0] = p->ss->max_cost_[0]; max_cost[1] = p->ss->max_cost_[1]; }
    if (dims) { dims[0] = p->ss->wid_; dims[1] = p->ss->hei_; dims[2] = p->ss->max_disp_; }
    return 1;
  }
  if (p->gp) {
    if (scale_wgt) scale_wgt[0] = 1.0;
    if (max_cost) { max_cost[0] = 2.8; max_cost[1] = 2.8; }
    if (dims) { dims[0] = p->gp->wid_; dims[1] = p->gp->hei_; dims[2] = p->gp->max_disp_; }
    return 1;
  }
  if (p->cp) {
    int S = p->cp->scale_num_;
    for (int s = 0; s < S; ++s) {
      if (scale_wgt) scale_wgt[s] = p->cp->scale_wgt_[s];
      if (max_cost) { max_cost[s] = 2.8; max_cost[S + s] = 2.8; }
      if (dims) { dims[3 * s] = p->cp->wid_[s]; dims[3 * s + 1] = p->cp->hei_[s]; dims[3 * s + 2] = p->cp->max_disp_[s]; }
    }
    return S;
  }
  int S = p->cs->scale_num_;
  for (int s = 0; s < S; ++s) {
    if (scale_wgt) scale_wgt[s] = p->cs->scale_wgt_[s];
    if (max_cost) { max_cost[s] = p->cs->max_cost_[0][s]; max_cost[S + s] = p->cs->max_cost_[1][s]; }
    if (dims) { dims[3 * s] = p->cs->wid_[s]; dims[3 * s + 1] = p->cs->hei_[s]; dims[3 * s + 2] = p->cs->max_disp_[s]; }
  }
  return S;
}
int ref_pc_scale_image(const void* p_, int s, int view, uint8_t* out) {
  const RefPC* p = (const RefPC*)p_;
  const Mat& m = p->ss ? p->ss->img_[view] : p->gp ? p->gp->img_[view] : p->cp ? p->cp->img_[view][s] : p->cs->img_[view][s];
  for (int y = 0; y < m.rows; ++y) memcpy(out + (size_t)y * 3 * m.cols, m.ptr<uchar>(y), (size_t)3 * m.cols);
  return 0;
}

void* ref_pm_create(const uint8_t* l, const uint8_t* r, int w, int h, int stride, int max_dis,
                    int dis_scale, uint64_t seed) {
  g_cspm_shim_seed = seed;
  g_cspm_shim_rows = (uint64_t)h;
  g_cspm_shim_time_calls = 0;  // the ctor's time(NULL) is call 0 (cs_patchmatch.cc:32)
  Mat lm(h, w, CV_8UC3, (void*)l, (size_t)stride), rm(h, w, CV_8UC3, (void*)r, (size_t)stride);
  return new CSPatchMatch(lm, rm, max_dis, dis_scale);
}
void ref_pm_destroy(void* pm) { delete (CSPatchMatch*)pm; }
void ref_pm_patchmatch(void* pm_, int iter_num, const void* pc_, int use_pp, int nthreads) {
  CSPatchMatch* pm = (CSPatchMatch*)pm_;
  // time-mode 0 (call counter) is only deterministic single-threaded
  omp_set_num_threads(g_cspm_shim_time_mode == 0 ? 1 : (nthreads > 0 ? nthreads : omp_get_num_procs()));
  g_cspm_shim_time_calls = 1;
  ++g_cspm_shim_epoch;
  Quiet q;
  pm->PatchMatch(iter_num, ((const RefPC*)pc_)->pc, use_pp != 0);
}
// 0 init, 1 spatial, 2 view, 3 refine, 4 plane2disp, 5 postprocess
void ref_pm_stage(void* pm_, int stage, int iter, const void* pc_, int nthreads) {
  CSPatchMatch* pm = (CSPatchMatch*)pm_;
  const IPlaneCost* pc = pc_ ? ((const RefPC*)pc_)->pc : NULL;
  omp_set_num_threads(g_cspm_shim_time_mode == 0 ? 1 : (nthreads > 0 ? nthreads : omp_get_num_procs()));
  const uint64_t H = pm->hei_;
  ++g_cspm_shim_epoch;
  Quiet q;
  switch (stage) {
    case 0: g_cspm_shim_time_calls = 1; pm->InitRandomPlane(pc); break;
    case 1: pm->SpatialPropagation(iter, pc); break;
    case 2: pm->ViewPropagation(iter, pc); break;
    case 3:
      g_cspm_shim_time_calls = 1 + 2 * H + (uint64_t)iter * n_ref_of(pm->max_dis_) * 2 * H;
      pm->PlaneRefinement(pm->max_dis_ / 2.0, pm->kMaxNorm_, pm->kZStopThres_, pc);
      break;
    case 4: pm->PlaneToDisp(); break;
    case 5: pm->PostProcessing(); break;
  }
}
void ref_pm_get_dis(const void* pm_, int view, uint8_t* out) {
  CSPatchMatch* pm = (CSPatchMatch*)pm_;
  const Mat& m = pm->dis_[view];
  for (int y = 0; y < m.rows; ++y) memcpy(out + (size_t)y * m.cols, m.ptr<uchar>(y), m.cols);
}
void ref_pm_set_dis(void* pm_, int view, const uint8_t* in) {
  CSPatchMatch* pm = (CSPatchMatch*)pm_;
  Mat& m = pm->dis_[view];
  for (int y = 0; y < m.rows; ++y) memcpy(m.ptr<uchar>(y), in + (size_t)y * m.cols, m.cols);
}
void ref_pm_get_planes(const void* pm_, int view, double* normal, double* param, double* min_cost) {
  const CSPatchMatch* pm = (const CSPatchMatch*)pm_;
  for (int y = 0; y < pm->hei_; ++y)
    for (int x = 0; x < pm->wid_; ++x) {
      size_t i = (size_t)y * pm->wid_ + x;
      Vec3d n = pm->plane_[view][y][x].norm(), p = pm->plane_[view][y][x].param();
      for (int k = 0; k < 3; ++k) {
        if (normal) normal[3 * i + k] = n[k];
        if (param) param[3 * i + k] = p[k];
      }
      if (min_cost) min_cost[i] = pm->min_cost_[view][y][x];
    }
}
void ref_pm_set_planes(void* pm_, int view, const double* normal, const double* param, const double* min_cost) {
  CSPatchMatch* pm = (CSPatchMatch*)pm_;
  for (int y = 0; y < pm->hei_; ++y)
    for (int x = 0; x < pm->wid_; ++x) {
      size_t i = (size_t)y * pm->wid_ + x;
      Plane& pl = pm->plane_[view][y][x];
      pl.norm_ = Vec3d(normal[3 * i], normal[3 * i + 1], normal[3 * i + 2]);
      pl.param_ = Vec3d(param[3 * i], param[3 * i + 1], param[3 * i + 2]);
      pm->min_cost_[view][y][x] = min_cost[i];
    }
}

}  // extern "C"
