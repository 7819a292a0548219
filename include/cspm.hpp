// cspm.hpp — header-only C++ adapters that re-expose the reference's plugin
// surface (same class and method names, argument order and error behaviour)
// on top of the C ABI in cspm_c.h.  A CSPM maintainer can include this header
// instead of cc/grd_cc.h, plane_cost/pre_ss_pc.h, plane_cost/pre_cs_pc.h and
// cs_patchmatch.h and keep main.cc:89-134 unchanged apart from the Mat type.
//
//   reference                                     here
//   CCMethod / GrdCC / CenCC  (cc_method.h:15-33) cspm::CCMethod / GrdCC / CenCC
//   CAMethod                  (ca_method.h:8-25)  cspm::CAMethod (interface only:
//                                                 nothing implements or calls it
//                                                 in the reference build)
//   IPlaneCost / PreSSPC / PreCSPC                cspm::IPlaneCost / PreSSPC / PreCSPC
//   Plane / RefView           (plane.h, commfunc.h:29)
//   CSPatchMatch              (cs_patchmatch.h:16-147)
//
// OpenCV is not a dependency: `cspm::Mat` is a minimal non-owning/owning view
// with the few members the call sites use (rows, cols, data, step, type()).
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "cspm_c.h"

namespace cspm {

enum RefView { kLeft = 0, kRight = 1 };  // commfunc.h:29
enum { CV_8UC1 = 0, CV_8UC3 = 16, CV_64FC1 = 6, CV_64FC3 = 22 };

struct Exception : std::runtime_error {  // stands in for cv::Exception (CV_Assert)
  int code;
  Exception(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
inline void check(int rc) {
  if (rc < 0) throw Exception(rc, cspm_last_error());
}
// a creator returned NULL: throw with the code the library recorded (EINVAL / ENODEV / ENOMEM / EUNSUPPORTED / ECUDA)
[[noreturn]] inline void throw_last() {
  const int code = cspm_last_error_code();
  throw Exception(code < 0 ? code : CSPM_EINVAL, cspm_last_error());
}

// Minimal cv::Mat stand-in.
class Mat {
 public:
  int rows = 0, cols = 0;
  uint8_t* data = nullptr;
  size_t step = 0;
  Mat() {}
  Mat(int r, int c, int type) : rows(r), cols(c), type_(type) {
    step = (size_t)c * elem();
    own_.assign((size_t)r * step, 0);
    data = own_.data();
  }
  Mat(int r, int c, int type, void* ext, size_t step_bytes) : rows(r), cols(c), data((uint8_t*)ext), step(step_bytes), type_(type) {}
  static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
  int type() const { return type_; }
  template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step); }
  template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step); }

 private:
  size_t elem() const { return type_ == CV_8UC1 ? 1 : type_ == CV_8UC3 ? 3 : type_ == CV_64FC1 ? 8 : 24; }
  int type_ = CV_8UC1;
  std::vector<uint8_t> own_;
};

// CV_Assert(l_img.type() == CV_8UC3 && r_img.type() == CV_8UC3) (pre_ss_pc.cc:24, cs_patchmatch.cc:8) — plus the
// geometry the reference takes for granted (it indexes the right image with the left image's sizes): one
// rows / cols / step for both images, because the C ABI receives a single (w, h, stride).
inline void require_pair_8uc3(const class Mat& l, const class Mat& r);

// plane.h:12-48
class Plane {
 public:
  Plane() {}
  Plane(const double norm[3], const double point[3]) {
    std::memcpy(norm_, norm, sizeof norm_);
    std::memcpy(point_, point, sizeof point_);
  }
  void set_point(double x, double y, double z) { point_[0] = x; point_[1] = y; point_[2] = z; }
  void set_norm(double x, double y, double z) { norm_[0] = x; norm_[1] = y; norm_[2] = z; }
  const double* norm() const { return norm_; }
  const double* point() const { return point_; }

 private:
  double norm_[3] = {0, 0, 0}, point_[3] = {0, 0, 0};
};

// cc_method.h:15-33
class CCMethod {
 public:
  virtual ~CCMethod() { cspm_cc_destroy(h_); }
  // costVol: maxDis Mats of H x W CV_64FC1, caller-allocated (pre_ss_pc.cc:40-43)
  virtual void buildCV(const Mat& lImg, const Mat& rImg, const int maxDis, Mat* costVol) { build(lImg, rImg, maxDis, costVol, 0); }
  virtual void buildRightCV(const Mat& lImg, const Mat& rImg, const int maxDis, Mat* rCostVol) { build(lImg, rImg, maxDis, rCostVol, 1); }
  cspm_cc_t* handle() const { return h_; }
  int device = 0;

 protected:
  explicit CCMethod(const char* name) : h_(cspm_cc_create(name)) {
    if (!h_) throw_last();
  }

 private:
  void build(const Mat& l, const Mat& r, int maxDis, Mat* vol, int right) {
    if (l.type() != CV_64FC3 || r.type() != CV_64FC3)  // CV_Assert, grd_cc.cpp:63
      throw Exception(CSPM_EINVAL, "CV_Assert(lImg.type() == CV_64FC3 && rImg.type() == CV_64FC3)");
    const size_t n = (size_t)l.rows * l.cols;
    std::vector<double> li(3 * n), ri(3 * n), out(n * maxDis);
    for (int y = 0; y < l.rows; ++y) {
      std::memcpy(&li[(size_t)y * 3 * l.cols], l.ptr<double>(y), sizeof(double) * 3 * l.cols);
      std::memcpy(&ri[(size_t)y * 3 * l.cols], r.ptr<double>(y), sizeof(double) * 3 * l.cols);
    }
    check(cspm_cc_build_cv(h_, li.data(), ri.data(), l.cols, l.rows, maxDis, right, out.data(), device));
    for (int d = 0; d < maxDis; ++d)
      for (int y = 0; y < l.rows; ++y)
        std::memcpy(vol[d].ptr<double>(y), &out[((size_t)d * l.rows + y) * l.cols], sizeof(double) * l.cols);
  }
  cspm_cc_t* h_;
};
class GrdCC : public CCMethod {  // cc/grd_cc.h:14-26
 public:
  GrdCC() : CCMethod("GRD") {}
};
class CenCC : public CCMethod {  // cc/cen_cc.h:11-23
 public:
  CenCC() : CCMethod("CEN") {}
};
// main.cc:39-55 — NULL for unknown names, like the reference
inline CCMethod* GetCCType(const std::string& name) {
  if (name == "GRD") return new GrdCC();
  if (name == "CEN") return new CenCC();
  return nullptr;
}

// ca_method.h:8-25 — kept so the surface is complete; no implementation exists
// in the reference build (ca_filter/* does not compile there).
class CAMethod {
 public:
  virtual ~CAMethod() {}
  virtual void aggreCV(const Mat& lImg, const Mat& rImg, const int maxDis, Mat* costVol) = 0;
};

// plane_cost/i_plane_cost.h:13-34
class IPlaneCost {
 public:
  virtual ~IPlaneCost() { cspm_pc_destroy(h_); }
  virtual double GetPlaneCost(const int& ref_x, const int& ref_y, const Plane& plane, const RefView& view) const {
    const int v = view;
    double out = 0.0;
    check(cspm_pc_get_plane_cost(h_, 1, &ref_x, &ref_y, plane.norm(), plane.point(), &v, &out));
    return out;
  }
  cspm_pc_t* handle() const { return h_; }

 protected:
  IPlaneCost() {}
  static void require_8uc3(const Mat& l, const Mat& r) { require_pair_8uc3(l, r); }
  cspm_pc_t* h_ = nullptr;
};
inline void require_pair_8uc3(const Mat& l, const Mat& r) {
  if (l.type() != CV_8UC3 || r.type() != CV_8UC3)
    throw Exception(CSPM_EINVAL, "CV_Assert(l_img.type() == CV_8UC3 && r_img.type() == CV_8UC3)");
  if (!l.data || !r.data || l.rows != r.rows || l.cols != r.cols || l.step != r.step)
    throw Exception(CSPM_EINVAL, "left and right image must have the same rows, cols and step");
}
class PreSSPC : public IPlaneCost {  // plane_cost/pre_ss_pc.h:18-52
 public:
  PreSSPC(const Mat& l_img, const Mat& r_img, const int& max_disp, const int& wnd_size, CCMethod* cc_method, int device = 0) {
    require_8uc3(l_img, r_img);
    h_ = cspm_pc_create_ss(l_img.data, r_img.data, l_img.cols, l_img.rows, (int)l_img.step, max_disp, wnd_size,
                           cc_method ? cc_method->handle() : nullptr, device);
    if (!h_) throw_last();
  }
};
class PreCSPC : public IPlaneCost {  // plane_cost/pre_cs_pc.h:19-62
 public:
  PreCSPC(const Mat& l_img, const Mat& r_img, const int& max_disp, const int& wnd_size, const int& scale_num,
          CCMethod* cc_method, const double& reg_lambda, int device = 0) {
    require_8uc3(l_img, r_img);
    h_ = cspm_pc_create_cs(l_img.data, r_img.data, l_img.cols, l_img.rows, (int)l_img.step, max_disp, wnd_size,
                           scale_num, cc_method ? cc_method->handle() : nullptr, reg_lambda, device);
    if (!h_) throw_last();
  }
};

class GrdPC : public IPlaneCost {  // plane_cost/grd_pc.h:19-71
 public:
  GrdPC(const Mat& l_img, const Mat& r_img, const int& max_disp, const int& wnd_size, int device = 0) {
    require_8uc3(l_img, r_img);
    h_ = cspm_pc_create_grdpc(l_img.data, r_img.data, l_img.cols, l_img.rows, (int)l_img.step, max_disp, wnd_size, device);
    if (!h_) throw_last();
  }
};
class CSPC : public IPlaneCost {  // plane_cost/cspc.h:17-62
 public:
  CSPC(const Mat& l_img, const Mat& r_img, const int& max_disp, const int& wnd_size, const int& scale_num,
       const double& reg_lambda, int device = 0) {
    require_8uc3(l_img, r_img);
    h_ = cspm_pc_create_cspc(l_img.data, r_img.data, l_img.cols, l_img.rows, (int)l_img.step, max_disp, wnd_size,
                             scale_num, reg_lambda, device);
    if (!h_) throw_last();
  }
};

// cs_patchmatch.h:16-147
class CSPatchMatch {
 public:
  CSPatchMatch(const Mat& l_img, const Mat& r_img, const int& max_dis, const int& dis_scale, uint64_t seed = 0x5EED,
               int schedule = CSPM_SCHEDULE_RASTER, int device = 0) {
    require_pair_8uc3(l_img, r_img);  // CV_Assert, cs_patchmatch.cc:8
    h_ = cspm_pm_create(l_img.data, r_img.data, l_img.cols, l_img.rows, (int)l_img.step, max_dis, dis_scale, seed,
                        schedule, device);
    if (!h_) throw_last();
    for (int v = 0; v < 2; ++v) dis_[v] = Mat::zeros(l_img.rows, l_img.cols, CV_8UC1);
  }
  ~CSPatchMatch() { cspm_pm_destroy(h_); }
  CSPatchMatch(const CSPatchMatch&) = delete;
  CSPatchMatch& operator=(const CSPatchMatch&) = delete;
  // The reference accepts any `const IPlaneCost*` and calls its virtual GetPlaneCost 43 x W x H x 2 times
  // (cs_patchmatch.h:65-66).  Here the plane cost is fused into the kernels, so only plane costs built by this
  // library (PreSSPC, PreCSPC, GrdPC, CSPC: handle() != NULL) can drive PatchMatch; a user-defined IPlaneCost
  // subclass is rejected with CSPM_EUNSUPPORTED instead of being evaluated on the host (no CPU fallback).
  void PatchMatch(const int& iter_num, const IPlaneCost* plane_cost, const bool& use_pp) {
    if (plane_cost && !plane_cost->handle())
      throw Exception(CSPM_EUNSUPPORTED, "PatchMatch needs a plane cost built by libcspm_b200 (PreSSPC / PreCSPC / GrdPC / CSPC); "
                                         "user-defined IPlaneCost subclasses cannot be fused into the kernels");
    check(cspm_pm_patchmatch(h_, iter_num, plane_cost ? plane_cost->handle() : nullptr, use_pp ? 1 : 0));
    for (int v = 0; v < 2; ++v) check(cspm_pm_get_dis(h_, v, dis_[v].data, (int)dis_[v].step));
  }
  Mat& dis(const RefView& view) { return dis_[view]; }
  cspm_pm_t* handle() const { return h_; }

 private:
  cspm_pm_t* h_ = nullptr;
  Mat dis_[2];
};

}  // namespace cspm
