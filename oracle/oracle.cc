// oracle/oracle.cc — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Line-by-line CPU restatement (C++17, no OpenCV) of the reference hot path
//   CSPM/plane.h:25-34                      Plane::update_param
//   CSPM/commfunc.h:117-145                 Round2Int, HandleBorder
//   CSPM/cc/grd_cc.cpp:4-35,60-154          GrdCC raw cost + buildCV/buildRightCV
//   CSPM/cc/cen_cc.cc:4-138                 CenCC raw cost (census)
//   CSPM/plane_cost/pre_ss_pc.cc:12-118     PreSSPC ctor + GetPlaneCost
//   CSPM/plane_cost/pre_cs_pc.cc:12-188     PreCSPC ctor + GetPlaneCost
//   CSPM/plane_cost/grd_pc.cc:12-178        GrdPC  (image-interpolating plane cost, single scale)
//   CSPM/plane_cost/cspc.cc:3-182           CSPC   (the same over a pyramid)
//   CSPM/cs_patchmatch.cc:3-34,51-345,590-602  CSPatchMatch driver
//   CSPM/cs_patchmatch.cc:347-588           post-processing (use_pp)
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may build, link or execute this file.
//
// PARITY PINNING: the reference ships no tests or golden vectors
// (SURVEY.md §4).  This restatement is pinned instead against
//   (1) oracle/_ref/libcspm_ref.so — the UNMODIFIED reference sources compiled
//       against oracle/cvshim (a minimal OpenCV stand-in), bit-exact on every
//       cost, plane and disparity (tests/test_oracle_vs_ref.py), and
//   (2) tests/golden/*.npz generated from (1) and from cv2-python 4.13.
// Build: g++ -O2 -ffp-contract=off -fopenmp (see oracle/Makefile).
#include <omp.h>

#include <algorithm>
#include <bitset>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "prims.h"

namespace {

using namespace cspm_prims;

// commfunc.h:24-29
const int kViewNum = 2;
const double kDoubleEps = 0.00000001;
const double kDoubleMax = std::numeric_limits<double>::max();
enum RefView { kLeft = 0, kRight = 1 };

// commfunc.h:117-121 — round-half-even via the 1.5*2^52 magic constant.
inline int Round2Int(double d) {
  d = d + 6755399441055744.0;
  int32_t lo;
  std::memcpy(&lo, &d, 4);
  return lo;
}
// commfunc.h:129-145 — single wrap-around, not clamp.
inline int HandleBorder(int loc, int size) {
  if (loc < 0) return loc + size;
  if (loc >= size) return loc - size;
  return loc;
}
inline uint8_t saturate_u8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

// plane.h:12-48
struct Plane {
  double norm[3] = {0, 0, 0};
  double point[3] = {0, 0, 0};
  double param[3] = {0, 0, 0};
  void set_point(double x, double y, double z) { point[0] = x; point[1] = y; point[2] = z; }
  void set_norm(double x, double y, double z) { norm[0] = x; norm[1] = y; norm[2] = z; }
  // plane.h:25-34.  norm_.dot(point_) is Matx::dot: s=0; s+=a[i]*b[i].
  void update_param() {
    double denom = std::max(std::fabs(norm[2]), kDoubleEps);
    if (norm[2] < 0.0) denom = -denom;
    param[0] = -norm[0] / denom;
    param[1] = -norm[1] / denom;
    double s = 0;
    for (int i = 0; i < 3; ++i) s += norm[i] * point[i];
    param[2] = s / denom;
  }
};
// cv::norm(Vec3d, NORM_L2): sqrt(normL2Sqr) with s += v*v left to right.
inline double norm_l2(const double v[3]) {
  double s = 0;
  for (int i = 0; i < 3; ++i) s += v[i] * v[i];
  return std::sqrt(s);
}

struct Img8 {  // CV_8UC3, BGR interleaved, dense
  int w = 0, h = 0;
  std::vector<uint8_t> d;
  const uint8_t* row(int y) const { return d.data() + (size_t)y * 3 * w; }
};
Img8 load_img(const uint8_t* p, int w, int h, int stride) {
  Img8 im; im.w = w; im.h = h; im.d.resize((size_t)w * h * 3);
  for (int y = 0; y < h; ++y) std::memcpy(&im.d[(size_t)y * 3 * w], p + (size_t)y * stride, (size_t)3 * w);
  return im;
}
// cvtColor(BGR2RGB) + convertTo(CV_64F)  (pre_ss_pc.cc:33-36)
std::vector<double> bgr8_to_rgb64(const Img8& im) {
  std::vector<double> o((size_t)im.w * im.h * 3);
  for (size_t i = 0; i < (size_t)im.w * im.h; ++i) {
    o[3 * i + 0] = im.d[3 * i + 2];
    o[3 * i + 1] = im.d[3 * i + 1];
    o[3 * i + 2] = im.d[3 * i + 0];
  }
  return o;
}

// ------------------------------------------------------------------ CCMethod
// cc_method.h:15-33.  Images are H x W x 3 f64 (RGB, 0..255); costVol is
// maxDis planar slices [d][y][x] f64, caller-allocated and zeroed.
struct CCMethod {
  virtual ~CCMethod() {}
  virtual void buildCV(const double* l, const double* r, int w, int h, int maxDis, double* vol) = 0;
  virtual void buildRightCV(const double* l, const double* r, int w, int h, int maxDis, double* vol) = 0;
};

// cc/grd_cc.h:6-9
const double BORDER_THRES = 3;
const double TAU_CLR = 10.0, TAU_GRD = 2.0, ALPHA = 0.1;

// grd_cc.cpp:4-19
inline double myCostGrd(const double* lC, const double* rC, const double* lG, const double* rG) {
  double clrDiff = 0;
  for (int c = 0; c < 3; c++) {
    double temp = std::fabs(lC[c] - rC[c]);
    clrDiff += temp;
  }
  clrDiff *= 0.3333333333;
  double grdDiff = std::fabs(lG[0] - rG[0]);
  clrDiff = clrDiff > TAU_CLR ? TAU_CLR : clrDiff;
  grdDiff = grdDiff > TAU_GRD ? TAU_GRD : grdDiff;
  return ALPHA * clrDiff + (1 - ALPHA) * grdDiff;
}
// grd_cc.cpp:21-35 — partner outside the image: compare against constant 3.
inline double myCostGrdBorder(const double* lC, const double* lG) {
  double clrDiff = 0;
  for (int c = 0; c < 3; c++) {
    double temp = std::fabs(lC[c] - BORDER_THRES);
    clrDiff += temp;
  }
  clrDiff *= 0.3333333333;
  double grdDiff = std::fabs(lG[0] - BORDER_THRES);
  clrDiff = clrDiff > TAU_CLR ? TAU_CLR : clrDiff;
  grdDiff = grdDiff > TAU_GRD ? TAU_GRD : grdDiff;
  return ALPHA * clrDiff + (1 - ALPHA) * grdDiff;
}

struct GrdCC : CCMethod {
  // grd_cc.cpp:70-77: convertTo(32F) -> RGB2GRAY -> Sobel(64F,1,0,1)
  static void gradients(const double* img, int w, int h, std::vector<double>* grd) {
    std::vector<float> gray((size_t)w * h);
    for (size_t i = 0; i < (size_t)w * h; ++i)
      gray[i] = rgb2gray_f32((float)img[3 * i], (float)img[3 * i + 1], (float)img[3 * i + 2]);
    grd->assign((size_t)w * h, 0.0);
    sobel_x_k1(gray.data(), w, h, grd->data());
  }
  // grd_cc.cpp:60-109
  void buildCV(const double* l, const double* r, int w, int h, int maxDis, double* vol) override {
    std::vector<double> lG, rG;
    gradients(l, w, h, &lG);
    gradients(r, w, h, &rG);
    for (int d = 0; d < maxDis; d++)
      for (int y = 0; y < h; y++) {
        const double* lData = l + (size_t)y * 3 * w;
        const double* rData = r + (size_t)y * 3 * w;
        const double* lGData = lG.data() + (size_t)y * w;
        const double* rGData = rG.data() + (size_t)y * w;
        double* cost = vol + ((size_t)d * h + y) * w;
        for (int x = 0; x < w; x++) {
          if (x - d >= 0)
            cost[x] = myCostGrd(lData + 3 * x, rData + 3 * (x - d), lGData + x, rGData + x - d);
          else
            cost[x] = myCostGrdBorder(lData + 3 * x, lGData + x);
        }
      }
  }
  // grd_cc.cpp:110-154
  void buildRightCV(const double* l, const double* r, int w, int h, int maxDis, double* vol) override {
    std::vector<double> lG, rG;
    gradients(l, w, h, &lG);
    gradients(r, w, h, &rG);
    for (int d = 0; d < maxDis; d++)
      for (int y = 0; y < h; y++) {
        const double* lData = l + (size_t)y * 3 * w;
        const double* rData = r + (size_t)y * 3 * w;
        const double* lGData = lG.data() + (size_t)y * w;
        const double* rGData = rG.data() + (size_t)y * w;
        double* cost = vol + ((size_t)d * h + y) * w;
        for (int x = 0; x < w; x++) {
          if (x + d < w)
            cost[x] = myCostGrd(lData + 3 * (x + d), rData + 3 * x, lGData + x + d, rGData + x);
          else
            cost[x] = myCostGrdBorder(rData + 3 * x, rGData + x);
        }
      }
  }
};

// cc/cen_cc.h:5-6, cen_cc.cc:4-138 — 9x9 census on 8-bit gray, wrap-around
// window border, Hamming distance, partner outside the image -> CENCUS_BIT.
const int CENCUS_WND = 9, CENCUS_BIT = 80;
struct CenCC : CCMethod {
  // cen_cc.cc:13-14: convertTo(CV_8U) (saturate_cast<uchar>(cvRound(v))), then
  // cvtColor(CV_RGB2GRAY) on 8U: OpenCV 2.4.8 fixed point with yuv_shift=14,
  // (R*4899 + G*9617 + B*1868 + 8192) >> 14.
  static void census(const double* img, int w, int h, std::vector<std::bitset<CENCUS_BIT>>* codes) {
    std::vector<uint8_t> gray((size_t)w * h);
    for (size_t i = 0; i < (size_t)w * h; ++i) {
      int ch[3];
      for (int c = 0; c < 3; ++c) {
        double v = img[3 * i + c];
        int iv = (int)std::nearbyint(v);
        ch[c] = iv < 0 ? 0 : (iv > 255 ? 255 : iv);
      }
      gray[i] = (uint8_t)((ch[0] * 4899 + ch[1] * 9617 + ch[2] * 1868 + 8192) >> 14);
    }
    const int H_WD = CENCUS_WND / 2;
    codes->assign((size_t)w * h, std::bitset<CENCUS_BIT>());
    for (int y = 0; y < h; y++)
      for (int x = 0; x < w; x++) {
        uint8_t c = gray[(size_t)y * w + x];
        int bitCnt = 0;
        std::bitset<CENCUS_BIT>& code = (*codes)[(size_t)y * w + x];
        for (int wy = -H_WD; wy <= H_WD; wy++) {
          int qy = (y + wy + h) % h;
          for (int wx = -H_WD; wx <= H_WD; wx++) {
            if (wy != 0 || wx != 0) {
              int qx = (x + wx + w) % w;
              code[bitCnt] = (c > gray[(size_t)qy * w + qx]);
              bitCnt++;
            }
          }
        }
      }
  }
  void buildCV(const double* l, const double* r, int w, int h, int maxDis, double* vol) override {
    std::vector<std::bitset<CENCUS_BIT>> lc, rc;
    census(l, w, h, &lc);
    census(r, w, h, &rc);
    for (int d = 0; d < maxDis; d++)
      for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
          double* cost = vol + ((size_t)d * h + y) * w;
          if (x - d >= 0)
            cost[x] = (double)(lc[(size_t)y * w + x] ^ rc[(size_t)y * w + x - d]).count();
          else
            cost[x] = CENCUS_BIT;
        }
  }
  void buildRightCV(const double* l, const double* r, int w, int h, int maxDis, double* vol) override {
    std::vector<std::bitset<CENCUS_BIT>> lc, rc;
    census(l, w, h, &lc);
    census(r, w, h, &rc);
    for (int d = 0; d < maxDis; d++)
      for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
          double* cost = vol + ((size_t)d * h + y) * w;
          if (x + d < w)
            cost[x] = (double)(rc[(size_t)y * w + x] ^ lc[(size_t)y * w + x + d]).count();
          else
            cost[x] = CENCUS_BIT;
        }
  }
};

CCMethod* GetCCType(const std::string& name) {  // main.cc:39-55
  if (name == "GRD") return new GrdCC();
  if (name == "CEN") return new CenCC();
  return nullptr;
}

// ---------------------------------------------------------------- IPlaneCost
const double WGT_GAMMA = 10.0;  // pre_ss_pc.h:16

struct IPlaneCost {  // plane_cost/i_plane_cost.h:13-34
  virtual ~IPlaneCost() {}
  virtual double GetPlaneCost(int ref_x, int ref_y, const Plane& plane, RefView view) const = 0;
};

struct ScaleData {
  int w = 0, h = 0, max_disp = 0;
  Img8 img[kViewNum];
  std::vector<double> vol[kViewNum];  // (max_disp+1) slices [d][y][x]
  double max_cost[kViewNum] = {0, 0};
  void build(CCMethod* cc) {
    // pre_ss_pc.cc:33-58 / pre_cs_pc.cc:57-84
    std::vector<double> tl = bgr8_to_rgb64(img[kLeft]);
    std::vector<double> tr = bgr8_to_rgb64(img[kRight]);
    for (int v = 0; v < kViewNum; ++v) {
      vol[v].assign((size_t)(max_disp + 1) * h * w, 0.0);
      if (v == kLeft) cc->buildCV(tl.data(), tr.data(), w, h, max_disp + 1, vol[v].data());
      else cc->buildRightCV(tl.data(), tr.data(), w, h, max_disp + 1, vol[v].data());
      max_cost[v] = -1.0;
      for (int d = 0; d < max_disp + 1; ++d) {
        double tmp = -DBL_MAX;
        const double* s = vol[v].data() + (size_t)d * h * w;
        for (size_t i = 0; i < (size_t)h * w; ++i) tmp = s[i] > tmp ? s[i] : tmp;
        if (tmp > max_cost[v]) max_cost[v] = tmp;
      }
    }
  }
  // pre_ss_pc.cc:84-116 == pre_cs_pc.cc:151-181 (the 35x35 slanted window sum)
  double window(int ref_x, int ref_y, double plane_a, double plane_b, double plane_c, int view,
                int half_wnd, const double* lookup_exp) const {
    double cost = 0.0;
    const uint8_t* I_p = img[view].row(ref_y) + 3 * ref_x;
    const double* V = vol[view].data();
    const size_t slice = (size_t)h * w;
    for (int dy = -half_wnd; dy <= half_wnd; ++dy) {
      int q_y = ref_y + dy;
      if (q_y >= 0 && q_y < h) {
        const uint8_t* I_q_y = img[view].row(q_y);
        const double q_disp_y = plane_b * q_y + plane_c;
        for (int dx = -half_wnd; dx <= half_wnd; ++dx) {
          int q_x = ref_x + dx;
          if (q_x >= 0 && q_x < w) {
            const uint8_t* I_q = I_q_y + 3 * q_x;
            int sum = std::abs(I_p[0] - I_q[0]) + std::abs(I_p[1] - I_q[1]) + std::abs(I_p[2] - I_q[2]);
            const double wgt = lookup_exp[sum];
            double q_disp = plane_a * q_x + q_disp_y;
            // static_cast<int>(double): x86 cvttsd2si -> INT_MIN when out of range / NaN
            int q_disp_floor;
            if (!(q_disp > -2147483649.0 && q_disp < 2147483648.0)) q_disp_floor = INT32_MIN;
            else q_disp_floor = static_cast<int>(q_disp);
            if (q_disp_floor <= 0 || q_disp_floor >= max_disp) {
              cost += wgt * max_cost[view];
            } else {
              int q_disp_ceil = q_disp_floor + 1;
              const double floor_wgt = q_disp_ceil - q_disp;
              double tmp = floor_wgt * V[(size_t)q_disp_floor * slice + (size_t)q_y * w + q_x] +
                           (1 - floor_wgt) * V[(size_t)q_disp_ceil * slice + (size_t)q_y * w + q_x];
              cost += wgt * tmp;
            }
          }
        }
      }
    }
    return cost;
  }
};

void make_exp_lut(std::vector<double>* lut) {  // pre_ss_pc.cc:61-64
  lut->resize(1000);
  for (int i = 0; i < 1000; ++i) (*lut)[i] = std::exp(-i * 1.0 / WGT_GAMMA);
}

struct PreSSPC : IPlaneCost {  // plane_cost/pre_ss_pc.cc
  ScaleData s0;
  int wnd_size, half_wnd;
  std::vector<double> lookup_exp;
  PreSSPC(const Img8& l, const Img8& r, int max_disp, int wnd, CCMethod* cc) : wnd_size(wnd) {
    s0.img[kLeft] = l; s0.img[kRight] = r;
    s0.w = l.w; s0.h = l.h; s0.max_disp = max_disp;
    half_wnd = wnd_size / 2;
    s0.build(cc);
    make_exp_lut(&lookup_exp);
  }
  double GetPlaneCost(int ref_x, int ref_y, const Plane& plane, RefView view) const override {
    return s0.window(ref_x, ref_y, plane.param[0], plane.param[1], plane.param[2], view, half_wnd,
                     lookup_exp.data());
  }
};

struct PreCSPC : IPlaneCost {  // plane_cost/pre_cs_pc.cc
  int wnd_size, scale_num;
  std::vector<ScaleData> sc;
  std::vector<double> scale_wgt, lookup_exp;
  PreCSPC(const Img8& l, const Img8& r, int max_disp, int wnd, int scale_num_, CCMethod* cc,
          double reg_lambda)
      : wnd_size(wnd), scale_num(scale_num_) {
    sc.resize(scale_num);
    sc[0].img[kLeft] = l; sc[0].img[kRight] = r;
    sc[0].w = l.w; sc[0].h = l.h; sc[0].max_disp = max_disp;
    for (int s = 1; s < scale_num; ++s) {  // pre_cs_pc.cc:42-49
      sc[s].h = (sc[s - 1].h + 1) / 2;
      sc[s].w = (sc[s - 1].w + 1) / 2;
      sc[s].max_disp = sc[s - 1].max_disp / 2;
      for (int v = 0; v < kViewNum; ++v) {
        Img8& d = sc[s].img[v];
        d.w = sc[s].w; d.h = sc[s].h; d.d.resize((size_t)d.w * d.h * 3);
        const Img8& p = sc[s - 1].img[v];
        pyrdown_u8c3(p.d.data(), p.w, p.h, 3 * p.w, d.d.data(), d.w, d.h, 3 * d.w);
      }
    }
    for (int s = 0; s < scale_num; ++s) sc[s].build(cc);
    // pre_cs_pc.cc:89-109
    std::vector<double> reg((size_t)scale_num * scale_num, 0.0), inv((size_t)scale_num * scale_num, 0.0);
    auto at = [&](int r_, int c_) -> double& { return reg[(size_t)r_ * scale_num + c_]; };
    for (int s = 0; s < scale_num; ++s) {
      if (s == 0) {
        at(s, s) = 1 + reg_lambda;
        at(s, s + 1) = -reg_lambda;
      } else if (s == scale_num - 1) {
        at(s, s) = 1 + reg_lambda;
        at(s, s - 1) = -reg_lambda;
      } else {
        at(s, s) = 1 + 2 * reg_lambda;
        at(s, s - 1) = -reg_lambda;
        at(s, s + 1) = -reg_lambda;
      }
    }
    invert_lu(reg.data(), scale_num, inv.data());
    scale_wgt.resize(scale_num);
    for (int s = 0; s < scale_num; ++s) scale_wgt[s] = inv[s];
    make_exp_lut(&lookup_exp);
  }
  // pre_cs_pc.cc:133-188
  double GetPlaneCost(int ref_x, int ref_y, const Plane& plane, RefView view) const override {
    const int half_wnd = wnd_size / 2;
    double cost = 0.0;
    double cur_disp = plane.param[0] * ref_x + plane.param[1] * ref_y + plane.param[2];
    int cur_y = ref_y, cur_x = ref_x;
    for (int s = 0; s < scale_num; ++s) {
      Plane cur_plane;
      cur_plane.set_norm(plane.norm[0], plane.norm[1], plane.norm[2]);
      cur_plane.set_point(cur_x, cur_y, cur_disp);
      cur_plane.update_param();
      double scale_cost = sc[s].window(cur_x, cur_y, cur_plane.param[0], cur_plane.param[1],
                                       cur_plane.param[2], view, half_wnd, lookup_exp.data());
      cost += scale_cost * scale_wgt[s];
      cur_y /= 2;
      cur_x /= 2;
      cur_disp /= 2.0;
    }
    return cost;
  }
};

// GrdPC (grd_pc.cc) and CSPC (cspc.cc): the raw cost is computed on the fly by linearly
// interpolating the partner IMAGE (colour and 8-bit-gray x-gradient) at the sub-pixel position
// q_x -+ d, with wrap-around (HandleBorder); scale_num == 0 -> GrdPC, else CSPC.
const double COST_ALPHA = 0.1;  // grd_pc.h:13-16
struct ImgPC : IPlaneCost {
  struct Level {
    int w, h, max_disp;
    Img8 img[kViewNum];
    std::vector<double> grd[kViewNum];
  };
  int wnd_size, scale_num;  // scale_num == 0: GrdPC
  std::vector<Level> lv;
  std::vector<double> scale_wgt, lookup_exp;
  ImgPC(const Img8& l, const Img8& r, int max_disp, int wnd, int scale_num_, double reg_lambda)
      : wnd_size(wnd), scale_num(scale_num_) {
    const int S = std::max(scale_num, 1);
    lv.resize(S);
    lv[0].img[kLeft] = l; lv[0].img[kRight] = r;
    lv[0].w = l.w; lv[0].h = l.h; lv[0].max_disp = max_disp;
    for (int s = 0; s < S; ++s) {
      if (s > 0) {  // cspc.cc:33-38
        lv[s].h = (lv[s - 1].h + 1) / 2;
        lv[s].w = (lv[s - 1].w + 1) / 2;
        lv[s].max_disp = lv[s - 1].max_disp / 2;
      }
      for (int v = 0; v < kViewNum; ++v) {
        Img8& d = lv[s].img[v];
        if (s > 0) {
          d.w = lv[s].w; d.h = lv[s].h; d.d.resize((size_t)d.w * d.h * 3);
          const Img8& p = lv[s - 1].img[v];
          pyrdown_u8c3(p.d.data(), p.w, p.h, 3 * p.w, d.d.data(), d.w, d.h, 3 * d.w);
        }
        // cvtColor(BGR2GRAY) 8U + Sobel(64F,1,0,1)  (grd_pc.cc:37-40, cspc.cc:44-47)
        std::vector<uint8_t> gray((size_t)d.w * d.h);
        for (size_t i = 0; i < gray.size(); ++i) gray[i] = gray8_fixed(d.d[3 * i + 2], d.d[3 * i + 1], d.d[3 * i]);
        lv[s].grd[v].assign(gray.size(), 0.0);
        sobel_x_k1_u8(gray.data(), d.w, d.h, lv[s].grd[v].data());
      }
    }
    if (scale_num > 0) {  // cspc.cc:51-72
      std::vector<double> reg((size_t)S * S, 0.0), inv((size_t)S * S, 0.0);
      auto at = [&](int r_, int c_) -> double& { return reg[(size_t)r_ * S + c_]; };
      for (int s = 0; s < S; ++s) {
        if (s == 0) { at(s, s) = 1 + reg_lambda; at(s, s + 1) = -reg_lambda; }
        else if (s == S - 1) { at(s, s) = 1 + reg_lambda; at(s, s - 1) = -reg_lambda; }
        else { at(s, s) = 1 + 2 * reg_lambda; at(s, s - 1) = -reg_lambda; at(s, s + 1) = -reg_lambda; }
      }
      invert_lu(reg.data(), S, inv.data());
      scale_wgt.assign(inv.begin(), inv.begin() + S);
    } else {
      scale_wgt.assign(1, 1.0);
    }
    make_exp_lut(&lookup_exp);
  }
  // grd_pc.cc:84-176 == cspc.cc:122-175
  double window(const Level& L, int ref_x, int ref_y, double plane_a, double plane_b, double plane_c, int view) const {
    const int half_wnd = wnd_size / 2;
    double cost = 0.0;
    const uint8_t* I_p = L.img[view].row(ref_y) + 3 * ref_x;
    for (int dy = -half_wnd; dy <= half_wnd; ++dy) {
      int q_y = ref_y + dy;
      if (q_y >= 0 && q_y < L.h) {
        const uint8_t* I_q_y = L.img[view].row(q_y);
        const double* G_q_y = L.grd[view].data() + (size_t)q_y * L.w;
        const uint8_t* I_ohter_y = L.img[1 - view].row(q_y);
        const double* G_other_y = L.grd[1 - view].data() + (size_t)q_y * L.w;
        const double q_disp_y = plane_b * q_y + plane_c;
        for (int dx = -half_wnd; dx <= half_wnd; ++dx) {
          int q_x = ref_x + dx;
          if (q_x >= 0 && q_x < L.w) {
            const uint8_t* I_q = I_q_y + 3 * q_x;
            int sum = std::abs(I_p[0] - I_q[0]) + std::abs(I_p[1] - I_q[1]) + std::abs(I_p[2] - I_q[2]);
            const double wgt = lookup_exp[sum];
            double q_disp = plane_a * q_x + q_disp_y;
            int q_disp_floor;
            if (!(q_disp > -2147483649.0 && q_disp < 2147483648.0)) q_disp_floor = INT32_MIN;
            else q_disp_floor = static_cast<int>(q_disp);
            if (q_disp_floor <= 0 || q_disp_floor >= L.max_disp) {
              cost += wgt * (COST_ALPHA * TAU_CLR + (1 - COST_ALPHA) * TAU_GRD);
            } else {
              const double other_x = q_x + (2 * view - 1) * q_disp;
              int floor_x = static_cast<int>(other_x);
              int ceil_x = floor_x + 1;
              const double floor_wgt = ceil_x - other_x;
              floor_x = HandleBorder(floor_x, L.w);
              ceil_x = HandleBorder(ceil_x, L.w);
              const uint8_t* I_floor = I_ohter_y + 3 * floor_x;
              const uint8_t* I_ceil = I_ohter_y + 3 * ceil_x;
              double clr_cost = std::fabs(I_q[0] - I_ceil[0] + floor_wgt * (I_ceil[0] - I_floor[0])) +
                                std::fabs(I_q[1] - I_ceil[1] + floor_wgt * (I_ceil[1] - I_floor[1])) +
                                std::fabs(I_q[2] - I_ceil[2] + floor_wgt * (I_ceil[2] - I_floor[2]));
              clr_cost *= 0.33333333333333;
              const double G_floor = G_other_y[floor_x];
              const double G_ceil = G_other_y[ceil_x];
              double grd_cost = std::fabs(G_q_y[q_x] - G_ceil + floor_wgt * (G_ceil - G_floor));
              cost += wgt * (COST_ALPHA * std::min(clr_cost, TAU_CLR) + (1 - COST_ALPHA) * std::min(grd_cost, TAU_GRD));
            }
          }
        }
      }
    }
    return cost;
  }
  double GetPlaneCost(int ref_x, int ref_y, const Plane& plane, RefView view) const override {
    if (scale_num == 0)  // GrdPC::GetPlaneCost, grd_pc.cc:71-178
      return window(lv[0], ref_x, ref_y, plane.param[0], plane.param[1], plane.param[2], view);
    double cost = 0.0;   // CSPC::GetPlaneCost, cspc.cc:107-182
    double cur_disp = plane.param[0] * ref_x + plane.param[1] * ref_y + plane.param[2];
    int cur_y = ref_y, cur_x = ref_x;
    for (int s = 0; s < scale_num; ++s) {
      Plane cur_plane;
      cur_plane.set_norm(plane.norm[0], plane.norm[1], plane.norm[2]);
      cur_plane.set_point(cur_x, cur_y, cur_disp);
      cur_plane.update_param();
      double scale_cost = window(lv[s], cur_x, cur_y, cur_plane.param[0], cur_plane.param[1], cur_plane.param[2], view);
      cost += scale_cost * scale_wgt[s];
      cur_y /= 2;
      cur_x /= 2;
      cur_disp /= 2.0;
    }
    return cost;
  }
};

// -------------------------------------------------------------- CSPatchMatch
const double WMF_GAMMA = 10.0;  // cs_patchmatch.h:14

struct CSPatchMatch {  // cs_patchmatch.{h,cc}
  int wid_, hei_, max_dis_, dis_scale_;
  uint64_t seed_;
  Img8 img_[kViewNum];
  std::vector<uint8_t> dis_[kViewNum];
  std::vector<Plane> plane_[kViewNum];
  std::vector<double> min_cost_[kViewNum];
  const double kMaxNorm_ = 1.0, kZStopThres_ = 0.1;  // cs_patchmatch.h:145-146

  CSPatchMatch(const Img8& l, const Img8& r, int max_dis, int dis_scale, uint64_t seed)
      : max_dis_(max_dis), dis_scale_(dis_scale), seed_(seed) {  // cs_patchmatch.cc:3-34
    img_[kLeft] = l; img_[kRight] = r;
    wid_ = l.w; hei_ = l.h;
    for (int v = 0; v < kViewNum; ++v) {
      dis_[v].assign((size_t)wid_ * hei_, 0);
      plane_[v].assign((size_t)wid_ * hei_, Plane());
      min_cost_[v].assign((size_t)wid_ * hei_, kDoubleMax);
    }
  }
  Plane& P(int v, int y, int x) { return plane_[v][(size_t)y * wid_ + x]; }
  double& C(int v, int y, int x) { return min_cost_[v][(size_t)y * wid_ + x]; }

  int n_ref() const {  // halvings of z_iter from max_dis/2 while >= 0.1 (:299-301,342-343)
    int n = 0;
    for (double z = max_dis_ / 2.0; z >= kZStopThres_; z /= 2.0) ++n;
    return n;
  }

  // cs_patchmatch.cc:115-148
  void InitRandomPlane(const IPlaneCost* pc) {
    const double norm_avg = 0.0, norm_std = 1.0;
    for (int v = kLeft; v <= kRight; ++v) {
#pragma omp parallel for schedule(dynamic, 4)
      for (int y = 0; y < hei_; ++y) {
        const uint64_t stream = stream_init(hei_, v, y);  // RNG rng(time(NULL)) (:130)
        uint32_t slot = 0;
        for (int x = 0; x < wid_; ++x) {
          double rand_dis = rng_uniform(seed_, stream, slot++, 0, kDoubleEps, (double)max_dis_);
          Plane& p = P(v, y, x);
          p.set_point(x, y, rand_dis);
          double rn[3];
          rng_fill_normal3(seed_, stream, slot++, norm_avg, norm_std, rn);
          double denom = std::max(norm_l2(rn), kDoubleEps);
          double inv = 1. / denom;  // Vec / double == Vec * (1./alpha) in OpenCV
          p.set_norm(rn[0] * inv, rn[1] * inv, rn[2] * inv);
          p.update_param();
          C(v, y, x) = pc->GetPlaneCost(x, y, p, (RefView)v);
        }
      }
    }
  }
  // cs_patchmatch.cc:163-216
  void SpatialPropagation(int cur_iter, const IPlaneCost* pc) {
    int x_st = wid_ - 2, x_ed = -1, x_inc = -1;
    int y_st = hei_ - 2, y_ed = -1, y_inc = -1;
    if (cur_iter % 2 == 0) {
      x_st = 1; x_ed = wid_; x_inc = 1;
      y_st = 1; y_ed = hei_; y_inc = 1;
    }
    for (int v = kLeft; v <= kRight; ++v) {
      for (int x = x_st; x != x_ed; x += x_inc) {
        const Plane nx_plane = P(v, y_st - y_inc, x - x_inc);
        const double nx_cost = pc->GetPlaneCost(x, y_st - y_inc, nx_plane, (RefView)v);
        if (nx_cost < C(v, y_st - y_inc, x)) {
          C(v, y_st - y_inc, x) = nx_cost;
          P(v, y_st - y_inc, x) = nx_plane;
        }
      }
      for (int y = y_st; y != y_ed; y += y_inc) {
        const Plane ny0 = P(v, y - y_inc, x_st - x_inc);
        const double ny0_cost = pc->GetPlaneCost(x_st - x_inc, y, ny0, (RefView)v);
        if (ny0_cost < C(v, y, x_st - x_inc)) {
          C(v, y, x_st - x_inc) = ny0_cost;
          P(v, y, x_st - x_inc) = ny0;
        }
        for (int x = x_st; x != x_ed; x += x_inc) {
          const Plane nx_plane = P(v, y, x - x_inc);
          const double nx_cost = pc->GetPlaneCost(x, y, nx_plane, (RefView)v);
          if (nx_cost < C(v, y, x)) {
            C(v, y, x) = nx_cost;
            P(v, y, x) = nx_plane;
          }
          const Plane ny_plane = P(v, y - y_inc, x);
          const double ny_cost = pc->GetPlaneCost(x, y, ny_plane, (RefView)v);
          if (ny_cost < C(v, y, x)) {
            C(v, y, x) = ny_cost;
            P(v, y, x) = ny_plane;
          }
        }
      }
    }
  }
  // Throughput schedule named in BASELINE.json's north_star (NOT in the reference; restated here so that the
  // kernel has a schedule-matched oracle, include/cspm_c.h CSPM_SCHEDULE_REDBLACK / _REDBLACK8): instead of the
  // raster Gauss-Seidel scan, iteration i makes two checkerboard half-sweeps, colour (i + half) & 1 first.  A pixel
  // (x, y) of the active colour tests, in this order, the planes of (x-1,y), (x+1,y), (x,y-1), (x,y+1) — and with
  // far8 also (x-5,y), (x+5,y), (x,y-5), (x,y+5) — against its running minimum with the reference's strict `<`.
  // All offsets are odd, so every neighbour has the other colour and is not modified during the half-sweep:
  // the result does not depend on the order in which the active pixels are processed.
  void SpatialPropagationRedBlack(int cur_iter, const IPlaneCost* pc, bool far8) {
    static const int kOff[8][2] = {{-1, 0}, {1, 0}, {0, -1}, {0, 1}, {-5, 0}, {5, 0}, {0, -5}, {0, 5}};
    const int noff = far8 ? 8 : 4;
    for (int half = 0; half < 2; ++half) {
      const int colour = (cur_iter + half) & 1;
      for (int v = kLeft; v <= kRight; ++v) {
#pragma omp parallel for schedule(dynamic, 4)
        for (int y = 0; y < hei_; ++y)
          for (int x = (y + colour) & 1; x < wid_; x += 2)
            for (int k = 0; k < noff; ++k) {
              const int nx = x + kOff[k][0], ny = y + kOff[k][1];
              if (nx < 0 || nx >= wid_ || ny < 0 || ny >= hei_) continue;
              const Plane cand = P(v, ny, nx);
              const double cost = pc->GetPlaneCost(x, y, cand, (RefView)v);
              if (cost < C(v, y, x)) {
                C(v, y, x) = cost;
                P(v, y, x) = cand;
              }
            }
      }
    }
  }
  int schedule_ = 0;  // 0 raster (the reference), 1 red-black 4-neighbour, 2 red-black 8-neighbour
  void Spatial(int cur_iter, const IPlaneCost* pc) {
    if (schedule_ == 0) SpatialPropagation(cur_iter, pc);
    else SpatialPropagationRedBlack(cur_iter, pc, schedule_ == 2);
  }
  // cs_patchmatch.cc:229-277
  void ViewPropagation(int cur_iter, const IPlaneCost* pc) {
    Plane cor_plane;
    int x_st = wid_ - 1, x_ed = -1, x_inc = -1;
    int y_st = hei_ - 1, y_ed = -1, y_inc = -1;
    if (cur_iter % 2 == 0) {
      x_st = 0; x_ed = wid_; x_inc = 1;
      y_st = 0; y_ed = hei_; y_inc = 1;
    }
    for (int v = kLeft; v <= kRight; ++v) {
      int other_view = 1 - v;
      for (int y = y_st; y != y_ed; y += y_inc) {
        for (int x = x_st; x != x_ed; x += x_inc) {
          const Plane& o = P(other_view, y, x);
          double disp = o.param[0] * x + o.param[1] * y + o.param[2];
          if (disp < 0.0) disp = 0.0;
          if (disp >= max_dis_) disp = max_dis_ - 1.0;
          int cor_x = 0;
          if (v == kLeft) cor_x = HandleBorder(x + Round2Int(disp), wid_);
          else cor_x = HandleBorder(x - Round2Int(disp), wid_);
          cor_plane.set_norm(o.norm[0], o.norm[1], o.norm[2]);
          cor_plane.set_point(cor_x, y, disp);
          cor_plane.update_param();
          const double cor_cost = pc->GetPlaneCost(cor_x, y, cor_plane, (RefView)v);
          if (cor_cost < C(v, y, cor_x)) {
            C(v, y, cor_x) = cor_cost;
            P(v, y, cor_x) = cor_plane;
          }
        }
      }
    }
  }
  // cs_patchmatch.cc:292-345; `iter` only selects the RNG stream.
  void PlaneRefinement(int iter, double z_max, double n_max, double z_thres, const IPlaneCost* pc) {
    double z_iter = z_max, n_iter = n_max;
    const int nref = n_ref();
    int step = 0;
    while (z_iter >= z_thres) {
      for (int v = kLeft; v <= kRight; ++v) {
#pragma omp parallel for schedule(dynamic, 4)
        for (int y = 0; y < hei_; ++y) {
          const uint64_t stream = stream_refine(hei_, nref, iter, step, v, y);  // (:309)
          uint32_t slot = 0;
          for (int x = 0; x < wid_; ++x) {
            Plane disturb_plane;
            const Plane& cur = P(v, y, x);
            double disturb_z = cur.param[0] * x + cur.param[1] * y + cur.param[2];
            double dz = rng_uniform(seed_, stream, slot++, 0, -z_iter, z_iter);
            disturb_plane.set_point(x, y, disturb_z + dz);
            double delta[3], dn[3];
            rng_fill_uniform3(seed_, stream, slot++, -n_iter, n_iter, delta);
            for (int i = 0; i < 3; ++i) dn[i] = cur.norm[i] + delta[i];
            double denom = std::max(norm_l2(dn), kDoubleEps);
            double inv = 1. / denom;
            disturb_plane.set_norm(dn[0] * inv, dn[1] * inv, dn[2] * inv);
            disturb_plane.update_param();
            const double cost = pc->GetPlaneCost(x, y, disturb_plane, (RefView)v);
            if (cost < C(v, y, x)) {
              P(v, y, x) = disturb_plane;
              C(v, y, x) = cost;
            }
          }
        }
      }
      z_iter /= 2.0;
      n_iter /= 2.0;
      ++step;
    }
  }
  // cs_patchmatch.cc:590-602
  void PlaneToDisp() {
    for (int v = kLeft; v <= kRight; ++v)
      for (int y = 0; y < hei_; ++y)
        for (int x = 0; x < wid_; ++x) {
          const Plane& p = P(v, y, x);
          double disp = 0;  // param().dot(Vec3d(x,y,1.0))
          disp += p.param[0] * x;
          disp += p.param[1] * y;
          disp += p.param[2] * 1.0;
          dis_[v][(size_t)y * wid_ + x] = saturate_u8(Round2Int(disp * dis_scale_));
        }
  }
  double plane_disp(int v, int y, int px, int x) {  // param().dot(Vec3d(x,y,1.0)) of pixel px's plane
    const Plane& p = P(v, y, px);
    double d = 0;
    d += p.param[0] * x;
    d += p.param[1] * y;
    d += p.param[2] * 1.0;
    return d;
  }
  // cs_patchmatch.cc:347-369
  void LeftRightCheck(std::vector<int>* valid) {
    for (int v = kLeft; v <= kRight; ++v) {
      for (int y = 0; y < hei_; y++) {
        for (int x = 0; x < wid_; x++) {
          int& cv = valid[v][(size_t)y * wid_ + x];
          cv = 0;
          double cur_dis = dis_[v][(size_t)y * wid_ + x] * 1.0 / dis_scale_;
          int other_x = x + (2 * v - 1) * Round2Int(cur_dis);
          if (other_x >= 0 && other_x < wid_) {
            double other_dis = dis_[1 - v][(size_t)y * wid_ + other_x] * 1.0 / dis_scale_;
            if (std::fabs(cur_dis - other_dis) <= 0.5 && cur_dis > 0.0) cv = 1;
          }
        }
      }
    }
  }
  // cs_patchmatch.cc:370-428
  void FillInvalid(std::vector<int>* valid) {
    for (int v = kLeft; v <= kRight; ++v) {
      for (int y = 0; y < hei_; ++y) {
        int* y_valid = valid[v].data() + (size_t)y * wid_;
        uint8_t* dis_data = dis_[v].data() + (size_t)y * wid_;
        for (int x = 0; x < wid_; ++x) {
          if (y_valid[x] == 0) {
            int l_first = x, l_find = 0;
            while (l_first >= 0) {
              if (y_valid[l_first]) { l_find = 1; break; }
              --l_first;
            }
            int r_find = 0, r_first = x;
            while (r_first < wid_) {
              if (y_valid[r_first]) { r_find = 1; break; }
              ++r_first;
            }
            if (l_find && r_find) {
              double l_d = plane_disp(v, y, l_first, x);
              double r_d = plane_disp(v, y, r_first, x);
              if (l_d <= r_d) dis_data[x] = saturate_u8(dis_scale_ * Round2Int(l_d));
              else dis_data[x] = saturate_u8(dis_scale_ * Round2Int(r_d));
            } else if (l_find) {
              dis_data[x] = saturate_u8(dis_scale_ * Round2Int(plane_disp(v, y, l_first, x)));
            } else if (r_find) {
              dis_data[x] = saturate_u8(dis_scale_ * Round2Int(plane_disp(v, y, r_first, x)));
            }
          }
        }
      }
    }
  }
  // cs_patchmatch.cc:430-506.  Note the in-place update: later pixels see
  // already-filtered disparities of earlier (invalid) pixels only through
  // q_dis_data of *valid* pixels, which are never rewritten -> order-free.
  void WeightedMedian(std::vector<int>* valid, int wnd_size, double gamma) {
    int half_wnd = wnd_size / 2;
    std::vector<double> lookup_exp(1000);
    for (int i = 0; i < 1000; ++i) lookup_exp[i] = std::exp(-i * 1.0 / gamma);
    for (int v = kLeft; v <= kRight; ++v) {
      for (int y = 0; y < hei_; y++) {
        uint8_t* cur_dis = dis_[v].data() + (size_t)y * wid_;
        const uint8_t* pL = img_[v].row(y);
        for (int x = 0; x < wid_; x++) {
          if (valid[v][(size_t)y * wid_ + x] == 0) {
            const uint8_t* pL_x = pL + 3 * x;
            double disp_hist[256];
            std::fill(disp_hist, disp_hist + 256, 0.0);
            double sum_wgt = 0.0;
            for (int wy = -half_wnd; wy <= half_wnd; wy++) {
              const int qy = y + wy;
              if (qy >= 0 && qy < hei_) {
                const int* qLValid = valid[v].data() + (size_t)qy * wid_;
                const uint8_t* qL = img_[v].row(qy);
                const uint8_t* q_dis_data = dis_[v].data() + (size_t)qy * wid_;
                for (int wx = -half_wnd; wx <= half_wnd; wx++) {
                  const int qx = x + wx;
                  if (qx >= 0 && qx < wid_) {
                    if (qLValid[qx]) {
                      const int q_disp = q_dis_data[qx];
                      const uint8_t* qL_x = qL + 3 * qx;
                      int clr_diff = std::abs(pL_x[0] - qL_x[0]) + std::abs(pL_x[1] - qL_x[1]) +
                                     std::abs(pL_x[2] - qL_x[2]);
                      double wgt = lookup_exp[clr_diff];
                      disp_hist[q_disp] += wgt;
                      sum_wgt += wgt;
                    }
                  }
                }
              }
            }
            double median_wgt = sum_wgt / 2.0;
            sum_wgt = 0.0;
            int median_disp = 0;
            for (int d = 0; d < 256; d++) {
              sum_wgt += disp_hist[d];
              if (sum_wgt >= median_wgt) { median_disp = d; break; }
            }
            if (median_wgt > 0.0) cur_dis[x] = (uint8_t)median_disp;
          }
        }
      }
    }
  }
  // cs_patchmatch.cc:508-588
  void PostProcessing() {
    std::vector<int> valid[kViewNum];
    for (int v = 0; v < kViewNum; ++v) valid[v].assign((size_t)wid_ * hei_, 0);
    LeftRightCheck(valid);
    FillInvalid(valid);
    WeightedMedian(valid, 35, WMF_GAMMA);
  }
  // cs_patchmatch.cc:51-109
  void PatchMatch(int iter_num, const IPlaneCost* pc, bool use_pp) {
    InitRandomPlane(pc);
    for (int i = 0; i < iter_num; ++i) {
      Spatial(i, pc);
      ViewPropagation(i, pc);
      PlaneRefinement(i, max_dis_ / 2.0, kMaxNorm_, kZStopThres_, pc);
    }
    PlaneToDisp();
    if (use_pp) PostProcessing();
  }
};

}  // namespace

// ------------------------------------------------------------------- C ABI
extern "C" {

// primitives (for the cv2 fixture tests)
void orc_gray_f32(const double* rgb, int n, float* out, int fma_variant) {
  for (int i = 0; i < n; ++i) {
    float r = (float)rgb[3 * i], g = (float)rgb[3 * i + 1], b = (float)rgb[3 * i + 2];
    out[i] = fma_variant ? rgb2gray_f32_fma(r, g, b) : rgb2gray_f32(r, g, b);
  }
}
void orc_sobel_x(const float* gray, int w, int h, double* out) { sobel_x_k1(gray, w, h, out); }
void orc_pyrdown(const uint8_t* src, int w, int h, uint8_t* dst) {
  pyrdown_u8c3(src, w, h, 3 * w, dst, (w + 1) / 2, (h + 1) / 2, 3 * ((w + 1) / 2));
}
int orc_invert(const double* a, int n, double* out) { return invert_lu(a, n, out) ? 1 : 0; }
int orc_round2int(double d) { return Round2Int(d); }
int orc_handle_border(int loc, int size) { return HandleBorder(loc, size); }
uint64_t orc_rng_draw(uint64_t seed, uint64_t stream, uint32_t slot, uint32_t sub) {
  return rng_draw(seed, stream, slot, sub);
}
double orc_rng_uniform(uint64_t seed, uint64_t stream, uint32_t slot, uint32_t sub, double a, double b) {
  return rng_uniform(seed, stream, slot, sub, a, b);
}
void orc_rng_normal3(uint64_t seed, uint64_t stream, uint32_t slot, double* v) {
  rng_fill_normal3(seed, stream, slot, 0.0, 1.0, v);
}
void orc_plane_param(const double* n, const double* p, double* prm) {
  Plane pl;
  pl.set_norm(n[0], n[1], n[2]);
  pl.set_point(p[0], p[1], p[2]);
  pl.update_param();
  prm[0] = pl.param[0]; prm[1] = pl.param[1]; prm[2] = pl.param[2];
}

// CCMethod::buildCV / buildRightCV
int orc_cc_build_cv(const char* name, const double* l_rgb, const double* r_rgb, int w, int h,
                    int max_dis, int right_view, double* vol) {
  CCMethod* cc = GetCCType(name);
  if (!cc) return -1;
  std::fill(vol, vol + (size_t)max_dis * h * w, 0.0);
  if (right_view) cc->buildRightCV(l_rgb, r_rgb, w, h, max_dis, vol);
  else cc->buildCV(l_rgb, r_rgb, w, h, max_dis, vol);
  delete cc;
  return 0;
}

// IPlaneCost (scale_num == 0 -> PreSSPC, else PreCSPC)
void* orc_pc_create(const uint8_t* l, const uint8_t* r, int w, int h, int stride, int max_disp,
                    int wnd_size, int scale_num, double reg_lambda, const char* cc_name) {
  if (std::string(cc_name) == "IMG")  // GrdPC (scale_num == 0) / CSPC: no CCMethod involved
    return (IPlaneCost*)new ImgPC(load_img(l, w, h, stride), load_img(r, w, h, stride), max_disp, wnd_size, scale_num, reg_lambda);
  CCMethod* cc = GetCCType(cc_name);
  if (!cc) return nullptr;
  Img8 li = load_img(l, w, h, stride), ri = load_img(r, w, h, stride);
  IPlaneCost* pc;
  if (scale_num == 0) pc = new PreSSPC(li, ri, max_disp, wnd_size, cc);
  else pc = new PreCSPC(li, ri, max_disp, wnd_size, scale_num, cc, reg_lambda);
  delete cc;
  return pc;
}
void orc_pc_destroy(void* pc) { delete (IPlaneCost*)pc; }
void orc_pc_get_plane_cost(const void* pc, int n, const int* x, const int* y, const double* normal,
                           const double* point, const int* view, double* out, int nthreads) {
  const IPlaneCost* p = (const IPlaneCost*)pc;
  if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 64)
  for (int i = 0; i < n; ++i) {
    Plane pl;
    pl.set_norm(normal[3 * i], normal[3 * i + 1], normal[3 * i + 2]);
    pl.set_point(point[3 * i], point[3 * i + 1], point[3 * i + 2]);
    pl.update_param();
    out[i] = p->GetPlaneCost(x[i], y[i], pl, (RefView)view[i]);
  }
}
// scale_wgt[S], max_cost[2][S], dims[S][3]=(w,h,max_disp).  S=1 for PreSSPC.
int orc_pc_info(const void* pc, double* scale_wgt, double* max_cost, int* dims) {
  const IPlaneCost* p = (const IPlaneCost*)pc;
  if (auto ss = dynamic_cast<const PreSSPC*>(p)) {
    if (scale_wgt) scale_wgt[0] = 1.0;
    if (max_cost) { max_cost[0] = ss->s0.max_cost[0]; max_cost[1] = ss->s0.max_cost[1]; }
    if (dims) { dims[0] = ss->s0.w; dims[1] = ss->s0.h; dims[2] = ss->s0.max_disp; }
    return 1;
  }
  if (auto ip = dynamic_cast<const ImgPC*>(p)) {
    const int S = (int)ip->lv.size();
    for (int s = 0; s < S; ++s) {
      if (scale_wgt) scale_wgt[s] = ip->scale_wgt[s];
      if (max_cost) { max_cost[s] = 2.8; max_cost[S + s] = 2.8; }
      if (dims) { dims[3 * s] = ip->lv[s].w; dims[3 * s + 1] = ip->lv[s].h; dims[3 * s + 2] = ip->lv[s].max_disp; }
    }
    return S;
  }
  auto cs = dynamic_cast<const PreCSPC*>(p);
  int S = cs->scale_num;
  for (int s = 0; s < S; ++s) {
    if (scale_wgt) scale_wgt[s] = cs->scale_wgt[s];
    if (max_cost) { max_cost[s] = cs->sc[s].max_cost[0]; max_cost[S + s] = cs->sc[s].max_cost[1]; }
    if (dims) { dims[3 * s] = cs->sc[s].w; dims[3 * s + 1] = cs->sc[s].h; dims[3 * s + 2] = cs->sc[s].max_disp; }
  }
  return S;
}
// pyramid image of scale s (BGR interleaved, dense)
int orc_pc_scale_image(const void* pc, int s, int view, uint8_t* out) {
  const IPlaneCost* p = (const IPlaneCost*)pc;
  const Img8* im = nullptr;
  if (auto ss = dynamic_cast<const PreSSPC*>(p)) im = &ss->s0.img[view];
  else if (auto ip = dynamic_cast<const ImgPC*>(p)) im = &ip->lv[s].img[view];
  else im = &dynamic_cast<const PreCSPC*>(p)->sc[s].img[view];
  std::memcpy(out, im->d.data(), im->d.size());
  return 0;
}

// CSPatchMatch
void* orc_pm_create(const uint8_t* l, const uint8_t* r, int w, int h, int stride, int max_dis,
                    int dis_scale, uint64_t seed) {
  return new CSPatchMatch(load_img(l, w, h, stride), load_img(r, w, h, stride), max_dis, dis_scale, seed);
}
void orc_pm_destroy(void* pm) { delete (CSPatchMatch*)pm; }
// 0 raster (reference order), 1 / 2 red-black with 4 / 8 neighbours (see SpatialPropagationRedBlack)
void orc_pm_set_schedule(void* pm, int schedule) { ((CSPatchMatch*)pm)->schedule_ = schedule; }
int orc_pm_n_ref(const void* pm) { return ((const CSPatchMatch*)pm)->n_ref(); }
void orc_pm_patchmatch(void* pm, int iter_num, const void* pc, int use_pp, int nthreads) {
  if (nthreads > 0) omp_set_num_threads(nthreads);
  ((CSPatchMatch*)pm)->PatchMatch(iter_num, (const IPlaneCost*)pc, use_pp != 0);
}
// stage-level entry points: 0 init, 1 spatial, 2 view, 3 refine, 4 plane2disp, 5 postprocess
void orc_pm_stage(void* pm_, int stage, int iter, const void* pc_, int nthreads) {
  CSPatchMatch* pm = (CSPatchMatch*)pm_;
  const IPlaneCost* pc = (const IPlaneCost*)pc_;
  if (nthreads > 0) omp_set_num_threads(nthreads);
  switch (stage) {
    case 0: pm->InitRandomPlane(pc); break;
    case 1: pm->Spatial(iter, pc); break;
    case 2: pm->ViewPropagation(iter, pc); break;
    case 3: pm->PlaneRefinement(iter, pm->max_dis_ / 2.0, pm->kMaxNorm_, pm->kZStopThres_, pc); break;
    case 4: pm->PlaneToDisp(); break;
    case 5: pm->PostProcessing(); break;
  }
}
void orc_pm_get_dis(const void* pm_, int view, uint8_t* out) {
  const CSPatchMatch* pm = (const CSPatchMatch*)pm_;
  std::memcpy(out, pm->dis_[view].data(), pm->dis_[view].size());
}
// normal[h*w*3], param[h*w*3], min_cost[h*w]
void orc_pm_get_planes(const void* pm_, int view, double* normal, double* param, double* min_cost) {
  const CSPatchMatch* pm = (const CSPatchMatch*)pm_;
  size_t n = (size_t)pm->wid_ * pm->hei_;
  for (size_t i = 0; i < n; ++i) {
    for (int k = 0; k < 3; ++k) {
      if (normal) normal[3 * i + k] = pm->plane_[view][i].norm[k];
      if (param) param[3 * i + k] = pm->plane_[view][i].param[k];
    }
    if (min_cost) min_cost[i] = pm->min_cost_[view][i];
  }
}
void orc_pm_set_planes(void* pm_, int view, const double* normal, const double* param, const double* min_cost) {
  CSPatchMatch* pm = (CSPatchMatch*)pm_;
  size_t n = (size_t)pm->wid_ * pm->hei_;
  for (size_t i = 0; i < n; ++i) {
    for (int k = 0; k < 3; ++k) {
      pm->plane_[view][i].norm[k] = normal[3 * i + k];
      pm->plane_[view][i].param[k] = param[3 * i + k];
    }
    pm->min_cost_[view][i] = min_cost[i];
  }
}
void orc_pm_set_dis(void* pm_, int view, const uint8_t* dis) {
  CSPatchMatch* pm = (CSPatchMatch*)pm_;
  std::memcpy(pm->dis_[view].data(), dis, pm->dis_[view].size());
}

}  // extern "C"
