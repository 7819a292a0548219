// oracle/cvshim/opencv2/opencv.hpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Minimal stand-in for the slice of OpenCV 2.4.8 that the reference hot path
// uses (CSPM/commfunc.h:9 includes <opencv2/opencv.hpp>; OpenCV C++ is not
// installed here and cannot be fetched).  With this header on the include
// path the UNMODIFIED reference sources
//   CSPM/cs_patchmatch.cc, plane_cost/pre_ss_pc.cc, plane_cost/pre_cs_pc.cc,
//   cc/grd_cc.cpp, cc/cen_cc.cc
// compile where they lie under /root/reference into oracle/_ref/libcspm_ref.so
// (recipe: oracle/Makefile).  Everything numerical in here delegates to
// oracle/prims.h, whose functions are checked against cv2-python 4.13
// (tests/golden/cv2_prims.npz).
//
// Deliberate difference from OpenCV: cv::RNG is the counter-based Philox
// generator of prims.h, and time() is interposed so that every
// `RNG rng(time(NULL))` in cs_patchmatch.cc:32,130,309 receives a
// deterministic stream id (the running call count) instead of the wall clock.
#pragma once
#include <omp.h>

#include <algorithm>
#include <atomic>
#include <bitset>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <iomanip>
#include <iostream>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../prims.h"

// ---- deterministic stand-in for time(NULL) --------------------------------
// mode 0: call counter — the RNG contract's stream id in a single-threaded run (parity, 1 thread)
// mode 1: constant, like a real wall clock within one second (timing runs)
// mode 2: the SAME stream ids as mode 0 with the reference's OpenMP row loops running in parallel.  Every
//         `#pragma omp parallel for` over rows (cs_patchmatch.cc:126,305) makes exactly H calls, one per row, so
//         the call count before a call, divided by H, is the index R of the loop it belongs to whatever the
//         interleaving; libgomp's default static schedule gives thread t the contiguous rows starting at
//         t*q + min(t, r) (q = H / nt, r = H % nt), in order, so row = start + (calls this thread made in loop R).
//         tests/test_oracle_vs_ref.py checks mode 2 on 1..N threads == mode 0 bit for bit.
inline uint64_t g_cspm_shim_seed = 0x5EED;
inline std::atomic<uint64_t> g_cspm_shim_time_calls{0};
inline int g_cspm_shim_time_mode = 0;
inline uint64_t g_cspm_shim_rows = 0;  // image height (mode 2), set by ref_pm_create
inline uint64_t g_cspm_shim_epoch = 0;  // bumped whenever the wrapper rewinds the call counter
inline uint64_t cspm_shim_time() {
  if (g_cspm_shim_time_mode == 1) return 1;
  const uint64_t c = g_cspm_shim_time_calls.fetch_add(1);
  if (g_cspm_shim_time_mode != 2 || !omp_in_parallel() || c == 0 || g_cspm_shim_rows == 0) return c;
  const uint64_t H = g_cspm_shim_rows, R = (c - 1) / H;
  static thread_local uint64_t loop = ~0ull, epoch = ~0ull, k = 0;
  if (loop != R || epoch != g_cspm_shim_epoch) { loop = R; epoch = g_cspm_shim_epoch; k = 0; }
  const uint64_t nt = (uint64_t)omp_get_num_threads(), t = (uint64_t)omp_get_thread_num();
  const uint64_t q = H / nt, r = H % nt, start = t * q + (t < r ? t : r);
  return 1 + R * H + start + k++;
}
#define time(x) cspm_shim_time()

#define CV_8U 0
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_64FC3 CV_MAKETYPE(CV_64F, 3)
#define CV_BGR2RGB 4
#define CV_BGR2GRAY 6
#define CV_RGB2GRAY 7
#define CV_BGR2Lab 44
#define CV_Assert(expr)                                                          \
  do {                                                                           \
    if (!(expr)) throw std::runtime_error(std::string("CV_Assert failed: ") + #expr); \
  } while (0)

typedef unsigned char uchar;
typedef uint64_t uint64;

namespace cv {

enum { NORM_L2 = 4 };

template <typename T>
static inline T saturate_cast(int v);
template <>
inline uchar saturate_cast<uchar>(int v) {
  return (uchar)((unsigned)v <= 255 ? v : v > 0 ? 255 : 0);
}
static inline int cvRound(double v) { return (int)std::nearbyint(v); }  // round-half-even (SSE2 cvtsd2si)

template <typename T> struct Point3_;

template <typename T, int cn>
struct Vec {
  T val[cn];
  Vec() { for (int i = 0; i < cn; ++i) val[i] = T(0); }
  Vec(T a, T b, T c) { static_assert(cn == 3, "3-vector"); val[0] = a; val[1] = b; val[2] = c; }
  T& operator[](int i) { return val[i]; }
  const T& operator[](int i) const { return val[i]; }
  // Matx::dot: s = 0; s += val[i]*M.val[i]
  T dot(const Vec& m) const {
    T s = 0;
    for (int i = 0; i < cn; ++i) s += val[i] * m.val[i];
    return s;
  }
};
typedef Vec<double, 3> Vec3d;

template <typename T, int cn>
static inline Vec<T, cn> operator+(const Vec<T, cn>& a, const Vec<T, cn>& b) {
  Vec<T, cn> r;
  for (int i = 0; i < cn; ++i) r.val[i] = a.val[i] + b.val[i];
  return r;
}
// OpenCV 2.4 operations.hpp: Vec / double == Vec(a, 1./alpha, Matx_ScaleOp())
template <typename T, int cn>
static inline Vec<T, cn> operator/(const Vec<T, cn>& a, double alpha) {
  Vec<T, cn> r;
  double s = 1. / alpha;
  for (int i = 0; i < cn; ++i) r.val[i] = (T)(a.val[i] * s);
  return r;
}
// cv::norm(Matx, NORM_L2) = sqrt(normL2Sqr): s += v*v left to right
template <typename T, int cn>
static inline double norm(const Vec<T, cn>& v, int /*normType*/) {
  double s = 0;
  for (int i = 0; i < cn; ++i) s += (double)v.val[i] * (double)v.val[i];
  return std::sqrt(s);
}

template <typename T>
struct Point3_ {
  T x, y, z;
  Point3_() : x(0), y(0), z(0) {}
  Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
  Point3_(const Vec<T, 3>& v) : x(v[0]), y(v[1]), z(v[2]) {}
  operator Vec<T, 3>() const { return Vec<T, 3>(x, y, z); }
};
typedef Point3_<double> Point3d;

class Mat {
 public:
  int rows = 0, cols = 0;
  uchar* data = nullptr;
  size_t step = 0;  // bytes per row
  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  void create(int r, int c, int type) {
    rows = r; cols = c; type_ = type;
    step = (size_t)c * elemSize();
    buf_ = std::make_shared<std::vector<uchar>>((size_t)r * step, (uchar)0);
    data = buf_->data();
  }
  static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
  // wrap external memory (no ownership)
  Mat(int r, int c, int type, void* ext, size_t step_) {
    rows = r; cols = c; type_ = type; step = step_; data = (uchar*)ext;
  }
  int type() const { return type_; }
  int depth() const { return type_ & 7; }
  int channels() const { return (type_ >> 3) + 1; }
  size_t elemSize1() const { int d = depth(); return d == CV_8U ? 1 : d == CV_32F ? 4 : 8; }
  size_t elemSize() const { return elemSize1() * channels(); }
  bool empty() const { return data == nullptr; }
  Mat clone() const {
    Mat m(rows, cols, type_);
    for (int y = 0; y < rows; ++y) std::memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, (size_t)cols * elemSize());
    return m;
  }
  template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step); }
  template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step); }
  template <typename T> T& at(int y, int x) { return ((T*)(data + (size_t)y * step))[x]; }
  template <typename T> const T& at(int y, int x) const { return ((const T*)(data + (size_t)y * step))[x]; }
  void convertTo(Mat& dst, int rtype) const {
    int ddepth = rtype & 7;
    Mat out(rows, cols, CV_MAKETYPE(ddepth, channels()));
    const int n = cols * channels();
    for (int y = 0; y < rows; ++y) {
      for (int i = 0; i < n; ++i) {
        double v;
        switch (depth()) {
          case CV_8U: v = ptr<uchar>(y)[i]; break;
          case CV_32F: v = ptr<float>(y)[i]; break;
          default: v = ptr<double>(y)[i]; break;
        }
        switch (ddepth) {
          case CV_8U: out.ptr<uchar>(y)[i] = saturate_cast<uchar>(cvRound(v)); break;
          case CV_32F: out.ptr<float>(y)[i] = (float)v; break;
          default: out.ptr<double>(y)[i] = v; break;
        }
      }
    }
    dst = out;
  }
  Mat inv() const {
    CV_Assert(type_ == CV_64FC1 && rows == cols);
    std::vector<double> a((size_t)rows * cols), o((size_t)rows * cols, 0.0);
    for (int y = 0; y < rows; ++y)
      for (int x = 0; x < cols; ++x) a[(size_t)y * cols + x] = at<double>(y, x);
    Mat r(rows, cols, CV_64FC1);
    if (cspm_prims::invert_lu(a.data(), rows, o.data()))
      for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) r.at<double>(y, x) = o[(size_t)y * cols + x];
    return r;
  }

 private:
  int type_ = 0;
  std::shared_ptr<std::vector<uchar>> buf_;
};

static inline void cvtColor(const Mat& src, Mat& dst, int code) {
  if (code == CV_BGR2RGB) {
    CV_Assert(src.channels() == 3 && src.depth() == CV_8U);
    Mat out(src.rows, src.cols, src.type());
    for (int y = 0; y < src.rows; ++y) {
      const uchar* s = src.ptr<uchar>(y);
      uchar* d = out.ptr<uchar>(y);
      for (int x = 0; x < src.cols; ++x) {
        d[3 * x] = s[3 * x + 2]; d[3 * x + 1] = s[3 * x + 1]; d[3 * x + 2] = s[3 * x];
      }
    }
    dst = out;
  } else if (code == CV_RGB2GRAY && src.depth() == CV_32F) {
    Mat out(src.rows, src.cols, CV_32FC1);
    for (int y = 0; y < src.rows; ++y) {
      const float* s = src.ptr<float>(y);
      float* d = out.ptr<float>(y);
      for (int x = 0; x < src.cols; ++x) d[x] = cspm_prims::rgb2gray_f32(s[3 * x], s[3 * x + 1], s[3 * x + 2]);
    }
    dst = out;
  } else if ((code == CV_RGB2GRAY || code == CV_BGR2GRAY) && src.depth() == CV_8U) {
    // OpenCV 2.4.8 RGB2Gray<uchar>: yuv_shift = 14, R2Y=4899, G2Y=9617, B2Y=1868
    Mat out(src.rows, src.cols, CV_8UC1);
    const int ri = code == CV_RGB2GRAY ? 0 : 2, bi = 2 - ri;
    for (int y = 0; y < src.rows; ++y) {
      const uchar* s = src.ptr<uchar>(y);
      uchar* d = out.ptr<uchar>(y);
      for (int x = 0; x < src.cols; ++x)
        d[x] = cspm_prims::gray8_fixed(s[3 * x + ri], s[3 * x + 1], s[3 * x + bi]);
    }
    dst = out;
  } else if (code == CV_BGR2Lab) {
    // grd_pc.cc:32 / cspc.cc:41 compute lab_ but never read it (USE_LAB_WGT is off): any image will do.
    dst = src.clone();
  } else {
    throw std::runtime_error("cvshim: unsupported cvtColor code");
  }
}

static inline void Sobel(const Mat& src, Mat& dst, int ddepth, int dx, int dy, int ksize) {
  if (src.type() == CV_8UC1) {  // grd_pc.cc:40, cspc.cc:47
    CV_Assert(ddepth == CV_64F && dx == 1 && dy == 0 && ksize == 1);
    std::vector<uchar> g8((size_t)src.rows * src.cols);
    for (int y = 0; y < src.rows; ++y) std::memcpy(&g8[(size_t)y * src.cols], src.ptr<uchar>(y), src.cols);
    Mat out8(src.rows, src.cols, CV_64FC1);
    cspm_prims::sobel_x_k1_u8(g8.data(), src.cols, src.rows, out8.ptr<double>(0));
    dst = out8;
    return;
  }
  CV_Assert(src.type() == CV_32FC1 && ddepth == CV_64F && dx == 1 && dy == 0 && ksize == 1);
  std::vector<float> g((size_t)src.rows * src.cols);
  for (int y = 0; y < src.rows; ++y) std::memcpy(&g[(size_t)y * src.cols], src.ptr<float>(y), sizeof(float) * src.cols);
  Mat out(src.rows, src.cols, CV_64FC1);
  cspm_prims::sobel_x_k1(g.data(), src.cols, src.rows, out.ptr<double>(0));
  dst = out;
}

static inline void pyrDown(const Mat& src, Mat& dst) {
  CV_Assert(src.type() == CV_8UC3);
  Mat out((src.rows + 1) / 2, (src.cols + 1) / 2, CV_8UC3);
  cspm_prims::pyrdown_u8c3(src.data, src.cols, src.rows, (int)src.step, out.data, out.cols, out.rows, (int)out.step);
  dst = out;
}

static inline void minMaxLoc(const Mat& m, double* minVal, double* maxVal) {
  CV_Assert(m.type() == CV_64FC1);
  double mn = std::numeric_limits<double>::max(), mx = -std::numeric_limits<double>::max();
  for (int y = 0; y < m.rows; ++y) {
    const double* p = m.ptr<double>(y);
    for (int x = 0; x < m.cols; ++x) {
      if (p[x] < mn) mn = p[x];
      if (p[x] > mx) mx = p[x];
    }
  }
  if (minVal) *minVal = mn;
  if (maxVal) *maxVal = mx;
}

class RNG {
 public:
  enum { UNIFORM = 0, NORMAL = 1 };
  RNG() : stream_(0xffffffffull), slot_(0) {}
  RNG(uint64 state) : stream_(state), slot_(0) {}
  double uniform(double a, double b) {
    return cspm_prims::rng_uniform(g_cspm_shim_seed, stream_, slot_++, 0, a, b);
  }
  void fill(Vec3d& v, int distType, double a, double b) {
    if (distType == NORMAL) cspm_prims::rng_fill_normal3(g_cspm_shim_seed, stream_, slot_++, a, b, v.val);
    else cspm_prims::rng_fill_uniform3(g_cspm_shim_seed, stream_, slot_++, a, b, v.val);
  }

 private:
  uint64 stream_;
  uint32_t slot_;
};

static inline int64_t getTickCount() { return 0; }
static inline double getTickFrequency() { return 1.0; }

}  // namespace cv
