// cspm_prep.cu — CCMethod / IPlaneCost side of libcspm_b200.so:
// image packing, float-gray x-gradient, pyrDown pyramid, raw-cost maximum,
// cross-scale weights, the batched GetPlaneCost entry point and the
// buildCV / buildRightCV parity export.
//
// Reference: CSPM/cc/grd_cc.cpp, CSPM/plane_cost/pre_ss_pc.cc:12-65,
// CSPM/plane_cost/pre_cs_pc.cc:12-115.
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include <math_constants.h>

#include "cspm_internal.h"

namespace cspm {

static thread_local std::string g_last_error;
static thread_local int g_last_error_code = CSPM_OK;
void set_last_error(const std::string& msg, int code) {
  g_last_error = msg;
  g_last_error_code = code;
}

static thread_local cudaStream_t g_thread_stream = nullptr;
static thread_local bool g_thread_stream_set = false;
cudaStream_t acquire_stream(bool* owned) {
  if (g_thread_stream_set) {
    *owned = false;
    return g_thread_stream;
  }
  cudaStream_t st;
  CSPM_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  *owned = true;
  return st;
}

void use_device(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    throw Error(CSPM_ENODEV, std::string("no CUDA device available (libcspm_b200 has no CPU fallback): ") +
                                 (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
  if (device < 0 || device >= n) throw Error(CSPM_EINVAL, "device index out of range");
  CSPM_CUDA(cudaSetDevice(device));
  static bool pool_set[64] = {};
  if (device < 64 && !pool_set[device]) {  // keep freed blocks in the pool (see DeviceArena)
    cudaMemPool_t pool;
    CSPM_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    unsigned long long keep = ~0ull;
    CSPM_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    pool_set[device] = true;
  }
}

// ------------------------------------------------------------------ kernels
// cvtColor(RGB2GRAY) on CV_32F (OpenCV 2.4.8 RGB2Gray<float>, no FMA).
__device__ __forceinline__ float gray_f32(float r, float g, float b) {
  return __fadd_rn(__fadd_rn(__fmul_rn(r, 0.299f), __fmul_rn(g, 0.587f)), __fmul_rn(b, 0.114f));
}
__device__ __forceinline__ float gray_of_bgr8(const uint8_t* p) {
  return gray_f32((float)p[2], (float)p[1], (float)p[0]);
}

// K1a: dense BGR8 -> padded packed colour + padded f64 x-gradient
// (Sobel ksize=1, REFLECT_101: exactly 0 in the first and last column).
// cg (GRD only, else null): the same two values in one 64-bit word, (G * 2^27 + 2^37) | colour << 40 — see
// cspm_device.cuh "Packed GRD pixel".  *inexact counts pixels whose gradient is not an integer multiple of
// 2^-27 (cannot happen for float gray of 8-bit images; the host refuses to continue if it ever does).
__global__ void k_pack(const uint8_t* __restrict__ bgr, int w, int h, int pad, int pitch,
                       uint32_t* __restrict__ color, double* __restrict__ grad, unsigned long long* __restrict__ cg,
                       unsigned long long* __restrict__ inexact) {
  const int xp = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (xp >= pitch) return;
  const int x = xp - pad;
  const size_t o = (size_t)y * pitch + xp;
  uint32_t c = 0x00030303u;  // BORDER_THRES for every channel (cc/grd_cc.cpp:21-35)
  double g = 3.0;
  if (x >= 0 && x < w) {
    const uint8_t* row = bgr + (size_t)y * 3 * w;
    const uint8_t* p = row + 3 * x;
    c = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
    g = 0.0;
    if (x > 0 && x < w - 1) g = __dsub_rn((double)gray_of_bgr8(p + 3), (double)gray_of_bgr8(p - 3));
  }
  color[o] = c;
  grad[o] = g;
  if (cg) {
    const double scaled = __dmul_rn(g, 134217728.0);  // * 2^27: exact
    const long long gi = __double2ll_rn(scaled);
    if (__ll2double_rn(gi) != scaled || gi >= (1ll << 37) || gi < -(1ll << 37)) atomicAdd(inexact, 1ull);
    cg[o] = (unsigned long long)(gi + (1ll << 37)) | ((unsigned long long)c << 40);
  }
}

// K1a': the same for GrdPC / CSPC (grd_pc.cc:37-40): cvtColor(BGR2GRAY) on 8 bits (14-bit fixed point)
// then Sobel(ksize=1) -> integer-valued f64 gradient.  Partner reads wrap around, so the pad is unused.
__device__ __forceinline__ int gray8_of_bgr8(const uint8_t* p) { return (p[2] * 4899 + p[1] * 9617 + p[0] * 1868 + 8192) >> 14; }
__global__ void k_pack_img(const uint8_t* __restrict__ bgr, int w, int h, int pad, int pitch,
                           uint32_t* __restrict__ color, double* __restrict__ grad) {
  const int xp = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (xp >= pitch) return;
  const int x = xp - pad;
  const size_t o = (size_t)y * pitch + xp;
  if (x < 0 || x >= w) {
    color[o] = 0u;
    grad[o] = 0.0;
    return;
  }
  const uint8_t* p = bgr + ((size_t)y * w + x) * 3;
  color[o] = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
  grad[o] = (x > 0 && x < w - 1) ? (double)(gray8_of_bgr8(p + 3) - gray8_of_bgr8(p - 3)) : 0.0;
}

// K1b: pyrDown on CV_8UC3 — separable [1 4 6 4 1], REFLECT_101, (sum+128)>>8.
__device__ __forceinline__ int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
  return p;
}
__global__ void k_pyrdown(const uint8_t* __restrict__ src, int w, int h, uint8_t* __restrict__ dst, int dw, int dh) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= dw || y >= dh) return;
  const int k[5] = {1, 4, 6, 4, 1};
  int sum[3] = {0, 0, 0};
  for (int j = -2; j <= 2; ++j) {
    const uint8_t* row = src + (size_t)reflect101(2 * y + j, h) * 3 * w;
    int r[3] = {0, 0, 0};
    for (int i = -2; i <= 2; ++i) {
      const uint8_t* p = row + 3 * reflect101(2 * x + i, w);
      r[0] += k[i + 2] * p[0];
      r[1] += k[i + 2] * p[1];
      r[2] += k[i + 2] * p[2];
    }
    sum[0] += k[j + 2] * r[0];
    sum[1] += k[j + 2] * r[1];
    sum[2] += k[j + 2] * r[2];
  }
  uint8_t* o = dst + ((size_t)y * dw + x) * 3;
  o[0] = (uint8_t)((sum[0] + 128) >> 8);
  o[1] = (uint8_t)((sum[1] + 128) >> 8);
  o[2] = (uint8_t)((sum[2] + 128) >> 8);
}

// CenCC (cc/cen_cc.cc:4-47): 8-bit gray (OpenCV 2.4.8 RGB2Gray<uchar>: 14-bit fixed point) and the
// 9x9 census transform with WRAP-AROUND window borders; bit k of the 80-bit code, k counting the
// window positions row-major without the centre, is (centre > neighbour).
__device__ __forceinline__ uint8_t gray8(int r, int g, int b) { return (uint8_t)((r * 4899 + g * 9617 + b * 1868 + 8192) >> 14); }
__global__ void k_gray8_bgr(const uint8_t* __restrict__ bgr, int n, uint8_t* __restrict__ gray) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) gray[i] = gray8(bgr[3 * i + 2], bgr[3 * i + 1], bgr[3 * i]);
}
__global__ void k_census(const uint8_t* __restrict__ gray, int w, int h, int pad, int pitch,
                         unsigned long long* __restrict__ lo, uint32_t* __restrict__ hi) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= w) return;
  const uint8_t c = gray[(size_t)y * w + x];
  unsigned long long l = 0;
  uint32_t hbits = 0;
  int k = 0;
  for (int wy = -4; wy <= 4; ++wy) {
    const int qy = (y + wy + h) % h;
    for (int wx = -4; wx <= 4; ++wx) {
      if (wy == 0 && wx == 0) continue;
      const int qx = (x + wx + w) % w;
      const unsigned long long bit = c > gray[(size_t)qy * w + qx] ? 1ull : 0ull;
      if (k < 64) l |= bit << k;
      else hbits |= (uint32_t)bit << (k - 64);
      ++k;
    }
  }
  const size_t o = (size_t)y * pitch + pad + x;
  lo[o] = l;
  hi[o] = hbits;
}
// buildCV / buildRightCV export for CenCC on arbitrary CV_64FC3 input: convertTo(CV_8U) is
// saturate_cast<uchar>(cvRound(v)) (round half to even), cen_cc.cc:13-17.
__global__ void k_gray8_f64(const double* __restrict__ rgb, int n, uint8_t* __restrict__ gray) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int ch[3];
  for (int k = 0; k < 3; ++k) ch[k] = min(max(__double2int_rn(rgb[3 * i + k]), 0), 255);
  gray[i] = gray8(ch[0], ch[1], ch[2]);
}
__global__ void k_build_cv_cen(const unsigned long long* __restrict__ llo, const uint32_t* __restrict__ lhi,
                               const unsigned long long* __restrict__ rlo, const uint32_t* __restrict__ rhi, int w, int h,
                               int right_view, double* __restrict__ vol) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, d = blockIdx.z;
  if (x >= w) return;
  const size_t row = (size_t)y * w;
  double c = 80.0;  // CENCUS_BIT
  if (!right_view) {
    if (x - d >= 0) c = (double)(__popcll(llo[row + x] ^ rlo[row + x - d]) + __popc((lhi[row + x] ^ rhi[row + x - d]) & 0xFFFFu));
  } else {
    if (x + d < w) c = (double)(__popcll(rlo[row + x] ^ llo[row + x + d]) + __popc((rhi[row + x] ^ lhi[row + x + d]) & 0xFFFFu));
  }
  vol[((size_t)d * h + y) * w + x] = c;
}

// K2: max over the whole raw cost volume d in [0, max_disp], all pixels
// (replaces the minMaxLoc loop pre_ss_pc.cc:51-58).  Costs are >= 0, so the
// f64 bit pattern orders like an unsigned integer.
template <int CC>
__global__ void k_max_raw_cost(PcDev pc, int s, int v, unsigned long long* out_bits) {
  __shared__ double s_lut_c[kLutC];
  for (int i = threadIdx.x; i < kLutC; i += blockDim.x) s_lut_c[i] = pc.lut_c[i];
  __syncthreads();
  const ScaleDev& sc = pc.sc[s];
  const TapCtx t = make_tap_ctx(pc, s, v, pc.lut_w, s_lut_c);
  double m = 0.0;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x < sc.w) {
    const int row_off = y * sc.pitch + sc.pad;
    const uint32_t cq = t.col_v[row_off + x];
    const double gq = t.grd_v[row_off + x];
    const uint32_t qhi = CC == CC_CEN ? t.hi_v[row_off + x] : 0u;
    for (int d = 0; d <= sc.max_disp; ++d) m = fmax(m, raw_cost<CC>(t.raw, row_off, x, cq, gq, qhi, d));
  }
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out_bits, (unsigned long long)__double_as_longlong(m));
}

// K3: batched IPlaneCost::GetPlaneCost, one thread per query.
template <int CC>
__global__ void k_plane_cost_batch(PcDev pc, int n, const int* __restrict__ x, const int* __restrict__ y,
                                   const double* __restrict__ normal, const double* __restrict__ point,
                                   const int* __restrict__ view, double* __restrict__ out) {
  __shared__ double s_lut_w[kLutW];
  __shared__ double s_lut_c[kLutC];
  stage_luts(pc, s_lut_w, s_lut_c);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  PlaneD p;
  p.nx = normal[3 * i]; p.ny = normal[3 * i + 1]; p.nz = normal[3 * i + 2];
  update_param(p, point[3 * i], point[3 * i + 1], point[3 * i + 2]);
  out[i] = plane_cost<CC>(pc, view[i], x[i], y[i], p, s_lut_w, s_lut_c, CUDART_INF);
}

// buildCV / buildRightCV export for arbitrary CV_64FC3 input (cc/grd_cc.cpp:60-154).
__global__ void k_gray_f64(const double* __restrict__ rgb, int n, float* __restrict__ gray) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) gray[i] = gray_f32((float)rgb[3 * i], (float)rgb[3 * i + 1], (float)rgb[3 * i + 2]);
}
__global__ void k_sobel_x(const float* __restrict__ gray, int w, int h, double* __restrict__ grd) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= w) return;
  double g = 0.0;
  if (x > 0 && x < w - 1) g = __dsub_rn((double)gray[(size_t)y * w + x + 1], (double)gray[(size_t)y * w + x - 1]);
  grd[(size_t)y * w + x] = g;
}
__device__ __forceinline__ double cost_grd_f64(const double* c, const double* o, double g, double go) {
  double clr = 0.0;
  for (int k = 0; k < 3; ++k) clr = dadd(clr, fabs(dsub(c[k], o[k])));
  clr = dmul(clr, 0.3333333333);
  double grd = fabs(dsub(g, go));
  clr = clr > 10.0 ? 10.0 : clr;
  grd = grd > 2.0 ? 2.0 : grd;
  return dadd(dmul(0.1, clr), dmul(1 - 0.1, grd));
}
__global__ void k_build_cv_grd(const double* __restrict__ l, const double* __restrict__ r,
                               const double* __restrict__ lg, const double* __restrict__ rg, int w, int h,
                               int max_dis, int right_view, double* __restrict__ vol) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, d = blockIdx.z;
  if (x >= w) return;
  const double three[3] = {3.0, 3.0, 3.0};
  const size_t row = (size_t)y * w;
  double c;
  if (!right_view) {
    if (x - d >= 0) c = cost_grd_f64(l + 3 * (row + x), r + 3 * (row + x - d), lg[row + x], rg[row + x - d]);
    else c = cost_grd_f64(l + 3 * (row + x), three, lg[row + x], 3.0);
  } else {
    if (x + d < w) c = cost_grd_f64(l + 3 * (row + x + d), r + 3 * (row + x), lg[row + x + d], rg[row + x]);
    else c = cost_grd_f64(r + 3 * (row + x), three, rg[row + x], 3.0);
  }
  vol[((size_t)d * h + y) * w + x] = c;
}

// ------------------------------------------------------------------- host
// Row 0 of the inverse of the tridiagonal cross-scale regulariser
// (pre_cs_pc.cc:89-109).  Mirrors OpenCV's invert(DECOMP_LU): closed forms
// for n <= 3, LU with partial pivoting above.
static void scale_weights(int n, double lambda, std::vector<double>* out) {
  std::vector<double> A((size_t)n * n, 0.0);
  for (int s = 0; s < n; ++s) {
    if (s == 0) {
      A[0] = 1 + lambda;
      A[1] = -lambda;
    } else if (s == n - 1) {
      A[(size_t)s * n + s] = 1 + lambda;
      A[(size_t)s * n + s - 1] = -lambda;
    } else {
      A[(size_t)s * n + s] = 1 + 2 * lambda;
      A[(size_t)s * n + s - 1] = -lambda;
      A[(size_t)s * n + s + 1] = -lambda;
    }
  }
  out->assign(n, 0.0);
  if (n == 2) {
    double d = A[0] * A[3] - A[1] * A[2];
    d = 1.0 / d;
    (*out)[0] = A[3] * d;
    (*out)[1] = -A[1] * d;
    return;
  }
  if (n == 3) {
    auto S = [&](int r, int c) { return A[(size_t)r * 3 + c]; };
    double d = S(0, 0) * (S(1, 1) * S(2, 2) - S(1, 2) * S(2, 1)) - S(0, 1) * (S(1, 0) * S(2, 2) - S(1, 2) * S(2, 0)) +
               S(0, 2) * (S(1, 0) * S(2, 1) - S(1, 1) * S(2, 0));
    d = 1.0 / d;
    (*out)[0] = (S(1, 1) * S(2, 2) - S(1, 2) * S(2, 1)) * d;
    (*out)[1] = (S(0, 2) * S(2, 1) - S(0, 1) * S(2, 2)) * d;
    (*out)[2] = (S(0, 1) * S(1, 2) - S(0, 2) * S(1, 1)) * d;
    return;
  }
  // Gauss elimination with partial pivoting on [A | I], then back substitution.
  std::vector<double> B((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) B[(size_t)i * n + i] = 1.0;
  for (int i = 0; i < n; ++i) {
    int k = i;
    for (int j = i + 1; j < n; ++j)
      if (std::fabs(A[(size_t)j * n + i]) > std::fabs(A[(size_t)k * n + i])) k = j;
    if (std::fabs(A[(size_t)k * n + i]) < 2.220446049250313e-14) throw Error(CSPM_EINVAL, "singular scale regulariser");
    if (k != i) {
      for (int j = i; j < n; ++j) std::swap(A[(size_t)i * n + j], A[(size_t)k * n + j]);
      for (int j = 0; j < n; ++j) std::swap(B[(size_t)i * n + j], B[(size_t)k * n + j]);
    }
    const double d = -1.0 / A[(size_t)i * n + i];
    for (int j = i + 1; j < n; ++j) {
      const double alpha = A[(size_t)j * n + i] * d;
      for (int c = i + 1; c < n; ++c) A[(size_t)j * n + c] += alpha * A[(size_t)i * n + c];
      for (int c = 0; c < n; ++c) B[(size_t)j * n + c] += alpha * B[(size_t)i * n + c];
    }
    A[(size_t)i * n + i] = -d;
  }
  for (int i = n - 1; i >= 0; --i)
    for (int j = 0; j < n; ++j) {
      double s = B[(size_t)i * n + j];
      for (int c = i + 1; c < n; ++c) s -= A[(size_t)i * n + c] * B[(size_t)c * n + j];
      B[(size_t)i * n + j] = s * A[(size_t)i * n + i];
    }
  for (int s = 0; s < n; ++s) (*out)[s] = B[s];
}

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

static cspm_pc* pc_create(const uint8_t* l, const uint8_t* r, bool on_device, int w, int h, int stride,
                          int max_disp, int wnd_size, int scale_num, bool cross_scale, const cspm_cc* cc,
                          double reg_lambda, int device) {
  CSPM_REQUIRE(l && r, "null image pointer");
  CSPM_REQUIRE(cc, "null cc_method (GetCCType returned NULL for this name)");
  CSPM_REQUIRE(w > 0 && h > 0 && stride >= 3 * w, "bad image geometry");
  CSPM_REQUIRE(max_disp >= 1, "max_disp must be >= 1");
  CSPM_REQUIRE(wnd_size >= 1 && wnd_size <= 63, "wnd_size out of range [1,63]");
  CSPM_REQUIRE(!cross_scale || (scale_num >= 2 && scale_num <= kMaxScales), "scale_num out of range [2,8]");
  // GrdPC / CSPC wrap the interpolated partner column once (HandleBorder): beyond max_disp <= width the reference
  // reads outside its rows
  CSPM_REQUIRE(cc->kind != CC_IMG || max_disp <= w, "GrdPC/CSPC need max_disp <= image width");
  use_device(device);

  cspm_pc* pc = new cspm_pc();
  static std::atomic<uint64_t> next_uid{1};
  pc->uid = next_uid.fetch_add(1);
  try {
    pc->device = device;
    pc->wnd_size = wnd_size;
    pc->scale_num = cross_scale ? scale_num : 1;
    pc->stream = acquire_stream(&pc->own_stream);
    pc->arena.stream = pc->stream;
    cudaStream_t st = pc->stream;
    PcDev& D = pc->dev;
    D.cc_kind = cc->kind;
    D.cross_scale = cross_scale ? 1 : 0;
    D.scale_num = pc->scale_num;
    D.half_wnd = wnd_size / 2;
    D.k_grd = 1 - 0.1;  // (1 - ALPHA), cc/grd_cc.cpp:18
    D.cg_exp_hi = 0x43300000u;
    D.k_grd27 = std::ldexp(D.k_grd, -27);  // exact; pairs with gradient differences carried as dG * 2^27

    // lookup tables (host libm exp, exactly as pre_ss_pc.cc:61-64 builds them)
    std::vector<double> lut_w(kLutW), lut_c(kLutC);
    for (int i = 0; i < kLutW; ++i) lut_w[i] = std::exp(-i * 1.0 / 10.0);
    for (int k = 0; k < kLutC; ++k) {
      double clr = (double)k;         // sum of three exact integer |differences|
      clr *= 0.3333333333;            // cc/grd_cc.cpp:13
      clr = clr > 10.0 ? 10.0 : clr;  // TAU_CLR
      lut_c[k] = 0.1 * clr;           // ALPHA * clrDiff
    }
    double* d_lut_w = pc->arena.alloc<double>(kLutW);
    double* d_lut_c = pc->arena.alloc<double>(kLutC);
    CSPM_CUDA(cudaMemcpyAsync(d_lut_w, lut_w.data(), sizeof(double) * kLutW, cudaMemcpyHostToDevice, st));
    CSPM_CUDA(cudaMemcpyAsync(d_lut_c, lut_c.data(), sizeof(double) * kLutC, cudaMemcpyHostToDevice, st));
    CSPM_CUDA(cudaStreamSynchronize(st));  // host vectors go out of scope below
    D.lut_w = d_lut_w;
    D.lut_c = d_lut_c;
    {  // 16x replicated weight table + colour table, contiguous, for the bulk kernels' single TMA bulk copy
      std::vector<double> rep((size_t)kBulkMbarOff);
      for (int i = 0; i < kLutWUsed; ++i)
        for (int r = 0; r < (1 << kLutRepShift); ++r) rep[((size_t)i << kLutRepShift) + r] = lut_w[i];
      for (int k = 0; k < kLutC; ++k)
        for (int r = 0; r < (1 << kLutCRepShift); ++r) rep[(size_t)kBulkLutCOff + ((size_t)k << kLutCRepShift) + r] = lut_c[k];
      double* d_rep = pc->arena.alloc<double>(rep.size());
      CSPM_CUDA(cudaMemcpyAsync(d_rep, rep.data(), sizeof(double) * rep.size(), cudaMemcpyHostToDevice, st));
      CSPM_CUDA(cudaStreamSynchronize(st));
      D.lut_rep = d_rep;
    }

    // scratch: raw-cost maxima [2 * kMaxScales] + the packed-layout exactness counter
    unsigned long long* d_max = pc->arena.alloc<unsigned long long>(2 * kMaxScales + 1);
    unsigned long long* d_inexact = d_max + 2 * kMaxScales;
    CSPM_CUDA(cudaMemsetAsync(d_max, 0, sizeof(unsigned long long) * (2 * kMaxScales + 1), st));

    // pyramid of dense BGR images (pre_cs_pc.cc:42-49)
    int sw = w, sh = h, sd = max_disp;
    for (int s = 0; s < pc->scale_num; ++s) {
      if (s > 0) {
        sh = (sh + 1) / 2;
        sw = (sw + 1) / 2;
        sd = sd / 2;
      }
      pc->dims[s][0] = sw; pc->dims[s][1] = sh; pc->dims[s][2] = sd;
      ScaleDev& S = D.sc[s];
      S.w = sw; S.h = sh; S.max_disp = sd;
      S.pad = round_up(sd + 2, 8);
      S.pitch = round_up(sw + 2 * S.pad, 8);
      for (int v = 0; v < 2; ++v) {
        pc->d_bgr[v][s] = pc->arena.alloc<uint8_t>((size_t)sw * sh * 3);
        if (s == 0) {
          const uint8_t* src = v == 0 ? l : r;
          CSPM_CUDA(cudaMemcpy2DAsync(pc->d_bgr[v][0], (size_t)3 * w, src, (size_t)stride, (size_t)3 * w, h,
                                      on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
        } else {
          dim3 b(32, 8), g((sw + 31) / 32, (sh + 7) / 8);
          k_pyrdown<<<g, b, 0, st>>>(pc->d_bgr[v][s - 1], pc->dims[s - 1][0], pc->dims[s - 1][1], pc->d_bgr[v][s], sw, sh);
          pc->stats.kernel_launches++;
        }
        uint32_t* color = pc->arena.alloc<uint32_t>((size_t)S.pitch * sh);
        double* grad = pc->arena.alloc<double>((size_t)S.pitch * sh);
        unsigned long long* cg = cc->kind == CC_GRD ? pc->arena.alloc<unsigned long long>((size_t)S.pitch * sh) : nullptr;
        dim3 g((S.pitch + 127) / 128, sh);
        if (cc->kind == CC_IMG) k_pack_img<<<g, 128, 0, st>>>(pc->d_bgr[v][s], sw, sh, S.pad, S.pitch, color, grad);
        else k_pack<<<g, 128, 0, st>>>(pc->d_bgr[v][s], sw, sh, S.pad, S.pitch, color, grad, cg, d_inexact);
        pc->stats.kernel_launches++;
        S.color[v] = color;
        S.grad[v] = grad;
        S.cg[v] = cg;
        S.cen_hi[v] = nullptr;
        if (cc->kind == CC_CEN) {  // census code: bits 0..63 replace the gradient array, bits 64..79 in cen_hi
          uint8_t* g8 = pc->arena.alloc<uint8_t>((size_t)sw * sh);
          uint32_t* hi = pc->arena.alloc<uint32_t>((size_t)S.pitch * sh);
          CSPM_CUDA(cudaMemsetAsync(hi, 0, sizeof(uint32_t) * (size_t)S.pitch * sh, st));
          k_gray8_bgr<<<(sw * sh + 255) / 256, 256, 0, st>>>(pc->d_bgr[v][s], sw * sh, g8);
          dim3 gc((sw + 127) / 128, sh);
          k_census<<<gc, 128, 0, st>>>(g8, sw, sh, S.pad, S.pitch, reinterpret_cast<unsigned long long*>(grad), hi);
          pc->stats.kernel_launches += 2;
          S.cen_hi[v] = hi;
        }
      }
    }
    CSPM_CUDA(cudaGetLastError());

    // raw-cost maxima (pre_ss_pc.cc:51-58 / pre_cs_pc.cc:75-82)
    for (int s = 0; s < pc->scale_num && cc->kind != CC_IMG; ++s)
      for (int v = 0; v < 2; ++v) {
        dim3 g((D.sc[s].w + 127) / 128, D.sc[s].h);
        if (cc->kind == CC_CEN) k_max_raw_cost<CC_CEN><<<g, 128, 0, st>>>(D, s, v, d_max + v * kMaxScales + s);
        else k_max_raw_cost<CC_GRD><<<g, 128, 0, st>>>(D, s, v, d_max + v * kMaxScales + s);
        pc->stats.kernel_launches++;
      }
    CSPM_CUDA(cudaGetLastError());
    unsigned long long h_max[2 * kMaxScales + 1];
    CSPM_CUDA(cudaMemcpyAsync(h_max, d_max, sizeof(h_max), cudaMemcpyDeviceToHost, st));
    CSPM_CUDA(cudaStreamSynchronize(st));
    if (h_max[2 * kMaxScales] != 0)
      throw Error(CSPM_EUNSUPPORTED, "internal: image gradient not representable in the packed GRD layout");
    for (int s = 0; s < pc->scale_num; ++s)
      for (int v = 0; v < 2; ++v) {
        double m;
        std::memcpy(&m, &h_max[v * kMaxScales + s], 8);
        // GrdPC / CSPC charge an impossible disparity COST_ALPHA*TAU_CLR + (1-COST_ALPHA)*TAU_GRD (grd_pc.cc:114-116)
        if (cc->kind == CC_IMG) m = 0.1 * 10.0 + (1 - 0.1) * 2.0;
        pc->max_cost[v][s] = m;
        D.sc[s].max_cost[v] = m;
      }

    // cross-scale weights (pre_cs_pc.cc:85-109)
    if (cross_scale) scale_weights(scale_num, reg_lambda, &pc->scale_wgt);
    else pc->scale_wgt.assign(1, 1.0);
    for (int s = 0; s < pc->scale_num; ++s) D.scale_wgt[s] = pc->scale_wgt[s];
    // reg_lambda < 0 (user-reachable, main.cc:34) can make a scale weight negative: window sums are then no longer
    // monotone, so the kernels neither terminate early nor certify comparisons from partial sums
    D.monotone = 1;
    for (int s = 0; s < pc->scale_num; ++s)
      if (!(pc->scale_wgt[s] >= 0.0)) D.monotone = 0;
    {  // measurement switch (bench.py's worst-case line): sum every window in full although the bound allows stopping
      const char* e = getenv("CSPM_NO_EARLY_TERM");
      D.early_term = (D.monotone && !(e && atoi(e) != 0)) ? 1 : 0;
    }

    // in-image window taps of one evaluation at every pixel of scale 0
    // (the centre of scale s is (x>>s, y>>s), pre_cs_pc.cc:183-184)
    const int hw = wnd_size / 2;
    uint64_t taps = 0;
    for (int s = 0; s < pc->scale_num; ++s) {
      const int ws = pc->dims[s][0], hs = pc->dims[s][1];
      uint64_t sx = 0, sy = 0;
      for (int x = 0; x < w; ++x) {
        const int c = x >> s;
        sx += (uint64_t)(std::min(c + hw, ws - 1) - std::max(c - hw, 0) + 1);
      }
      for (int y = 0; y < h; ++y) {
        const int c = y >> s;
        sy += (uint64_t)(std::min(c + hw, hs - 1) - std::max(c - hw, 0) + 1);
      }
      taps += sx * sy;
    }
    pc->taps_per_image_eval = taps;
    return pc;
  } catch (...) {
    pc->arena.release();
    if (pc->stream && pc->own_stream) cudaStreamDestroy(pc->stream);
    delete pc;
    throw;
  }
}

}  // namespace cspm

using namespace cspm;

extern "C" {

const char* cspm_last_error(void) { return g_last_error.c_str(); }
int cspm_last_error_code(void) { return g_last_error_code; }
const char* cspm_version(void) { return "cspm-b200 0.1 (sm_100a)"; }
int cspm_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

void cspm_set_thread_stream(void* cuda_stream) {
  g_thread_stream = (cudaStream_t)cuda_stream;
  g_thread_stream_set = cuda_stream != nullptr;
}
uint64_t cspm_pc_taps_per_image_eval(const cspm_pc_t* pc) { return pc ? pc->taps_per_image_eval : 0; }

cspm_cc_t* cspm_cc_create(const char* name) {
  if (!name) {
    set_last_error("null cc name");
    return nullptr;
  }
  std::string n(name);
  if (n == "GRD") return new cspm_cc{CC_GRD, n};
  if (n == "CEN") return new cspm_cc{CC_CEN, n};
  set_last_error("unknown cc_method name '" + n + "' (GetCCType returns NULL, main.cc:47-54)");
  return nullptr;
}
void cspm_cc_destroy(cspm_cc_t* cc) { delete cc; }
const char* cspm_cc_name(const cspm_cc_t* cc) { return cc ? cc->name.c_str() : ""; }

int cspm_cc_build_cv(const cspm_cc_t* cc, const double* l_rgb, const double* r_rgb, int w, int h, int max_dis,
                     int right_view, double* cost_vol, int device) {
  return guard([&] {
    CSPM_REQUIRE(cc && l_rgb && r_rgb && cost_vol, "null argument");
    CSPM_REQUIRE(w > 0 && h > 0 && max_dis > 0, "bad geometry");
    use_device(device);
    DeviceArena A;
    try {
      const size_t n = (size_t)w * h;
      double* d_l = A.alloc<double>(3 * n);
      double* d_r = A.alloc<double>(3 * n);
      if (cc->kind == CC_CEN) {
        uint8_t* g8 = A.alloc<uint8_t>(n);
        unsigned long long* lo[2] = {A.alloc<unsigned long long>(n), A.alloc<unsigned long long>(n)};
        uint32_t* hi[2] = {A.alloc<uint32_t>(n), A.alloc<uint32_t>(n)};
        double* d_vol = A.alloc<double>(n * max_dis);
        CSPM_CUDA(cudaMemcpy(d_l, l_rgb, sizeof(double) * 3 * n, cudaMemcpyHostToDevice));
        CSPM_CUDA(cudaMemcpy(d_r, r_rgb, sizeof(double) * 3 * n, cudaMemcpyHostToDevice));
        dim3 gc((w + 127) / 128, h);
        for (int v = 0; v < 2; ++v) {
          k_gray8_f64<<<(unsigned)((n + 255) / 256), 256>>>(v == 0 ? d_l : d_r, (int)n, g8);
          k_census<<<gc, 128>>>(g8, w, h, 0, w, lo[v], hi[v]);
        }
        dim3 gv((w + 127) / 128, h, max_dis);
        k_build_cv_cen<<<gv, 128>>>(lo[0], hi[0], lo[1], hi[1], w, h, right_view, d_vol);
        CSPM_CUDA(cudaGetLastError());
        CSPM_CUDA(cudaMemcpy(cost_vol, d_vol, sizeof(double) * n * max_dis, cudaMemcpyDeviceToHost));
        CSPM_CUDA(cudaDeviceSynchronize());
        A.release();
        return;
      }
      float* d_gray = A.alloc<float>(n);
      double* d_lg = A.alloc<double>(n);
      double* d_rg = A.alloc<double>(n);
      double* d_vol = A.alloc<double>(n * max_dis);
      CSPM_CUDA(cudaMemcpy(d_l, l_rgb, sizeof(double) * 3 * n, cudaMemcpyHostToDevice));
      CSPM_CUDA(cudaMemcpy(d_r, r_rgb, sizeof(double) * 3 * n, cudaMemcpyHostToDevice));
      dim3 gs((w + 127) / 128, h);
      k_gray_f64<<<(unsigned)((n + 255) / 256), 256>>>(d_l, (int)n, d_gray);
      k_sobel_x<<<gs, 128>>>(d_gray, w, h, d_lg);
      k_gray_f64<<<(unsigned)((n + 255) / 256), 256>>>(d_r, (int)n, d_gray);
      k_sobel_x<<<gs, 128>>>(d_gray, w, h, d_rg);
      dim3 gv((w + 127) / 128, h, max_dis);
      k_build_cv_grd<<<gv, 128>>>(d_l, d_r, d_lg, d_rg, w, h, max_dis, right_view, d_vol);
      CSPM_CUDA(cudaGetLastError());
      CSPM_CUDA(cudaMemcpy(cost_vol, d_vol, sizeof(double) * n * max_dis, cudaMemcpyDeviceToHost));
      CSPM_CUDA(cudaDeviceSynchronize());
    } catch (...) {
      A.release();
      throw;
    }
    A.release();
  });
}

cspm_pc_t* cspm_pc_create_ss(const uint8_t* l, const uint8_t* r, int w, int h, int stride, int max_disp,
                             int wnd_size, const cspm_cc_t* cc, int device) {
  return guard_ptr<cspm_pc>([&] { return pc_create(l, r, false, w, h, stride, max_disp, wnd_size, 1, false, cc, 0.0, device); });
}
cspm_pc_t* cspm_pc_create_cs(const uint8_t* l, const uint8_t* r, int w, int h, int stride, int max_disp,
                             int wnd_size, int scale_num, const cspm_cc_t* cc, double reg_lambda, int device) {
  return guard_ptr<cspm_pc>([&] { return pc_create(l, r, false, w, h, stride, max_disp, wnd_size, scale_num, true, cc, reg_lambda, device); });
}
cspm_pc_t* cspm_pc_create_ss_dev(const uint8_t* l, const uint8_t* r, int w, int h, int stride, int max_disp,
                                 int wnd_size, const cspm_cc_t* cc, int device) {
  return guard_ptr<cspm_pc>([&] { return pc_create(l, r, true, w, h, stride, max_disp, wnd_size, 1, false, cc, 0.0, device); });
}
cspm_pc_t* cspm_pc_create_cs_dev(const uint8_t* l, const uint8_t* r, int w, int h, int stride, int max_disp,
                                 int wnd_size, int scale_num, const cspm_cc_t* cc, double reg_lambda, int device) {
  return guard_ptr<cspm_pc>([&] { return pc_create(l, r, true, w, h, stride, max_disp, wnd_size, scale_num, true, cc, reg_lambda, device); });
}
cspm_pc_t* cspm_pc_create_grdpc(const uint8_t* l, const uint8_t* r, int w, int h, int stride, int max_disp, int wnd_size,
                                int device) {
  const cspm_cc img{CC_IMG, "IMG"};
  return guard_ptr<cspm_pc>([&] { return pc_create(l, r, false, w, h, stride, max_disp, wnd_size, 1, false, &img, 0.0, device); });
}
cspm_pc_t* cspm_pc_create_cspc(const uint8_t* l, const uint8_t* r, int w, int h, int stride, int max_disp, int wnd_size,
                               int scale_num, double reg_lambda, int device) {
  const cspm_cc img{CC_IMG, "IMG"};
  return guard_ptr<cspm_pc>([&] { return pc_create(l, r, false, w, h, stride, max_disp, wnd_size, scale_num, true, &img, reg_lambda, device); });
}
void cspm_pc_destroy(cspm_pc_t* pc) {
  if (!pc) return;
  cudaSetDevice(pc->device);
  pc->arena.release();
  cudaStreamSynchronize(pc->stream);
  if (pc->own_stream) cudaStreamDestroy(pc->stream);
  delete pc;
}

int cspm_pc_get_plane_cost(const cspm_pc_t* pc, int n, const int* x, const int* y, const double* normal_xyz,
                           const double* point_xyz, const int* view, double* out_cost) {
  return guard([&] {
    CSPM_REQUIRE(pc && x && y && normal_xyz && point_xyz && view && out_cost, "null argument");
    CSPM_REQUIRE(n >= 0, "negative count");
    if (n == 0) return;
    for (int i = 0; i < n; ++i) {
      CSPM_REQUIRE(view[i] == 0 || view[i] == 1, "view must be 0 or 1");
      CSPM_REQUIRE(x[i] >= 0 && x[i] < pc->dims[0][0] && y[i] >= 0 && y[i] < pc->dims[0][1], "reference pixel outside the image");
    }
    use_device(pc->device);
    DeviceArena A;
    A.stream = pc->stream;
    try {
      int* d_x = A.alloc<int>(n);
      int* d_y = A.alloc<int>(n);
      int* d_v = A.alloc<int>(n);
      double* d_n = A.alloc<double>(3 * (size_t)n);
      double* d_p = A.alloc<double>(3 * (size_t)n);
      double* d_o = A.alloc<double>(n);
      cudaStream_t st = pc->stream;
      CSPM_CUDA(cudaMemcpyAsync(d_x, x, sizeof(int) * n, cudaMemcpyHostToDevice, st));
      CSPM_CUDA(cudaMemcpyAsync(d_y, y, sizeof(int) * n, cudaMemcpyHostToDevice, st));
      CSPM_CUDA(cudaMemcpyAsync(d_v, view, sizeof(int) * n, cudaMemcpyHostToDevice, st));
      CSPM_CUDA(cudaMemcpyAsync(d_n, normal_xyz, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, st));
      CSPM_CUDA(cudaMemcpyAsync(d_p, point_xyz, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, st));
      CSPM_DISPATCH_CC(pc->dev.cc_kind, k_plane_cost_batch<CC><<<(n + 127) / 128, 128, 0, st>>>(pc->dev, n, d_x, d_y, d_n, d_p, d_v, d_o));
      CSPM_CUDA(cudaGetLastError());
      pc->stats.kernel_launches++;
      pc->stats.plane_evals += n;
      CSPM_CUDA(cudaMemcpyAsync(out_cost, d_o, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
      CSPM_CUDA(cudaStreamSynchronize(st));
    } catch (...) {
      A.release();
      throw;
    }
    A.release();
  });
}

int cspm_pc_info(const cspm_pc_t* pc, double* scale_wgt, double* max_cost, int* dims) {
  if (!pc) {
    set_last_error("null plane cost");
    return CSPM_EINVAL;
  }
  const int S = pc->scale_num;
  for (int s = 0; s < S; ++s) {
    if (scale_wgt) scale_wgt[s] = pc->scale_wgt[s];
    if (max_cost) {
      max_cost[s] = pc->max_cost[0][s];
      max_cost[S + s] = pc->max_cost[1][s];
    }
    if (dims) {
      dims[3 * s] = pc->dims[s][0];
      dims[3 * s + 1] = pc->dims[s][1];
      dims[3 * s + 2] = pc->dims[s][2];
    }
  }
  return S;
}

int cspm_pc_scale_image(const cspm_pc_t* pc, int s, int view, uint8_t* out) {
  return guard([&] {
    CSPM_REQUIRE(pc && out, "null argument");
    CSPM_REQUIRE(s >= 0 && s < pc->scale_num && (view == 0 || view == 1), "bad scale/view");
    use_device(pc->device);
    CSPM_CUDA(cudaMemcpyAsync(out, pc->d_bgr[view][s], (size_t)pc->dims[s][0] * pc->dims[s][1] * 3, cudaMemcpyDeviceToHost, pc->stream));
    CSPM_CUDA(cudaStreamSynchronize(pc->stream));
  });
}

void* cspm_pc_stream(const cspm_pc_t* pc) { return pc ? (void*)pc->stream : nullptr; }
int cspm_pc_stats(const cspm_pc_t* pc, cspm_stats_t* out) {
  if (!pc || !out) return CSPM_EINVAL;
  *out = pc->stats;
  return CSPM_OK;
}

}  // extern "C"
