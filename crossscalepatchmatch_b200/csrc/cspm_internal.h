// cspm_internal.h — host-side handle types shared by the translation units of
// libcspm_b200.so.  Not installed; the public surface is include/cspm_c.h.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/cspm_c.h"
#include "cspm_device.cuh"

namespace cspm {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string& msg, int code = CSPM_EINVAL);

#define CSPM_CUDA(call)                                                                          \
  do {                                                                                           \
    cudaError_t e__ = (call);                                                                    \
    if (e__ != cudaSuccess)                                                                      \
      throw ::cspm::Error(e__ == cudaErrorMemoryAllocation ? CSPM_ENOMEM                         \
                          : (e__ == cudaErrorNoDevice || e__ == cudaErrorInsufficientDriver ||  \
                             e__ == cudaErrorInvalidDevice) ? CSPM_ENODEV : CSPM_ECUDA,          \
                          std::string(#call) + ": " + cudaGetErrorString(e__));                  \
  } while (0)

#define CSPM_REQUIRE(cond, msg)                                    \
  do {                                                             \
    if (!(cond)) throw ::cspm::Error(CSPM_EINVAL, std::string(msg)); \
  } while (0)

// Run body, translate exceptions into the C error convention.
template <typename F>
int guard(F&& f) {
  try {
    f();
    return CSPM_OK;
  } catch (const Error& e) {
    set_last_error(e.what(), e.code);
    return e.code;
  } catch (const std::bad_alloc& e) {
    set_last_error(e.what(), CSPM_ENOMEM);
    return CSPM_ENOMEM;
  } catch (const std::exception& e) {
    set_last_error(e.what(), CSPM_ECUDA);
    return CSPM_ECUDA;
  }
}
// Creators return NULL; the reason stays readable through cspm_last_error() / cspm_last_error_code().
template <typename T, typename F>
T* guard_ptr(F&& f) {
  T* out = nullptr;
  guard([&] { out = f(); });
  return out;
}

// Selects the device for the current thread and checks it exists.  This is
// where a box without a GPU fails loudly (no CPU fallback anywhere).
void use_device(int device);
// Stream for a new handle: the thread's injected stream (cspm_set_thread_stream)
// or a fresh private one.  *owned tells the destructor whether to destroy it.
cudaStream_t acquire_stream(bool* owned);

// Stream-ordered allocations from the device's default memory pool.  The pool
// keeps freed blocks (release threshold = unlimited, set in use_device), so the
// create / destroy churn of per-pair handles costs no cudaMalloc / cudaFree
// after the first pair.
struct DeviceArena {
  std::vector<void*> ptrs;
  cudaStream_t stream = nullptr;
  template <typename T>
  T* alloc(size_t n) {
    void* p = nullptr;
    CSPM_CUDA(cudaMallocAsync(&p, n * sizeof(T) + 256, stream));  // slack: vector loads may touch the tail
    ptrs.push_back(p);
    return (T*)p;
  }
  void release() {
    for (void* p : ptrs) cudaFreeAsync(p, stream);
    ptrs.clear();
  }
};

}  // namespace cspm

// Instantiate `body` with `constexpr int CC` bound to the plane cost's raw-cost kind.
#define CSPM_DISPATCH_CC(kind, ...)                                   \
  do {                                                                \
    switch (kind) {                                                   \
      case ::cspm::CC_CEN: { constexpr int CC = ::cspm::CC_CEN; __VA_ARGS__; } break; \
      case ::cspm::CC_IMG: { constexpr int CC = ::cspm::CC_IMG; __VA_ARGS__; } break; \
      default: { constexpr int CC = ::cspm::CC_GRD; __VA_ARGS__; } break;             \
    }                                                                 \
  } while (0)

// The same with `constexpr bool CS` bound to the plane cost being cross-scale (PreCSPC / CSPC): the bulk
// kernels are compiled once per (cost kind, single / cross scale) so the single-scale code carries none of
// the pyramid state in registers.
#define CSPM_DISPATCH_CC_CS(kind, cross, ...)                                            \
  do {                                                                                   \
    if (cross) { constexpr bool CS = true; CSPM_DISPATCH_CC(kind, __VA_ARGS__); }        \
    else { constexpr bool CS = false; CSPM_DISPATCH_CC(kind, __VA_ARGS__); }             \
  } while (0)

struct cspm_cc {
  int kind;  // cspm::CcKind
  std::string name;
};

struct cspm_pc {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  cspm::PcDev dev{};
  int wnd_size = 35;
  int scale_num = 1;  // 1 for PreSSPC
  std::vector<double> scale_wgt;
  double max_cost[2][cspm::kMaxScales] = {};
  int dims[cspm::kMaxScales][3] = {};
  uint8_t* d_bgr[2][cspm::kMaxScales] = {};  // dense 8-bit BGR pyramid (device)
  cspm::DeviceArena arena;
  mutable cspm_stats_t stats{};
  // Sum over pixels of the in-image window size, per scale (for tap counting).
  uint64_t taps_per_image_eval = 0;
  uint64_t uid = 0;  // unique per created plane cost (never reused, unlike the pointer)
};

struct cspm_pm {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  int w = 0, h = 0, max_dis = 0, dis_scale = 1, schedule = 0, n_ref = 0;
  uint64_t seed = 0;
  double* st[2] = {nullptr, nullptr};  // 7 planes of w*h: nx ny nz a b c cost
  uint8_t* dis[2] = {nullptr, nullptr};
  uint8_t* d_bgr[2] = {nullptr, nullptr};  // dense BGR copies (post-processing weights)
  // view-propagation scratch
  double* cand_cost = nullptr;
  int* cand_tx = nullptr;
  unsigned long long* tmin = nullptr;
  unsigned int* win = nullptr;
  unsigned long long* tap_counter = nullptr;
  unsigned int* grid_barrier = nullptr;  // persistent spatial sweep: one progress word (own 128-B line) per CTA
  // wavefront sweep (cspm_pm_wave.cuh): pixels that adopted a plane on the fast path, their compacted lists,
  // list lengths [2] and the sweep counters
  uint8_t* wave_changed = nullptr;       // [2][w*h]
  int* wave_list = nullptr;              // [2][w*h]
  unsigned int* wave_count = nullptr;    // [2]
  unsigned long long* wave_stats = nullptr;  // [4], accumulated over the sweeps of a call
  void* wave_pairs = nullptr;            // device array of per-pair descriptors for batched sweeps (kMaxBatch entries)
  // post-processing scratch
  int* valid[2] = {nullptr, nullptr};
  uint8_t* dis_tmp = nullptr;
  cspm::DeviceArena arena;
  cspm_stats_t stats{};
  float stage_ms[6] = {};
  // uid of the plane cost ALL stored min-costs were computed with (0: unknown — before InitRandomPlane,
  // after cspm_pm_set_planes, or after stages run with different plane costs).  Only when it matches the
  // plane cost of the current call may the spatial sweep assume cost == GetPlaneCost(stored plane).
  uint64_t cost_owner = 0;
};
