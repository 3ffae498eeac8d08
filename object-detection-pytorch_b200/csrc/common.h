// Host-side helpers shared by every translation unit of libssdbox.so.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstddef>
#include <cstdint>
#include <cstdio>

#include "ssdbox.h"

namespace ssdbox {

void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

struct DevInfo {
  int device;
  int sm_count;
  int max_smem_optin;  // bytes of dynamic shared memory a CTA may opt in to
};
// attributes of the current device (cached per device id)
int get_dev_info(DevInfo* out);

// Preferred shared-memory carve-out (percent) applied to every kernel of a step, or -1 to leave the driver's
// heuristic alone.  Consecutive kernels that ask for different carve-outs make every SM drain and re-partition
// its L1 / shared memory between them; one common value removes that from the kernel boundaries.
int smem_carveout_pct();

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- opt-in per-kernel timing (bench / profiling only; off by default) -------------------------
enum KernelId {
  KID_INIT = 0, KID_MATCH, KID_LOSS_STREAM, KID_MINE, KID_LOSS_BWD, KID_DET_STREAM, KID_DET_SEGMENT,
  KID_DET_OVERFLOW, KID_MATERIALIZE, KID_DET_SEGMENT_BIG, KID_COUNT
};
void timer_begin(int kid, cudaStream_t st);
void timer_end(int kid, cudaStream_t st);
struct TimerScope {
  int kid;
  cudaStream_t st;
  TimerScope(int k, cudaStream_t s) : kid(k), st(s) { timer_begin(k, s); }
  ~TimerScope() { timer_end(kid, st); }
};

// carve consecutive 256-byte aligned regions out of the caller's workspace
struct Carver {
  char* base;
  size_t off;
  explicit Carver(void* p) : base(static_cast<char*>(p)), off(0) {}
  template <typename T>
  T* take(size_t count) {
    T* r = reinterpret_cast<T*>(base + off);
    off = align_up(off + count * sizeof(T));
    return r;
  }
};

}  // namespace ssdbox

#define SSDBOX_REQUIRE(cond, code, ...)                   \
  do {                                                    \
    if (!(cond)) return ::ssdbox::fail((code), __VA_ARGS__); \
  } while (0)

#define SSDBOX_CUDA(call)                                            \
  do {                                                               \
    cudaError_t e__ = (call);                                        \
    if (e__ != cudaSuccess) return ::ssdbox::cuda_fail(e__, #call);  \
  } while (0)

#define SSDBOX_CARVE(kern)                                                                              \
  do {                                                                                                  \
    int pct__ = ::ssdbox::smem_carveout_pct();                                                          \
    if (pct__ >= 0) SSDBOX_CUDA(cudaFuncSetAttribute((kern), cudaFuncAttributePreferredSharedMemoryCarveout, pct__)); \
  } while (0)

#define SSDBOX_LAUNCH_OK(name)                                        \
  do {                                                                \
    cudaError_t e__ = cudaGetLastError();                             \
    if (e__ != cudaSuccess) return ::ssdbox::cuda_fail(e__, name);    \
  } while (0)
