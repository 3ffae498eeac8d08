// Error reporting, device-attribute cache, ABI version, workspace sizing.
#include <cstring>
#include <mutex>
#include <vector>

#include "common.h"
#include "sizes.h"

namespace ssdbox {

static thread_local char g_err[512] = {0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int cuda_fail(cudaError_t e, const char* what) {
  snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return SSDBOX_ECUDA;
}

int get_dev_info(DevInfo* out) {
  static std::mutex mu;
  static DevInfo cache[64];
  static bool have[64] = {false};
  int dev = -1;
  SSDBOX_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(SSDBOX_ECUDA, "device id %d out of range", dev);
  std::lock_guard<std::mutex> lk(mu);
  if (!have[dev]) {
    DevInfo d;
    d.device = dev;
    SSDBOX_CUDA(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev));
    SSDBOX_CUDA(cudaDeviceGetAttribute(&d.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    cache[dev] = d;
    have[dev] = true;
  }
  *out = cache[dev];
  return SSDBOX_OK;
}

int smem_carveout_pct() {
#ifdef SSDBOX_EXPERIMENTS
  static int v = [] {
    const char* e = getenv("SSDBOX_CARVEOUT");
    return e ? atoi(e) : -1;
  }();
  return v;
#else
  return -1;
#endif
}

// ---- opt-in kernel timers ---------------------------------------------------------------------
struct TimerSlot {
  std::vector<cudaEvent_t> start, stop;   // pending pairs
  double total_ms = 0.0;
  long long launches = 0;
};
static std::mutex g_timer_mu;
static bool g_timer_on = false;
static TimerSlot g_timers[KID_COUNT];

void timer_begin(int kid, cudaStream_t st) {
  if (!g_timer_on) return;
  std::lock_guard<std::mutex> lk(g_timer_mu);
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, st);
  g_timers[kid].start.push_back(e);
}

void timer_end(int kid, cudaStream_t st) {
  if (!g_timer_on) return;
  std::lock_guard<std::mutex> lk(g_timer_mu);
  if (g_timers[kid].start.size() == g_timers[kid].stop.size()) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) {
    cudaEventDestroy(g_timers[kid].start.back());
    g_timers[kid].start.pop_back();
    return;
  }
  cudaEventRecord(e, st);
  g_timers[kid].stop.push_back(e);
}

static void timer_collect() {
  for (int k = 0; k < KID_COUNT; ++k) {
    TimerSlot& t = g_timers[k];
    for (size_t i = 0; i < t.stop.size(); ++i) {
      float ms = 0.f;
      if (cudaEventSynchronize(t.stop[i]) == cudaSuccess && cudaEventElapsedTime(&ms, t.start[i], t.stop[i]) == cudaSuccess) {
        t.total_ms += ms;
        t.launches += 1;
      }
      cudaEventDestroy(t.start[i]);
      cudaEventDestroy(t.stop[i]);
    }
    t.start.clear();
    t.stop.clear();
  }
}

}  // namespace ssdbox

extern "C" int ssdbox_timers_enable(int on) {
  std::lock_guard<std::mutex> lk(ssdbox::g_timer_mu);
  ssdbox::timer_collect();
  if (on) {
    for (int k = 0; k < ssdbox::KID_COUNT; ++k) {
      ssdbox::g_timers[k].total_ms = 0.0;
      ssdbox::g_timers[k].launches = 0;
    }
  }
  ssdbox::g_timer_on = on != 0;
  return SSDBOX_OK;
}

extern "C" int ssdbox_timers_read(int kernel_id, double* total_ms, int64_t* launches) {
  if (kernel_id < 0 || kernel_id >= ssdbox::KID_COUNT || !total_ms || !launches)
    return ssdbox::fail(SSDBOX_EINVAL, "timers_read: bad argument");
  std::lock_guard<std::mutex> lk(ssdbox::g_timer_mu);
  ssdbox::timer_collect();
  *total_ms = ssdbox::g_timers[kernel_id].total_ms;
  *launches = ssdbox::g_timers[kernel_id].launches;
  return SSDBOX_OK;
}

extern "C" int ssdbox_abi_version(void) { return SSDBOX_ABI_VERSION; }

extern "C" int ssdbox_last_error(char* buf, size_t n) {
  size_t len = strlen(ssdbox::g_err);
  if (buf && n) {
    size_t c = len < n - 1 ? len : n - 1;
    memcpy(buf, ssdbox::g_err, c);
    buf[c] = 0;
  }
  return (int)len;
}

extern "C" size_t ssdbox_workspace_bytes(int op, int B, int P, int C, int gmax, int top_k) {
  using namespace ssdbox;
  if (B < 0 || P < 0 || C < 0 || gmax < 0 || top_k < 0) return 0;
  switch (op) {
    case SSDBOX_OP_MATCH: return match_ws_bytes(B, P, gmax) + 256;
    case SSDBOX_OP_LOSS_FWD: return loss_ws_bytes(B, P, C, gmax) + 256;
    case SSDBOX_OP_DETECT: return detect_ws_bytes(B, P, C, top_k) + 256;
    case SSDBOX_OP_NMS: return nms_ws_bytes(P, top_k) + 256;
    case SSDBOX_OP_LSE: return 256;
    case SSDBOX_OP_MINE: return mine_ws_bytes(B, P) + 256;
    case SSDBOX_OP_COMPACT: return compact_ws_bytes(B, C) + 256;
    case SSDBOX_OP_VOC_EVAL: return voc_eval_ws_bytes(P, gmax, C) + 256;
    default: return 0;
  }
}
