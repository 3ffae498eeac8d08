// Error reporting, device-attribute cache, ABI version, workspace sizing.
#include <cstring>
#include <mutex>

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

}  // namespace ssdbox

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
    default: return 0;
  }
}
