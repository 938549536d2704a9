// C-ABI plumbing: error text, version, device gate.
#include "api_util.h"

#include <cstring>
#include <mutex>

namespace lrb {

char* last_error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

namespace {
struct DevInfo {
  int known = 0;
  int sms = 0;
  int cc = 0;
};
DevInfo g_dev[64];
std::mutex g_dev_mu;

int query(DevInfo& out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return set_error(static_cast<int>(e), "cudaGetDevice: %s", cudaGetErrorString(e));
  if (dev < 0 || dev >= 64) return set_error(LRB_ERR_BAD_ARG, "device index %d out of range", dev);
  std::lock_guard<std::mutex> lk(g_dev_mu);
  if (!g_dev[dev].known) {
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess)
      return set_error(static_cast<int>(e), "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    g_dev[dev].sms = prop.multiProcessorCount;
    g_dev[dev].cc = prop.major * 10 + prop.minor;
    g_dev[dev].known = 1;
  }
  out = g_dev[dev];
  return LRB_OK;
}
}  // namespace

int check_arch() {
  DevInfo d;
  int rc = query(d);
  if (rc != LRB_OK) return rc;
  if (d.cc / 10 != 10)
    return set_error(LRB_ERR_ARCH, "llamarec_b200 kernels are built for sm_100a only; device is sm_%d", d.cc);
  return LRB_OK;
}

int device_sm_count() {
  DevInfo d;
  if (query(d) != LRB_OK) return 0;
  return d.sms;
}

}  // namespace lrb

extern "C" {

const char* lrb_last_error(void) { return lrb::last_error_buffer(); }

int lrb_version(void) { return 100; }

int lrb_device_info(int* num_sms, int* cc) {
  lrb::DevInfo d;
  int rc = lrb::query(d);
  if (rc != LRB_OK) return rc;
  if (num_sms) *num_sms = d.sms;
  if (cc) *cc = d.cc;
  return LRB_OK;
}

}  // extern "C"
