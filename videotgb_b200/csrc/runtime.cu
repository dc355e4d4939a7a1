// Error plumbing and device checks shared by every C-ABI entry point.
#include "common.cuh"

#include <cstdarg>
#include <cstdio>

namespace tgb {

static thread_local char g_err[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_status(cudaError_t e, const char* what) {
  set_last_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return TGB_ECUDA;
}

// per host thread: the "current device" is per-thread state, and so is this cache of its properties
static thread_local int g_dev_checked = -1, g_sms = 0, g_cc_major = 0;

static int probe() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_status(e, "cudaGetDevice");
  if (dev == g_dev_checked) return TGB_OK;
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) return cuda_status(e, "cudaGetDeviceProperties");
  g_sms = prop.multiProcessorCount;
  g_cc_major = prop.major;
  g_dev_checked = dev;
  return TGB_OK;
}

int check_device() {
  int st = probe();
  if (st != TGB_OK) return st;
  if (g_cc_major != 10) {
    set_last_error("videotgb_b200 needs an sm_100 (B200) device, found compute capability %d.x -- no fallback",
                   g_cc_major);
    return TGB_EARCH;
  }
  return TGB_OK;
}

int sm_count() {
  probe();
  return g_sms > 0 ? g_sms : 148;
}

}  // namespace tgb

extern "C" const char* tgb_last_error(void) { return tgb::g_err; }
extern "C" int tgb_version(void) { return 200; }
extern "C" int tgb_device_check(void) { return tgb::check_device(); }
extern "C" int tgb_sm_count(void) { return tgb::sm_count(); }
