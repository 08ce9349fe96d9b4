// Library-wide state of the C ABI: error text, ABI version, launch counter, device properties.
#include "ff_common.cuh"
#include "../../include/ffb200.h"
#include <string.h>
#include <stdlib.h>

static thread_local char g_err[512] = "";
long long g_ff_launches = 0;

void ff_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int ff_num_sms() {
  static int cache[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  int& n = cache[dev & 63];
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

bool ff_pdl_enabled() {
  static const bool on = []() { const char* e = getenv("FFB200_PDL"); return e && e[0] == '1'; }();
  return on;
}

extern "C" int ff_abi_version(void) { return FFB200_ABI_VERSION; }
extern "C" const char* ff_last_error(void) { return g_err; }
extern "C" long long ff_launch_count(void) { return g_ff_launches; }
