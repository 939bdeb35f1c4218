// C-ABI plumbing: error text, device properties, version.  The operator entry points live next to
// their kernels (sampling.cu, render_ops.cu, grid_ops.cu, stencil.cu, adam.cu); all are declared in
// include/voxurf_b200.h.
#include "common.cuh"

#include <stdio.h>
#include <string.h>

static thread_local char g_err[512] = "";

void vx_set_error(const char* where, const char* what) { snprintf(g_err, sizeof(g_err), "%s: %s", where, what); }

static unsigned long long g_launches = 0;

int vx_check_launch(const char* where) {
  ++g_launches;
  const cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return 0;
  vx_set_error(where, cudaGetErrorString(e));
  return (int)e;
}

int vx_num_sms() {
  static int sms[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (sms[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    sms[dev] = n;
  }
  return sms[dev];
}

VX_API const char* vx_last_error() { return g_err; }
VX_API int vx_abi_version() { return 1; }
VX_API int vx_sm_count() { return vx_num_sms(); }

// number of launches issued through this library since load (bench.py's gpu_launches claim)
VX_API unsigned long long vx_launch_count() { return g_launches; }
