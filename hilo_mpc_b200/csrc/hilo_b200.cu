// libhilo_b200.so — core C-ABI library (see include/hilo_b200.h).
//
// Generated problem modules (NMPC/MHE transcriptions, model-specific estimator kernels) are separate shared objects;
// this library loads them with dlopen and forwards the calls, and it hosts the model-independent kernels.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/hilo_b200.h"

namespace {
thread_local char g_err[1024] = "";
void set_err(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace

struct hb_ocp {
  void* dl;
  void (*dims)(hb_ocp_dims*);
  size_t (*work)(int, int);
  int (*solve)(const hb_ocp_call*);
  hb_ocp_dims d;
};

extern "C" {

int hb_version(void) { return 1; }

const char* hb_last_error(void) { return g_err; }

int hb_device_info(int device, int* sm_count, size_t* total_mem, int* cc_major, int* cc_minor) {
  cudaDeviceProp p;
  cudaError_t e = cudaGetDeviceProperties(&p, device);
  if (e != cudaSuccess) { set_err("cudaGetDeviceProperties: %s", cudaGetErrorString(e)); return (int)e; }
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (total_mem) *total_mem = p.totalGlobalMem;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return 0;
}

int hb_ocp_open(const char* module_path, hb_ocp_t* out) {
  if (module_path == nullptr || out == nullptr) { set_err("hb_ocp_open: null argument"); return -1; }
  void* dl = dlopen(module_path, RTLD_NOW | RTLD_LOCAL);
  if (dl == nullptr) { set_err("hb_ocp_open: %s", dlerror()); return -2; }
  hb_ocp* h = new hb_ocp();
  h->dl = dl;
  h->dims = (void (*)(hb_ocp_dims*))dlsym(dl, "hbm_ocp_dims");
  h->work = (size_t(*)(int, int))dlsym(dl, "hbm_ocp_work_doubles");
  h->solve = (int (*)(const hb_ocp_call*))dlsym(dl, "hbm_ocp_solve");
  if (h->dims == nullptr || h->work == nullptr || h->solve == nullptr) {
    set_err("hb_ocp_open: %s is not a hilo_b200 OCP module (missing hbm_ocp_* symbols)", module_path);
    dlclose(dl);
    delete h;
    return -3;
  }
  h->dims(&h->d);
  *out = h;
  return 0;
}

int hb_ocp_get_dims(hb_ocp_t h, hb_ocp_dims* dims) {
  if (h == nullptr || dims == nullptr) { set_err("hb_ocp_get_dims: null argument"); return -1; }
  *dims = h->d;
  return 0;
}

size_t hb_ocp_work_doubles(hb_ocp_t h, int B, int n_alpha) {
  if (h == nullptr || B <= 0 || n_alpha <= 0) return 0;
  return h->work(B, n_alpha);
}

int hb_ocp_solve(hb_ocp_t h, const hb_ocp_call* call) {
  if (h == nullptr || call == nullptr) { set_err("hb_ocp_solve: null argument"); return -1; }
  const int rc = h->solve(call);
  if (rc < 0) set_err("hb_ocp_solve: invalid call (code %d: -1 batch, -2 workspace too small, -3 n_alpha)", rc);
  if (rc > 0) set_err("hb_ocp_solve: CUDA error %d (%s)", rc, cudaGetErrorString((cudaError_t)rc));
  return rc;
}

int hb_ocp_close(hb_ocp_t h) {
  if (h == nullptr) return -1;
  if (h->dl) dlclose(h->dl);
  delete h;
  return 0;
}

}  // extern "C"
