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
#include "est_kernels.cuh"
#include "lmpc_qp.cuh"

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
  size_t (*ric_work)(int);
  int (*ric)(const hb_ocp_riccati_call*);
  hb_ocp_dims d;
};

struct hb_est {
  void* dl;
  void (*dims)(hb_est_dims*);
  size_t (*work)(int, int);
  int (*sim)(const hb_est_call*);
  int (*ekf)(const hb_est_call*);
  int (*ukf)(const hb_est_call*);
  int (*pf)(const hb_est_call*);
  hb_est_dims d;
};

extern "C" {

int hb_version(void) { return HB_ABI_VERSION; }

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
  h->ric_work = (size_t(*)(int))dlsym(dl, "hbm_ocp_riccati_work_doubles");
  h->ric = (int (*)(const hb_ocp_riccati_call*))dlsym(dl, "hbm_ocp_riccati");
  if (h->dims == nullptr || h->work == nullptr || h->solve == nullptr) {
    set_err("hb_ocp_open: %s is not a hilo_b200 OCP module (missing hbm_ocp_* symbols)", module_path);
    dlclose(dl);
    delete h;
    return -3;
  }
  {
    int (*abi)() = (int (*)())dlsym(dl, "hbm_abi_version");
    const int v = abi ? abi() : 0;
    if (v != HB_ABI_VERSION) {
      set_err("hb_ocp_open: %s was built for ABI version %d, the library is version %d (stale module: rebuild it)",
              module_path, v, HB_ABI_VERSION);
      dlclose(dl);
      delete h;
      return -5;
    }
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

size_t hb_ocp_riccati_work_doubles(hb_ocp_t h, int B) {
  if (h == nullptr || h->ric_work == nullptr || B <= 0) return 0;
  return h->ric_work(B);
}

int hb_ocp_riccati(hb_ocp_t h, const hb_ocp_riccati_call* call) {
  if (h == nullptr || call == nullptr) { set_err("hb_ocp_riccati: null argument"); return -1; }
  if (h->ric == nullptr) { set_err("hb_ocp_riccati: the module does not export the sweep"); return -4; }
  const int rc = h->ric(call);
  if (rc < 0) set_err("hb_ocp_riccati: invalid call (code %d: -1 argument, -2 workspace too small)", rc);
  if (rc > 0) set_err("hb_ocp_riccati: CUDA error %d (%s)", rc, cudaGetErrorString((cudaError_t)rc));
  return rc;
}

int hb_ocp_solve(hb_ocp_t h, const hb_ocp_call* call) {
  if (h == nullptr || call == nullptr) { set_err("hb_ocp_solve: null argument"); return -1; }
  const int rc = h->solve(call);
  if (rc < 0) set_err("hb_ocp_solve: invalid call (code %d: -1 batch, -2 workspace too small, -3 n_alpha)", rc);
  if (rc > 0) set_err("hb_ocp_solve: CUDA error %d (%s)", rc, cudaGetErrorString((cudaError_t)rc));
  return rc;
}

int hb_est_open(const char* module_path, hb_est_t* out) {
  if (module_path == nullptr || out == nullptr) { set_err("hb_est_open: null argument"); return -1; }
  void* dl = dlopen(module_path, RTLD_NOW | RTLD_LOCAL);
  if (dl == nullptr) { set_err("hb_est_open: %s", dlerror()); return -2; }
  hb_est* h = new hb_est();
  h->dl = dl;
  h->dims = (void (*)(hb_est_dims*))dlsym(dl, "hbm_est_dims");
  h->work = (size_t(*)(int, int))dlsym(dl, "hbm_est_work_doubles");
  h->sim = (int (*)(const hb_est_call*))dlsym(dl, "hbm_sim_step");
  h->ekf = (int (*)(const hb_est_call*))dlsym(dl, "hbm_ekf_step");
  h->ukf = (int (*)(const hb_est_call*))dlsym(dl, "hbm_ukf_step");
  h->pf = (int (*)(const hb_est_call*))dlsym(dl, "hbm_pf_step");
  if (!h->dims || !h->work || !h->sim || !h->ekf || !h->ukf || !h->pf) {
    set_err("hb_est_open: %s is not a hilo_b200 estimator module (missing hbm_* symbols)", module_path);
    dlclose(dl);
    delete h;
    return -3;
  }
  {
    int (*abi)() = (int (*)())dlsym(dl, "hbm_abi_version");
    const int v = abi ? abi() : 0;
    if (v != HB_ABI_VERSION) {
      set_err("hb_est_open: %s was built for ABI version %d, the library is version %d (stale module: rebuild it)",
              module_path, v, HB_ABI_VERSION);
      dlclose(dl);
      delete h;
      return -5;
    }
  }
  h->dims(&h->d);
  *out = h;
  return 0;
}

int hb_est_get_dims(hb_est_t h, hb_est_dims* dims) {
  if (h == nullptr || dims == nullptr) { set_err("hb_est_get_dims: null argument"); return -1; }
  *dims = h->d;
  return 0;
}

size_t hb_est_work_doubles(hb_est_t h, int B, int n_particles) {
  if (h == nullptr || B <= 0) return 0;
  return h->work(B, n_particles > 0 ? n_particles : 0);
}

static int est_forward(const char* what, hb_est_t h, const hb_est_call* c, int (*fn)(const hb_est_call*)) {
  if (h == nullptr || c == nullptr) { set_err("%s: null argument", what); return -1; }
  const int rc = fn(c);
  if (rc < 0) set_err("%s: invalid call (code %d: -1 batch, -2 workspace/particles, -4 missing array)", what, rc);
  if (rc > 0) set_err("%s: CUDA error %d (%s)", what, rc, cudaGetErrorString((cudaError_t)rc));
  return rc;
}
int hb_sim_step(hb_est_t h, const hb_est_call* c) { return est_forward("hb_sim_step", h, c, h ? h->sim : nullptr); }
int hb_ekf_step(hb_est_t h, const hb_est_call* c) { return est_forward("hb_ekf_step", h, c, h ? h->ekf : nullptr); }
int hb_ukf_step(hb_est_t h, const hb_est_call* c) { return est_forward("hb_ukf_step", h, c, h ? h->ukf : nullptr); }
int hb_pf_step(hb_est_t h, const hb_est_call* c) { return est_forward("hb_pf_step", h, c, h ? h->pf : nullptr); }

int hb_est_close(hb_est_t h) {
  if (h == nullptr) return -1;
  if (h->dl) dlclose(h->dl);
  delete h;
  return 0;
}

int hb_pf_resample(int B, int n, const double* q, const double* unif, double* cdf_work, long long* idx, void* stream) {
  if (B <= 0 || n <= 0 || !q || !unif || !cdf_work || !idx) { set_err("hb_pf_resample: invalid argument"); return -1; }
  hb::EstArgs a;
  memset(&a, 0, sizeof(a));
  a.B = B; a.n_particles = n; a.q = const_cast<double*>(q); a.unif = unif; a.cdf = cdf_work; a.idx = idx;
  hb::launch_1d<hb::PfResampleOnly>(a, B, (cudaStream_t)stream);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_err("hb_pf_resample: %s", cudaGetErrorString(e)); return (int)e; }
  return 0;
}

size_t hb_lmpc_work_doubles(int B, int n, int mx) {
  if (B <= 0 || n <= 0 || mx < 0) return 0;
  return (size_t)B * hb::lmpc_work_per_instance(n, mx);
}

int hb_lmpc_solve(int B, int n, int mx, int nx, const double* H, const double* F, const double* Gam, const double* Phi,
                  const double* ulb, const double* uub, const double* xlb, const double* xub, const double* x0,
                  double* u, double* xpred, int* status, int* iters, double tol, int max_iter, double* work,
                  size_t work_doubles, void* stream) {
  if (B <= 0 || n <= 0 || n > 256 || mx < 0 || nx <= 0 || !H || !F || !ulb || !uub || !x0 || !u || !status || !iters ||
      (mx > 0 && (!Gam || !Phi || !xlb || !xub))) {
    set_err("hb_lmpc_solve: invalid argument");
    return -1;
  }
  if (work == nullptr || work_doubles < hb_lmpc_work_doubles(B, n, mx)) { set_err("hb_lmpc_solve: workspace too small"); return -2; }
  hb::LmpcArgs a;
  a.B = B; a.n = n; a.mx = mx; a.nx = nx; a.max_iter = max_iter > 0 ? max_iter : 100; a.tol = tol > 0.0 ? tol : 1e-10;
  a.H = H; a.F = F; a.Gam = Gam; a.Phi = Phi; a.ulb = ulb; a.uub = uub; a.xlb = xlb; a.xub = xub; a.x0 = x0;
  a.u = u; a.xpred = xpred; a.status = status; a.iters = iters; a.work = work;
  hb::launch_1d<hb::LmpcDenseQp>(a, B, (cudaStream_t)stream);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_err("hb_lmpc_solve: %s", cudaGetErrorString(e)); return (int)e; }
  return 0;
}

int hb_ocp_close(hb_ocp_t h) {
  if (h == nullptr) return -1;
  if (h->dl) dlclose(h->dl);
  delete h;
  return 0;
}

}  // extern "C"
