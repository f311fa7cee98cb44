// Warp-level primitives used by the multi-lane kernels, with a host emulation for the CPU logic tests.
//
// Device build: thin wrappers over the sm_100a intrinsics / PTX (cp.async = LDGSTS, shuffles, __syncwarp).
// HB_HOSTSIM build (tests only): one OS thread per lane, pthread barriers per warp, and cp.async modelled as a
// per-lane queue of pending 16-byte copies that lands either immediately or as late as wait_group allows
// (HB_EMU_CPASYNC_LATE=1), so that both kinds of ring-buffer races show up on a machine without a GPU.
#pragma once
#include "hb_common.cuh"

#if HB_DEVICE_BUILD
namespace hb {
namespace wp {
__device__ __forceinline__ unsigned s_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ int tid() { return (int)threadIdx.x; }
__device__ __forceinline__ int bid() { return (int)blockIdx.x; }
__device__ __forceinline__ void syncwarp() { __syncwarp(); }
__device__ __forceinline__ bool any(bool p) { return __any_sync(0xffffffffu, p) != 0; }
__device__ __forceinline__ double shfl(double v, int src, int width) { return __shfl_sync(0xffffffffu, v, src, width); }
__device__ __forceinline__ double shfl_xor(double v, int m, int width) { return __shfl_xor_sync(0xffffffffu, v, m, width); }
__device__ __forceinline__ void cp16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NPENDING>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(NPENDING) : "memory"); }
#ifndef HB_RCP_NEWTON_STEPS
#define HB_RCP_NEWTON_STEPS 2
#endif
__device__ __forceinline__ double rcp(double x) {   // x > 0: hardware approximation + two Newton steps
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
#if HB_RCP_NEWTON_STEPS >= 2
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
#endif
  return r;
}
__device__ __forceinline__ double* dyn_smem() {
  extern __shared__ __align__(128) double hb_dyn_smem[];
  return hb_dyn_smem;
}
// ---- mbarrier + 1-D bulk async copy (TMA), transaction-count completion
typedef unsigned long long bar_t;
__device__ __forceinline__ void bar_init(bar_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count));
}
__device__ __forceinline__ void bar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void bar_expect_tx(bar_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(bar_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "HB_WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra HB_WAIT_DONE;\n"
      "bra HB_WAIT_LOOP;\n"
      "HB_WAIT_DONE:\n"
      "}\n" ::"r"(s_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, bar_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s_u32(dst)),
               "l"(src), "r"(bytes), "r"(s_u32(bar))
               : "memory");
}
// fp64 tensor-core product D(8x8) = A(8x4) B(4x8) + C, fragments: a = A[lane/4][lane%4], b = B[lane%4][lane/4],
// c0/c1 = C[lane/4][2*(lane%4) + 0/1]
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
// hides how a per-lane constant was derived: the compiler must keep it in a register instead of recomputing it from the
// thread index inside the stage loop (it rematerialises freely under register pressure: ~10 instructions per stage)
__device__ __forceinline__ int keep(int v) { asm volatile("" : "+r"(v)); return v; }
__device__ __forceinline__ unsigned keep(unsigned v) { asm volatile("" : "+r"(v)); return v; }
__device__ __forceinline__ long long keep(long long v) { asm volatile("" : "+l"(v)); return v; }
// same product with distinct output and accumulator-input registers: (d0, d1) = A B + (c0, c1)
__device__ __forceinline__ void dmma4(double& d0, double& d1, double a, double b, double c0, double c1) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
               : "=d"(d0), "=d"(d1)
               : "d"(a), "d"(b), "d"(c0), "d"(c1));
}
// prefetch of the 128-byte line that holds *p into L2 (per-lane address; no register, no shared memory);
// KEEP: mark the line evict-last (it is read again before the kernel ends)
template <bool KEEP>
__device__ __forceinline__ void prefetch_l2(const void* p) {
  if (KEEP) asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(p));
  else asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
// ---- accesses predicated on one bit of a per-lane flag word, with 32-bit shared addresses and immediate offsets.
// Written as PTX so that the flag word stays ONE register and every access is "test bit + predicated op" (the
// compiler otherwise splits the word into one register per bit or rebuilds it inside the loop).  OFF in doubles.
typedef unsigned sa_t;
__device__ __forceinline__ sa_t sa(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ sa_t sadd(sa_t a, int doubles) { return a + 8 * doubles; }
template <int OFF>
__device__ __forceinline__ double lds(sa_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(a), "n"(OFF * 8) : "memory");
  return v;
}
template <int OFF>
__device__ __forceinline__ void lds2(sa_t a, double& v0, double& v1) {
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2+%3];" : "=d"(v0), "=d"(v1) : "r"(a), "n"(OFF * 8) : "memory");
}
template <int BIT, int OFF>
__device__ __forceinline__ void sts_if(int f, sa_t a, double v) {
  asm volatile("{\n.reg .pred p;\n.reg .b32 t;\nand.b32 t, %0, %1;\nsetp.ne.b32 p, t, 0;\n@p st.shared.f64 [%2+%3], %4;\n}\n"
               ::"r"(f), "n"(BIT), "r"(a), "n"(OFF * 8), "d"(v) : "memory");
}
template <int BIT, int OFF>
__device__ __forceinline__ void sts2_if(int f, sa_t a, double v0, double v1) {
  asm volatile("{\n.reg .pred p;\n.reg .b32 t;\nand.b32 t, %0, %1;\nsetp.ne.b32 p, t, 0;\n@p st.shared.v2.f64 [%2+%3], {%4,%5};\n}\n"
               ::"r"(f), "n"(BIT), "r"(a), "n"(OFF * 8), "d"(v0), "d"(v1) : "memory");
}
template <int BIT>
__device__ __forceinline__ void stg_if(int f, double* p, double v) {
  asm volatile("{\n.reg .pred p;\n.reg .b32 t;\nand.b32 t, %0, %1;\nsetp.ne.b32 p, t, 0;\n@p st.global.f64 [%2], %3;\n}\n"
               ::"r"(f), "n"(BIT), "l"(p), "d"(v) : "memory");
}
// L2 eviction policies for the cache_hint operand of loads (warp-uniform 64-bit descriptors)
typedef unsigned long long l2pol_t;
__device__ __forceinline__ l2pol_t l2_evict_first() {
  l2pol_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ l2pol_t l2_evict_last() {
  l2pol_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// predicated read-only 128-bit load with an L2 policy; the destination keeps its value in the other lanes
template <int BIT>
__device__ __forceinline__ void ldg2_if(int f, const double* p, double& v0, double& v1, l2pol_t pol) {
  asm volatile("{\n.reg .pred p;\n.reg .b32 t;\nand.b32 t, %2, %3;\nsetp.ne.b32 p, t, 0;\n@p ld.global.nc.L2::cache_hint.v2.f64 {%0,%1}, [%4], %5;\n}\n"
               : "+d"(v0), "+d"(v1) : "r"(f), "n"(BIT), "l"(p), "l"(pol));
}
__device__ __forceinline__ void ldg2_pol(const double* p, double& v0, double& v1, l2pol_t pol) {
  asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v0), "=d"(v1) : "l"(p), "l"(pol));
}
__device__ __forceinline__ double ldg_pol(const double* p, l2pol_t pol) {
  double v;
  asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
  return v;
}
template <int BIT>
__device__ __forceinline__ void ldg_if(int f, const double* p, double& v) {
  asm volatile("{\n.reg .pred p;\n.reg .b32 t;\nand.b32 t, %1, %2;\nsetp.ne.b32 p, t, 0;\n@p ld.global.nc.f64 %0, [%3];\n}\n"
               : "+d"(v) : "r"(f), "n"(BIT), "l"(p));
}
template <int BIT>
__device__ __forceinline__ void prefetch_l2_if(int f, const void* p) {
  asm volatile("{\n.reg .pred p;\n.reg .b32 t;\nand.b32 t, %0, %1;\nsetp.ne.b32 p, t, 0;\n@p prefetch.global.L2 [%2];\n}\n" ::"r"(f), "n"(BIT), "l"(p));
}
template <int BIT>
__device__ __forceinline__ double sel(int f, double a, double b) {   // bit set ? a : b
  double v;
  asm("{\n.reg .pred p;\n.reg .b32 t;\nand.b32 t, %1, %2;\nsetp.ne.b32 p, t, 0;\nselp.f64 %0, %3, %4, p;\n}\n" : "=d"(v) : "r"(f), "n"(BIT), "d"(a), "d"(b));
  return v;
}
template <int BIT>
__device__ __forceinline__ void zero_if(int f, double& v) {
  asm("{\n.reg .pred p;\n.reg .b32 t;\nand.b32 t, %1, %2;\nsetp.ne.b32 p, t, 0;\n@p mov.f64 %0, 0d0000000000000000;\n}\n" : "+d"(v) : "r"(f), "n"(BIT));
}
template <int BIT>
__device__ __forceinline__ void zero_if(int f, double& v0, double& v1) {
  asm("{\n.reg .pred p;\n.reg .b32 t;\nand.b32 t, %2, %3;\nsetp.ne.b32 p, t, 0;\n@p mov.f64 %0, 0d0000000000000000;\n@p mov.f64 %1, 0d0000000000000000;\n}\n"
      : "+d"(v0), "+d"(v1) : "r"(f), "n"(BIT));
}
}  // namespace wp
}  // namespace hb
#define HB_WP __device__ __forceinline__
#define HB_WP_COLD __device__ __noinline__   // rare paths: kept out of the register allocation of the hot one
#else
#include <pthread.h>
#include <thread>
#include <utility>
#include <vector>
namespace hb {
namespace wp {
typedef unsigned long long bar_t;
struct PendingBulk { bar_t* bar; void* dst; const void* src; size_t bytes; };
struct WarpState {
  pthread_barrier_t bar;
  double dbuf[32];
  double abuf[32], bbuf[32];
  int ibuf[32];
  std::vector<PendingBulk> bulk;   // bulk copies in flight (late mode), touched by one lane at a time
};
struct LaneCtx {
  int tid = 0, bid = 0;
  WarpState* warp = nullptr;
  double* smem = nullptr;
  bool late = false;
  std::vector<std::vector<std::pair<void*, const void*>>> groups;  // committed cp.async groups, oldest first
  std::vector<std::pair<void*, const void*>> open;                 // copies issued since the last commit
};
inline LaneCtx*& ctx() { static thread_local LaneCtx* c = nullptr; return c; }
inline int tid() { return ctx()->tid; }
inline int bid() { return ctx()->bid; }
inline void syncwarp() { pthread_barrier_wait(&ctx()->warp->bar); }
inline bool any(bool p) {
  LaneCtx* c = ctx();
  c->warp->ibuf[c->tid & 31] = p ? 1 : 0;
  syncwarp();
  bool r = false;
  for (int i = 0; i < 32; ++i) r = r || c->warp->ibuf[i] != 0;
  syncwarp();
  return r;
}
inline double shfl(double v, int src, int width) {
  LaneCtx* c = ctx();
  const int lane = c->tid & 31;
  c->warp->dbuf[lane] = v;
  syncwarp();
  const double r = c->warp->dbuf[(lane & ~(width - 1)) | (src & (width - 1))];
  syncwarp();
  return r;
}
inline double shfl_xor(double v, int m, int width) {
  LaneCtx* c = ctx();
  const int lane = c->tid & 31;
  c->warp->dbuf[lane] = v;
  syncwarp();
  const double r = c->warp->dbuf[(lane & ~(width - 1)) | ((lane ^ m) & (width - 1))];
  syncwarp();
  return r;
}
inline void cp16(void* dst, const void* src) {
  LaneCtx* c = ctx();
  if (c->late) c->open.emplace_back(dst, src);
  else memcpy(dst, src, 16);
}
inline void cp_commit() {
  LaneCtx* c = ctx();
  if (c->late) { c->groups.push_back(c->open); c->open.clear(); }
}
template <int NPENDING>
inline void cp_wait() {
  LaneCtx* c = ctx();
  while ((int)c->groups.size() > NPENDING) {
    for (auto& pr : c->groups.front()) memcpy(pr.first, pr.second, 16);
    c->groups.erase(c->groups.begin());
  }
}
inline double rcp(double x) { return 1.0 / x; }
inline double* dyn_smem() { return ctx()->smem; }
inline void dmma(double& c0, double& c1, double a, double b) {   // same fragment layout as mma.sync.m8n8k4.f64
  LaneCtx* c = ctx();
  const int lane = c->tid & 31, gid = lane >> 2, tig = lane & 3;
  c->warp->abuf[lane] = a;
  c->warp->bbuf[lane] = b;
  syncwarp();
  for (int e = 0; e < 2; ++e) {
    const int n = 2 * tig + e;
    double acc = e ? c1 : c0;
    for (int k = 0; k < 4; ++k) acc += c->warp->abuf[gid * 4 + k] * c->warp->bbuf[n * 4 + k];
    (e ? c1 : c0) = acc;
  }
  syncwarp();
}
inline int keep(int v) { return v; }
inline double* keep(double* v) { return v; }
inline long long keep(long long v) { return v; }
inline void dmma4(double& d0, double& d1, double a, double b, double c0, double c1) {
  double t0 = c0, t1 = c1;
  dmma(t0, t1, a, b);
  d0 = t0; d1 = t1;
}
template <bool KEEP>
inline void prefetch_l2(const void*) {}
typedef double* sa_t;
inline sa_t sa(const void* p) { return const_cast<double*>(static_cast<const double*>(p)); }
inline sa_t sadd(sa_t a, int doubles) { return a + doubles; }
template <int OFF> inline double lds(sa_t a) { return a[OFF]; }
template <int OFF> inline void lds2(sa_t a, double& v0, double& v1) { v0 = a[OFF]; v1 = a[OFF + 1]; }
template <int BIT, int OFF> inline void sts_if(int f, sa_t a, double v) { if (f & BIT) a[OFF] = v; }
template <int BIT, int OFF> inline void sts2_if(int f, sa_t a, double v0, double v1) { if (f & BIT) { a[OFF] = v0; a[OFF + 1] = v1; } }
template <int BIT> inline void stg_if(int f, double* p, double v) { if (f & BIT) *p = v; }
typedef int l2pol_t;
inline l2pol_t l2_evict_first() { return 0; }
inline l2pol_t l2_evict_last() { return 0; }
template <int BIT> inline void ldg2_if(int f, const double* p, double& v0, double& v1, l2pol_t) { if (f & BIT) { v0 = p[0]; v1 = p[1]; } }
inline double ldg_pol(const double* p, l2pol_t) { return *p; }
inline void ldg2_pol(const double* p, double& v0, double& v1, l2pol_t) { v0 = p[0]; v1 = p[1]; }
template <int BIT> inline void ldg_if(int f, const double* p, double& v) { if (f & BIT) v = *p; }
template <int BIT> inline void prefetch_l2_if(int, const void*) {}
template <int BIT> inline double sel(int f, double a, double b) { return (f & BIT) ? a : b; }
template <int BIT> inline void zero_if(int f, double& v) { if (f & BIT) v = 0.0; }
template <int BIT> inline void zero_if(int f, double& v0, double& v1) { if (f & BIT) { v0 = 0.0; v1 = 0.0; } }
// mbarrier / bulk copy: immediate mode copies at issue; late mode copies when somebody waits on the barrier
// (parities are not modelled - only data movement and its earliest/latest visibility)
inline void bar_init(bar_t* bar, unsigned) { *bar = 0; }
inline void bar_init_fence() {}
inline void bar_expect_tx(bar_t*, unsigned) {}
inline void bulk_g2s(void* dst, const void* src, unsigned bytes, bar_t* bar) {
  LaneCtx* c = ctx();
  if (c->late) c->warp->bulk.push_back(PendingBulk{bar, dst, src, (size_t)bytes});
  else memcpy(dst, src, bytes);
}
inline void bar_wait(bar_t* bar, unsigned) {   // warp-collective in the emulation (every lane of the warp must call it)
  LaneCtx* c = ctx();
  syncwarp();
  if ((c->tid & 31) == 0) {
    auto& v = c->warp->bulk;
    for (size_t i = 0; i < v.size();) {
      if (v[i].bar == bar) { memcpy(v[i].dst, v[i].src, v[i].bytes); v.erase(v.begin() + i); }
      else ++i;
    }
  }
  syncwarp();
}

// run Body::run(args) for a grid of CTAs, CTA after CTA, one OS thread per lane
template <class Body, class Args>
inline void launch(const Args& a, int grid, int threads, size_t smem_bytes) {
  const char* env = getenv("HB_EMU_CPASYNC_LATE");
  const bool late = env != nullptr && env[0] == '1';
  const int nwarps = (threads + 31) / 32;
  for (int b = 0; b < grid; ++b) {
    std::vector<double> smem(smem_bytes / 8 + 16, 0.0);
    double* sm = smem.data();
    while ((reinterpret_cast<uintptr_t>(sm) & 127) != 0) ++sm;
    std::vector<WarpState> ws(nwarps);
    for (int w = 0; w < nwarps; ++w) {
      const int cnt = (w == nwarps - 1) ? threads - 32 * w : 32;
      pthread_barrier_init(&ws[w].bar, nullptr, (unsigned)cnt);
    }
    std::vector<LaneCtx> lc(threads);
    std::vector<std::thread> th;
    th.reserve(threads);
    for (int t = 0; t < threads; ++t) {
      lc[t].tid = t; lc[t].bid = b; lc[t].warp = &ws[t / 32]; lc[t].smem = sm; lc[t].late = late;
      th.emplace_back([&a, &lc, t]() { ctx() = &lc[t]; Body::run(a); ctx() = nullptr; });
    }
    for (auto& t : th) t.join();
    for (int w = 0; w < nwarps; ++w) pthread_barrier_destroy(&ws[w].bar);
  }
}
}  // namespace wp
}  // namespace hb
#define HB_WP inline
#define HB_WP_COLD inline
#endif
