// Internal helpers shared by the kernels of libf2d_b200.so (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <utility>

#include "../../include/f2d_b200.h"

struct f2d_comm;

namespace f2d {

extern char g_err[512];
extern long long g_launches;

// per-kernel accounting (f2d_prof_begin / f2d_prof_report): while it is on, every launch site
// records a CUDA event behind its kernel on the profiled stream; the time between two
// consecutive events is charged to the later kernel under the name given by prof_tag() (or
// the launching function's name).  Meant for runs without CUDA graphs (graph replays launch
// nothing from the host).
extern bool g_prof;
void prof_mark(const char *fallback);
void prof_tag(const char *fmt, ...);

inline int fail(int code, const char *msg) {
  snprintf(g_err, sizeof g_err, "%s", msg);
  return code;
}
inline int cuda_fail(cudaError_t e, const char *where) {
  snprintf(g_err, sizeof g_err, "CUDA error at %s: %s", where, cudaGetErrorString(e));
  return F2D_ERR_CUDA;
}

#define F2D_CUDA(call)                                  \
  do {                                                  \
    cudaError_t e__ = (call);                           \
    if (e__ != cudaSuccess) return f2d::cuda_fail(e__, #call); \
  } while (0)

// call after every kernel launch
#define F2D_LAUNCHED()                                          \
  do {                                                          \
    ++f2d::g_launches;                                          \
    if (f2d::g_prof) f2d::prof_mark(__func__);                  \
    cudaError_t e__ = cudaPeekAtLastError();                    \
    if (e__ != cudaSuccess) return f2d::cuda_fail(e__, __func__); \
  } while (0)

inline cudaStream_t S(f2d_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- programmatic dependent launch (sm_90+) ------------------------------------------------
// A kernel launched through launch_pdl() may become resident while the kernel in front of it on
// the stream (or in the captured graph) is still running: its CTAs set up their barriers and
// indices, then block in pdl_wait() until that kernel has completed and its stores are visible.
// The launch gap between two small kernels of a multigrid cycle (most of a 256^2 level's cost)
// is hidden that way.  Rules: a kernel launched by launch_pdl() calls pdl_wait() before its
// first access to global memory another kernel may have written, and pdl_trigger() as early as
// it likes (here: at entry).  F2D_PDL=1 turns the attribute on (default: plain stream order).
extern int g_pdl;   // -1: not read yet
bool pdl_enabled();
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                              Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// IEEE ops that the compiler may not contract into FMAs
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }

// periodic source index of a halo cell (identity in the interior); n includes halos
__device__ __forceinline__ int wrap_src(int j, int n, int nh) {
  int m2 = n - 2 * nh;
  return j < nh ? j + m2 : (j >= n - nh ? j - m2 : j);
}

// For an interior cell (j,i) enumerate the halo images it owns under the doubly
// periodic fill (fortran_multigrid.f90:365-412).  Usually 0, 1 or 3 locations; up to
// 8 when the interior is narrower than 2*nh (coarsest multigrid levels).
// Calls f(jj,ii) for each image.
// ywrap = false (y-slab decomposition): only the x images are local, the y halo rows
// belong to the neighbouring ranks and are filled by the exchange kernel (f2d_comm.cu).
// (Written without index arrays: local arrays indexed by a run-time count end up as
// predicated select chains or local memory, which cost the one-CTA tail kernel 35 %.)
template <bool YWRAP, class F>
__device__ __forceinline__ void for_each_halo_image_t(int j, int i, int ny, int nx, int nh, F f) {
  const int m2 = ny - 2 * nh, n2 = nx - 2 * nh;
  const bool xlo = i < 2 * nh, xhi = i >= n2;
  if (xlo) f(j, i + n2);
  if (xhi) f(j, i - n2);
  if (YWRAP) {
    if (j < 2 * nh) {   // image in the top halo rows ny-nh..ny-1
      const int jj = j + m2;
      f(jj, i);
      if (xlo) f(jj, i + n2);
      if (xhi) f(jj, i - n2);
    }
    if (j >= m2) {      // image in the bottom halo rows 0..nh-1
      const int jj = j - m2;
      f(jj, i);
      if (xlo) f(jj, i + n2);
      if (xhi) f(jj, i - n2);
    }
  }
}
// doubly periodic (single GPU / replicated levels / shared-memory tail)
template <class F>
__device__ __forceinline__ void for_each_halo_image(int j, int i, int ny, int nx, int nh, F f) {
  for_each_halo_image_t<true>(j, i, ny, nx, nh, f);
}
// run-time choice (kernels shared by slab and full levels; rim cells only)
template <class F>
__device__ __forceinline__ void for_each_halo_image(int j, int i, int ny, int nx, int nh, F f, bool ywrap) {
  if (ywrap) for_each_halo_image_t<true>(j, i, ny, nx, nh, f);
  else for_each_halo_image_t<false>(j, i, ny, nx, nh, f);
}

// ---- multi-GPU: lock-step protocol over peer memory (f2d_comm.cu) ---------------------
// Every rank keeps one control block at the start of its arena.  `done` counts the
// synchronisation points this rank has completed; slot[r] is the count rank r last
// published here.  All ranks run the same kernel sequence (SPMD), so "slot[r] >= my done"
// means: rank r has finished at least every synchronising kernel I have finished.
constexpr int MAXRANKS = 16;
constexpr int RED_SLOTS = 32;
struct Ctrl {
  unsigned long long done;                 // synchronisation points completed by this rank
  unsigned long long slot[MAXRANKS];       // slot[r]: last count published here by rank r
  unsigned int blocks_done;                // last-block detection of the current kernel
  unsigned int pad;
  double red[2][MAXRANKS][RED_SLOTS];      // all-reduce staging, double buffered
  unsigned long long red_epoch;
};

// Kernel argument of the kernels that fill their neighbours' halo rows themselves (fused
// compute + exchange): the CTAs that own the 3 top / bottom interior rows of the output
// also store them into the north / south rank's halo rows (peer stores over NVLink), and
// the last of these boundary CTAs publishes the new epoch.  The boundary CTAs of the
// consuming kernel wait for the neighbours' epoch before they read their halo rows.
// me == nullptr: single GPU, no peers.
struct Peer {
  Ctrl *me;
  Ctrl *const *peers;
  long long north_off, south_off;   // bytes from a local arena address to the same buffer on rank+1 / rank-1
  int rank, nranks;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// Boundary CTAs call this before touching halo rows (read) or the neighbour's memory
// (write): wait until the neighbour has completed every synchronising kernel this rank
// has.  All threads of the CTA must call it.
__device__ __forceinline__ void peer_wait(const Peer &P, bool south, bool north) {
  if (threadIdx.x == 0) {
    const unsigned long long d = *((volatile unsigned long long *)&P.me->done);
    const int rn = (P.rank + 1) % P.nranks, rs = (P.rank + P.nranks - 1) % P.nranks;
    if (south) while (ld_acquire_sys(&P.me->slot[rs]) < d) {}
    if (north) while (ld_acquire_sys(&P.me->slot[rn]) < d) {}
  }
  __syncthreads();
}
// Boundary CTAs call this after their stores (local and peer): the last of the `nbound`
// boundary CTAs bumps this rank's count and publishes it to both neighbours.
__device__ __forceinline__ void peer_done(const Peer &P, unsigned nbound) {
  __syncthreads();   // the CTA's stores happen-before thread 0's system fence (cumulativity)
  if (threadIdx.x == 0) {
    __threadfence_system();
    unsigned t = atomicAdd(&P.me->blocks_done, 1u);
    if (t == nbound - 1) {
      P.me->blocks_done = 0;
      __threadfence_system();   // acquire the other CTAs' fences, release the flag stores
      const unsigned long long D = P.me->done + 1;
      *((volatile unsigned long long *)&P.me->done) = D;
      // published to every rank (not only the neighbours), so that "slot[r] >= done" is a
      // valid test of any rank's progress (the gather pushes into all ranks)
      for (int r = 0; r < P.nranks; r++)
        if (r != P.rank) *((volatile unsigned long long *)&P.peers[r]->slot[P.rank]) = D;
    }
  }
}
// Boundary tile rows first: hardware block row 0 -> south boundary, 1 -> north boundary,
// the rest -> interior, so the halo rows leave early and the transfer overlaps the interior.
__device__ __forceinline__ int peer_tile_row(int by, int gy) {
  if (gy < 3) return by;
  return by == 0 ? 0 : (by == 1 ? gy - 1 : by - 1);
}
template <class T>
__device__ __forceinline__ T *peer_addr(T *p, long long off) {
  return reinterpret_cast<T *>(reinterpret_cast<char *>(p) + off);
}

// Runge-Kutta stage update of the velocities by f2d_ts_xpay / f2d_ts_xpay2 (f2d_operators.cu)
int uv_stage(const double *u, const double *v, const double *ub, const double *vb, const double *ue,
             const double *ve, double *uo, double *vo, double c, size_t n, f2d_stream_t s);

// ---- multi-GPU plumbing (f2d_comm.cu) ------------------------------------------------
Peer comm_peer(const f2d_comm *c);
int comm_drain(f2d_comm *c, cudaStream_t s);
int comm_rank(const f2d_comm *c);
int comm_size(const f2d_comm *c);
void *comm_alloc(f2d_comm *c, size_t nbytes);
bool comm_owns(const f2d_comm *c, const void *p);
int comm_exchange(f2d_comm *c, double *const *arrs, int narr, int nh, int ny, int nx, cudaStream_t s);
int comm_allreduce(f2d_comm *c, double *vals, int n, unsigned maxmask, cudaStream_t s);
int comm_barrier(f2d_comm *c, int all, cudaStream_t s);
int comm_gather(f2d_comm *c, const double *slab, double *full, int ny_loc, int nx, int nh, cudaStream_t s,
                bool barrier_first = true, bool yimages = false);
int comm_gather_i8(f2d_comm *c, const int8_t *slab, int8_t *full, int ny_loc, int nx, int nh, cudaStream_t s);

}  // namespace f2d
