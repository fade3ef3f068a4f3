// Internal helpers shared by the kernels of libf2d_b200.so (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/f2d_b200.h"

struct f2d_comm;

namespace f2d {

extern char g_err[512];
extern long long g_launches;

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
    cudaError_t e__ = cudaPeekAtLastError();                    \
    if (e__ != cudaSuccess) return f2d::cuda_fail(e__, __func__); \
  } while (0)

inline cudaStream_t S(f2d_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

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
template <class F>
__device__ __forceinline__ void for_each_halo_image(int j, int i, int ny, int nx, int nh,
                                                    F f, bool ywrap = true) {
  int m2 = ny - 2 * nh, n2 = nx - 2 * nh;
  int jr[3], ic[3];
  int nj = 0, ni = 0;
  jr[nj++] = j;
  if (ywrap && j < 2 * nh) jr[nj++] = j + m2;   // image in the top halo rows ny-nh..ny-1
  if (ywrap && j >= m2) jr[nj++] = j - m2;      // image in the bottom halo rows 0..nh-1
  ic[ni++] = i;
  if (i < 2 * nh) ic[ni++] = i + n2;
  if (i >= n2) ic[ni++] = i - n2;
  for (int a = 0; a < nj; a++)
    for (int b = 0; b < ni; b++)
      if (a | b) f(jr[a], ic[b]);
}

// ---- multi-GPU plumbing (f2d_comm.cu) ------------------------------------------------
int comm_rank(const f2d_comm *c);
int comm_size(const f2d_comm *c);
void *comm_alloc(f2d_comm *c, size_t nbytes);
bool comm_owns(const f2d_comm *c, const void *p);
int comm_exchange(f2d_comm *c, double *const *arrs, int narr, int nh, int ny, int nx, cudaStream_t s);
int comm_allreduce(f2d_comm *c, double *vals, int n, unsigned maxmask, cudaStream_t s);
int comm_barrier(f2d_comm *c, int all, cudaStream_t s);
int comm_gather(f2d_comm *c, const double *slab, double *full, int ny_loc, int nx, int nh, cudaStream_t s);
int comm_gather_i8(f2d_comm *c, const int8_t *slab, int8_t *full, int ny_loc, int nx, int nh, cudaStream_t s);

}  // namespace f2d
