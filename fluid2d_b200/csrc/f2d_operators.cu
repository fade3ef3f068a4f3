// Stencil operators of core/fortran_operators.f90, the periodic halo fill of
// gmg/fortran_multigrid.f90:365-412, and the whole-state combinations of
// core/timescheme.py -- memory-bound elementwise / 5-point kernels.
#include "f2d_common.cuh"

namespace f2d {
char g_err[512] = "";
long long g_launches = 0;
}  // namespace f2d

using namespace f2d;

extern "C" int f2d_abi_version(void) { return F2D_ABI_VERSION; }
extern "C" const char *f2d_last_error(void) { return f2d::g_err; }
extern "C" long long f2d_launch_count(void) { return f2d::g_launches; }
extern "C" void f2d_launch_count_reset(void) { f2d::g_launches = 0; }

extern "C" int f2d_copy(void *dst, const void *src, size_t nbytes, f2d_stream_t s) {
  if (!dst || !src) return fail(F2D_ERR_ARG, "copy: null");
  F2D_CUDA(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToDevice, S(s)));
  return F2D_OK;
}
extern "C" int f2d_zero(void *dst, size_t nbytes, f2d_stream_t s) {
  if (!dst) return fail(F2D_ERR_ARG, "zero: null");
  F2D_CUDA(cudaMemsetAsync(dst, 0, nbytes, S(s)));
  return F2D_OK;
}

// ---------------------------------------------------------------------------
// halo fill: one thread per halo cell, pulls from the periodic interior source
// ---------------------------------------------------------------------------
template <typename T>
__global__ void k_fill_halo(T *__restrict__ x, int ny, int nx, int nh, int ywrap) {
  // halo cells: 2*nh full rows + (ny-2nh) rows x 2*nh columns; without ywrap (y-slab
  // decomposition) only the side columns of the interior rows are local
  long long nrowcells = 2LL * nh * nx;
  long long total = nrowcells + 2LL * nh * (ny - 2 * nh);
  for (long long t = (ywrap ? 0 : nrowcells) + blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    int j, i;
    if (t < nrowcells) {
      int r = (int)(t / nx);
      i = (int)(t % nx);
      j = r < nh ? r : ny - 2 * nh + r;  // r in [nh,2nh) -> top rows
    } else {
      long long q = t - nrowcells;
      int r = (int)(q / (2 * nh));
      int c = (int)(q % (2 * nh));
      j = nh + r;
      i = c < nh ? c : nx - 2 * nh + c;
    }
    x[(size_t)j * nx + i] = x[(size_t)wrap_src(j, ny, nh) * nx + wrap_src(i, nx, nh)];
  }
}

template <typename T>
static int fill_halo_t(T *x, int nh, int ny, int nx, cudaStream_t s, int ywrap = 1) {
  if (!x || ny <= 2 * nh || nx <= 2 * nh || nh < 1) return fail(F2D_ERR_ARG, "fill_halo: bad shape");
  if (ny - 2 * nh < nh || nx - 2 * nh < nh) return fail(F2D_ERR_ARG, "fill_halo: interior narrower than halo");
  long long total = 2LL * nh * nx + 2LL * nh * (ny - 2 * nh);
  int threads = 256;
  int blocks = cdiv(total, threads);
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_fill_halo<T><<<blocks, threads, 0, s>>>(x, ny, nx, nh, ywrap);
  F2D_LAUNCHED();
  return F2D_OK;
}

extern "C" int f2d_fill_halo(double *x, int nh, int ny, int nx, f2d_stream_t s) {
  return fill_halo_t<double>(x, nh, ny, nx, S(s));
}
extern "C" int f2d_fill_halo_i8(int8_t *x, int nh, int ny, int nx, f2d_stream_t s) {
  return fill_halo_t<int8_t>(x, nh, ny, nx, S(s));
}
extern "C" int f2d_fill_halo_x(double *x, int nh, int ny, int nx, f2d_stream_t s) {
  return fill_halo_t<double>(x, nh, ny, nx, S(s), 0);
}

// ---------------------------------------------------------------------------
// 2-D launch helper: thread (i,j) over a [ny][nx] array, 32x8 blocks
// ---------------------------------------------------------------------------
static inline dim3 grid2d(int ny, int nx, dim3 b) { return dim3(cdiv(nx, b.x), cdiv(ny, b.y)); }
#define IJ()                                       \
  int i = blockIdx.x * blockDim.x + threadIdx.x;   \
  int j = blockIdx.y * blockDim.y + threadIdx.y;   \
  if (i >= nx || j >= ny) return;                  \
  size_t c = (size_t)j * nx + i

// celltocorner: fortran_operators.f90:44-64, xp(1..m-1,1..n-1)
__global__ void k_celltocorner(const double *__restrict__ xr, double *__restrict__ xp, int ny, int nx) {
  IJ();
  if (j > ny - 2 || i > nx - 2) return;
  xp[c] = 0.25 * (((xr[c] + xr[c + 1]) + xr[c + nx]) + xr[c + nx + 1]);
}
extern "C" int f2d_celltocorner(const double *xr, double *xp, int ny, int nx, f2d_stream_t s) {
  if (!xr || !xp || ny < 2 || nx < 2) return fail(F2D_ERR_ARG, "celltocorner: bad args");
  dim3 b(32, 8);
  k_celltocorner<<<grid2d(ny, nx, b), b, 0, S(s)>>>(xr, xp, ny, nx);
  F2D_LAUNCHED();
  return F2D_OK;
}

// cornertocell: fortran_operators.f90:102-122, xr(2..m,2..n)
__global__ void k_cornertocell(const double *__restrict__ xp, double *__restrict__ xr, int ny, int nx) {
  IJ();
  if (j < 1 || i < 1) return;
  xr[c] = 0.25 * (((xp[c] + xp[c - 1]) + xp[c - nx]) + xp[c - nx - 1]);
}
extern "C" int f2d_cornertocell(const double *xp, double *xr, int ny, int nx, f2d_stream_t s) {
  if (!xr || !xp || ny < 2 || nx < 2) return fail(F2D_ERR_ARG, "cornertocell: bad args");
  dim3 b(32, 8);
  k_cornertocell<<<grid2d(ny, nx, b), b, 0, S(s)>>>(xp, xr, ny, nx);
  F2D_LAUNCHED();
  return F2D_OK;
}

// computeorthogradient: fortran_operators.f90:2-39, rows/cols 2..m-1 / 2..n-1
__global__ void k_orthogradient(const int8_t *__restrict__ msk, const double *__restrict__ psi,
                                double zdx, double zdy, double *__restrict__ u,
                                double *__restrict__ v, int ny, int nx) {
  IJ();
  if (j < 1 || j > ny - 2 || i < 1 || i > nx - 2) return;
  int m0 = msk[c];
  double p = psi[c];
  u[c] = (m0 + msk[c + 1] == 2) ? zdy * (psi[c - nx] - p) : 0.;
  v[c] = (m0 + msk[c + nx] == 2) ? zdx * (p - psi[c - 1]) : 0.;
}
extern "C" int f2d_orthogradient(const int8_t *msk, const double *psi, double dx, double dy, int nh,
                                 double *u, double *v, int ny, int nx, f2d_stream_t s) {
  (void)nh;
  if (!msk || !psi || !u || !v || ny < 3 || nx < 3) return fail(F2D_ERR_ARG, "orthogradient: bad args");
  dim3 b(32, 8);
  k_orthogradient<<<grid2d(ny, nx, b), b, 0, S(s)>>>(msk, psi, 1. / dx, 1. / dy, u, v, ny, nx);
  F2D_LAUNCHED();
  return F2D_OK;
}

// psi *= mskp, then computeorthogradient, in one pass (operators.py:481,493 without
// islands).  A thread needs the masked psi of its south and west neighbours; it forms
// them itself from psi and mskp.  psi is updated in place: a neighbour may already have
// been masked when it is read, which is harmless because masking is idempotent
// (x*1 = x, x*0 = +-0) -- this is why the island case (psi += psi_island) is not fused.
__global__ void k_mask_orthogradient(const int8_t *__restrict__ msk, const int8_t *__restrict__ mskp,
                                     double *psi, double zdx, double zdy, double *__restrict__ u,
                                     double *__restrict__ v, int ny, int nx) {
  IJ();
  double p = __dmul_rn(psi[c], (double)mskp[c]);
  psi[c] = p;
  if (j < 1 || j > ny - 2 || i < 1 || i > nx - 2) return;
  int m0 = msk[c];
  double ps = __dmul_rn(psi[c - nx], (double)mskp[c - nx]);
  double pw = __dmul_rn(psi[c - 1], (double)mskp[c - 1]);
  u[c] = (m0 + msk[c + 1] == 2) ? zdy * (ps - p) : 0.;
  v[c] = (m0 + msk[c + nx] == 2) ? zdx * (p - pw) : 0.;
}
extern "C" int f2d_mask_orthogradient(const int8_t *msk, const int8_t *mskp, double *psi, double dx, double dy,
                                      int nh, double *u, double *v, int ny, int nx, f2d_stream_t s) {
  (void)nh;
  if (!msk || !mskp || !psi || !u || !v || ny < 3 || nx < 3) return fail(F2D_ERR_ARG, "mask_orthogradient: bad args");
  dim3 b(32, 8);
  k_mask_orthogradient<<<grid2d(ny, nx, b), b, 0, S(s)>>>(msk, mskp, psi, 1. / dx, 1. / dy, u, v, ny, nx);
  F2D_LAUNCHED();
  return F2D_OK;
}

// add_diffusion: fortran_operators.f90:125-156, rows/cols 2..m-1 where msk==1
__global__ void k_add_diffusion(const int8_t *__restrict__ msk, const double *__restrict__ t,
                                double coef, double *__restrict__ d, int ny, int nx) {
  IJ();
  if (j < 1 || j > ny - 2 || i < 1 || i > nx - 2) return;
  if (msk[c] != 1) return;
  double tc = t[c];
  double acc = msk[c - 1] * (t[c - 1] - tc);
  acc = acc + msk[c + 1] * (t[c + 1] - tc);
  acc = acc + msk[c - nx] * (t[c - nx] - tc);
  acc = acc + msk[c + nx] * (t[c + nx] - tc);
  d[c] = d[c] + coef * acc;
}
extern "C" int f2d_add_diffusion(const int8_t *msk, const double *trac, double dx, int nh, double Kdiff,
                                 double *dtrac, int ny, int nx, int fill, f2d_stream_t s) {
  if (!msk || !trac || !dtrac || ny < 3 || nx < 3) return fail(F2D_ERR_ARG, "add_diffusion: bad args");
  dim3 b(32, 8);
  k_add_diffusion<<<grid2d(ny, nx, b), b, 0, S(s)>>>(msk, trac, Kdiff / (dx * dx), dtrac, ny, nx);
  F2D_LAUNCHED();
  if (fill == 2) return f2d_fill_halo_x(dtrac, nh, ny, nx, s);
  if (fill) return f2d_fill_halo(dtrac, nh, ny, nx, s);
  return F2D_OK;
}

// add_torque: fortran_operators.f90:330-381.  ml/mr = max(1, pair sums); the update is
// applied where ml+mr == 4, i.e. msk(i-1)=msk(i)=msk(i+1)=1.  premask: y *= msk first
// on the WHOLE array (operators.py:311).
__global__ void k_add_torque(const int8_t *__restrict__ msk, const double *__restrict__ b, double coef,
                             double *__restrict__ d, int ny, int nx, int nh, int premask) {
  IJ();
  double y = d[c];
  int m0 = msk[c];
  bool touched = false;
  if (premask) { y = y * (double)m0; touched = true; }
  if (j >= nh && j < ny - nh && i >= nh && i < nx - nh) {
    int ml = m0 + msk[c - 1];
    if (ml < 1) ml = 1;
    int mr = msk[c + 1] + m0;
    if (mr < 1) mr = 1;
    if (ml + mr == 4) { y = y + ((b[c + 1] - b[c - 1]) * coef) * (double)m0; touched = true; }
  }
  if (touched) d[c] = y;
}
extern "C" int f2d_add_torque(const int8_t *msk, const double *buoy, double dx, int nh, double gravity,
                              double *domega, int ny, int nx, int premask, int fill, f2d_stream_t s) {
  if (!msk || !buoy || !domega || ny <= 2 * nh || nx <= 2 * nh) return fail(F2D_ERR_ARG, "add_torque: bad args");
  dim3 b(32, 8);
  k_add_torque<<<grid2d(ny, nx, b), b, 0, S(s)>>>(msk, buoy, 0.5 * gravity / dx, domega, ny, nx, nh, premask);
  F2D_LAUNCHED();
  if (fill == 2) return f2d_fill_halo_x(domega, nh, ny, nx, s);
  if (fill) return f2d_fill_halo(domega, nh, ny, nx, s);
  return F2D_OK;
}

// computenoslipsourceterm: fortran_operators.f90:221-277 as a gather.
// The Fortran visits (j,i), j=nh+1..m-nh+1, i=nh+1..n-nh+1 in row-major order; at each
// visit it sets y(j,i)=0, then adds face terms either to y(j,i) (cell is fluid) or to
// the already-visited y(j,i-1) / y(j-1,i) (cell is solid).  Gathered per target cell
// T=(j,i), in the order the Fortran performs the additions:
//   1. own west face  (msk(T)!=0, msk(j,i-1)+msk(T)==1)      : y -= vW(j,i)
//   2. own south face (msk(T)!=0, msk(j-1,i)+msk(T)==1)      : y += uS(j,i)
//   3. east neighbour E=(j,i+1) solid, visited, msk(T)+msk(E)==1 : y += vW(j,i+1)
//   4. north neighbour N=(j+1,i) solid, visited, msk(T)+msk(N)==1: y -= uS(j+1,i)
// Cells with row<=nh or (visited row and col<=nh) are zeroed and then only receive 3./4.
// vW(j,i) = (x(j,i)+x(j-1,i)-x(j,i-2)-x(j-1,i-2))*cff ; uS(j,i) = -(x(j,i)+x(j,i-1)-x(j-2,i)-x(j-2,i-1))*cff
__global__ void k_noslip_source(const int8_t *__restrict__ msk, const double *__restrict__ x,
                                double *__restrict__ y, double cff, int ny, int nx, int nh) {
  IJ();
  // 1-based coordinates of the Fortran
  int J = j + 1, I = i + 1;
  int jlo = nh + 1, jhi = ny - nh + 1, ilo = nh + 1, ihi = nx - nh + 1;
  bool visited = (J >= jlo && J <= jhi && I >= ilo && I <= ihi);
  bool zeroed = visited || (J <= nh) || (J >= jlo && J <= jhi && I <= nh);
  // scatter targets can be (j,i-1) with i-1 = nh (zeroed column) and (j-1,i) with j-1 = nh
  if (!zeroed) return;
  double acc = 0.;
  int mT = msk[c];
  if (visited && mT != 0) {
    if (msk[c - 1] + mT == 1) {
      double vW = (((x[c] + x[c - nx]) - x[c - 2]) - x[c - nx - 2]) * cff;
      acc = acc - vW;
    }
    if (msk[c - nx] + mT == 1) {
      double uS = -((((x[c] + x[c - 1]) - x[c - 2 * nx]) - x[c - 2 * nx - 1]) * cff);
      acc = acc + uS;
    }
  }
  // east neighbour visited?
  if (J >= jlo && J <= jhi && (I + 1) >= ilo && (I + 1) <= ihi) {
    int mE = msk[c + 1];
    if (mE == 0 && mT + mE == 1) {
      size_t e = c + 1;
      double vW = (((x[e] + x[e - nx]) - x[e - 2]) - x[e - nx - 2]) * cff;
      acc = acc + vW;
    }
  }
  if ((J + 1) >= jlo && (J + 1) <= jhi && I >= ilo && I <= ihi) {
    int mN = msk[c + nx];
    if (mN == 0 && mT + mN == 1) {
      size_t n_ = c + nx;
      double uS = -((((x[n_] + x[n_ - 1]) - x[n_ - 2 * nx]) - x[n_ - 2 * nx - 1]) * cff);
      acc = acc - uS;
    }
  }
  y[c] = acc;
}
extern "C" int f2d_noslip_source(const int8_t *msk, const double *psi, double *y, double dx, double dy,
                                 int nh, int ny, int nx, f2d_stream_t s) {
  if (!msk || !psi || !y || ny <= 2 * nh || nx <= 2 * nh) return fail(F2D_ERR_ARG, "noslip_source: bad args");
  dim3 b(32, 8);
  k_noslip_source<<<grid2d(ny, nx, b), b, 0, S(s)>>>(msk, psi, y, 1. / (2 * dx * dy), ny, nx, nh);
  F2D_LAUNCHED();
  return F2D_OK;
}

// ---------------------------------------------------------------------------
// elementwise: time-scheme combinations and model glue.  __dmul_rn/__dadd_rn keep
// numpy's rounding sequence (a product is rounded before it is added).
// ---------------------------------------------------------------------------
template <class F>
__global__ void k_elementwise(size_t n, F f) {
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) f(k);
}
template <class F>
static int elementwise(size_t n, cudaStream_t s, F f) {
  if (n == 0) return F2D_OK;
  int threads = 256;
  long long blocks = (long long)((n + threads - 1) / threads);
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  k_elementwise<<<(int)blocks, threads, 0, s>>>(n, f);
  F2D_LAUNCHED();
  return F2D_OK;
}

extern "C" int f2d_ts_axpy(double *y, double c, const double *a, size_t n, f2d_stream_t s) {
  return elementwise(n, S(s), [=] __device__(size_t k) { y[k] = add_rn(y[k], mul_rn(c, a[k])); });
}
extern "C" int f2d_ts_xpay(double *out, const double *x, double c, const double *a, size_t n, f2d_stream_t s) {
  return elementwise(n, S(s), [=] __device__(size_t k) { out[k] = add_rn(x[k], mul_rn(c, a[k])); });
}
extern "C" int f2d_ts_xpay2(double *out, const double *x, double c, const double *a, const double *b, size_t n,
                            f2d_stream_t s) {
  return elementwise(n, S(s), [=] __device__(size_t k) { out[k] = add_rn(x[k], mul_rn(c, add_rn(a[k], b[k]))); });
}
extern "C" int f2d_ts_rk3ssp_final(double *x, double c, const double *a, const double *b, const double *d,
                                   size_t n, f2d_stream_t s) {
  return elementwise(n, S(s), [=] __device__(size_t k) {
    x[k] = add_rn(x[k], mul_rn(c, add_rn(add_rn(a[k], b[k]), mul_rn(4., d[k]))));
  });
}
extern "C" int f2d_ts_ab2(double *x, double c0, const double *a, double c1, const double *b, size_t n,
                          f2d_stream_t s) {
  return elementwise(n, S(s), [=] __device__(size_t k) {
    x[k] = add_rn(x[k], add_rn(mul_rn(c0, a[k]), -mul_rn(c1, b[k])));
  });
}
extern "C" int f2d_ts_ab3(double *x, double c0, const double *a, double c1, const double *b, double c2,
                          const double *d, size_t n, f2d_stream_t s) {
  return elementwise(n, S(s), [=] __device__(size_t k) {
    x[k] = add_rn(x[k], add_rn(add_rn(mul_rn(c0, a[k]), -mul_rn(c1, b[k])), mul_rn(c2, d[k])));
  });
}
extern "C" int f2d_ts_set_xpay(double *x, const double *xb, double c, const double *a, size_t n, f2d_stream_t s) {
  return elementwise(n, S(s), [=] __device__(size_t k) { x[k] = add_rn(xb[k], mul_rn(c, a[k])); });
}
extern "C" int f2d_ts_asselin(double *xs, double c, const double *x, const double *xb, size_t n, f2d_stream_t s) {
  return elementwise(n, S(s), [=] __device__(size_t k) {
    xs[k] = add_rn(xs[k], mul_rn(c, add_rn(add_rn(x[k], xb[k]), -mul_rn(2., xs[k]))));
  });
}
extern "C" int f2d_ts_am3(double *x, const double *xs, const double *xb, size_t n, f2d_stream_t s) {
  const double w = 1. / 12.;
  return elementwise(n, S(s), [=] __device__(size_t k) {
    x[k] = mul_rn(w, add_rn(add_rn(mul_rn(5., x[k]), mul_rn(8., xs[k])), -xb[k]));
  });
}
extern "C" int f2d_mul_field(double *y, const double *a, size_t n, f2d_stream_t s) {
  return elementwise(n, S(s), [=] __device__(size_t k) { y[k] = mul_rn(y[k], a[k]); });
}
extern "C" int f2d_mul_mask(double *y, const int8_t *a, size_t n, f2d_stream_t s) {
  return elementwise(n, S(s), [=] __device__(size_t k) { y[k] = mul_rn(y[k], (double)a[k]); });
}
extern "C" int f2d_scale(double *y, double alpha, size_t n, f2d_stream_t s) {
  return elementwise(n, S(s), [=] __device__(size_t k) { y[k] = mul_rn(y[k], alpha); });
}
extern "C" int f2d_add_scaled(double *y, double alpha, const double *a, size_t n, f2d_stream_t s) {
  return elementwise(n, S(s), [=] __device__(size_t k) { y[k] = add_rn(y[k], mul_rn(alpha, a[k])); });
}
extern "C" int f2d_add_scaled_mask(double *y, double alpha, const int8_t *a, size_t n, f2d_stream_t s) {
  return elementwise(n, S(s), [=] __device__(size_t k) { y[k] = add_rn(y[k], mul_rn(alpha, (double)a[k])); });
}
extern "C" int f2d_set_sum(double *y, const double *a, double alpha, const double *b, size_t n, f2d_stream_t s) {
  return elementwise(n, S(s), [=] __device__(size_t k) { y[k] = add_rn(a[k], mul_rn(alpha, b[k])); });
}
// core/fluxes.py:120-127: cell-centred velocities uc = 0.5*(u + roll(u,1,axis=1)),
// vc = 0.5*(v + roll(v,1,axis=0)) on the interior, periodic images stored by the
// owning cell (the fill_halo that follows each in the reference)
__global__ void k_flx_cellvel(const double *__restrict__ u, const double *__restrict__ v, double *__restrict__ uc,
                              double *__restrict__ vc, int nh, int ny, int nx, int fill) {
  int i = blockIdx.x * blockDim.x + threadIdx.x + nh;
  int j = blockIdx.y * blockDim.y + threadIdx.y + nh;
  if (i >= nx - nh || j >= ny - nh) return;
  size_t c = (size_t)j * nx + i;
  double a = mul_rn(0.5, add_rn(u[c], u[c - 1]));
  double b = mul_rn(0.5, add_rn(v[c], v[c - nx]));
  uc[c] = a;
  vc[c] = b;
  if (fill)
    for_each_halo_image(j, i, ny, nx, nh, [&](int jj, int ii) {
      uc[(size_t)jj * nx + ii] = a;
      vc[(size_t)jj * nx + ii] = b;
    }, fill == 1);
}
extern "C" int f2d_flx_cellvel(const double *u, const double *v, double *uc, double *vc, int nh, int ny, int nx,
                               int fill_halo, f2d_stream_t s) {
  if (nh < 1 || ny <= 2 * nh || nx <= 2 * nh) return fail(F2D_ERR_ARG, "flx_cellvel: bad shape");
  dim3 blk(32, 8), grd(cdiv(nx - 2 * nh, 32), cdiv(ny - 2 * nh, 8));
  k_flx_cellvel<<<grd, blk, 0, S(s)>>>(u, v, uc, vc, nh, ny, nx, fill_halo);
  F2D_LAUNCHED();
  return F2D_OK;
}
// core/fluxes.py:160-177: rev = cff*(fwd + sign*bwd), irr = cff*(fwd - sign*bwd), sign = +-1
extern "C" int f2d_flx_split(double *rev, double *irr, const double *fwd, const double *bwd, double cff, double sign,
                             size_t n, f2d_stream_t s) {
  return elementwise(n, S(s), [=] __device__(size_t k) {
    double sb = mul_rn(sign, bwd[k]);
    rev[k] = mul_rn(cff, add_rn(fwd[k], sb));
    irr[k] = mul_rn(cff, add_rn(fwd[k], -sb));
  });
}
extern "C" int f2d_div_scalar(double *y, double d, size_t n, f2d_stream_t s) {
  return elementwise(n, S(s), [=] __device__(size_t k) { y[k] = __ddiv_rn(y[k], d); });
}
extern "C" int f2d_sub_lin2_mask(double *y, double pa, const double *a, double pb, const double *b,
                                 const int8_t *mask, size_t n, f2d_stream_t s) {
  return elementwise(n, S(s), [=] __device__(size_t k) {
    y[k] = add_rn(y[k], -mul_rn(add_rn(mul_rn(pa, a[k]), mul_rn(pb, b[k])), (double)mask[k]));
  });
}
extern "C" int f2d_sub_lin2(double *y, double pa, const double *a, double pb, const double *b, size_t n,
                            f2d_stream_t s) {
  return elementwise(n, S(s), [=] __device__(size_t k) {
    y[k] = add_rn(y[k], -add_rn(mul_rn(a[k], pa), mul_rn(b[k], pb)));
  });
}
extern "C" int f2d_sub_devscalar(double *y, const double *dev_scalar, double denom, size_t n, f2d_stream_t s) {
  return elementwise(n, S(s), [=] __device__(size_t k) { y[k] = add_rn(y[k], -__ddiv_rn(dev_scalar[0], denom)); });
}
extern "C" int f2d_sub_devscalar_mask(double *y, const double *dev_scalar, double denom, const int8_t *a, size_t n,
                                      f2d_stream_t s) {
  return elementwise(n, S(s), [=] __device__(size_t k) {
    y[k] = add_rn(y[k], -mul_rn(__ddiv_rn(dev_scalar[0], denom), (double)a[k]));
  });
}
