// Interior masked reductions: core/fortran_diag.f90, computenorm/computeinner of
// gmg/fortran_multigrid.f90:813-867, Grid.domain_integration (grid.py:132-140) and
// the fused Euler diagnostics (euler.py:185-223).
//
// Two deterministic stages: (1) RB blocks sweep the interior rows (block b takes rows
// b, b+RB, ...), threads accumulate privately, warp-shuffle + shared-memory tree per
// block -> scratch[b][o]; (2) one block folds the RB partials in a fixed order.
#include "f2d_common.cuh"

using namespace f2d;

namespace {

constexpr int RB = 148 * 4;  // stage-1 blocks
constexpr int RT = 256;      // threads per block
constexpr int MAXOUT = 8;

template <int NOUT, unsigned MAXMASK>
__device__ __forceinline__ void combine(double *a, const double *b) {
#pragma unroll
  for (int o = 0; o < NOUT; o++) a[o] = ((MAXMASK >> o) & 1u) ? fmax(a[o], b[o]) : a[o] + b[o];
}

template <int NOUT, unsigned MAXMASK>
__device__ __forceinline__ void block_reduce(double *acc, double *out) {
  __shared__ double sh[RT / 32][NOUT];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    double other[NOUT];
#pragma unroll
    for (int o = 0; o < NOUT; o++) other[o] = __shfl_down_sync(0xffffffffu, acc[o], off);
    combine<NOUT, MAXMASK>(acc, other);
  }
  if (lane == 0)
#pragma unroll
    for (int o = 0; o < NOUT; o++) sh[w][o] = acc[o];
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < RT / 32; k++) combine<NOUT, MAXMASK>(acc, sh[k]);
#pragma unroll
    for (int o = 0; o < NOUT; o++) out[o] = acc[o];
  }
}

template <int NOUT, unsigned MAXMASK, class F>
__global__ void __launch_bounds__(RT) k_reduce1(int ny, int nx, int nh, double *__restrict__ scratch, F f) {
  double acc[NOUT];
#pragma unroll
  for (int o = 0; o < NOUT; o++) acc[o] = 0.;
  for (int j = nh + blockIdx.x; j < ny - nh; j += gridDim.x)
    for (int i = nh + threadIdx.x; i < nx - nh; i += RT) f((size_t)j * nx + i, acc);
  block_reduce<NOUT, MAXMASK>(acc, scratch + (size_t)blockIdx.x * NOUT);
}

template <int NOUT, unsigned MAXMASK, class G>
__global__ void __launch_bounds__(RT) k_reduce2(int nblk, const double *__restrict__ scratch,
                                                double *__restrict__ out, G finish) {
  double acc[NOUT];
#pragma unroll
  for (int o = 0; o < NOUT; o++) acc[o] = 0.;
  for (int b = threadIdx.x; b < nblk; b += RT) combine<NOUT, MAXMASK>(acc, scratch + (size_t)b * NOUT);
  __shared__ double res[NOUT];
  block_reduce<NOUT, MAXMASK>(acc, res);
  __syncthreads();
  if (threadIdx.x == 0) finish(res, out);
}

template <int NOUT, unsigned MAXMASK, class F, class G>
int reduce(int nh, int ny, int nx, double *out, double *scratch, cudaStream_t s, F f, G finish) {
  static_assert(NOUT <= MAXOUT, "too many outputs");
  if (!out || !scratch) return fail(F2D_ERR_ARG, "reduce: null out/scratch");
  if (ny <= 2 * nh || nx <= 2 * nh) return fail(F2D_ERR_ARG, "reduce: bad shape");
  int nblk = ny - 2 * nh < RB ? ny - 2 * nh : RB;
  prof_tag("k_reduce1<nout%d> %dx%d", NOUT, nx - 2 * nh, ny - 2 * nh);
  k_reduce1<NOUT, MAXMASK><<<nblk, RT, 0, s>>>(ny, nx, nh, scratch, f);
  F2D_LAUNCHED();
  k_reduce2<NOUT, MAXMASK><<<1, RT, 0, s>>>(nblk, scratch, out, finish);
  F2D_LAUNCHED();
  return F2D_OK;
}

struct Copy1 { __device__ void operator()(const double *r, double *o) const { o[0] = r[0]; } };
struct Copy2 { __device__ void operator()(const double *r, double *o) const { o[0] = r[0]; o[1] = r[1]; } };

}  // namespace

extern "C" size_t f2d_reduce_scratch_len(void) { return (size_t)RB * MAXOUT; }

extern "C" int f2d_computedotprod(const int8_t *msk, const double *x, const double *y, int nh, int ny, int nx,
                                  double *out, double *scratch, f2d_stream_t s) {
  if (!msk || !x || !y) return fail(F2D_ERR_ARG, "computedotprod: null");
  return reduce<1, 0u>(nh, ny, nx, out, scratch, S(s),
                       [=] __device__(size_t c, double *a) { if (msk[c] != 0) a[0] += x[c] * y[c]; }, Copy1());
}
extern "C" int f2d_computemax(const int8_t *msk, const double *x, int nh, int ny, int nx, double *out,
                              double *scratch, f2d_stream_t s) {
  if (!msk || !x) return fail(F2D_ERR_ARG, "computemax: null");
  return reduce<1, 1u>(nh, ny, nx, out, scratch, S(s),
                       [=] __device__(size_t c, double *a) { if (msk[c] != 0) a[0] = fmax(a[0], fabs(x[c])); },
                       Copy1());
}
extern "C" int f2d_computesum(const int8_t *msk, const double *x, int nh, int ny, int nx, double *out,
                              double *scratch, f2d_stream_t s) {
  if (!msk || !x) return fail(F2D_ERR_ARG, "computesum: null");
  return reduce<1, 0u>(nh, ny, nx, out, scratch, S(s),
                       [=] __device__(size_t c, double *a) { if (msk[c] != 0) a[0] += x[c]; }, Copy1());
}
extern "C" int f2d_computesumandnorm(const int8_t *msk, const double *x, int nh, int ny, int nx, double *out,
                                     double *scratch, f2d_stream_t s) {
  if (!msk || !x) return fail(F2D_ERR_ARG, "computesumandnorm: null");
  return reduce<2, 0u>(nh, ny, nx, out, scratch, S(s),
                       [=] __device__(size_t c, double *a) {
                         if (msk[c] == 1) { double t = x[c]; a[0] += t; a[1] += t * t; }
                       }, Copy2());
}
extern "C" int f2d_computenormmaxu(const int8_t *msk, const double *x, int nh, int ny, int nx, double *out,
                                   double *scratch, f2d_stream_t s) {
  if (!msk || !x) return fail(F2D_ERR_ARG, "computenormmaxu: null");
  return reduce<2, 2u>(nh, ny, nx, out, scratch, S(s),
                       [=] __device__(size_t c, double *a) {
                         if (msk[c] + msk[c + 1] == 2) { double t = x[c]; a[0] += t * t; a[1] = fmax(a[1], fabs(t)); }
                       }, Copy2());
}
extern "C" int f2d_computekemaxu(const int8_t *msk, const double *u, const double *v, int nh, int ny, int nx,
                                 double *out, double *scratch, f2d_stream_t s) {
  if (!msk || !u || !v) return fail(F2D_ERR_ARG, "computekemaxu: null");
  return reduce<2, 2u>(nh, ny, nx, out, scratch, S(s),
                       [=] __device__(size_t c, double *a) {
                         if (msk[c] == 1) {
                           double ue = u[c], uw = u[c - 1], vn = v[c], vs = v[c - nx];
                           a[0] += (ue * ue + uw * uw) + (vn * vn + vs * vs);
                           a[1] = fmax(a[1], fabs(ue + uw) + fabs(vn + vs));
                         }
                       },
                       [] __device__(const double *r, double *o) { o[0] = r[0] * 0.25; o[1] = r[1] * 0.5; });
}
extern "C" int f2d_computekemaxuv(const int8_t *msk, const double *u, const double *v, int nh, int ny, int nx,
                                  double *out, double *scratch, f2d_stream_t s) {
  if (!msk || !u || !v) return fail(F2D_ERR_ARG, "computekemaxuv: null");
  return reduce<3, 6u>(nh, ny, nx, out, scratch, S(s),
                       [=] __device__(size_t c, double *a) {
                         if (msk[c] == 1) {
                           double ue = u[c], uw = u[c - 1], vn = v[c], vs = v[c - nx];
                           double zu = ue * ue + uw * uw, zv = vn * vn + vs * vs;
                           a[0] += zu + zv;
                           a[1] = fmax(a[1], zu);
                           a[2] = fmax(a[2], zv);
                         }
                       },
                       [] __device__(const double *r, double *o) {
                         o[0] = r[0] * 0.25; o[1] = sqrt(r[1] / 2.); o[2] = sqrt(r[2] / 2.);
                       });
}
extern "C" int f2d_computekewithpsi(const int8_t *msk, const double *omega, const double *psi, int nh, int ny,
                                    int nx, double *out, double *scratch, f2d_stream_t s) {
  if (!msk || !omega || !psi) return fail(F2D_ERR_ARG, "computekewithpsi: null");
  return reduce<1, 0u>(nh, ny, nx, out, scratch, S(s),
                       [=] __device__(size_t c, double *a) {
                         if (msk[c] == 1)
                           a[0] -= 0.125 * (((psi[c] + psi[c - nx]) + psi[c - nx - 1]) + psi[c - 1]) * omega[c];
                       }, Copy1());
}
extern "C" int f2d_computenorm(const int8_t *msk, const double *x, int nh, int ny, int nx, double *out,
                               double *scratch, f2d_stream_t s) {
  if (!msk || !x) return fail(F2D_ERR_ARG, "computenorm: null");
  return reduce<1, 0u>(nh, ny, nx, out, scratch, S(s),
                       [=] __device__(size_t c, double *a) { if (msk[c] != 0) { double t = x[c]; a[0] += t * t; } },
                       Copy1());
}
extern "C" int f2d_domain_sum(const double *x, int nh, int ny, int nx, double *out, double *scratch,
                              f2d_stream_t s) {
  if (!x) return fail(F2D_ERR_ARG, "domain_sum: null");
  return reduce<1, 0u>(nh, ny, nx, out, scratch, S(s), [=] __device__(size_t c, double *a) { a[0] += x[c]; },
                       Copy1());
}
extern "C" int f2d_diag_euler(const int8_t *msk, const double *u, const double *v, const double *w,
                              const double *psi, const double *source, const double *xr, const double *yr,
                              int nh, int ny, int nx, double *out, double *scratch, f2d_stream_t s) {
  if (!msk || !u || !v || !w || !psi || !source || !xr || !yr) return fail(F2D_ERR_ARG, "diag_euler: null");
  return reduce<8, 1u>(nh, ny, nx, out, scratch, S(s),
                       [=] __device__(size_t c, double *a) {
                         int m = msk[c];
                         if (m != 0) {
                           double wc = w[c];
                           if (m == 1) {
                             double ue = u[c], uw = u[c - 1], vn = v[c], vs = v[c - nx];
                             a[0] = fmax(a[0], fabs(ue + uw) + fabs(vn + vs));
                             a[1] += (ue * ue + uw * uw) + (vn * vn + vs * vs);
                             a[2] += wc;
                             a[3] += wc * wc;
                           }
                           a[4] += wc * xr[c];
                           a[5] += wc * yr[c];
                           a[6] += psi[c];
                           a[7] += wc * source[c];
                         }
                       },
                       [] __device__(const double *r, double *o) {
                         o[0] = r[0] * 0.5; o[1] = r[1] * 0.25;
                         for (int k = 2; k < 8; k++) o[k] = r[k];
                       });
}
