// Flux-form advection: core/fortran_advection.f90 (adv_upwind :2-165, adv_centered
// :169-284) and the flux-storing variants of core/fortran_fluxes.f90.
//
// One CTA computes a TX x TY tile of dq = -div(U q).  The q tile (halo 3) and the mask
// tile are staged in shared memory; every east-face flux of the tile is computed once
// and shared through shared memory, the north-face flux is carried in a register
// while the thread marches up its column (the Fortran's fym).  The periodic halo fill
// that Operators.rhs_adv performs next (operators.py:231) is fused: a thread that
// owns a rim cell also stores its halo images.
//
// The running mask sums of the Fortran (mx5/mx3, my5/my3; :62-70,96-97,132-133) are
// window sums here -- the same integers.
#include "f2d_common.cuh"

using namespace f2d;

namespace {

constexpr int NH = 3;
constexpr int TX = 128;
constexpr int TY = 16;
constexpr int SW = TX + 2 * NH;  // shared tile width
constexpr int SH = TY + 2 * NH;

struct AdvC {
  double d1, d2, d3, d4, d5, c1, c2, c3;  // upwind weights (float32 literals)
  double e1, e2, e3, f1, f2, g1;          // centred weights
  double zdx, zdy, u1, aa, bb;
  int method;
};

__device__ __forceinline__ double split_speed(const AdvC &k, double vel) {
  double UU = fabs(vel);
  if (k.method == 1 && UU < k.u1) UU = k.aa * (vel * vel) + k.bb;
  return UU;
}

// upwind face flux from the 6 values q[-2..3] and masks m[-2..3] along the direction
template <int ORDER, bool MASKED>
__device__ __forceinline__ double upw_flux(const AdvC &k, double vel, double qm2, double qm1, double q0,
                                           double qp1, double qp2, double qp3, int mm2, int mm1, int m0,
                                           int mp1, int mp2, int mp3) {
  if (MASKED && (m0 + mp1 != 2)) return 0.;
  double UU = split_speed(k, vel);
  double up = 0.5 * (vel + UU);
  double um = 0.5 * (vel - UU);
  bool p5 = (ORDER == 5) && (!MASKED || (mm2 + mm1 + m0 + mp1 + mp2 == 5));
  bool p3 = (ORDER >= 3) && (!MASKED || (mm1 + m0 + mp1 == 3));
  bool n5 = (ORDER == 5) && (!MASKED || (mm1 + m0 + mp1 + mp2 + mp3 == 5));
  bool n3 = (ORDER >= 3) && (!MASKED || (m0 + mp1 + mp2 == 3));
  double qp, qm;
  if (p5)
    qp = k.d1 * qm2 + k.d2 * qm1 + k.d3 * q0 + k.d4 * qp1 + k.d5 * qp2;
  else if (p3)
    qp = k.c1 * qm1 + k.c2 * q0 + k.c3 * qp1;
  else
    qp = q0;
  if (n5)
    qm = k.d5 * qm1 + k.d4 * q0 + k.d3 * qp1 + k.d2 * qp2 + k.d1 * qp3;
  else if (n3)
    qm = k.c3 * q0 + k.c2 * qp1 + k.c1 * qp2;
  else
    qm = qp1;
  return up * qp + um * qm;
}

// centred face flux (fortran_advection.f90:229-262); note that order 6 falls straight
// to order 2 when the 6-window is not all fluid, as the Fortran's elif chain does
template <int ORDER, bool MASKED>
__device__ __forceinline__ double cen_flux(const AdvC &k, double vel, double qm2, double qm1, double q0,
                                           double qp1, double qp2, double qp3, int mm2, int mm1, int m0,
                                           int mp1, int mp2, int mp3) {
  if (MASKED && (m0 + mp1 != 2)) return 0.;
  double qp = 0.;
  if ((ORDER == 6) && (!MASKED || (mm2 + mm1 + m0 + mp1 + mp2 + mp3 == 6))) {
    qp = k.e1 * (qm2 + qp3) + k.e2 * (qm1 + qp2);
    qp = qp + k.e3 * (q0 + qp1);
  } else if ((ORDER == 4) && (!MASKED || (mm1 + m0 + mp1 + mp2 == 4))) {
    qp = k.f1 * (qm1 + qp2) + k.f2 * (q0 + qp1);
  } else if (ORDER >= 2) {
    qp = k.g1 * (q0 + qp1);
  }
  return vel * qp;
}

template <bool UPW, int ORDER, bool MASKED>
__device__ __forceinline__ double face_flux(const AdvC &k, double vel, const double *q, int qs,
                                            const int8_t *m, int ms) {
  // q points at the cell on the low side of the face; qs/ms = stride along the direction
  int mm2 = 1, mm1 = 1, m0 = 1, mp1 = 1, mp2 = 1, mp3 = 1;
  if (MASKED) {
    mm2 = m[-2 * ms]; mm1 = m[-ms]; m0 = m[0]; mp1 = m[ms]; mp2 = m[2 * ms]; mp3 = m[3 * ms];
  }
  if (UPW)
    return upw_flux<ORDER, MASKED>(k, vel, q[-2 * qs], q[-qs], q[0], q[qs], q[2 * qs], q[3 * qs], mm2, mm1,
                                   m0, mp1, mp2, mp3);
  else
    return cen_flux<ORDER, MASKED>(k, vel, q[-2 * qs], q[-qs], q[0], q[qs], q[2 * qs], q[3 * qs], mm2, mm1,
                                   m0, mp1, mp2, mp3);
}

template <bool UPW, int ORDER, bool MASKED, bool FLX>
__global__ void __launch_bounds__(TX)
k_adv(const int8_t *__restrict__ msk, const double *__restrict__ q, double *__restrict__ dq,
      const double *__restrict__ u, const double *__restrict__ v, double *__restrict__ xflx,
      double *__restrict__ yflx, AdvC k, int ny, int nx, int fill) {
  __shared__ double sq[SH][SW];
  __shared__ double sfx[TY][TX + 1];
  __shared__ int8_t sm[MASKED ? SH : 1][MASKED ? SW : 1];
  const int tx = threadIdx.x;
  const int i0 = NH + blockIdx.x * TX;  // first output column of the tile
  const int j0 = NH + blockIdx.y * TY;
  // stage q (and msk) rows j0-3 .. j0+TY+2, cols i0-3 .. i0+TX+2
  for (int r = 0; r < SH; r++) {
    int j = j0 - NH + r;
    for (int cidx = tx; cidx < SW; cidx += TX) {
      int i = i0 - NH + cidx;
      bool in = (j < ny) && (i < nx);
      sq[r][cidx] = in ? q[(size_t)j * nx + i] : 0.;
      if (MASKED) sm[r][cidx] = in ? msk[(size_t)j * nx + i] : (int8_t)0;
    }
  }
  __syncthreads();
  const int i = i0 + tx;
  const bool col_ok = i < nx - NH;
  // east-face fluxes: thread tx -> face of column i (all rows); threads 0..TY-1 also
  // compute the west face of the tile (column i0-1) for row tx
  for (int r = 0; r < TY; r++) {
    int j = j0 + r;
    double f = 0.;
    if (col_ok && j < ny - NH)
      f = face_flux<UPW, ORDER, MASKED>(k, u[(size_t)j * nx + i], &sq[r + NH][tx + NH], 1,
                                        MASKED ? &sm[r + NH][tx + NH] : nullptr, 1);
    sfx[r][tx + 1] = f;
  }
  if (tx < TY) {
    int j = j0 + tx;
    double f = 0.;
    if (j < ny - NH)
      f = face_flux<UPW, ORDER, MASKED>(k, u[(size_t)j * nx + (i0 - 1)], &sq[tx + NH][NH - 1], 1,
                                        MASKED ? &sm[tx + NH][NH - 1] : nullptr, 1);
    sfx[tx][0] = f;
    if (FLX && blockIdx.x == 0 && j < ny - NH) xflx[(size_t)j * nx + (i0 - 1)] = f;
  }
  __syncthreads();
  if (!col_ok) return;
  // north-face flux of the row below the tile, then march up
  double fym = face_flux<UPW, ORDER, MASKED>(k, v[(size_t)(j0 - 1) * nx + i], &sq[NH - 1][tx + NH], SW,
                                             MASKED ? &sm[NH - 1][tx + NH] : nullptr, SW);
  if (FLX && blockIdx.y == 0) yflx[(size_t)(j0 - 1) * nx + i] = fym;
#pragma unroll 4
  for (int r = 0; r < TY; r++) {
    int j = j0 + r;
    if (j >= ny - NH) break;
    size_t c = (size_t)j * nx + i;
    double fy = face_flux<UPW, ORDER, MASKED>(k, v[c], &sq[r + NH][tx + NH], SW,
                                              MASKED ? &sm[r + NH][tx + NH] : nullptr, SW);
    double fxe = sfx[r][tx + 1], fxw = sfx[r][tx];
    double y = -k.zdx * (fxe - fxw) - k.zdy * (fy - fym);
    dq[c] = y;
    if (FLX) {
      xflx[c] = fxe;
      yflx[c] = fy;
    }
    if (fill)
      for_each_halo_image(j, i, ny, nx, NH, [&](int jj, int ii) { dq[(size_t)jj * nx + ii] = y; });
    fym = fy;
  }
}

template <bool UPW, int ORDER>
int launch_adv(const int8_t *msk, const double *q, double *dq, const double *u, const double *v,
               double *xflx, double *yflx, const AdvC &k, int ny, int nx, int fill, bool masked,
               cudaStream_t s) {
  dim3 grid(cdiv(nx - 2 * NH, TX), cdiv(ny - 2 * NH, TY));
  bool flx = xflx != nullptr;
#define GO(M, F) k_adv<UPW, ORDER, M, F><<<grid, TX, 0, s>>>(msk, q, dq, u, v, xflx, yflx, k, ny, nx, fill)
  if (masked) {
    if (flx) GO(true, true); else GO(true, false);
  } else {
    if (flx) GO(false, true); else GO(false, false);
  }
#undef GO
  F2D_LAUNCHED();
  return F2D_OK;
}

int adv_common(bool upw, const int8_t *msk, const double *q, double *dq, const double *u, const double *v,
               double *xflx, double *yflx, const double *cst, int nh, int method, int order, int ny, int nx,
               int fill, cudaStream_t s) {
  if (nh != NH) return fail(F2D_ERR_NH, "NHALO = 3 is compulsory with UP5");
  if (!q || !dq || !u || !v || !cst) return fail(F2D_ERR_ARG, "adv: null pointer");
  if ((xflx == nullptr) != (yflx == nullptr)) return fail(F2D_ERR_ARG, "adv: xflx and yflx go together");
  if (ny < 2 * NH + NH || nx < 2 * NH + NH) return fail(F2D_ERR_ARG, "adv: grid too small");
  if (method != 0 && method != 1) return fail(F2D_ERR_ARG, "adv: flux splitting method must be 0 or 1");
  AdvC k;
  k.d1 = (double)(1.f / 30.f);  k.d2 = (double)(-13.f / 60.f); k.d3 = (double)(47.f / 60.f);
  k.d4 = (double)(9.f / 20.f);  k.d5 = (double)(-1.f / 20.f);
  k.c1 = (double)(-1.f / 6.f);  k.c2 = (double)(5.f / 6.f);    k.c3 = (double)(2.f / 6.f);
  k.e1 = (double)(1.f / 60.f);  k.e2 = (double)(-2.f / 15.f);  k.e3 = (double)(37.f / 60.f);
  k.f1 = (double)(-1.f / 12.f); k.f2 = (double)(7.f / 12.f);   k.g1 = (double)(1.f / 2.f);
  double dx = cst[0], dy = cst[1], umax = cst[3], aparab = cst[4];
  k.zdx = 1. / dx;
  k.zdy = 1. / dy;
  k.u1 = aparab * umax;
  k.aa = 1. / (2. * k.u1);  // +inf when umax == 0: the parabolic branch is then never taken
  k.bb = k.u1 * 0.5;
  k.method = method;
  bool masked = msk != nullptr;
  if (upw) {
    switch (order) {
      case 1: return launch_adv<true, 1>(msk, q, dq, u, v, xflx, yflx, k, ny, nx, fill, masked, s);
      case 3: return launch_adv<true, 3>(msk, q, dq, u, v, xflx, yflx, k, ny, nx, fill, masked, s);
      case 5: return launch_adv<true, 5>(msk, q, dq, u, v, xflx, yflx, k, ny, nx, fill, masked, s);
    }
    return fail(F2D_ERR_ARG, "adv_upwind: order must be 1, 3 or 5");
  }
  switch (order) {
    case 2: return launch_adv<false, 2>(msk, q, dq, u, v, xflx, yflx, k, ny, nx, fill, masked, s);
    case 4: return launch_adv<false, 4>(msk, q, dq, u, v, xflx, yflx, k, ny, nx, fill, masked, s);
    case 6: return launch_adv<false, 6>(msk, q, dq, u, v, xflx, yflx, k, ny, nx, fill, masked, s);
  }
  return fail(F2D_ERR_ARG, "adv_centered: order must be 2, 4 or 6");
}

}  // namespace

// msk == NULL selects the all-fluid specialisation (no mask reads); callers pass NULL
// only when every cell of msk, halo included, is 1 (geometry 'perio', grid.py:82-84).
extern "C" int f2d_adv_upwind(const int8_t *msk, const double *q, double *dq, const double *u, const double *v,
                              double *xflx, double *yflx, const double *cst5, int nh, int method, int order,
                              int ny, int nx, int fill_halo, f2d_stream_t s) {
  return adv_common(true, msk, q, dq, u, v, xflx, yflx, cst5, nh, method, order, ny, nx, fill_halo, S(s));
}
extern "C" int f2d_adv_centered(const int8_t *msk, const double *q, double *dq, const double *u,
                                const double *v, double *xflx, double *yflx, const double *cst5, int nh,
                                int method, int order, int ny, int nx, int fill_halo, f2d_stream_t s) {
  (void)method;
  return adv_common(false, msk, q, dq, u, v, xflx, yflx, cst5, nh, 0, order, ny, nx, fill_halo, S(s));
}
