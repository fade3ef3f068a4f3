// Flux-form advection: core/fortran_advection.f90 (adv_upwind :2-165, adv_centered
// :169-284) and the flux-storing variants of core/fortran_fluxes.f90.
//
// One CTA computes a TX x TY tile of dq = -div(U q).  The q tile (halo 3) and the u, v
// tiles are staged in shared memory with asynchronous copies (LDGSTS); every east-face
// flux of the tile is computed once (in place over the u tile) and shared through
// shared memory, the north-face flux is carried in a register while a thread marches
// up its strip of rows (the Fortran's fym).  The periodic halo fill
// that Operators.rhs_adv performs next (operators.py:231) is fused: a thread that
// owns a rim cell also stores its halo images.
//
// The running mask sums of the Fortran (mx5/mx3, my5/my3; :62-70,96-97,132-133) are
// window sums here -- the same integers.
#include <cstddef>
#include "f2d_common.cuh"

using namespace f2d;

namespace {

constexpr int NH = 3;
constexpr int TX = 64;    // outputs per tile in x
constexpr int TY = 32;    // outputs per tile in y
constexpr int NT = 256;   // threads per CTA
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

// 8-byte asynchronous global->shared copy (LDGSTS); !pred zero-fills the destination
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem, bool pred) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem);
  int sz = pred ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

struct AdvSmem {
  double q[SH][SW];       // tracer tile, halo 3
  double u[TY][TX + 1];   // u on the east faces i0-1 .. i0+TX-1; overwritten by the x fluxes
  double v[TY + 1][TX];   // v on the north faces of rows j0-1 .. j0+TY-1
  int8_t m[SH][SW];       // mask tile (MASKED only)
};

// mask-free instantiations: launched without the mask tile (54.8 KB) and compiled for 4 CTAs
// per SM (<= 64 registers)
template <bool UPW, int ORDER, bool MASKED>
__global__ void __launch_bounds__(NT, MASKED ? 3 : 4)
k_adv(const int8_t *__restrict__ msk, const double *__restrict__ q, double *__restrict__ dq,
      const double *__restrict__ u, const double *__restrict__ v, double *__restrict__ xflx,
      double *__restrict__ yflx, AdvC k, int ny, int nx, int fill) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  AdvSmem &S = *reinterpret_cast<AdvSmem *>(smem_raw);
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int i0 = NH + blockIdx.x * TX;  // first output column of the tile
  const int j0 = NH + blockIdx.y * TY;
  // ---- stage q (halo 3), u, v with asynchronous copies: one tile row per warp and pass
  for (int r = warp; r < SH; r += NT / 32) {
    int j = j0 - NH + r;
    const double *row = q + (size_t)j * nx + (i0 - NH);
#pragma unroll
    for (int c = lane; c < SW; c += 32) {
      bool in = j < ny && (i0 - NH + c) < nx;
      cp_async8(&S.q[r][c], in ? row + c : q, in);
      if (MASKED) S.m[r][c] = in ? msk[(size_t)j * nx + i0 - NH + c] : (int8_t)0;
    }
  }
  for (int r = warp; r < TY; r += NT / 32) {
    int j = j0 + r;
    const double *row = u + (size_t)j * nx + (i0 - 1);
#pragma unroll
    for (int c = lane; c < TX + 1; c += 32) {
      bool in = j < ny && (i0 - 1 + c) < nx;
      cp_async8(&S.u[r][c], in ? row + c : u, in);
    }
  }
  for (int r = warp; r < TY + 1; r += NT / 32) {
    int j = j0 - 1 + r;
    const double *row = v + (size_t)j * nx + i0;
#pragma unroll
    for (int c = lane; c < TX; c += 32) {
      bool in = j < ny && (i0 + c) < nx;
      cp_async8(&S.v[r][c], in ? row + c : v, in);
    }
  }
  cp_async_wait_all();
  __syncthreads();
  const int tx = t & (TX - 1), tg = t >> 6;   // column, row group (TX == 64)
  const int r0 = tg * (TY / 4);
  const int i = i0 + tx;
  const bool col_ok = i < nx - NH;
  const bool flx = xflx != nullptr;
  // ---- east-face fluxes, in place over the u tile: thread (tx,tg) -> face of column i for
  // its rows; threads 0..TY-1 also take the west face of the tile (column i0-1) of row t
#pragma unroll 2
  for (int r = r0; r < r0 + TY / 4; r++) {
    int j = j0 + r;
    double f = 0.;
    if (col_ok && j < ny - NH)
      f = face_flux<UPW, ORDER, MASKED>(k, S.u[r][tx + 1], &S.q[r + NH][tx + NH], 1,
                                        MASKED ? &S.m[r + NH][tx + NH] : nullptr, 1);
    S.u[r][tx + 1] = f;
  }
  if (t < TY) {
    int j = j0 + t;
    double f = 0.;
    if (j < ny - NH)
      f = face_flux<UPW, ORDER, MASKED>(k, S.u[t][0], &S.q[t + NH][NH - 1], 1,
                                        MASKED ? &S.m[t + NH][NH - 1] : nullptr, 1);
    S.u[t][0] = f;
    if (flx && blockIdx.x == 0 && j < ny - NH) xflx[(size_t)j * nx + (i0 - 1)] = f;
  }
  __syncthreads();
  if (!col_ok) return;
  // ---- north-face fluxes marching up the strip (the Fortran's fym), divergence, stores
  const bool rim = fill && ((j0 < 2 * NH) || (i0 < 2 * NH) || (j0 + TY > ny - 2 * NH) || (i0 + TX > nx - 2 * NH));
  if (j0 + r0 >= ny - NH) return;
  double fym = face_flux<UPW, ORDER, MASKED>(k, S.v[r0][tx], &S.q[r0 + NH - 1][tx + NH], SW,
                                             MASKED ? &S.m[r0 + NH - 1][tx + NH] : nullptr, SW);
  if (flx && blockIdx.y == 0 && r0 == 0) yflx[(size_t)(j0 - 1) * nx + i] = fym;
#pragma unroll 2
  for (int r = r0; r < r0 + TY / 4; r++) {
    int j = j0 + r;
    if (j >= ny - NH) break;
    size_t c = (size_t)j * nx + i;
    double fy = face_flux<UPW, ORDER, MASKED>(k, S.v[r + 1][tx], &S.q[r + NH][tx + NH], SW,
                                              MASKED ? &S.m[r + NH][tx + NH] : nullptr, SW);
    double fxe = S.u[r][tx + 1], fxw = S.u[r][tx];
    double y = -k.zdx * (fxe - fxw) - k.zdy * (fy - fym);
    dq[c] = y;
    if (flx) {
      xflx[c] = fxe;
      yflx[c] = fy;
    }
    if (rim)
      for_each_halo_image(j, i, ny, nx, NH, [&](int jj, int ii) { dq[(size_t)jj * nx + ii] = y; }, fill != 2);
    fym = fy;
  }
}

template <bool UPW, int ORDER, bool MASKED>
int launch_adv_m(const int8_t *msk, const double *q, double *dq, const double *u, const double *v,
                 double *xflx, double *yflx, const AdvC &k, int ny, int nx, int fill, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    F2D_CUDA(cudaFuncSetAttribute(k_adv<UPW, ORDER, MASKED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)sizeof(AdvSmem)));
    attr_set = true;
  }
  dim3 grid(cdiv(nx - 2 * NH, TX), cdiv(ny - 2 * NH, TY));
  const size_t sm = MASKED ? sizeof(AdvSmem) : offsetof(AdvSmem, m);
  prof_tag("k_adv<upw%d,order%d,masked%d> %dx%d", (int)UPW, ORDER, (int)MASKED, nx - 2 * NH, ny - 2 * NH);
  k_adv<UPW, ORDER, MASKED><<<grid, NT, sm, s>>>(msk, q, dq, u, v, xflx, yflx, k, ny, nx, fill);
  F2D_LAUNCHED();
  return F2D_OK;
}

template <bool UPW, int ORDER>
int launch_adv(const int8_t *msk, const double *q, double *dq, const double *u, const double *v,
               double *xflx, double *yflx, const AdvC &k, int ny, int nx, int fill, bool masked,
               cudaStream_t s) {
  if (masked) return launch_adv_m<UPW, ORDER, true>(msk, q, dq, u, v, xflx, yflx, k, ny, nx, fill, s);
  return launch_adv_m<UPW, ORDER, false>(msk, q, dq, u, v, xflx, yflx, k, ny, nx, fill, s);
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
  // core/fortran_fluxes.f90's adv_centered (flux outputs) has no 6th-order branch: order = 6
  // falls through to its `order.ge.2` two-point mean
  if (xflx != nullptr && order == 6) order = 2;
  return adv_common(false, msk, q, dq, u, v, xflx, yflx, cst5, nh, 0, order, ny, nx, fill_halo, S(s));
}
