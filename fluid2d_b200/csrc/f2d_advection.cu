// Flux-form advection: core/fortran_advection.f90 (adv_upwind :2-165, adv_centered
// :169-284) and the flux-storing variants of core/fortran_fluxes.f90.
//
// One CTA computes a TX x TY tile of dq = -div(U q).  The q tile (halo 3) and the u, v
// tiles arrive in shared memory as three TMA boxes (one instruction each; per-element
// asynchronous copies, LDGSTS, on arrays too small for a box or with F2D_ADV_TMA=0).  All
// tracers of a model go through ONE launch (f2d_adv_multi: Operators.rhs_adv's loop over the
// tracer list, operators.py:214-236), and the same kernel can write the Runge-Kutta stage
// state x + coef*dq (timescheme.py:172-176) while the tendency is in registers.  Every east-face
// flux of the tile is computed once (in place over the u tile) and shared through
// shared memory, the north-face flux is carried in a register while a thread marches
// up its strip of rows (the Fortran's fym).  The periodic halo fill
// that Operators.rhs_adv performs next (operators.py:231) is fused: a thread that
// owns a rim cell also stores its halo images.
//
// The running mask sums of the Fortran (mx5/mx3, my5/my3; :62-70,96-97,132-133) are
// window sums here -- the same integers.
#include <cstddef>
#include <map>
#include <tuple>
#include "f2d_common.cuh"
#include "f2d_tma.cuh"

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

// Tiles in shared memory.  TMA: dense boxes (128-byte aligned, even widths; the v tile starts
// one column early because a box must start at an even column: the tile proper sits at column
// offset 1).  Otherwise (odd nx) the rows are staged by LDGSTS with the same layout.
constexpr int UW = TX + 2;     // u tile width: faces i0-1 .. i0+TX (the last one unused)
constexpr int VW = TX + 2;     // v tile width: columns i0-1 .. i0+TX
struct AdvSmem {
  alignas(128) double q[SH][SW];     // tracer tile, halo 3
  alignas(128) double u[TY][UW];     // u on the east faces i0-1 .. ; overwritten by the x fluxes
  alignas(128) double v[TY + 1][VW]; // v on the north faces of rows j0-1 .. j0+TY-1, column i0+c at [.][c+1]
  alignas(8) uint64_t bar;
  int8_t m[SH][SW];                  // mask tile (MASKED only)
};

// One launch advects up to ADV_MAXT tracers of a model (operators.py:214-236 is a loop over the
// tracer list): the CTAs of the tracers of one tile are neighbours in the grid (blockIdx.x =
// tile_x * n + tracer), so the velocity tiles they all read come from L2 after the first one.
// Optionally the kernel also writes the Runge-Kutta stage state xo = xb + coef * dq
// (timescheme.py:172-176; product rounded before the sum, as numpy does).
constexpr int ADV_MAXT = 4;
struct AdvBatch {
  const double *q[ADV_MAXT];
  double *dq[ADV_MAXT];
  const double *xb[ADV_MAXT];
  double *xo[ADV_MAXT];
  double coef;
  int n;
};
struct AdvMaps {
  CUtensorMap q[ADV_MAXT], u, v;
};

// mask-free instantiations: launched without the mask tile and compiled for 4 CTAs per SM
// (<= 64 registers)
template <bool UPW, int ORDER, bool MASKED, bool TMA>
__global__ void __launch_bounds__(NT, MASKED ? 3 : 4)
k_adv(const int8_t *__restrict__ msk, const AdvBatch B, const double *__restrict__ u, const double *__restrict__ v,
      double *__restrict__ xflx, double *__restrict__ yflx, AdvC k, int ny, int nx, int fill,
      const __grid_constant__ AdvMaps M) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  AdvSmem &S = *reinterpret_cast<AdvSmem *>(smem_raw);
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int tr = B.n > 1 ? (int)(blockIdx.x % (unsigned)B.n) : 0;
  const int bx = B.n > 1 ? (int)(blockIdx.x / (unsigned)B.n) : (int)blockIdx.x;
  const double *__restrict__ q = B.q[tr];
  double *__restrict__ dq = B.dq[tr];
  const double *__restrict__ xb = B.xb[tr];
  double *__restrict__ xo = B.xo[tr];
  const int i0 = NH + bx * TX;  // first output column of the tile
  const int j0 = NH + blockIdx.y * TY;
  if (TMA) {
    // ---- three boxes, one instruction each (zero fill outside the array)
    if (t == 0) f2d::mbar_init(&S.bar, 1);
    __syncthreads();
    if (t == 0) {
      f2d::mbar_expect_tx(&S.bar, (SH * SW + TY * UW + (TY + 1) * VW) * 8);
      f2d::tma_load_2d(&S.q[0][0], &M.q[tr], &S.bar, i0 - NH, j0 - NH);
      f2d::tma_load_2d(&S.u[0][0], &M.u, &S.bar, i0 - 1, j0);
      f2d::tma_load_2d(&S.v[0][0], &M.v, &S.bar, i0 - 1, j0 - 1);
    }
    if (MASKED) {
      // every byte load issued before the first store: one L2 round trip for the mask tile
      constexpr int NW = NT / 32, NK = (SH + NW - 1) / NW, NCC = (SW + 31) / 32;
      int8_t mv[NK][NCC];
#pragma unroll
      for (int kk = 0; kk < NK; kk++) {
        const int r = warp + kk * NW, j = j0 - NH + r;
#pragma unroll
        for (int cc = 0; cc < NCC; cc++) {
          const int c = lane + cc * 32;
          mv[kk][cc] = (r < SH && c < SW && j < ny && (i0 - NH + c) < nx) ? msk[(size_t)j * nx + i0 - NH + c] : (int8_t)0;
        }
      }
#pragma unroll
      for (int kk = 0; kk < NK; kk++) {
        const int r = warp + kk * NW;
#pragma unroll
        for (int cc = 0; cc < NCC; cc++) {
          const int c = lane + cc * 32;
          if (r < SH && c < SW) S.m[r][c] = mv[kk][cc];
        }
      }
    }
    f2d::mbar_wait(&S.bar, 0);
  } else {
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
        cp_async8(&S.v[r][c + 1], in ? row + c : v, in);
      }
    }
    cp_async_wait_all();
  }
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
    if (flx && bx == 0 && j < ny - NH) xflx[(size_t)j * nx + (i0 - 1)] = f;
  }
  __syncthreads();
  if (!col_ok) return;
  // ---- north-face fluxes marching up the strip (the Fortran's fym), divergence, stores
  const bool rim = fill && ((j0 < 2 * NH) || (i0 < 2 * NH) || (j0 + TY > ny - 2 * NH) || (i0 + TX > nx - 2 * NH));
  if (j0 + r0 >= ny - NH) return;
  // The 6-point window of the north face marches in registers: the face above row r reads the
  // rows r-2 .. r+3, five of which the face below already holds (one shared load per face
  // instead of six; the same for the mask window).
  const double *qc = &S.q[r0 + NH - 1][tx + NH];        // cell below the first face (row r0-1)
  double w0 = qc[-2 * SW], w1 = qc[-SW], w2 = qc[0], w3 = qc[SW], w4 = qc[2 * SW], w5 = qc[3 * SW];
  int n0 = 1, n1 = 1, n2 = 1, n3 = 1, n4 = 1, n5 = 1;
  const int8_t *mc = &S.m[r0 + NH - 1][tx + NH];
  if (MASKED) { n0 = mc[-2 * SW]; n1 = mc[-SW]; n2 = mc[0]; n3 = mc[SW]; n4 = mc[2 * SW]; n5 = mc[3 * SW]; }
  auto north_flux = [&](double vel) {
    return UPW ? upw_flux<ORDER, MASKED>(k, vel, w0, w1, w2, w3, w4, w5, n0, n1, n2, n3, n4, n5)
               : cen_flux<ORDER, MASKED>(k, vel, w0, w1, w2, w3, w4, w5, n0, n1, n2, n3, n4, n5);
  };
  double fym = north_flux(S.v[r0][tx + 1]);
  if (flx && blockIdx.y == 0 && r0 == 0) yflx[(size_t)(j0 - 1) * nx + i] = fym;
  const double coef = B.coef;
#pragma unroll 2
  for (int r = r0; r < r0 + TY / 4; r++) {
    int j = j0 + r;
    if (j >= ny - NH) break;
    size_t c = (size_t)j * nx + i;
    // the window moves one row north: its new top is row r + 3
    w0 = w1; w1 = w2; w2 = w3; w3 = w4; w4 = w5;
    w5 = S.q[r + NH + 3][tx + NH];
    if (MASKED) { n0 = n1; n1 = n2; n2 = n3; n3 = n4; n4 = n5; n5 = S.m[r + NH + 3][tx + NH]; }
    double fy = north_flux(S.v[r + 1][tx + 1]);
    double fxe = S.u[r][tx + 1], fxw = S.u[r][tx];
    double y = -k.zdx * (fxe - fxw) - k.zdy * (fy - fym);
    dq[c] = y;
    if (xo) xo[c] = add_rn(xb[c], mul_rn(coef, y));
    if (flx) {
      xflx[c] = fxe;
      yflx[c] = fy;
    }
    if (rim)
      for_each_halo_image(j, i, ny, nx, NH, [&](int jj, int ii) {
        dq[(size_t)jj * nx + ii] = y;
        // the halo of the stage state is the reference's whole-array sum evaluated THERE: xb's
        // halo need not hold the images of its interior (masked domains, sponge, user edits)
        if (xo) xo[(size_t)jj * nx + ii] = add_rn(xb[(size_t)jj * nx + ii], mul_rn(coef, y));
      }, fill != 2);
    fym = fy;
  }
}

// tensor maps of the advected fields, cached by (pointer, shape, box)
bool adv_tmap(const double *base, int ny, int nx, int boxh, int boxw, CUtensorMap *out) {
  static std::map<std::tuple<const void *, int, int, int, int>, CUtensorMap> cache;
  if (!base || nx < boxw || ny < boxh) return false;
  auto key = std::make_tuple((const void *)base, ny, nx, boxh, boxw);
  auto it = cache.find(key);
  if (it == cache.end()) {
    CUtensorMap tm;
    if (f2d::make_tmap_2d(&tm, base, ny, nx, boxh, boxw) != 0) return false;
    if (cache.size() > 256) cache.clear();
    it = cache.emplace(key, tm).first;
  }
  *out = it->second;
  return true;
}

template <bool UPW, int ORDER, bool MASKED>
int launch_adv_m(const int8_t *msk, const AdvBatch &B, const double *u, const double *v,
                 double *xflx, double *yflx, const AdvC &k, int ny, int nx, int fill, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    F2D_CUDA(cudaFuncSetAttribute(k_adv<UPW, ORDER, MASKED, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)sizeof(AdvSmem)));
    F2D_CUDA(cudaFuncSetAttribute(k_adv<UPW, ORDER, MASKED, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)sizeof(AdvSmem)));
    attr_set = true;
  }
  static AdvMaps M;
  static const char *notma = getenv("F2D_ADV_TMA");
  bool tma = !(notma && notma[0] == '0');
  for (int t = 0; tma && t < B.n; t++) tma = adv_tmap(B.q[t], ny, nx, SH, SW, &M.q[t]);
  if (tma) tma = adv_tmap(u, ny, nx, TY, UW, &M.u) && adv_tmap(v, ny, nx, TY + 1, VW, &M.v);
  dim3 grid(cdiv(nx - 2 * NH, TX) * B.n, cdiv(ny - 2 * NH, TY));
  const size_t sm = MASKED ? sizeof(AdvSmem) : offsetof(AdvSmem, m);
  prof_tag("k_adv<upw%d,order%d,masked%d> %dx%d x%d%s", (int)UPW, ORDER, (int)MASKED, nx - 2 * NH, ny - 2 * NH, B.n,
           B.xo[0] ? " +stage" : "");
  if (tma) k_adv<UPW, ORDER, MASKED, true><<<grid, NT, sm, s>>>(msk, B, u, v, xflx, yflx, k, ny, nx, fill, M);
  else k_adv<UPW, ORDER, MASKED, false><<<grid, NT, sm, s>>>(msk, B, u, v, xflx, yflx, k, ny, nx, fill, M);
  F2D_LAUNCHED();
  return F2D_OK;
}

template <bool UPW, int ORDER>
int launch_adv(const int8_t *msk, const AdvBatch &B, const double *u, const double *v,
               double *xflx, double *yflx, const AdvC &k, int ny, int nx, int fill, bool masked,
               cudaStream_t s) {
  if (masked) return launch_adv_m<UPW, ORDER, true>(msk, B, u, v, xflx, yflx, k, ny, nx, fill, s);
  return launch_adv_m<UPW, ORDER, false>(msk, B, u, v, xflx, yflx, k, ny, nx, fill, s);
}

int adv_common(bool upw, const int8_t *msk, const AdvBatch &B, const double *u, const double *v,
               double *xflx, double *yflx, const double *cst, int nh, int method, int order, int ny, int nx,
               int fill, cudaStream_t s) {
  if (nh != NH) return fail(F2D_ERR_NH, "NHALO = 3 is compulsory with UP5");
  if (B.n < 1 || B.n > ADV_MAXT || !u || !v || !cst) return fail(F2D_ERR_ARG, "adv: null pointer");
  for (int t = 0; t < B.n; t++) {
    if (!B.q[t] || !B.dq[t]) return fail(F2D_ERR_ARG, "adv: null pointer");
    if ((B.xo[t] == nullptr) != (B.xb[t] == nullptr) || (B.xo[t] == nullptr) != (B.xo[0] == nullptr))
      return fail(F2D_ERR_ARG, "adv: the stage output needs xbase and xout for every tracer of the batch");
  }
  if (B.n > 1 && xflx) return fail(F2D_ERR_ARG, "adv: flux outputs are for one tracer at a time");
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
      case 1: return launch_adv<true, 1>(msk, B, u, v, xflx, yflx, k, ny, nx, fill, masked, s);
      case 3: return launch_adv<true, 3>(msk, B, u, v, xflx, yflx, k, ny, nx, fill, masked, s);
      case 5: return launch_adv<true, 5>(msk, B, u, v, xflx, yflx, k, ny, nx, fill, masked, s);
    }
    return fail(F2D_ERR_ARG, "adv_upwind: order must be 1, 3 or 5");
  }
  switch (order) {
    case 2: return launch_adv<false, 2>(msk, B, u, v, xflx, yflx, k, ny, nx, fill, masked, s);
    case 4: return launch_adv<false, 4>(msk, B, u, v, xflx, yflx, k, ny, nx, fill, masked, s);
    case 6: return launch_adv<false, 6>(msk, B, u, v, xflx, yflx, k, ny, nx, fill, masked, s);
  }
  return fail(F2D_ERR_ARG, "adv_centered: order must be 2, 4 or 6");
}

}  // namespace

static AdvBatch one_tracer(const double *q, double *dq) {
  AdvBatch B = {};
  B.q[0] = q;
  B.dq[0] = dq;
  B.n = 1;
  return B;
}

// msk == NULL selects the all-fluid specialisation (no mask reads); callers pass NULL
// only when every cell of msk, halo included, is 1 (geometry 'perio', grid.py:82-84).
extern "C" int f2d_adv_upwind(const int8_t *msk, const double *q, double *dq, const double *u, const double *v,
                              double *xflx, double *yflx, const double *cst5, int nh, int method, int order,
                              int ny, int nx, int fill_halo, f2d_stream_t s) {
  return adv_common(true, msk, one_tracer(q, dq), u, v, xflx, yflx, cst5, nh, method, order, ny, nx, fill_halo, S(s));
}
extern "C" int f2d_adv_centered(const int8_t *msk, const double *q, double *dq, const double *u,
                                const double *v, double *xflx, double *yflx, const double *cst5, int nh,
                                int method, int order, int ny, int nx, int fill_halo, f2d_stream_t s) {
  (void)method;
  // core/fortran_fluxes.f90's adv_centered (flux outputs) has no 6th-order branch: order = 6
  // falls through to its `order.ge.2` two-point mean
  if (xflx != nullptr && order == 6) order = 2;
  return adv_common(false, msk, one_tracer(q, dq), u, v, xflx, yflx, cst5, nh, 0, order, ny, nx, fill_halo, S(s));
}
// Operators.rhs_adv (operators.py:214-236) in one launch: the tracers of the model share the
// velocity tiles; xbase / xout (both NULL, or one pointer per tracer): the kernel also writes the
// Runge-Kutta stage state xout[t] = xbase[t] + coef * dq[t] (timescheme.py:172-176) with its halo
extern "C" int f2d_adv_multi(const int8_t *msk, const double *const *q, double *const *dq, int ntracers,
                             const double *u, const double *v, const double *cst5, int nh, int upwind, int method,
                             int order, const double *const *xbase, double *const *xout, double coef, int ny,
                             int nx, int fill_halo, f2d_stream_t s) {
  if (!q || !dq || ntracers < 1) return fail(F2D_ERR_ARG, "adv_multi: null pointer");
  if ((xbase == nullptr) != (xout == nullptr)) return fail(F2D_ERR_ARG, "adv_multi: xbase and xout go together");
  for (int t0 = 0; t0 < ntracers; t0 += ADV_MAXT) {
    AdvBatch B = {};
    B.n = ntracers - t0 < ADV_MAXT ? ntracers - t0 : ADV_MAXT;
    B.coef = coef;
    for (int t = 0; t < B.n; t++) {
      B.q[t] = q[t0 + t];
      B.dq[t] = dq[t0 + t];
      B.xb[t] = xbase ? xbase[t0 + t] : nullptr;
      B.xo[t] = xout ? xout[t0 + t] : nullptr;
    }
    int rc = adv_common(upwind != 0, msk, B, u, v, nullptr, nullptr, cst5, nh, upwind ? method : 0, order, ny, nx,
                        fill_halo, S(s));
    if (rc != F2D_OK) return rc;
  }
  return F2D_OK;
}
