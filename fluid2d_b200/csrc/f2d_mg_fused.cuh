// Fused, shared-memory-tiled multigrid kernels (included by f2d_multigrid.cu).
//
//   k_smooth2        : Grid.smooth = two damped-Jacobi sweeps (smoothtwicewithA,
//                      fortran_multigrid.f90:2-127) + halo fill, one pass over HBM.
//                      Variants: the input may be zero (first visit of a coarse level in
//                      a V-cycle, hierarchy.py:101-102) or xin + I(xcoarse) (interpolate
//                      :415-498 and the `x += r` of hierarchy.py:123-126 fused in).
//   k_resid_restrict : computeresidualwithA (:320-362) + restrict (:501-546) + halo
//                      fills; the fine residual never goes to HBM.
//
// Coefficient classes (template parameters MASKED, STORED):
//   !MASKED,!STORED  the level is all fluid and its matrix is one constant 9-point
//                    stencil (doubly periodic domains): no mask, no matrix traffic;
//   MASKED,!STORED   matrix = constant stencil x mask products (finest level of any
//                    domain, level.py:288-298): 1 byte/cell of mask traffic;
//   MASKED,STORED    general stored coefficients (Galerkin levels next to walls).
// f2d_mg_create verifies on the device, entry by entry, that the class it selects
// reproduces the stored matrix exactly (k_check_const).
//
// Arithmetic: each value is computed by the same expression, in the same order, as the
// Fortran; only the traversal is different.
#pragma once

namespace fused {

constexpr int NH = 3;
constexpr int TX = 64;   // outputs per tile in x
constexpr int TY = 32;   // outputs per tile in y
constexpr int NT = 256;  // threads per CTA

struct LevelK {
  int ny, nx;
  const int8_t *msk;
  const double *A;   // 5 planes (STORED)
  double c[5];       // SW,S,SE,W,C (constant classes)
  double c1, c2, c3; // omega, 1-omega, omega/|C|
};

// coefficients of the 9-point operator at cell `g` (global index) / mask window
template <bool MASKED, bool STORED>
struct Coefs {
  double sw, s, se, w, e, nw, n, ne, c3;
  // m: pointer to the centre of a mask window with row stride ms (MASKED only)
  __device__ __forceinline__ void load(const LevelK &L, size_t g, const int8_t *m, int ms) {
    if (STORED) {
      size_t pl = (size_t)L.ny * L.nx;
      const double *A1 = L.A, *A2 = L.A + pl, *A3 = L.A + 2 * pl, *A4 = L.A + 3 * pl, *A5 = L.A + 4 * pl;
      int nx = L.nx;
      sw = A1[g]; s = A2[g]; se = A3[g]; w = A4[g];
      e = A4[g + 1]; nw = A3[g + nx - 1]; n = A2[g + nx]; ne = A1[g + nx + 1];
      c3 = L.c1 / fabs(A5[g]);
    } else if (MASKED) {
      sw = m[-ms - 1] ? L.c[0] : 0.; s = m[-ms] ? L.c[1] : 0.; se = m[-ms + 1] ? L.c[2] : 0.;
      w = m[-1] ? L.c[3] : 0.;       e = m[1] ? L.c[3] : 0.;
      nw = m[ms - 1] ? L.c[2] : 0.;  n = m[ms] ? L.c[1] : 0.;   ne = m[ms + 1] ? L.c[0] : 0.;
      c3 = L.c3;
    } else {
      sw = L.c[0]; s = L.c[1]; se = L.c[2]; w = L.c[3]; e = L.c[3]; nw = L.c[2]; n = L.c[1]; ne = L.c[0];
      c3 = L.c3;
    }
  }
};

// damped-Jacobi value from a 3x3 window (rows lo/mid/hi, columns l/c/r)
template <bool MASKED, bool STORED>
__device__ __forceinline__ double jacobi_val(const LevelK &L, const Coefs<MASKED, STORED> &k, double ll, double lc,
                                             double lr, double ml, double mc, double mr, double hl, double hc,
                                             double hr, double b) {
  double acc = k.sw * ll;
  acc = acc + k.s * lc;
  acc = acc + k.se * lr;
  acc = acc + k.w * ml;
  acc = acc + k.e * mr;
  acc = acc + k.nw * hl;
  acc = acc + k.n * hc;
  acc = acc + k.ne * hr;
  return mc * L.c2 + k.c3 * (acc - b);
}

// residual value b - A x from a 3x3 window
template <bool MASKED, bool STORED>
__device__ __forceinline__ double resid_val(const LevelK &L, const Coefs<MASKED, STORED> &k, double cdiag,
                                            double ll, double lc, double lr, double ml, double mc, double mr,
                                            double hl, double hc, double hr, double b) {
  double val = b - k.sw * ll;
  val = val - k.s * lc;
  val = val - k.se * lr;
  val = val - k.w * ml;
  val = val - cdiag * mc;
  val = val - k.e * mr;
  val = val - k.nw * hl;
  val = val - k.n * hc;
  val = val - k.ne * hr;
  return val;
}

__device__ __forceinline__ double interp_w2(int s) { return s == 2 ? 0.5 : (s == 1 ? 1. : 0.); }
__device__ __forceinline__ double interp_w4(int s) {
  const double third = (double)0.3333333333333333333333333333f;
  return s == 4 ? 0.25 : (s == 3 ? third : (s == 2 ? 0.5 : (s == 1 ? 1. : 0.)));
}

// ---------------------------------------------------------------------------
// k_smooth2
// ---------------------------------------------------------------------------
constexpr int XW = TX + 4, XH = TY + 4;  // x tile (halo 2)
constexpr int YW = TX + 2, YH = TY + 2;  // sweep-1 tile (halo 1)
constexpr int CW = XW / 2 + 2, CH = XH / 2 + 2;  // coarse tile for the fused interpolation

struct Smooth2Smem {
  double xs[XH][XW];
  double y1[YH][YW];
  double bs[YH][YW];
  double cs[CH][CW];
  int8_t ms[XH][XW];
  int8_t cm[CH][CW];
};

// 8-byte asynchronous global->shared copy (LDGSTS); !pred zero-fills the destination
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem, bool pred) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem);
  int sz = pred ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// INPUT: 0 xin, 1 zero, 2 I(xc), 3 xin + I(xc)
template <bool MASKED, bool STORED, int INPUT>
__global__ void __launch_bounds__(NT)
k_smooth2(LevelK L, const double *__restrict__ xin, const double *__restrict__ b, double *__restrict__ xout,
          const double *__restrict__ xc, const int8_t *__restrict__ mskc, int nxc, int nyc, double *acc) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smooth2Smem &S = *reinterpret_cast<Smooth2Smem *>(smem_raw);
  const int ny = L.ny, nx = L.nx;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int i0 = NH + blockIdx.x * TX, j0 = NH + blockIdx.y * TY;
  constexpr bool INTERP = INPUT >= 2;
  constexpr bool HAVE_X = (INPUT == 0 || INPUT == 3);
  const int cj0 = ((j0 - 2) >> 1) + 1, ci0 = ((i0 - 2) >> 1) + 1;
  // ---- stage everything with asynchronous copies: one tile row per warp and pass
  if (HAVE_X) {
    for (int r = warp; r < XH; r += NT / 32) {
      int j = j0 - 2 + r;
      const double *row = xin + (size_t)j * nx + (i0 - 2);
#pragma unroll
      for (int q = lane; q < XW; q += 32) {
        bool in = j < ny && (i0 - 2 + q) < nx;
        cp_async8(&S.xs[r][q], in ? row + q : xin, in);
      }
    }
  }
  for (int r = warp; r < YH; r += NT / 32) {
    int j = j0 - 1 + r;
    const double *row = b + (size_t)j * nx + (i0 - 1);
#pragma unroll
    for (int q = lane; q < YW; q += 32) {
      bool in = j < ny && (i0 - 1 + q) < nx;
      cp_async8(&S.bs[r][q], in ? row + q : b, in);
    }
  }
  if (INTERP) {
    for (int r = warp; r < CH; r += NT / 32) {
      int j = cj0 + r;
      const double *row = xc + (size_t)j * nxc + ci0;
      for (int q = lane; q < CW; q += 32) {
        bool in = j < nyc && (ci0 + q) < nxc;
        cp_async8(&S.cs[r][q], in ? row + q : xc, in);
        if (MASKED) S.cm[r][q] = in ? mskc[(size_t)j * nxc + ci0 + q] : (int8_t)0;
      }
    }
  }
  if (MASKED) {
    for (int r = warp; r < XH; r += NT / 32) {
      int j = j0 - 2 + r;
      for (int q = lane; q < XW; q += 32) {
        int i = i0 - 2 + q;
        S.ms[r][q] = (j < ny && i < nx) ? L.msk[(size_t)j * nx + i] : (int8_t)0;
      }
    }
  }
  cp_async_wait_all();
  __syncthreads();
  // ---- fused interpolation: xs = [xin +] I(xc)  (fortran_multigrid.f90:415-498)
  if (INTERP) {
    for (int r = warp; r < XH; r += NT / 32) {
      int j = j0 - 2 + r;
      int lj = (j >> 1) + 1 - cj0, pj = j & 1;
      for (int q = lane; q < XW; q += 32) {
        int i = i0 - 2 + q;
        double iv = 0.;
        if (j < ny && i < nx && (!MASKED || S.ms[r][q] > 0)) {
          int li = (i >> 1) + 1 - ci0, pi = i & 1;
          if (!pj && !pi) {
            iv = S.cs[lj][li];
          } else if (!pj) {
            int sm = MASKED ? S.cm[lj][li] + S.cm[lj][li + 1] : 2;
            iv = (S.cs[lj][li] + S.cs[lj][li + 1]) * interp_w2(sm);
          } else if (!pi) {
            int sm = MASKED ? S.cm[lj][li] + S.cm[lj + 1][li] : 2;
            iv = (S.cs[lj][li] + S.cs[lj + 1][li]) * interp_w2(sm);
          } else {
            int sm = MASKED ? S.cm[lj][li] + S.cm[lj][li + 1] + S.cm[lj + 1][li] + S.cm[lj + 1][li + 1] : 4;
            iv = interp_w4(sm) * (((S.cs[lj][li] + S.cs[lj][li + 1]) + S.cs[lj + 1][li]) + S.cs[lj + 1][li + 1]);
          }
        }
        S.xs[r][q] = (INPUT == 3) ? S.xs[r][q] + iv : iv;
      }
    }
    __syncthreads();
  }
  // ---- sweep 1 on the tile + ring 1, restricted to [2, n-3] (all that sweep 2 reads).
  // Thread (tx, tg) takes column tx of the y1 tile and a run of rows, carrying the 3x3
  // window in registers (3 shared-memory loads per point); the two extra columns of
  // the ring are done point-wise afterwards.
  const int tx = t & (TX - 1), tg = t >> 6;  // TX == 64
  Coefs<MASKED, STORED> kc;
  if (!MASKED && !STORED) kc.load(L, 0, nullptr, 0);
  auto sweep1_point = [&](int r, int q) {   // r,q index the y1 tile
    int j = j0 - 1 + r, i = i0 - 1 + q;
    double val = 0.;
    if (j >= 2 && j <= ny - 3 && i >= 2 && i <= nx - 3) {
      int xr = r + 1, xq = q + 1;  // same point in the x tile
      if (!MASKED || S.ms[xr][xq] != 0) {
        Coefs<MASKED, STORED> k;
        if (MASKED || STORED) k.load(L, (size_t)j * nx + i, MASKED ? &S.ms[xr][xq] : nullptr, XW); else k = kc;
        double x00 = 0., x01 = 0., x02 = 0., x10 = 0., x11 = 0., x12 = 0., x20 = 0., x21 = 0., x22 = 0.;
        if (INPUT != 1) {
          x00 = S.xs[xr - 1][xq - 1]; x01 = S.xs[xr - 1][xq]; x02 = S.xs[xr - 1][xq + 1];
          x10 = S.xs[xr][xq - 1];     x11 = S.xs[xr][xq];     x12 = S.xs[xr][xq + 1];
          x20 = S.xs[xr + 1][xq - 1]; x21 = S.xs[xr + 1][xq]; x22 = S.xs[xr + 1][xq + 1];
        }
        val = jacobi_val<MASKED, STORED>(L, k, x00, x01, x02, x10, x11, x12, x20, x21, x22, S.bs[r][q]);
      }
    }
    S.y1[r][q] = val;
  };
  {
    const int r0 = tg * 8 + (tg < 2 ? tg : 2), nr = tg < 2 ? 9 : 8;
    const int q = tx, xq = q + 1;
    const int i = i0 - 1 + q;
    const bool colok = (i >= 2 && i <= nx - 3);
    double a0 = 0., a1 = 0., a2 = 0., m0 = 0., m1 = 0., m2 = 0.;
    if (INPUT != 1) {
      a0 = S.xs[r0][xq - 1]; a1 = S.xs[r0][xq]; a2 = S.xs[r0][xq + 1];
      m0 = S.xs[r0 + 1][xq - 1]; m1 = S.xs[r0 + 1][xq]; m2 = S.xs[r0 + 1][xq + 1];
    }
#pragma unroll 3
    for (int r = r0; r < r0 + nr; r++) {
      double h0 = 0., h1 = 0., h2 = 0.;
      if (INPUT != 1) { h0 = S.xs[r + 2][xq - 1]; h1 = S.xs[r + 2][xq]; h2 = S.xs[r + 2][xq + 1]; }
      int j = j0 - 1 + r;
      double val = 0.;
      if (colok && j >= 2 && j <= ny - 3 && (!MASKED || S.ms[r + 1][xq] != 0)) {
        Coefs<MASKED, STORED> k;
        if (MASKED || STORED) k.load(L, (size_t)j * nx + i, MASKED ? &S.ms[r + 1][xq] : nullptr, XW); else k = kc;
        val = jacobi_val<MASKED, STORED>(L, k, a0, a1, a2, m0, m1, m2, h0, h1, h2, S.bs[r][q]);
      }
      S.y1[r][q] = val;
      a0 = m0; a1 = m1; a2 = m2;
      m0 = h0; m1 = h1; m2 = h2;
    }
    if (t < YH * 2) sweep1_point(t >> 1, TX + (t & 1));
  }
  __syncthreads();
  // ---- sweep 2 on the tile interior; tiles that touch the rim also store halo images
  {
    const bool rim = (j0 < 2 * NH) || (i0 < 2 * NH) || (j0 + TY > ny - 2 * NH) || (i0 + TX > nx - 2 * NH);
    const int r0 = tg * 8;
    const int i = i0 + tx;
    const int yq = tx + 1;
    if (i <= nx - 1 - NH) {
      double a0 = S.y1[r0][yq - 1], a1 = S.y1[r0][yq], a2 = S.y1[r0][yq + 1];
      double m0 = S.y1[r0 + 1][yq - 1], m1 = S.y1[r0 + 1][yq], m2 = S.y1[r0 + 1][yq + 1];
#pragma unroll 4
      for (int r = r0; r < r0 + 8; r++) {
        int j = j0 + r;
        if (j > ny - 1 - NH) break;
        double h0 = S.y1[r + 2][yq - 1], h1 = S.y1[r + 2][yq], h2 = S.y1[r + 2][yq + 1];
        double val = 0.;
        if (!MASKED || S.ms[r + 2][tx + 2] != 0) {
          Coefs<MASKED, STORED> k;
          if (MASKED || STORED) k.load(L, (size_t)j * nx + i, MASKED ? &S.ms[r + 2][tx + 2] : nullptr, XW); else k = kc;
          val = jacobi_val<MASKED, STORED>(L, k, a0, a1, a2, m0, m1, m2, h0, h1, h2, S.bs[r + 1][yq]);
        }
        if (acc) {
          // solve(): `x += self.x[0]` (hierarchy.py:171) fused into the last kernel of the
          // F-cycle: the correction is added to psi instead of being stored
          acc[(size_t)j * nx + i] = acc[(size_t)j * nx + i] + val;
          if (rim)
            f2d::for_each_halo_image(j, i, ny, nx, NH, [&](int jj, int ii) {
              acc[(size_t)jj * nx + ii] = acc[(size_t)jj * nx + ii] + val;
            });
        } else {
          xout[(size_t)j * nx + i] = val;
          if (rim)
            f2d::for_each_halo_image(j, i, ny, nx, NH, [&](int jj, int ii) { xout[(size_t)jj * nx + ii] = val; });
        }
        a0 = m0; a1 = m1; a2 = m2;
        m0 = h0; m1 = h1; m2 = h2;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// k_resid_restrict: coarse tile RTX x RTY, fine residual tile (2RTX+1) x (2RTY+1)
// ---------------------------------------------------------------------------
constexpr int RTX = 32, RTY = 16;
constexpr int RW = 2 * RTX + 1, RH = 2 * RTY + 1;  // residual tile
constexpr int RXW = RW + 2, RXH = RH + 2;          // x tile

// residual at an arbitrary fine cell straight from global memory (ring cells whose
// periodic source lies far from the tile)
template <bool MASKED, bool STORED>
__device__ double resid_global(const LevelK &L, const double *__restrict__ x, const double *__restrict__ b, int j,
                               int i) {
  int nx = L.nx;
  size_t g = (size_t)j * nx + i;
  if (MASKED && L.msk[g] == 0) return 0.;
  Coefs<MASKED, STORED> k;
  k.load(L, g, MASKED ? L.msk + g : nullptr, nx);
  double cdiag = STORED ? L.A[4 * (size_t)L.ny * nx + g] : L.c[4];
  return resid_val<MASKED, STORED>(L, k, cdiag, x[g - nx - 1], x[g - nx], x[g - nx + 1], x[g - 1], x[g], x[g + 1],
                                   x[g + nx - 1], x[g + nx], x[g + nx + 1], b[g]);
}

struct ResidSmem {
  double xs[RXH][RXW];
  double rs[RH][RW];
  double bs[RH][RW];
  int8_t ms[RXH][RXW];
};

template <bool MASKED, bool STORED>
__global__ void __launch_bounds__(NT)
k_resid_restrict(LevelK L, const double *__restrict__ x, const double *__restrict__ b, double *__restrict__ bc,
                 const int8_t *__restrict__ mskc, int nyc, int nxc) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ResidSmem &S = *reinterpret_cast<ResidSmem *>(smem_raw);
  const int ny = L.ny, nx = L.nx;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int ci0 = NH + blockIdx.x * RTX, cj0 = NH + blockIdx.y * RTY;  // first coarse output
  const int fi0 = 2 * ci0 - 3, fj0 = 2 * cj0 - 3;                      // first fine residual point
  for (int r = warp; r < RXH; r += NT / 32) {
    int j = fj0 - 1 + r;
    const double *row = x + (size_t)j * nx + (fi0 - 1);
#pragma unroll
    for (int q = lane; q < RXW; q += 32) {
      bool in = j < ny && (fi0 - 1 + q) < nx;
      cp_async8(&S.xs[r][q], in ? row + q : x, in);
      if (MASKED) S.ms[r][q] = in ? L.msk[(size_t)j * nx + fi0 - 1 + q] : (int8_t)0;
    }
  }
  for (int r = warp; r < RH; r += NT / 32) {
    int j = fj0 + r;
    const double *row = b + (size_t)j * nx + fi0;
#pragma unroll
    for (int q = lane; q < RW; q += 32) {
      bool in = j < ny && (fi0 + q) < nx;
      cp_async8(&S.bs[r][q], in ? row + q : b, in);
    }
  }
  cp_async_wait_all();
  __syncthreads();
  // ---- fine residual on the (2RTX+1) x (2RTY+1) tile: column strips with the 3x3 window
  // in registers; column 2RTX point-wise
  Coefs<MASKED, STORED> kc;
  if (!MASKED && !STORED) kc.load(L, 0, nullptr, 0);
  auto resid_point = [&](int r, int q, double x00, double x01, double x02, double x10, double x11, double x12,
                         double x20, double x21, double x22) -> double {
    int j = fj0 + r, i = fi0 + q;
    double val = 0.;
    if (j <= ny - NH && i <= nx - NH) {
      if (j == ny - NH || i == nx - NH) {
        // first halo ring on the high side: the reference reads the halo-filled residual
        // there, i.e. the residual of the periodic source cell
        val = resid_global<MASKED, STORED>(L, x, b, f2d::wrap_src(j, ny, NH), f2d::wrap_src(i, nx, NH));
      } else if (!MASKED || S.ms[r + 1][q + 1] != 0) {
        size_t g = (size_t)j * nx + i;
        Coefs<MASKED, STORED> k;
        if (MASKED || STORED) k.load(L, g, MASKED ? &S.ms[r + 1][q + 1] : nullptr, RXW); else k = kc;
        double cdiag = STORED ? L.A[4 * (size_t)ny * nx + g] : L.c[4];
        val = resid_val<MASKED, STORED>(L, k, cdiag, x00, x01, x02, x10, x11, x12, x20, x21, x22, S.bs[r][q]);
      }
    }
    return val;
  };
  {
    const int tx = t & 63, tg = t >> 6;
    const int r0 = tg * 8 + (tg < 1 ? 0 : 1), nr = tg < 1 ? 9 : 8;   // 9,8,8,8 = 33 rows
    const int xq = tx + 1;
    double a0 = S.xs[r0][xq - 1], a1 = S.xs[r0][xq], a2 = S.xs[r0][xq + 1];
    double m0 = S.xs[r0 + 1][xq - 1], m1 = S.xs[r0 + 1][xq], m2 = S.xs[r0 + 1][xq + 1];
#pragma unroll 3
    for (int r = r0; r < r0 + nr; r++) {
      double h0 = S.xs[r + 2][xq - 1], h1 = S.xs[r + 2][xq], h2 = S.xs[r + 2][xq + 1];
      S.rs[r][tx] = resid_point(r, tx, a0, a1, a2, m0, m1, m2, h0, h1, h2);
      a0 = m0; a1 = m1; a2 = m2;
      m0 = h0; m1 = h1; m2 = h2;
    }
    if (t < RH) {
      int r = t, q = RW - 1, xr = r + 1, xq2 = q + 1;
      S.rs[r][q] = resid_point(r, q, S.xs[xr - 1][xq2 - 1], S.xs[xr - 1][xq2], S.xs[xr - 1][xq2 + 1],
                               S.xs[xr][xq2 - 1], S.xs[xr][xq2], S.xs[xr][xq2 + 1], S.xs[xr + 1][xq2 - 1],
                               S.xs[xr + 1][xq2], S.xs[xr + 1][xq2 + 1]);
    }
  }
  __syncthreads();
  // ---- full-weighting restriction of the residual tile, halo images stored on rim tiles
  const bool rim = (cj0 < 2 * NH) || (ci0 < 2 * NH) || (cj0 + RTY > nyc - 2 * NH) || (ci0 + RTX > nxc - 2 * NH);
#pragma unroll
  for (int p = t; p < RTY * RTX; p += NT) {
    int r = p / RTX, q = p % RTX;
    int j = cj0 + r, i = ci0 + q;
    if (j > nyc - 1 - NH || i > nxc - 1 - NH) continue;
    size_t g = (size_t)j * nxc + i;
    double val = 0.;
    if (!MASKED || mskc[g] != 0) {
      int fr = 2 * r + 1, fq = 2 * q + 1;  // centre in the residual tile
      val = 0.25 * S.rs[fr][fq] +
            0.125 * (((S.rs[fr][fq - 1] + S.rs[fr][fq + 1]) + S.rs[fr - 1][fq]) + S.rs[fr + 1][fq]) +
            0.0625 * (((S.rs[fr - 1][fq - 1] + S.rs[fr - 1][fq + 1]) + S.rs[fr + 1][fq - 1]) + S.rs[fr + 1][fq + 1]);
    }
    bc[g] = val;
    if (rim)
      f2d::for_each_halo_image(j, i, nyc, nxc, NH, [&](int jj, int ii) { bc[(size_t)jj * nxc + ii] = val; });
  }
}

// ---------------------------------------------------------------------------
// k_check_const: does "constant stencil x mask products" reproduce the stored matrix
// on every entry the kernels read (cells [2, n-3] compute; they read coefficients and
// masks of [1, n-2])?  flag[0] is cleared on the first mismatch; flag[1] is cleared when
// some cell of [1, n-2] is solid.
// ---------------------------------------------------------------------------
__global__ void k_check_const(LevelK L, int *flag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int j = blockIdx.y * blockDim.y + threadIdx.y;
  int ny = L.ny, nx = L.nx;
  // masks are consulted on [1, n-2] (gates on [2, n-3] + their neighbours); the outer
  // ring never enters a value that survives the halo fill
  if (j < 1 || j > ny - 2 || i < 1 || i > nx - 2) return;
  size_t g = (size_t)j * nx + i;
  if (L.msk[g] == 0) {
    flag[1] = 0;
    return;
  }
  if (j < 2 || j > ny - 3 || i < 2 || i > nx - 3) return;
  Coefs<true, true> st;
  st.load(L, g, nullptr, 0);
  Coefs<true, false> cm;
  cm.load(L, g, L.msk + g, nx);
  double diag = L.A[4 * (size_t)ny * nx + g];
  bool ok = st.sw == cm.sw && st.s == cm.s && st.se == cm.se && st.w == cm.w && st.e == cm.e && st.nw == cm.nw &&
            st.n == cm.n && st.ne == cm.ne && diag == L.c[4];
  if (!ok) flag[0] = 0;
}

// first cell (lowest index) whose 3x3 neighbourhood is all fluid -> idx[0]
__global__ void k_find_interior(const int8_t *__restrict__ msk, int ny, int nx, unsigned long long *idx) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (j < NH || j > ny - 1 - NH || i < NH || i > nx - 1 - NH) return;
  size_t g = (size_t)j * nx + i;
  for (int dj = -1; dj <= 1; dj++)
    for (int di = -1; di <= 1; di++)
      if (msk[g + (ptrdiff_t)dj * nx + di] == 0) return;
  atomicMin(idx, (unsigned long long)g);
}

}  // namespace fused
