// Coarse-tail multigrid kernel (included by f2d_multigrid.cu).
//
// The levels whose arrays fit together in one SM's shared memory (local size <= 64, i.e.
// <= 70x70 with the halo) are about 1.5 % of the cells of a V-cycle but most of its
// operator applications (~130 per V-cycle, 56 halo fills); as
// separate launches they cost a few microseconds each.  Here ONE CTA runs a whole
// V-cycle (hierarchy.py:98-127) or F-cycle (hierarchy.py:131-151) of that sub-hierarchy
// with x, b and one scratch array per level resident in shared memory, __syncthreads()
// between operator applications, and the periodic halo fills done in shared memory.
//
// Arithmetic is the same expression, in the same order, as the Fortran kernels
// (through fused::jacobi_val / resid_val); the matrix comes either from the constant
// stencil class (all levels of the tail in class 1) or from the stored coefficients
// read through L1/L2.
#pragma once

namespace tail {

constexpr int NH = 3;
constexpr int NT = 1024;      // upper bound of the block size (the launch picks blockDim.x)
constexpr int MAXL = 8;       // levels 64,32,16,8,4 at most in practice
constexpr int MAXN = 64;      // largest interior size handled

struct Params {
  int nlev;                   // number of tail levels
  fused::LevelK lv[MAXL];     // geometry / matrix of each tail level (index 0 = finest of the tail)
  int off[MAXL];              // offset (in doubles) of the level inside each shared array
  int total;                  // sum of cells
  int ndeepest;
  const double *b_in;         // global rhs of the finest tail level
  const double *x_in;         // global first guess (V program, xmode 0) or nullptr
  double *x_out;              // global result of the finest tail level
  double *acc;                // if set: acc += result instead of storing it (solve(), hierarchy.py:171)
};

struct Ctx {
  double *X, *B, *T;          // shared arrays, all levels concatenated
  int t;
};

// iterate over the cells of a level: rows by warp, columns by lane
template <class F>
__device__ __forceinline__ void for_cells(int ny, int nx, int jlo, int jhi, int ilo, int ihi, F f) {
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nw = blockDim.x >> 5;
  for (int j = jlo + warp; j <= jhi; j += nw)
    for (int i = ilo + lane; i <= ihi; i += 32) f(j, i);
}

template <bool MASKED, bool STORED>
__device__ __forceinline__ double jacobi_at(const fused::LevelK &L, const fused::Coefs<MASKED, STORED> &kc,
                                            const double *__restrict__ s, const double *__restrict__ b, int j, int i) {
  int nx = L.nx;
  size_t g = (size_t)j * nx + i;
  if (MASKED && L.msk[g] == 0) return 0.;
  fused::Coefs<MASKED, STORED> k;
  if (MASKED || STORED) k.load(L, g, MASKED ? L.msk + g : nullptr, nx); else k = kc;
  const double *p = s + g;
  return fused::jacobi_val<MASKED, STORED>(L, k, p[-nx - 1], p[-nx], p[-nx + 1], p[-1], p[0], p[1], p[nx - 1], p[nx],
                                           p[nx + 1], b[g]);
}

// two damped-Jacobi sweeps + halo fill, x in place (scratch t)
template <bool MASKED, bool STORED>
__device__ void smooth2(const fused::LevelK &L, double *x, const double *b, double *t, bool xzero) {
  const int ny = L.ny, nx = L.nx;
  fused::Coefs<MASKED, STORED> kc;
  if (!MASKED && !STORED) kc.load(L, 0, nullptr, 0);
  (void)xzero;
  for_cells(ny, nx, 2, ny - 3, 2, nx - 3,
            [&](int j, int i) { t[j * nx + i] = jacobi_at<MASKED, STORED>(L, kc, x, b, j, i); });
  __syncthreads();
  for_cells(ny, nx, NH, ny - 1 - NH, NH, nx - 1 - NH, [&](int j, int i) {
    double val = jacobi_at<MASKED, STORED>(L, kc, t, b, j, i);
    x[j * nx + i] = val;
    f2d::for_each_halo_image(j, i, ny, nx, NH, [&](int jj, int ii) { x[jj * nx + ii] = val; });
  });
  __syncthreads();
}

// r = b - A x on the interior + halo fill
template <bool MASKED, bool STORED>
__device__ void residual(const fused::LevelK &L, const double *x, const double *b, double *r) {
  const int ny = L.ny, nx = L.nx;
  fused::Coefs<MASKED, STORED> kc;
  if (!MASKED && !STORED) kc.load(L, 0, nullptr, 0);
  for_cells(ny, nx, NH, ny - 1 - NH, NH, nx - 1 - NH, [&](int j, int i) {
    size_t g = (size_t)j * nx + i;
    double val = 0.;
    if (!MASKED || L.msk[g] != 0) {
      fused::Coefs<MASKED, STORED> k;
      if (MASKED || STORED) k.load(L, g, MASKED ? L.msk + g : nullptr, nx); else k = kc;
      double cdiag = STORED ? L.A[4 * (size_t)ny * nx + g] : L.c[4];
      const double *p = x + g;
      val = fused::resid_val<MASKED, STORED>(L, k, cdiag, p[-nx - 1], p[-nx], p[-nx + 1], p[-1], p[0], p[1],
                                             p[nx - 1], p[nx], p[nx + 1], b[g]);
    }
    r[g] = val;
    f2d::for_each_halo_image(j, i, ny, nx, NH, [&](int jj, int ii) { r[jj * nx + ii] = val; });
  });
  __syncthreads();
}

// full-weighting restriction fine -> coarse (coarse interior + halo fill)
template <bool MASKED>
__device__ void restrict_to(const fused::LevelK &Lc, const double *xf, int nxf, double *xc) {
  const int ny = Lc.ny, nx = Lc.nx;
  for_cells(ny, nx, NH, ny - 1 - NH, NH, nx - 1 - NH, [&](int j, int i) {
    int g = j * nx + i;
    double val = 0.;
    if (!MASKED || Lc.msk[g] != 0) {
      const double *f = xf + (2 * j - 2) * nxf + (2 * i - 2);
      val = 0.25 * f[0] + 0.125 * (((f[-1] + f[1]) + f[-nxf]) + f[nxf]) +
            0.0625 * (((f[-nxf - 1] + f[-nxf + 1]) + f[nxf - 1]) + f[nxf + 1]);
    }
    xc[g] = val;
    f2d::for_each_halo_image(j, i, ny, nx, NH, [&](int jj, int ii) { xc[jj * nx + ii] = val; });
  });
  __syncthreads();
}

// xf = [xf +] I(xc) over the whole fine array
template <bool MASKED>
__device__ void interpolate(const fused::LevelK &Lf, const fused::LevelK &Lc, const double *xc, double *xf, bool add) {
  const int ny = Lf.ny, nx = Lf.nx, nxc = Lc.nx;
  for_cells(ny, nx, 0, ny - 1, 0, nx - 1, [&](int j, int i) {
    int g = j * nx + i;
    double iv = 0.;
    if (!MASKED || Lf.msk[g] > 0) {
      int k = ((j >> 1) + 1) * nxc + (i >> 1) + 1;
      int pj = j & 1, pi = i & 1;
      const int8_t *mc = Lc.msk;
      if (!pj && !pi) {
        iv = xc[k];
      } else if (!pj) {
        int s = MASKED ? mc[k] + mc[k + 1] : 2;
        iv = (xc[k] + xc[k + 1]) * fused::interp_w2(s);
      } else if (!pi) {
        int s = MASKED ? mc[k] + mc[k + nxc] : 2;
        iv = (xc[k] + xc[k + nxc]) * fused::interp_w2(s);
      } else {
        int s = MASKED ? mc[k] + mc[k + 1] + mc[k + nxc] + mc[k + nxc + 1] : 4;
        iv = fused::interp_w4(s) * (((xc[k] + xc[k + 1]) + xc[k + nxc]) + xc[k + nxc + 1]);
      }
    }
    xf[g] = add ? xf[g] + iv : iv;
  });
  __syncthreads();
}

__device__ __forceinline__ void fill_zero(double *x, int n) {
  for (int p = threadIdx.x; p < n; p += blockDim.x) x[p] = 0.;
  __syncthreads();
}

// Coarsest level of an all-fluid doubly periodic hierarchy (constant-stencil class): every
// halo cell is a periodic image, so the m x n interior values are the only unknowns, and the
// value smoothtwicewithA computes on the ring around the interior is, bit for bit, the image
// of an interior value (same expression, same operands).  The ndeepest double sweeps
// (hierarchy.py:114-116, x = 0 first) therefore run in ONE warp on two m*n arrays with
// periodic indexing and __syncwarp() -- no block barrier, no halo fill -- and the result is
// expanded to the full array (interior + images) at the end.  u: scratch of >= 2*m*n doubles.
constexpr int MAXU = 64;   // unknowns handled (two per lane)
__device__ __forceinline__ bool coarsest_periodic_ok(const fused::LevelK &L) {
  return (L.ny - 2 * NH) * (L.nx - 2 * NH) <= MAXU && L.ny * L.nx >= 2 * (L.ny - 2 * NH) * (L.nx - 2 * NH);
}
__device__ inline void coarsest_periodic(const fused::LevelK &L, double *x, const double *b, double *u, int ndeepest) {
  const int ny = L.ny, nx = L.nx, m = ny - 2 * NH, n = nx - 2 * NH, U = m * n;
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    fused::Coefs<false, false> kc;
    kc.load(L, 0, nullptr, 0);
    double *src = u, *dst = u + U;
    int nb[2][8];
    double bq[2];
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const int p = lane + 32 * q;
      bq[q] = 0.;
      if (p < U) {
        const int pj = p / n, pi = p - pj * n;
        const int jm = (pj + m - 1) % m, jp = (pj + 1) % m, im = (pi + n - 1) % n, ip = (pi + 1) % n;
        nb[q][0] = jm * n + im; nb[q][1] = jm * n + pi; nb[q][2] = jm * n + ip;
        nb[q][3] = pj * n + im;                         nb[q][4] = pj * n + ip;
        nb[q][5] = jp * n + im; nb[q][6] = jp * n + pi; nb[q][7] = jp * n + ip;
        bq[q] = b[(pj + NH) * nx + pi + NH];
        src[p] = 0.;
      }
    }
    __syncwarp();
    for (int s = 0; s < 2 * ndeepest; s++) {
#pragma unroll
      for (int q = 0; q < 2; q++) {
        const int p = lane + 32 * q;
        if (p < U)
          dst[p] = fused::jacobi_val<false, false>(L, kc, src[nb[q][0]], src[nb[q][1]], src[nb[q][2]], src[nb[q][3]],
                                                   src[p], src[nb[q][4]], src[nb[q][5]], src[nb[q][6]], src[nb[q][7]],
                                                   bq[q]);
      }
      __syncwarp();
      double *tmp = src; src = dst; dst = tmp;
    }
    // 2*ndeepest sweeps: the result is back in u[0..U)
  }
  __syncthreads();
  for (int c = threadIdx.x; c < ny * nx; c += blockDim.x) {
    const int j = c / nx, i = c - j * nx;
    x[c] = u[((j - NH + 4 * m) % m) * n + (i - NH + 4 * n) % n];
  }
  __syncthreads();
}

// V-cycle of the tail levels [l1, nlev-1]; x of level l1 is whatever the shared array holds
template <bool MASKED, bool STORED>
__device__ void vcycle(const Params &P, double *X, double *B, double *T, int l1) {
  const int last = P.nlev - 1;
  for (int l = l1; l < last; l++) {
    const fused::LevelK &L = P.lv[l];
    double *x = X + P.off[l], *b = B + P.off[l], *t = T + P.off[l];
    if (l > l1) fill_zero(x, L.ny * L.nx);
    smooth2<MASKED, STORED>(L, x, b, t, false);
    residual<MASKED, STORED>(L, x, b, t);
    restrict_to<MASKED>(P.lv[l + 1], t, L.nx, B + P.off[l + 1]);
  }
  {
    const fused::LevelK &L = P.lv[last];
    double *x = X + P.off[last], *b = B + P.off[last], *t = T + P.off[last];
    if (!MASKED && !STORED && coarsest_periodic_ok(L)) {
      coarsest_periodic(L, x, b, t, P.ndeepest);
    } else {
      fill_zero(x, L.ny * L.nx);
      for (int k = 0; k < P.ndeepest; k++) smooth2<MASKED, STORED>(L, x, b, t, false);
    }
  }
  for (int l = last - 1; l >= l1; l--) {
    const fused::LevelK &L = P.lv[l];
    double *x = X + P.off[l], *b = B + P.off[l], *t = T + P.off[l];
    interpolate<MASKED>(L, P.lv[l + 1], X + P.off[l + 1], x, true);
    smooth2<MASKED, STORED>(L, x, b, t, false);
  }
}

// PROGRAM 0: V-cycle from the finest tail level, x = 0 initially
//         1: V-cycle, first guess read from x_in
//         2: F-cycle of the tail (restrict b down, coarsest solve, interpolate + V-cycle up)
template <bool MASKED, bool STORED>
__global__ void __launch_bounds__(NT, 1) k_mg_tail(const __grid_constant__ Params P, int program) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *X = reinterpret_cast<double *>(smem_raw);
  double *B = X + P.total;
  double *T = B + P.total;
  const int n0 = P.lv[0].ny * P.lv[0].nx;
  for (int p = threadIdx.x; p < n0; p += blockDim.x) {
    B[p] = P.b_in[p];
    X[p] = (program == 1) ? P.x_in[p] : 0.;
  }
  __syncthreads();
  if (program == 2) {
    const int last = P.nlev - 1;
    for (int l = 0; l < last; l++) restrict_to<MASKED>(P.lv[l + 1], B + P.off[l], P.lv[l].nx, B + P.off[l + 1]);
    {
      const fused::LevelK &L = P.lv[last];
      double *x = X + P.off[last];
      if (!MASKED && !STORED && coarsest_periodic_ok(L)) {
        coarsest_periodic(L, x, B + P.off[last], T + P.off[last], P.ndeepest);
      } else {
        fill_zero(x, L.ny * L.nx);
        for (int k = 0; k < P.ndeepest; k++) smooth2<MASKED, STORED>(L, x, B + P.off[last], T + P.off[last], false);
      }
    }
    for (int l = last - 1; l >= 0; l--) {
      interpolate<MASKED>(P.lv[l], P.lv[l + 1], X + P.off[l + 1], X + P.off[l], false);
      vcycle<MASKED, STORED>(P, X, B, T, l);
    }
  } else {
    vcycle<MASKED, STORED>(P, X, B, T, 0);
  }
  if (P.acc)
    for (int p = threadIdx.x; p < n0; p += blockDim.x) P.acc[p] = P.acc[p] + X[p];
  else
    for (int p = threadIdx.x; p < n0; p += blockDim.x) P.x_out[p] = X[p];
}

}  // namespace tail
