// Geometric multigrid: gmg/hierarchy.py (Gmg), gmg/level.py (Grid, Gridinfo) and the
// kernels of gmg/fortran_multigrid.f90, single rank.
//
// Device data per level: int8 corner mask, the 5 stored diagonals of the symmetric
// 9-point matrix as 5 planes A[k][ny][nx] (k = 0..4: SW,S,SE,W,C), fields x,b,r and a
// scratch t.  Every operator application is followed in the reference by a periodic
// halo fill (level.py:365,384,493); here the fill is fused into the producing kernel
// (threads owning rim cells also store the halo images).
#include <stdlib.h>
#include <algorithm>
#include <vector>
#include <map>
#include <tuple>

#include "f2d_common.cuh"
#include "f2d_tma.cuh"
#include "f2d_mg_fused.cuh"
#include "f2d_mg_tail.cuh"
#include "f2d_mg_ctail.cuh"
#include "f2d_mg_ptail.cuh"

using namespace f2d;

namespace {
constexpr int NH = 3;

struct Level {
  int ny = 0, nx = 0;       // with halos
  int8_t *msk = nullptr;
  double *A = nullptr;      // 5 planes
  double *x = nullptr, *b = nullptr, *r = nullptr, *t = nullptr;
  int mode = 0;             // 0 stored, 1 constant, 2 constant x mask products
  double cst[5] = {0, 0, 0, 0, 0};  // SW,S,SE,W,C of the constant classes
  int ywrap = 1;            // 0 on slab (distributed) levels
  bool arena = false;       // x,b,t come from the communicator's symmetric heap
  size_t n() const { return (size_t)ny * nx; }
};
}  // namespace

struct f2d_mg {
  std::vector<Level> L;
  double omega = 8. / 9.;
  int npre = 1, npost = 1, ndeepest = 16, nvcyc = 1;  // hierarchy.py:29-32
  int relax = 0;              // 0 damped Jacobi (smoothtwicewithA), 1 line relaxation (smoothtridiag)
  double *scratch = nullptr;  // reductions
  double *partials = nullptr; // per-block sums of k_resid_sumsq
  size_t npartials = 0;
  double *dscal = nullptr;    // device scalars [4]
  double *hscal = nullptr;    // pinned host mirror [4]
  cudaStream_t cap = nullptr; // capture stream for CUDA graphs
  bool graphs = true;
  f2d_comm *comm = nullptr;   // y-slab decomposition (nullptr: single GPU)
  int lg = 0;                 // number of distributed (slab) levels; L[0] is global level lg
  std::vector<Level> S;       // slab levels 0..lg (S[lg]: slab-shaped view of L[0])
  int tail0 = -1;             // first level of the shared-memory tail (-1: no tail kernel)
  bool tail_const = false;    // every tail level is in the constant-stencil class
  size_t tail_smem = 0;
  // one-shot Runge-Kutta stage update of the velocities for the next f2d_invert_vorticity
  struct UVStage { bool on = false; const double *ub, *vb, *ue, *ve; double *uo, *vo; double c; } uvs;
  int tail_nt = tail::NT;     // threads of the one-CTA tail kernel (F2D_TAIL_NT)
  bool ctail = false;         // the tail runs on a thread-block cluster (f2d_mg_ctail.cuh), from 256^2 / 128^2 down
  int ctail_nc = 1;           // CTAs of that cluster (1 when no level of the tail is distributed)
  ctail::Params ctp;          // its level table (pointers / program filled per launch)
  bool ptail = false;         // all-fluid doubly periodic square tail: the interior-only kernel (f2d_mg_ptail.cuh)
  int ptail_top = 0;          // log2 of its finest level
  long long *trace = nullptr; // f2d_mg_set_trace
  int trace_cap = 0;
  struct G { cudaGraphExec_t exec; long long kernels; };
  std::map<std::tuple<int, int, const void *, const void *, const void *>, G> cache;
  // Gmg.solve as ONE graph: norms, a device-side WHILE node around (F-cycle, residual,
  // convergence test), result read back once
  struct SolveState { double normb, res0, res; int nite, ndiv, diverged, pad; };
  struct SG { cudaGraphExec_t exec; long long pre, body; };
  std::map<std::tuple<const void *, const void *, double, int>, SG> solve_cache;
  bool tma = true;                 // stage tiles with TMA (F2D_MG_TMA=0: per-element cp.async everywhere)
  std::map<std::tuple<const void *, int, int, int, int>, CUtensorMap> tmaps;
  cudaStream_t cap2 = nullptr;     // capture stream of the WHILE body
  SolveState *dstate = nullptr, *hstate = nullptr;
};

namespace {

// ---------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------
#define IJ2()                                      \
  int i = blockIdx.x * blockDim.x + threadIdx.x;   \
  int j = blockIdx.y * blockDim.y + threadIdx.y;   \
  if (i >= nx || j >= ny) return;                  \
  size_t c = (size_t)j * nx + i

inline dim3 grid2d(int ny, int nx, dim3 b) { return dim3(cdiv(nx, b.x), cdiv(ny, b.y)); }

// sum over the 8 off-diagonal neighbours of A*s (fortran_multigrid.f90:68-77)
__device__ __forceinline__ double offdiag(const double *__restrict__ A, size_t pl, const double *__restrict__ s,
                                          size_t c, int nx) {
  const double *A1 = A, *A2 = A + pl, *A3 = A + 2 * pl, *A4 = A + 3 * pl;
  double acc = A1[c] * s[c - nx - 1];
  acc = acc + A2[c] * s[c - nx];
  acc = acc + A3[c] * s[c - nx + 1];
  acc = acc + A4[c] * s[c - 1];
  acc = acc + A4[c + 1] * s[c + 1];
  acc = acc + A3[c + nx - 1] * s[c + nx - 1];
  acc = acc + A2[c + nx] * s[c + nx];
  acc = acc + A1[c + nx + 1] * s[c + nx + 1];
  return acc;
}

// one damped-Jacobi sweep on [lo, n-1-lo]^2 (fortran_multigrid.f90:62-86 / :96-117);
// fill != 0: the range is the interior and halo images are stored too
__global__ void k_jacobi(const int8_t *__restrict__ msk, const double *__restrict__ A, const double *__restrict__ xin,
                         const double *__restrict__ b, double *__restrict__ xout, double c1, double c2, int ny,
                         int nx, int lo, int fill, int ywrap = 1) {
  IJ2();
  if (j < lo || j > ny - 1 - lo || i < lo || i > nx - 1 - lo) return;
  size_t pl = (size_t)ny * nx;
  double val = 0.;
  if (msk[c] != 0) {
    double c3 = c1 / fabs(A[4 * pl + c]);
    val = xin[c] * c2 + c3 * (offdiag(A, pl, xin, c, nx) - b[c]);
  }
  xout[c] = val;
  if (fill) for_each_halo_image(j, i, ny, nx, NH, [&](int jj, int ii) { xout[(size_t)jj * nx + ii] = val; }, ywrap != 0);
}

// residual on the interior + halo images (fortran_multigrid.f90:320-362 + fill)
__global__ void k_residual(const int8_t *__restrict__ msk, const double *__restrict__ A, const double *__restrict__ x,
                           const double *__restrict__ b, double *__restrict__ r, int ny, int nx, int ywrap = 1) {
  IJ2();
  if (j < NH || j > ny - 1 - NH || i < NH || i > nx - 1 - NH) return;
  size_t pl = (size_t)ny * nx;
  double val = 0.;
  if (msk[c] != 0) {
    const double *A1 = A, *A2 = A + pl, *A3 = A + 2 * pl, *A4 = A + 3 * pl, *A5 = A + 4 * pl;
    val = b[c] - A1[c] * x[c - nx - 1];
    val = val - A2[c] * x[c - nx];
    val = val - A3[c] * x[c - nx + 1];
    val = val - A4[c] * x[c - 1];
    val = val - A5[c] * x[c];
    val = val - A4[c + 1] * x[c + 1];
    val = val - A3[c + nx - 1] * x[c + nx - 1];
    val = val - A2[c + nx] * x[c + nx];
    val = val - A1[c + nx + 1] * x[c + nx + 1];
  }
  r[c] = val;
  for_each_halo_image(j, i, ny, nx, NH, [&](int jj, int ii) { r[(size_t)jj * nx + ii] = val; }, ywrap != 0);
}

// full-weighting restriction on the coarse interior + halo images
// (fortran_multigrid.f90:501-546 + fill); msk2 == nullptr means all ones
// PEER: y-slab levels -- the block rows holding the 3 bottom / top interior rows also fill
// the neighbouring ranks' halo rows (see f2d::Peer)
template <bool PEER>
__global__ void k_restrict(const int8_t *__restrict__ msk2, const double *__restrict__ x1, double *__restrict__ x2,
                           int ny, int nx /*coarse*/, int nx1, int ywrap, f2d::Peer P) {
  f2d::pdl_trigger();
  f2d::pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const size_t c = (size_t)j * nx + i;
  const int m2 = ny - 2 * NH;
  const int rown = m2 / (int)blockDim.y;   // block row of the top interior rows m2..m2+NH-1
  const bool bsouth = PEER && blockIdx.y == 0, bnorth = PEER && (int)blockIdx.y == rown;
  if (PEER && (bsouth || bnorth)) {
    // peer_wait / peer_done use threadIdx.x == 0 as the CTA's leader: 2-D blocks here
    if (threadIdx.x == 0 && threadIdx.y == 0) {
      const unsigned long long d = *((volatile unsigned long long *)&P.me->done);
      const int rn = (P.rank + 1) % P.nranks, rs = (P.rank + P.nranks - 1) % P.nranks;
      if (bsouth) while (f2d::ld_acquire_sys(&P.me->slot[rs]) < d) {}
      if (bnorth) while (f2d::ld_acquire_sys(&P.me->slot[rn]) < d) {}
    }
    __syncthreads();
  }
  const bool inside = !(j < NH || j > ny - 1 - NH || i < NH || i > nx - 1 - NH);
  if (!PEER && !inside) return;
  if (inside) {
  double val = 0.;
  if (!msk2 || msk2[c] != 0) {
    // coarse 0-based J2 <-> fine 0-based J1 = 2*J2 - 2
    size_t f = (size_t)(2 * j - 2) * nx1 + (2 * i - 2);
    val = 0.25 * x1[f] +
          0.125 * (((x1[f - 1] + x1[f + 1]) + x1[f - nx1]) + x1[f + nx1]) +
          0.0625 * (((x1[f - nx1 - 1] + x1[f - nx1 + 1]) + x1[f + nx1 - 1]) + x1[f + nx1 + 1]);
  }
  x2[c] = val;
  for_each_halo_image(j, i, ny, nx, NH, [&](int jj, int ii) { x2[(size_t)jj * nx + ii] = val; }, ywrap != 0);
  if (PEER) {
    if (j < 2 * NH) {
      double *q = f2d::peer_addr(x2, P.south_off);
      q[(size_t)(j + m2) * nx + i] = val;
      for_each_halo_image(j + m2, i, ny, nx, NH, [&](int jj, int ii) { q[(size_t)jj * nx + ii] = val; }, false);
    }
    if (j >= m2) {
      double *q = f2d::peer_addr(x2, P.north_off);
      q[(size_t)(j - m2) * nx + i] = val;
      for_each_halo_image(j - m2, i, ny, nx, NH, [&](int jj, int ii) { q[(size_t)jj * nx + ii] = val; }, false);
    }
  }
  }
  if (PEER && (bsouth || bnorth)) {
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
      __threadfence_system();
      const unsigned nbound = gridDim.x * (rown == 0 ? 1u : 2u);
      unsigned t = atomicAdd(&P.me->blocks_done, 1u);
      if (t == nbound - 1) {
        P.me->blocks_done = 0;
        __threadfence_system();
        const unsigned long long D = P.me->done + 1;
        *((volatile unsigned long long *)&P.me->done) = D;
        for (int r = 0; r < P.nranks; r++)
          if (r != P.rank) *((volatile unsigned long long *)&P.peers[r]->slot[P.rank]) = D;
      }
    }
  }
}

// mask-aware bilinear interpolation over the whole fine array
// (fortran_multigrid.f90:415-498); add: x1 += I(x2)
__global__ void k_interpolate(const int8_t *__restrict__ msk1, const int8_t *__restrict__ msk2,
                              const double *__restrict__ x2, double *__restrict__ x1, int ny, int nx /*fine*/,
                              int nx2, int add) {
  IJ2();
  const double third = (double)0.3333333333333333333333333333f;
  double val = 0.;
  if (msk1[c] > 0) {
    int j2 = (j >> 1) + 1, i2 = (i >> 1) + 1;
    size_t k = (size_t)j2 * nx2 + i2;
    int pj = j & 1, pi = i & 1;
    if (!pj && !pi) {
      val = x2[k];
    } else if (!pj) {
      int s = msk2[k] + msk2[k + 1];
      double w = s == 2 ? 0.5 : (s == 1 ? 1. : 0.);
      val = (x2[k] + x2[k + 1]) * w;
    } else if (!pi) {
      int s = msk2[k] + msk2[k + nx2];
      double w = s == 2 ? 0.5 : (s == 1 ? 1. : 0.);
      val = (x2[k] + x2[k + nx2]) * w;
    } else {
      int s = msk2[k] + msk2[k + 1] + msk2[k + nx2] + msk2[k + nx2 + 1];
      double w = s == 4 ? 0.25 : (s == 3 ? third : (s == 2 ? 0.5 : (s == 1 ? 1. : 0.)));
      val = w * (((x2[k] + x2[k + 1]) + x2[k + nx2]) + x2[k + nx2 + 1]);
    }
  }
  x1[c] = add ? x1[c] + val : val;
}

// ---- set-up kernels ---------------------------------------------------------
// Line relaxation (fortran_multigrid.f90:215-317 smoothtridiag + tridiag): for each column
// i = 2..n-1 whose first interior corner is fluid, x(4:m-4, i) := V^{-1} (b - H x), V the
// vertical part of the operator (diagonal A5, off-diagonal A2), H the six other neighbours.
// The columns are swept west to east IN PLACE, so column i reads the new values of column
// i-1: a Gauss-Seidel order that leaves no parallelism across columns if the result is to be
// the reference's.  ONE CTA: the right-hand side and the two diagonals of a column are
// evaluated by all threads into shared memory, one thread runs the Thomas recurrences
// (forward elimination with the pivots, back substitution), all threads store the column.
// Shared memory: 4 arrays of ny doubles (rhs/solution, d, ud, gam).
__global__ void __launch_bounds__(256)
k_smooth_tridiag(const int8_t *__restrict__ msk, const double *__restrict__ A, double *x,
                 const double *__restrict__ b, int ny, int nx) {
  extern __shared__ double tri_smem[];
  double *y = tri_smem, *d = y + ny, *ud = d + ny, *gam = ud + ny;
  const size_t pl = (size_t)ny * nx;
  const double *A1 = A, *A2 = A + pl, *A3 = A + 2 * pl, *A4 = A + 3 * pl, *A5 = A + 4 * pl;
  const int k0 = NH, k1 = ny - 5;          // Fortran rows 4 .. m-4
  for (int i = 1; i <= nx - 2; i++) {      // Fortran columns 2 .. n-1
    if (msk[(size_t)NH * nx + i] == 0) continue;   // uniform over the block
    for (int j = k0 + (int)threadIdx.x; j <= k1; j += (int)blockDim.x) {
      size_t c = (size_t)j * nx + i;
      double r = b[c] - A1[c] * x[c - nx - 1];
      r = r - A3[c] * x[c - nx + 1];
      r = r - A4[c] * x[c - 1];
      r = r - A4[c + 1] * x[c + 1];
      r = r - A3[c + nx - 1] * x[c + nx - 1];
      r = r - A1[c + nx + 1] * x[c + nx + 1];
      y[j] = r;
      d[j] = A5[c];
      ud[j] = A2[c + nx];
    }
    __syncthreads();
    if (threadIdx.x == 0 && k1 >= k0) {
      double bet = 1. / d[k0];
      double prev = y[k0] * bet;
      y[k0] = prev;
      for (int k = k0 + 1; k <= k1; k++) {
        double dd = ud[k - 1];
        double g = dd * bet;
        gam[k] = g;
        bet = 1. / (d[k] - dd * g);
        prev = (y[k] - dd * prev) * bet;
        y[k] = prev;
      }
      for (int k = k1 - 1; k >= k0; k--) {
        prev = y[k] - gam[k + 1] * prev;
        y[k] = prev;
      }
    }
    __syncthreads();
    for (int j = k0 + (int)threadIdx.x; j <= k1; j += (int)blockDim.x) x[(size_t)j * nx + i] = y[j];
    __syncthreads();   // the next column reads this one from global memory
  }
}

__global__ void k_mask_from_double(const double *__restrict__ a, int8_t *__restrict__ m, size_t n) {
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
    m[k] = (int8_t)a[k];
}
__global__ void k_mask_to_double(const int8_t *__restrict__ m, double *__restrict__ a, size_t n) {
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
    a[k] = (double)m[k];
}
__global__ void k_fill_const(double *__restrict__ a, double v, size_t n) {
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) a[k] = v;
}
__global__ void k_threshold_mask(const double *__restrict__ w, int8_t *__restrict__ m, size_t n) {
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
    m[k] = w[k] <= 0.5 ? 0 : 1;
}
struct Stencil9 { double v[9]; };
// level.py:288-298: A_k(J,I) = stencil_k*coef*msk(J+j,I+i)*msk(J,I) on 1..n-2, 0 on the outer ring
__global__ void k_finest_matrix(const int8_t *__restrict__ msk, double *__restrict__ A9, Stencil9 st, int ny, int nx) {
  IJ2();
  size_t pl = (size_t)ny * nx;
  bool inner = (j >= 1 && j <= ny - 2 && i >= 1 && i <= nx - 2);
  for (int k = 0; k < 9; k++) {
    double val = 0.;
    if (inner) {
      int di = (k % 3) - 1, dj = (k / 3) - 1;
      val = (st.v[k] * (double)msk[c + (ptrdiff_t)dj * nx + di]) * (double)msk[c];
    }
    A9[k * pl + c] = val;
  }
}
// coarsenmatrix: fortran_multigrid.f90:706-811 on the coarse range nh..m2-nh (1-based)
__global__ void k_coarsenmatrix(const double *__restrict__ Af, double *__restrict__ Ac, const int8_t *__restrict__ msk1,
                                const int8_t *__restrict__ msk2, int ny2, int nx2, int ny1, int nx1) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int j = blockIdx.y * blockDim.y + threadIdx.y;
  // 1-based j2 = j+1 in [nh, m2-nh]
  if (j + 1 < NH || j + 1 > ny2 - NH || i + 1 < NH || i + 1 > nx2 - NH) return;
  size_t c2 = (size_t)j * nx2 + i;
  size_t pl1 = (size_t)ny1 * nx1, pl2 = (size_t)ny2 * nx2;
  if (msk2[c2] != 1) {
    for (int l = 0; l < 9; l++) Ac[l * pl2 + c2] = 0.;
    return;
  }
  const double coefv[3] = {0.125, 0.25, 0.125};  // coef(ki,kj) = cw(ki)*cw(kj)*... see below
  // coef table 0.125,0.25,0.125 / 0.25,0.5,0.25 / 0.125,0.25,0.125
  auto coef = [&](int ki, int kj) -> double {
    double a = coefv[ki + 1];
    return kj == 0 ? 2. * a : a;
  };
  // fine centre, 0-based: 1-based i1 = 2*(i2-nh)+nh  ->  0-based = 2*(i+1-nh)+nh-1
  int i1 = 2 * (i + 1 - NH) + NH - 1;
  int j1 = 2 * (j + 1 - NH) + NH - 1;
  for (int l = 0; l < 9; l++) {
    int di2 = l % 3 - 1, dj2 = l / 3 - 1;
    double z5[5][5];
    for (int a = 0; a < 5; a++)
      for (int b2 = 0; b2 < 5; b2++) z5[a][b2] = 0.;
    for (int kj = -1; kj <= 1; kj++)
      for (int ki = -1; ki <= 1; ki++) {
        int ii = 2 * di2 + ki, jj = 2 * dj2 + kj;
        if (abs(ii) <= 2 && abs(jj) <= 2)
          if (msk1[(size_t)(j1 + jj) * nx1 + (i1 + ii)] == 1) z5[jj + 2][ii + 2] = 2. * coef(ki, kj);
      }
    double w = 0.;
    for (int jj = -1; jj <= 1; jj++)
      for (int ii = -1; ii <= 1; ii++) {
        double z3 = 0.;
        size_t f = (size_t)(j1 + jj) * nx1 + (i1 + ii);
        if (msk1[f] == 1)
          for (int kj = -1; kj <= 1; kj++)
            for (int ki = -1; ki <= 1; ki++) {
              int k = (ki + 1) + (kj + 1) * 3;
              z3 = z3 + Af[k * pl1 + f] * z5[jj + kj + 2][ii + ki + 2];
            }
        w = w + (0.5 * coef(ii, jj)) * z3;
      }
    Ac[l * pl2 + c2] = w;
  }
}
// hierarchy.py:80-84
__global__ void k_helmholtz(double *__restrict__ A5, double shift, size_t n) {
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
    if (A5[k] != 0.) A5[k] = A5[k] - shift;
}
// 6th matrix plane: omega / |centre coefficient|, the factor of the damped-Jacobi update
// (fortran_multigrid.f90:2-127 divides at every point of every sweep; the quotient is the same)
__global__ void k_inverse_diagonal(const double *__restrict__ A5, double *__restrict__ A6, double omega, size_t n) {
  size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (k < n) A6[k] = omega / fabs(A5[k]);
}
__global__ void k_add_inplace(double *__restrict__ y, const double *__restrict__ a, size_t n) {
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
    y[k] = y[k] + a[k];
}

// solve(): residual of the finest level + its squared norm in one pass
// (hierarchy.py:159-162,172-174: g.residual(x, b, self.b[0]); g.norm(self.b[0])).
// r = b - A x on the interior with its halo images;
// the per-block sums of r^2 go to `partial` and are folded by k_fold_partials
// (deterministic two-stage sum; r is 0 on solid corners, so the unmasked sum of r^2
// equals computenorm's masked one).
constexpr int RST = 256;

__device__ __forceinline__ double block_sum(double v) {
  __shared__ double sh[RST / 32];
  for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0)
    for (int k = 1; k < RST / 32; k++) v += sh[k];
  return v;
}

// Block (bx, by): columns NH + 256*bx .. of rows NH + RSR*by ..; a thread marches its column
// north with the 3x3 window of x in registers (3 loads per point, neighbouring lanes share the
// cache lines), stores r with its halo images and accumulates r^2 in row order.
constexpr int RSR = 16;
template <bool MASKED, bool STORED>
__global__ void __launch_bounds__(RST)
k_resid_sumsq(fused::LevelK L, const double *__restrict__ x, const double *__restrict__ b, double *__restrict__ r,
              double *__restrict__ partial) {
  const int ny = L.ny, nx = L.nx;
  f2d::pdl_trigger();
  f2d::pdl_wait();
  fused::Coefs<MASKED, STORED> kc;
  if (!MASKED && !STORED) kc.load(L, 0, nullptr, 0);
  const int i = NH + blockIdx.x * RST + threadIdx.x;
  const int j0 = NH + blockIdx.y * RSR;
  double acc = 0.;
  if (i <= nx - 1 - NH) {
    const bool rimcol = i < 2 * NH || i >= nx - 2 * NH;
    const double *p = x + (size_t)j0 * nx + i;
    double a0 = p[-nx - 1], a1 = p[-nx], a2 = p[-nx + 1];
    double m0 = p[-1], m1 = p[0], m2 = p[1];
#pragma unroll 4
    for (int k = 0; k < RSR; k++) {
      const int j = j0 + k;
      if (j > ny - 1 - NH) break;
      const size_t g = (size_t)j * nx + i;
      const double h0 = x[g + nx - 1], h1 = x[g + nx], h2 = x[g + nx + 1];
      double val = 0.;
      if (!MASKED || L.msk[g] != 0) {
        fused::Coefs<MASKED, STORED> kk;
        if (MASKED || STORED) kk.load(L, g, MASKED ? L.msk + g : nullptr, nx); else kk = kc;
        const double cdiag = STORED ? __ldg(L.A + 4 * (size_t)ny * nx + g) : L.c[4];
        val = fused::resid_val<MASKED, STORED>(L, kk, cdiag, a0, a1, a2, m0, m1, m2, h0, h1, h2, b[g]);
      }
      r[g] = val;
      if (rimcol || j < 2 * NH || j >= ny - 2 * NH)
        for_each_halo_image(j, i, ny, nx, NH, [&](int jj, int ii) { r[(size_t)jj * nx + ii] = val; }, L.ywrap != 0);
      acc += val * val;
      a0 = m0; a1 = m1; a2 = m2;
      m0 = h0; m1 = h1; m2 = h2;
    }
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) partial[blockIdx.y * gridDim.x + blockIdx.x] = acc;
}
// The same operator on 64 x 32 tiles whose x tile (halo 1) arrives by TMA: the nine values of a
// point come from shared memory (3 loads per point with the marching window), b and the output go
// straight between registers and HBM.  Used whenever a tensor map of x exists (even nx, level at
// least one box large); the strip kernel above is the fallback.
constexpr int QTX = 64, QTY = 32, QW = QTX + 2, QH = QTY + 2;
struct SumsqSmem {
  alignas(128) double xs[QH][QW];
  alignas(8) uint64_t bar;
};
template <bool MASKED, bool STORED>
__global__ void __launch_bounds__(RST, 4)
k_resid_sumsq_tma(fused::LevelK L, const double *__restrict__ b, double *__restrict__ r, double *__restrict__ partial,
                  const __grid_constant__ CUtensorMap tmx) {
  __shared__ SumsqSmem S;
  const int ny = L.ny, nx = L.nx, t = threadIdx.x;
  const int i0 = NH + blockIdx.x * QTX, j0 = NH + blockIdx.y * QTY;
  f2d::pdl_trigger();
  if (t == 0) {
    f2d::mbar_init(&S.bar, 1);
    f2d::tma_prefetch_desc(&tmx);
  }
  f2d::pdl_wait();
  __syncthreads();
  if (t == 0) {
    f2d::mbar_expect_tx(&S.bar, QH * QW * 8);
    f2d::tma_load_2d(&S.xs[0][0], &tmx, &S.bar, i0 - 1, j0 - 1);
  }
  const int tx = t & (QTX - 1), tg = t >> 6;
  const int i = i0 + tx, jf = j0 + tg * 8;
  const bool col_ok = i <= nx - 1 - NH;
  double bv[8];
#pragma unroll
  for (int k = 0; k < 8; k++) bv[k] = (col_ok && jf + k <= ny - 1 - NH) ? b[(size_t)(jf + k) * nx + i] : 0.;
  fused::Coefs<MASKED, STORED> kc;
  if (!MASKED && !STORED) kc.load(L, 0, nullptr, 0);
  f2d::mbar_wait(&S.bar, 0);
  double acc = 0.;
  if (col_ok) {
    const bool rimcol = i < 2 * NH || i >= nx - 2 * NH;
    const double *sp = &S.xs[tg * 8 + 1][tx + 1];       // centre of the first point
    double a0 = sp[-QW - 1], a1 = sp[-QW], a2 = sp[-QW + 1];
    double m0 = sp[-1], m1 = sp[0], m2 = sp[1];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int j = jf + k;
      if (j <= ny - 1 - NH) {
        const double h0 = sp[(k + 1) * QW - 1], h1 = sp[(k + 1) * QW], h2 = sp[(k + 1) * QW + 1];
        const size_t g = (size_t)j * nx + i;
        double val = 0.;
        if (!MASKED || L.msk[g] != 0) {
          fused::Coefs<MASKED, STORED> kk;
          if (MASKED || STORED) kk.load(L, g, MASKED ? L.msk + g : nullptr, nx); else kk = kc;
          const double cdiag = STORED ? __ldg(L.A + 4 * (size_t)ny * nx + g) : L.c[4];
          val = fused::resid_val<MASKED, STORED>(L, kk, cdiag, a0, a1, a2, m0, m1, m2, h0, h1, h2, bv[k]);
        }
        r[g] = val;
        if (rimcol || j < 2 * NH || j >= ny - 2 * NH)
          for_each_halo_image(j, i, ny, nx, NH, [&](int jj, int ii) { r[(size_t)jj * nx + ii] = val; }, L.ywrap != 0);
        acc += val * val;
        a0 = m0; a1 = m1; a2 = m2;
        m0 = h0; m1 = h1; m2 = h2;
      }
    }
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) partial[blockIdx.y * gridDim.x + blockIdx.x] = acc;
}
__global__ void __launch_bounds__(RST) k_fold_partials(const double *__restrict__ partial, int n, double *out) {
  f2d::pdl_trigger();
  f2d::pdl_wait();
  double acc = 0.;
  for (int k = threadIdx.x; k < n; k += RST) acc += partial[k];
  acc = block_sum(acc);
  if (threadIdx.x == 0) out[0] = acc;
}

inline int nblocks1d(size_t n) {
  long long b = (long long)((n + 255) / 256);
  return (int)(b > 148LL * 16 ? 148LL * 16 : (b < 1 ? 1 : b));
}

// ---------------------------------------------------------------------------
// per-level operators
// ---------------------------------------------------------------------------
fused::LevelK level_k(f2d_mg *mg, const Level &l) {
  fused::LevelK k;
  k.ywrap = l.ywrap;
  k.ny = l.ny; k.nx = l.nx; k.msk = l.msk; k.A = l.A;
  for (int q = 0; q < 5; q++) k.c[q] = l.cst[q];
  k.c1 = mg->omega;
  k.c2 = 1. - mg->omega;
  k.c3 = l.cst[4] != 0. ? mg->omega / fabs(l.cst[4]) : 0.;
  return k;
}
fused::LevelK level_k(f2d_mg *mg, int lev) { return level_k(mg, mg->L[lev]); }

// tensor map (cached) of a level array for boxh x boxw tiles; false: TMA not usable here
bool get_tmap(f2d_mg *mg, const double *base, int ny, int nx, int boxh, int boxw, CUtensorMap *out) {
  if (!mg->tma || !base || nx < boxw || ny < boxh) return false;
  auto key = std::make_tuple((const void *)base, ny, nx, boxh, boxw);
  auto it = mg->tmaps.find(key);
  if (it == mg->tmaps.end()) {
    CUtensorMap tm;
    if (make_tmap_2d(&tm, base, ny, nx, boxh, boxw) != 0) return false;
    it = mg->tmaps.emplace(key, tm).first;
  }
  *out = it->second;
  return true;
}

// Levels that give the standard 64 x 32 tiles fewer CTAs than the device has SMs run on 64 x 8
// tiles: four times as many CTAs, a quarter of the rows per thread (the kernels of such levels
// are bound by the instruction latency of one CTA, with most SMs idle).
// F2D_SMALL_TILE_MAXCTAS: largest standard-tile CTA count that still switches (0: never).
// Measured on a B200 (profiles/r02_vcycle_by_level_4096_v7.txt, r02_bench_*_v7): 256^2 (32 standard
// CTAs) 12.3 -> 8.1 us per level visit, 512^2 (128 CTAs) 8.2 -> 9.0 us (slower: stays standard);
// stored-coefficient levels gain up to 128 CTAs (256 x 64: 13 -> 7 us per kernel, 1024 x 256: 14.3 -> 13.2).
bool small_tiles(const Level &l) {
  static long long maxctas = -1;
  if (maxctas < 0) {
    maxctas = 96;
    if (const char *e = getenv("F2D_SMALL_TILE_MAXCTAS")) maxctas = atoll(e);
  }
  const long long ctas = (long long)cdiv(l.nx - 2 * NH, fused::TX) * cdiv(l.ny - 2 * NH, fused::TY);
  const long long lim = (l.mode == 0 && maxctas > 0) ? (maxctas * 3) / 2 : maxctas;   // stored: the short strips also prefetch their coefficients
  return ctas <= lim && l.ny - 2 * NH >= fused::TYS;
}

// fused double sweep: xout = S2(input), input = xin | 0 | I(xc) | xin + I(xc)
template <int INPUT, int TYP>
int launch_smooth2(f2d_mg *mg, Level &l, Level *cl, const double *xin, const double *b, double *xout,
                   const double *xc, cudaStream_t s, double *acc) {
  fused::LevelK k = level_k(mg, l);
  dim3 grid(cdiv(l.nx - 2 * NH, fused::TX), cdiv(l.ny - 2 * NH, TYP));
  const size_t sm = l.mode == 1 ? fused::smooth2_smem_nomask<TYP>() : sizeof(fused::Smooth2SmemT<TYP>);
  const int8_t *mskc = nullptr;
  int nxc = 0, nyc = 0;
  if (INPUT >= 2) {
    Level &c = *cl;
    mskc = c.msk; nxc = c.nx; nyc = c.ny;
  }
  // slab levels of a multi-GPU hierarchy: the kernel also fills the neighbours' halo rows
  const bool peer = mg->comm != nullptr && l.ywrap == 0;
  f2d::Peer P = comm_peer(peer ? mg->comm : nullptr);
  if (peer && !comm_owns(mg->comm, acc ? acc : xout))
    return fail(F2D_ERR_ARG, "smooth: the output of a slab level must live in the symmetric heap");
  CUtensorMap tmx, tmc;
  memset(&tmx, 0, sizeof tmx); memset(&tmc, 0, sizeof tmc);
  int use_tma = mg->tma ? 1 : 0;   // (INPUT == 1 stages nothing but the masks: either path does)
  if (use_tma && (INPUT == 0 || INPUT == 3)) use_tma = get_tmap(mg, xin, l.ny, l.nx, fused::smooth2_xh<TYP>(), fused::XP, &tmx);
  if (use_tma && INPUT >= 2) use_tma = get_tmap(mg, xc, nyc, nxc, fused::smooth2_ch<TYP>(), fused::CP, &tmc);
  prof_tag("k_smooth2<mode%d,input%d%s> %dx%d", l.mode, INPUT, peer ? ",peer" : "", l.nx - 2 * NH, l.ny - 2 * NH);
#define F2D_SM2(M, St)                                                                                              \
  do {                                                                                                              \
    if (peer) F2D_CUDA(f2d::launch_pdl(fused::k_smooth2<M, St, INPUT, true, TYP>, grid, dim3(fused::NT), sm, s, k, xin, b, xout, xc, mskc, nxc, nyc, acc, P, use_tma, tmx, tmc)); \
    else F2D_CUDA(f2d::launch_pdl(fused::k_smooth2<M, St, INPUT, false, TYP>, grid, dim3(fused::NT), sm, s, k, xin, b, xout, xc, mskc, nxc, nyc, acc, P, use_tma, tmx, tmc));     \
  } while (0)
  switch (l.mode) {
    case 1: F2D_SM2(false, false); break;
    case 2: F2D_SM2(true, false); break;
    default: F2D_SM2(true, true); break;
  }
#undef F2D_SM2
  F2D_LAUNCHED();
  return F2D_OK;
}
int smooth2_L(f2d_mg *mg, Level &l, Level *cl, int input, const double *xin, const double *b, double *xout,
              const double *xc, cudaStream_t s, double *acc = nullptr) {
  if (small_tiles(l)) {
    switch (input) {
      case 0: return launch_smooth2<0, fused::TYS>(mg, l, cl, xin, b, xout, xc, s, acc);
      case 1: return launch_smooth2<1, fused::TYS>(mg, l, cl, xin, b, xout, xc, s, acc);
      case 2: return launch_smooth2<2, fused::TYS>(mg, l, cl, xin, b, xout, xc, s, acc);
      default: return launch_smooth2<3, fused::TYS>(mg, l, cl, xin, b, xout, xc, s, acc);
    }
  }
  switch (input) {
    case 0: return launch_smooth2<0, fused::TY>(mg, l, cl, xin, b, xout, xc, s, acc);
    case 1: return launch_smooth2<1, fused::TY>(mg, l, cl, xin, b, xout, xc, s, acc);
    case 2: return launch_smooth2<2, fused::TY>(mg, l, cl, xin, b, xout, xc, s, acc);
    default: return launch_smooth2<3, fused::TY>(mg, l, cl, xin, b, xout, xc, s, acc);
  }
}
int smooth2(f2d_mg *mg, int lev, int input, const double *xin, const double *b, double *xout, const double *xc,
            cudaStream_t s, double *acc = nullptr) {
  Level *cl = lev + 1 < (int)mg->L.size() ? &mg->L[lev + 1] : nullptr;
  return smooth2_L(mg, mg->L[lev], cl, input, xin, b, xout, xc, s, acc);
}
template <bool M, bool St, int I>
cudaError_t set_smem_one() {
  cudaError_t e = cudaFuncSetAttribute(fused::k_smooth2<M, St, I, false, fused::TY>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(fused::Smooth2Smem));
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(fused::k_smooth2<M, St, I, true, fused::TY>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)sizeof(fused::Smooth2Smem));
  // (the small-tile instantiations stay under the 48 KB that need no opt-in)
}
static_assert(sizeof(fused::Smooth2SmemT<fused::TYS>) <= 48 * 1024 && sizeof(fused::ResidSmemT<fused::RTYS>) <= 48 * 1024,
              "small tiles: dynamic shared memory without opt-in");
template <int I>
cudaError_t set_smem_input() {
  cudaError_t e = set_smem_one<false, false, I>();
  if (e == cudaSuccess) e = set_smem_one<true, false, I>();
  if (e == cudaSuccess) e = set_smem_one<true, true, I>();
  return e;
}
template <bool M, bool St>
cudaError_t set_smem_resid() {
  cudaError_t e = cudaFuncSetAttribute(fused::k_resid_restrict<M, St, false, fused::RTY>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(fused::ResidSmem));
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(fused::k_resid_restrict<M, St, true, fused::RTY>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)sizeof(fused::ResidSmem));
}
cudaError_t set_smem_all() {
  cudaError_t e = set_smem_input<0>();
  if (e == cudaSuccess) e = set_smem_resid<false, false>();
  if (e == cudaSuccess) e = set_smem_resid<true, false>();
  if (e == cudaSuccess) e = set_smem_resid<true, true>();
  if (e == cudaSuccess) e = set_smem_input<1>();
  if (e == cudaSuccess) e = set_smem_input<2>();
  if (e == cudaSuccess) e = set_smem_input<3>();
  return e;
}

// Grid.smooth: nite x (double sweep + fill); ping-pong through the level scratch t
int op_smooth(f2d_mg *mg, int lev, double *x, const double *b, int nite, cudaStream_t s) {
  Level &l = mg->L[lev];
  if (mg->relax == 1) {
    // level.py:340-349: three line relaxations (+ halo fill) per requested iteration
    const size_t sm = 4 * (size_t)l.ny * sizeof(double);
    for (int k = 0; k < 3 * nite; k++) {
      k_smooth_tridiag<<<1, 256, sm, s>>>(l.msk, l.A, x, b, l.ny, l.nx);
      F2D_LAUNCHED();
      int rc = f2d_fill_halo(x, NH, l.ny, l.nx, (f2d_stream_t)s);
      if (rc != F2D_OK) return rc;
    }
    return F2D_OK;
  }
  double *cur = x, *other = l.t;
  for (int k = 0; k < nite; k++) {
    int rc = smooth2(mg, lev, 0, cur, b, other, nullptr, s);
    if (rc != F2D_OK) return rc;
    double *tmp = cur; cur = other; other = tmp;
  }
  if (cur != x) F2D_CUDA(cudaMemcpyAsync(x, cur, l.n() * sizeof(double), cudaMemcpyDeviceToDevice, s));
  return F2D_OK;
}
int op_residual(f2d_mg *mg, int lev, const double *x, const double *b, double *r, cudaStream_t s) {
  Level &l = mg->L[lev];
  dim3 blk(32, 8);
  k_residual<<<grid2d(l.ny, l.nx, blk), blk, 0, s>>>(l.msk, l.A, x, b, r, l.ny, l.nx);
  F2D_LAUNCHED();
  return F2D_OK;
}
// residual of level 0 + sum of its squares -> out[0] (device)
int op_resid_sumsq_L(f2d_mg *mg, Level &l, const double *x, const double *b, double *r, double *out,
                     cudaStream_t s) {
  fused::LevelK k = level_k(mg, l);
  CUtensorMap tmx;
  memset(&tmx, 0, sizeof tmx);
  if (get_tmap(mg, x, l.ny, l.nx, QH, QW, &tmx)) {
    dim3 tg(cdiv(l.nx - 2 * NH, QTX), cdiv(l.ny - 2 * NH, QTY));
    const int ntb = (int)(tg.x * tg.y);
    if ((size_t)ntb > mg->npartials) return fail(F2D_ERR_ARG, "resid_sumsq: partial-sum buffer too small");
    prof_tag("k_resid_sumsq<mode%d> %dx%d", l.mode, l.nx - 2 * NH, l.ny - 2 * NH);
    switch (l.mode) {
      case 1: F2D_CUDA(f2d::launch_pdl(k_resid_sumsq_tma<false, false>, tg, dim3(RST), 0, s, k, b, r, mg->partials, tmx)); break;
      case 2: F2D_CUDA(f2d::launch_pdl(k_resid_sumsq_tma<true, false>, tg, dim3(RST), 0, s, k, b, r, mg->partials, tmx)); break;
      default: F2D_CUDA(f2d::launch_pdl(k_resid_sumsq_tma<true, true>, tg, dim3(RST), 0, s, k, b, r, mg->partials, tmx)); break;
    }
    F2D_LAUNCHED();
    F2D_CUDA(f2d::launch_pdl(k_fold_partials, dim3(1), dim3(RST), 0, s, mg->partials, ntb, out));
    F2D_LAUNCHED();
    return F2D_OK;
  }
  dim3 grid(cdiv(l.nx - 2 * NH, RST), cdiv(l.ny - 2 * NH, RSR));
  const int nb = (int)(grid.x * grid.y);
  if ((size_t)nb > mg->npartials) return fail(F2D_ERR_ARG, "resid_sumsq: partial-sum buffer too small");
  prof_tag("k_resid_sumsq<mode%d> %dx%d", l.mode, l.nx - 2 * NH, l.ny - 2 * NH);
  switch (l.mode) {
    case 1: F2D_CUDA(f2d::launch_pdl(k_resid_sumsq<false, false>, grid, dim3(RST), 0, s, k, x, b, r, mg->partials)); break;
    case 2: F2D_CUDA(f2d::launch_pdl(k_resid_sumsq<true, false>, grid, dim3(RST), 0, s, k, x, b, r, mg->partials)); break;
    default: F2D_CUDA(f2d::launch_pdl(k_resid_sumsq<true, true>, grid, dim3(RST), 0, s, k, x, b, r, mg->partials)); break;
  }
  F2D_LAUNCHED();
  F2D_CUDA(f2d::launch_pdl(k_fold_partials, dim3(1), dim3(RST), 0, s, mg->partials, nb, out));
  F2D_LAUNCHED();
  return F2D_OK;
}
int op_resid_sumsq(f2d_mg *mg, const double *x, const double *b, double *r, double *out, cudaStream_t s) {
  return op_resid_sumsq_L(mg, mg->L[0], x, b, r, out, s);
}
int op_restrict_L(f2d_mg *mg, Level &f, Level &c, const double *xf, double *xc, cudaStream_t s) {
  dim3 blk(32, 8);
  const bool peer = mg->comm != nullptr && c.ywrap == 0;
  f2d::Peer P = comm_peer(peer ? mg->comm : nullptr);
  if (peer && !comm_owns(mg->comm, xc))
    return fail(F2D_ERR_ARG, "restrict: the output of a slab level must live in the symmetric heap");
  prof_tag("k_restrict%s %dx%d", peer ? "<peer>" : "", f.nx - 2 * NH, f.ny - 2 * NH);
  if (peer) F2D_CUDA(f2d::launch_pdl(k_restrict<true>, grid2d(c.ny, c.nx, blk), blk, 0, s, c.msk, xf, xc, c.ny, c.nx, f.nx, c.ywrap, P));
  else F2D_CUDA(f2d::launch_pdl(k_restrict<false>, grid2d(c.ny, c.nx, blk), blk, 0, s, c.msk, xf, xc, c.ny, c.nx, f.nx, c.ywrap, P));
  F2D_LAUNCHED();
  return F2D_OK;
}
int op_restrict(f2d_mg *mg, int lev, const double *xf, double *xc, cudaStream_t s) {
  return op_restrict_L(mg, mg->L[lev], mg->L[lev + 1], xf, xc, s);
}
// residual + restriction fused: bc = R(b - A x), the fine residual stays on chip
int op_resid_restrict_L(f2d_mg *mg, Level &l, Level &c, const double *x, const double *b, double *bc,
                        cudaStream_t s);
int op_resid_restrict(f2d_mg *mg, int lev, const double *x, const double *b, double *bc, cudaStream_t s) {
  return op_resid_restrict_L(mg, mg->L[lev], mg->L[lev + 1], x, b, bc, s);
}
template <int RTYP>
int launch_resid_restrict(f2d_mg *mg, Level &l, Level &c, const double *x, const double *b, double *bc, cudaStream_t s);
int op_resid_restrict_L(f2d_mg *mg, Level &l, Level &c, const double *x, const double *b, double *bc,
                        cudaStream_t s) {
  if (small_tiles(l)) return launch_resid_restrict<fused::RTYS>(mg, l, c, x, b, bc, s);
  return launch_resid_restrict<fused::RTY>(mg, l, c, x, b, bc, s);
}
template <int RTYP>
int launch_resid_restrict(f2d_mg *mg, Level &l, Level &c, const double *x, const double *b, double *bc, cudaStream_t s) {
  fused::LevelK k = level_k(mg, l);
  dim3 grid(cdiv(c.nx - 2 * NH, fused::RTX), cdiv(c.ny - 2 * NH, RTYP));
  const size_t sm = l.mode == 1 ? fused::resid_smem_nomask<RTYP>() : sizeof(fused::ResidSmemT<RTYP>);
  const bool peer = mg->comm != nullptr && l.ywrap == 0;
  f2d::Peer P = comm_peer(peer ? mg->comm : nullptr);
  if (peer && !comm_owns(mg->comm, bc))
    return fail(F2D_ERR_ARG, "restrict: the output of a slab level must live in the symmetric heap");
  CUtensorMap tmx, tmb;
  memset(&tmx, 0, sizeof tmx); memset(&tmb, 0, sizeof tmb);
  int use_tma = get_tmap(mg, x, l.ny, l.nx, 2 * RTYP + 3, fused::RXP, &tmx);
  if (use_tma) use_tma = get_tmap(mg, b, l.ny, l.nx, 2 * RTYP + 1, fused::RBP, &tmb);
  prof_tag("k_resid_restrict<mode%d%s> %dx%d", l.mode, peer ? ",peer" : "", l.nx - 2 * NH, l.ny - 2 * NH);
#define F2D_RR(M, St)                                                                                         \
  do {                                                                                                        \
    if (peer) F2D_CUDA(f2d::launch_pdl(fused::k_resid_restrict<M, St, true, RTYP>, grid, dim3(fused::NT), sm, s, k, x, b, bc, c.msk, c.ny, c.nx, P, use_tma, tmx, tmb)); \
    else F2D_CUDA(f2d::launch_pdl(fused::k_resid_restrict<M, St, false, RTYP>, grid, dim3(fused::NT), sm, s, k, x, b, bc, c.msk, c.ny, c.nx, P, use_tma, tmx, tmb));     \
  } while (0)
  switch (l.mode) {
    case 1: F2D_RR(false, false); break;
    case 2: F2D_RR(true, false); break;
    default: F2D_RR(true, true); break;
  }
#undef F2D_RR
  F2D_LAUNCHED();
  return F2D_OK;
}
// A level's whole visit on the way down from a zero first guess in one kernel
// (fused::k_zsmooth_resid_restrict): t = S2(0, b), bc = R(b - A t).  All-fluid doubly periodic
// levels whose coarse grid tiles exactly; F2D_MG_NO_ZRR=1 keeps the two-kernel form.
bool zrr_ok(f2d_mg *mg, const Level &l, const Level &c) {
  static int off = -1;
  if (off < 0) {
    const char *e = getenv("F2D_MG_NO_ZRR");
    off = (e && e[0] == '1') ? 1 : 0;
  }
  // (y-slab levels: the halo cells evaluated in place hold the neighbouring rank's data)
  if (off || !mg->tma || mg->relax != 0 || l.mode != 1 || (!l.ywrap && !mg->comm)) return false;
  const int rty = small_tiles(l) ? fused::RTYS : fused::RTY;
  if ((c.nx - 2 * NH) % fused::RTX || (c.ny - 2 * NH) % rty) return false;
  return l.nx - 2 * NH == 2 * (c.nx - 2 * NH) && l.ny - 2 * NH == 2 * (c.ny - 2 * NH);
}
template <int RTYP>
int launch_zsmooth_rr(f2d_mg *mg, Level &l, Level &c, const double *b, double *t, double *bc, cudaStream_t s) {
  fused::LevelK k = level_k(mg, l);
  CUtensorMap tmb;
  memset(&tmb, 0, sizeof tmb);
  if (!get_tmap(mg, b, l.ny, l.nx, fused::zrr_bh<RTYP>(), fused::ZBP, &tmb)) return fail(F2D_ERR_ARG, "zsmooth_rr: no tensor map");
  dim3 grid((c.nx - 2 * NH) / fused::RTX, (c.ny - 2 * NH) / RTYP);
  const bool peer = mg->comm != nullptr && l.ywrap == 0;
  f2d::Peer P = comm_peer(peer ? mg->comm : nullptr);
  if (peer && (!comm_owns(mg->comm, t) || !comm_owns(mg->comm, bc)))
    return fail(F2D_ERR_ARG, "zsmooth_rr: the outputs of a slab level must live in the symmetric heap");
  prof_tag("k_zsmooth_rr<mode1%s> %dx%d", peer ? ",peer" : "", l.nx - 2 * NH, l.ny - 2 * NH);
  if (peer)
    F2D_CUDA(f2d::launch_pdl(fused::k_zsmooth_resid_restrict<RTYP, true>, grid, dim3(fused::NT),
                             sizeof(fused::ZrrSmemT<RTYP>), s, k, t, bc, c.ny, c.nx, P, tmb));
  else
    F2D_CUDA(f2d::launch_pdl(fused::k_zsmooth_resid_restrict<RTYP, false>, grid, dim3(fused::NT),
                             sizeof(fused::ZrrSmemT<RTYP>), s, k, t, bc, c.ny, c.nx, P, tmb));
  F2D_LAUNCHED();
  return F2D_OK;
}
int op_zsmooth_rr_L(f2d_mg *mg, Level &l, Level &c, const double *b, double *t, double *bc, cudaStream_t s) {
  if (small_tiles(l)) return launch_zsmooth_rr<fused::RTYS>(mg, l, c, b, t, bc, s);
  return launch_zsmooth_rr<fused::RTY>(mg, l, c, b, t, bc, s);
}
int op_interpolate(f2d_mg *mg, int lev, const double *xc, double *xf, int add, cudaStream_t s) {
  Level &f = mg->L[lev];
  dim3 blk(32, 8);
  k_interpolate<<<grid2d(f.ny, f.nx, blk), blk, 0, s>>>(f.msk, mg->L[lev + 1].msk, xc, xf, f.ny, f.nx,
                                                        mg->L[lev + 1].nx, add);
  F2D_LAUNCHED();
  return F2D_OK;
}

#define TRY(call)               \
  do {                          \
    int rc__ = (call);          \
    if (rc__ != F2D_OK) return rc__; \
  } while (0)

// deepest level: x = 0, then ndeepest double sweeps (hierarchy.py:114-116); the first
// sweep takes the zero input implicitly, the result ends in X (ndeepest is even)
int coarsest_enqueue(f2d_mg *mg, double *X, const double *B, cudaStream_t s) {
  int last = (int)mg->L.size() - 1;
  Level &l = mg->L[last];
  if (mg->relax == 1) {
    F2D_CUDA(cudaMemsetAsync(X, 0, l.n() * sizeof(double), s));
    return op_smooth(mg, last, X, B, mg->ndeepest, s);
  }
  double *cur = X, *other = l.t;
  for (int k = 0; k < mg->ndeepest; k++) {
    TRY(smooth2(mg, last, k == 0 ? 1 : 0, cur, B, other, nullptr, s));
    double *tmp = cur; cur = other; other = tmp;
  }
  if (cur != X) F2D_CUDA(cudaMemcpyAsync(X, cur, l.n() * sizeof(double), cudaMemcpyDeviceToDevice, s));
  return F2D_OK;
}

// instantiations of the periodic tail: (cluster size, log2 of the finest level)
template <int TOP, int NC>
int ptail_launch_t(const ptail::Params &P, int program, cudaStream_t s, bool probe, size_t *smem_out) {
  constexpr size_t smem = ptail::smem_bytes<TOP, NC>();
  if (smem_out) *smem_out = smem;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(NC);
  cfg.blockDim = dim3(ptail::NT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = NC;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = (!probe && f2d::pdl_enabled()) ? 2 : 1;
  if (probe) {   // set-up: attributes, and can the device co-schedule such a cluster?
    if (NC > 8 && cudaFuncSetAttribute(ptail::k_mg_ptail<TOP, NC>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
      cudaGetLastError();
      return F2D_ERR_CUDA;
    }
    if (cudaFuncSetAttribute(ptail::k_mg_ptail<TOP, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      cudaGetLastError();
      return F2D_ERR_CUDA;
    }
    if (NC > 1) {
      int nclusters = 0;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, ptail::k_mg_ptail<TOP, NC>, &cfg);
      if (e != cudaSuccess || nclusters < 1) { cudaGetLastError(); return F2D_ERR_CUDA; }
    }
    return F2D_OK;
  }
  F2D_CUDA(cudaLaunchKernelEx(&cfg, ptail::k_mg_ptail<TOP, NC>, P, program));
  return F2D_OK;
}
int ptail_launch(int top, int nc, const ptail::Params &P, int program, cudaStream_t s, bool probe, size_t *smem_out) {
#define F2D_PT(T, N) if (top == T && nc == N) return ptail_launch_t<T, N>(P, program, s, probe, smem_out)
  F2D_PT(8, 16); F2D_PT(7, 16); F2D_PT(6, 16);
  F2D_PT(7, 8); F2D_PT(6, 8);
  F2D_PT(6, 1); F2D_PT(5, 1); F2D_PT(4, 1); F2D_PT(3, 1); F2D_PT(2, 1);
#undef F2D_PT
  return F2D_ERR_ARG;
}

// one launch of the shared-memory tail: program 0/1 = V-cycle (x = 0 / x = x_in first),
// 2 = F-cycle of the levels tail0..last; rhs b_in, result x_out (both of level tail0)
int tail_launch(f2d_mg *mg, int program, const double *b_in, const double *x_in, double *x_out, cudaStream_t s,
                double *acc = nullptr) {
  {
    Level &t0l = mg->L[mg->tail0];
    prof_tag("%s<program%d> %dx%d", mg->ptail ? "k_mg_ptail" : (mg->ctail ? "k_mg_ctail" : "k_mg_tail"), program,
             t0l.nx - 2 * NH, t0l.ny - 2 * NH);
  }
  if (mg->ptail) {
    ptail::Params P;
    const int last = (int)mg->L.size() - 1;
    for (int lg = ptail::LGMIN; lg <= mg->ptail_top; lg++) P.k[lg - ptail::LGMIN] = level_k(mg, last - (lg - ptail::LGMIN));
    P.ndeepest = mg->ndeepest;
    P.b_in = b_in;
    P.x_in = x_in;
    P.x_out = x_out;
    P.acc = acc;
    P.trace = mg->trace;
    P.trace_cap = mg->trace_cap;
    TRY(ptail_launch(mg->ptail_top, mg->ctail_nc, P, program, s, false, nullptr));
    F2D_LAUNCHED();
    return F2D_OK;
  }
  if (mg->ctail) {
    ctail::Params P = mg->ctp;
    P.b_in = b_in;
    P.x_in = x_in;
    P.x_out = x_out;
    P.acc = acc;
    P.trace = mg->trace;
    P.trace_cap = mg->trace_cap;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(mg->ctail_nc);
    cfg.blockDim = dim3(ctail::NT);
    cfg.dynamicSmemBytes = mg->tail_smem;
    cfg.stream = s;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = mg->ctail_nc;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = f2d::pdl_enabled() ? 2 : 1;
    if (mg->tail_const) F2D_CUDA(cudaLaunchKernelEx(&cfg, ctail::k_mg_ctail<false, false>, P, program));
    else F2D_CUDA(cudaLaunchKernelEx(&cfg, ctail::k_mg_ctail<true, true>, P, program));
    F2D_LAUNCHED();
    return F2D_OK;
  }
  tail::Params P;
  int n = (int)mg->L.size() - mg->tail0;
  P.nlev = n;
  int off = 0;
  for (int k = 0; k < n; k++) {
    P.lv[k] = level_k(mg, mg->tail0 + k);
    P.off[k] = off;
    off += mg->L[mg->tail0 + k].ny * mg->L[mg->tail0 + k].nx;
  }
  P.total = off;
  P.ndeepest = mg->ndeepest;
  P.b_in = b_in;
  P.x_in = x_in;
  P.x_out = x_out;
  P.acc = acc;
  if (mg->tail_const)
    tail::k_mg_tail<false, false><<<1, mg->tail_nt, mg->tail_smem, s>>>(P, program);
  else
    tail::k_mg_tail<true, true><<<1, mg->tail_nt, mg->tail_smem, s>>>(P, program);
  F2D_LAUNCHED();
  return F2D_OK;
}

// hierarchy.py:98-127 with npre = npost = 1; x0/b0 stand for self.x[lev1], self.b[lev1].
// Going down, the pre-smoothed field of a level lives in its scratch t; coming up,
// interpolation, correction and post-smoothing are one kernel that writes x again.
// Levels >= tail0 are done by one launch of the shared-memory tail kernel.
// first_input: 0 = start from x0 as it is; 2 = x0 := I(x[lev1+1]) first (the F-cycle's
// coarsetofine, hierarchy.py:145-146, fused into the pre-smoothing).
int vcycle_enqueue(f2d_mg *mg, int lev1, double *x0, double *b0, cudaStream_t s, int first_input = 0,
                   double *acc = nullptr) {
  int last = (int)mg->L.size() - 1;
  auto X = [&](int lev) { return lev == lev1 ? x0 : mg->L[lev].x; };
  auto B = [&](int lev) { return lev == lev1 ? b0 : mg->L[lev].b; };
  if (mg->relax == 1) {
    // line relaxation: the operators one by one, as hierarchy.py:98-127 lists them (the fused
    // kernels below are built around the Jacobi double sweep)
    if (first_input == 1) F2D_CUDA(cudaMemsetAsync(X(lev1), 0, mg->L[lev1].n() * sizeof(double), s));
    if (first_input == 2) TRY(op_interpolate(mg, lev1, X(lev1 + 1), X(lev1), 0, s));
    for (int lev = lev1; lev < last; lev++) {
      Level &l = mg->L[lev];
      if (lev > lev1) F2D_CUDA(cudaMemsetAsync(X(lev), 0, l.n() * sizeof(double), s));
      TRY(op_smooth(mg, lev, X(lev), B(lev), mg->npre, s));
      TRY(op_residual(mg, lev, X(lev), B(lev), l.r, s));
      TRY(op_restrict(mg, lev, l.r, B(lev + 1), s));
    }
    TRY(coarsest_enqueue(mg, X(last), B(last), s));
    for (int lev = last - 1; lev >= lev1; lev--) {
      TRY(op_interpolate(mg, lev, X(lev + 1), X(lev), 1, s));
      TRY(op_smooth(mg, lev, X(lev), B(lev), mg->npost, s));
    }
    if (acc) {
      k_add_inplace<<<nblocks1d(mg->L[lev1].n()), 256, 0, s>>>(acc, X(lev1), mg->L[lev1].n());
      F2D_LAUNCHED();
    }
    return F2D_OK;
  }
  const int t0 = mg->tail0;
  if (t0 >= 0 && lev1 == t0 && first_input == 0) return tail_launch(mg, 1, b0, x0, x0, s);
  if (t0 >= 0 && lev1 == t0 && first_input == 1) return tail_launch(mg, 0, b0, nullptr, x0, s);
  const bool use_tail = t0 >= 0 && lev1 < t0;
  const int bottom = use_tail ? t0 : last;   // first level NOT handled by the big kernels
  if (lev1 == last) return coarsest_enqueue(mg, X(last), B(last), s);
  for (int lev = lev1; lev < bottom; lev++) {
    Level &l = mg->L[lev];
    int input = lev > lev1 ? 1 : first_input;
    if (input == 1 && zrr_ok(mg, l, mg->L[lev + 1])) {
      TRY(op_zsmooth_rr_L(mg, l, mg->L[lev + 1], B(lev), l.t, B(lev + 1), s));
      continue;
    }
    TRY(smooth2(mg, lev, input, X(lev), B(lev), l.t, input == 2 ? X(lev + 1) : nullptr, s));
    TRY(op_resid_restrict(mg, lev, l.t, B(lev), B(lev + 1), s));
  }
  if (use_tail)
    TRY(tail_launch(mg, 0, B(t0), nullptr, X(t0), s));
  else
    TRY(coarsest_enqueue(mg, X(last), B(last), s));
  for (int lev = bottom - 1; lev >= lev1; lev--)
    TRY(smooth2(mg, lev, 3, mg->L[lev].t, B(lev), X(lev), X(lev + 1), s, lev == lev1 ? acc : nullptr));
  return F2D_OK;
}

// hierarchy.py:131-151
// acc != nullptr (only with lev1 below the coarsest level): the result is added to acc
// instead of being stored in x0
int fcycle_enqueue(f2d_mg *mg, int lev1, double *x0, double *b0, cudaStream_t s, double *acc = nullptr) {
  int last = (int)mg->L.size() - 1;
  auto X = [&](int lev) { return lev == lev1 ? x0 : mg->L[lev].x; };
  auto B = [&](int lev) { return lev == lev1 ? b0 : mg->L[lev].b; };
  const int t0 = mg->tail0;
  if (t0 >= 0 && lev1 <= t0) {
    for (int lev = lev1; lev < t0; lev++) TRY(op_restrict(mg, lev, B(lev), B(lev + 1), s));
    TRY(tail_launch(mg, 2, B(t0), nullptr, X(t0), s, lev1 == t0 ? acc : nullptr));
    for (int lev = t0 - 1; lev >= lev1; lev--)
      for (int k = 0; k < mg->nvcyc; k++)
        TRY(vcycle_enqueue(mg, lev, X(lev), B(lev), s, k == 0 ? 2 : 0,
                           (lev == lev1 && k == mg->nvcyc - 1) ? acc : nullptr));
    return F2D_OK;
  }
  for (int lev = lev1; lev < last; lev++) TRY(op_restrict(mg, lev, B(lev), B(lev + 1), s));
  TRY(coarsest_enqueue(mg, X(last), B(last), s));
  for (int lev = last - 1; lev >= lev1; lev--)
    for (int k = 0; k < mg->nvcyc; k++)
      TRY(vcycle_enqueue(mg, lev, X(lev), B(lev), s, k == 0 ? 2 : 0,
                         (lev == lev1 && k == mg->nvcyc - 1) ? acc : nullptr));
  return F2D_OK;
}

int slab_cycle_enqueue(f2d_mg *mg, int kind, int lev1, double *x0, double *b0, cudaStream_t s, double *acc);

// run `kind` (0 two V-cycles from level 0, 1 F-cycle, 2 single V-cycle) through a cached graph
int run_cycle(f2d_mg *mg, int kind, int lev1, double *x0, double *b0, cudaStream_t s, double *acc = nullptr) {
  auto enqueue = [&](cudaStream_t st) -> int {
    if (mg->comm) return slab_cycle_enqueue(mg, kind, lev1, x0, b0, st, acc);
    if (kind == 0) {
      TRY(vcycle_enqueue(mg, 0, x0, b0, st));
      return vcycle_enqueue(mg, 0, x0, b0, st);
    }
    if (kind == 1) return fcycle_enqueue(mg, lev1, x0, b0, st, acc);
    return vcycle_enqueue(mg, lev1, x0, b0, st);
  };
  if (!mg->graphs) return enqueue(s);
  auto key = std::make_tuple(kind, lev1, (const void *)x0, (const void *)b0, (const void *)acc);
  auto it = mg->cache.find(key);
  if (it == mg->cache.end()) {
    long long before = g_launches;
    F2D_CUDA(cudaStreamBeginCapture(mg->cap, cudaStreamCaptureModeThreadLocal));
    int rc = enqueue(mg->cap);
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(mg->cap, &graph);
    long long nk = g_launches - before;
    g_launches = before;
    if (rc != F2D_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return cuda_fail(e, "cudaStreamEndCapture");
    f2d_mg::G g;
    g.kernels = nk;
    e = cudaGraphInstantiate(&g.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGraphInstantiate");
    it = mg->cache.emplace(key, g).first;
  }
  F2D_CUDA(cudaGraphLaunch(it->second.exec, s));
  g_launches += it->second.kernels;
  return F2D_OK;
}

int read_scalars(f2d_mg *mg, int n, cudaStream_t s) {
  F2D_CUDA(cudaMemcpyAsync(mg->hscal, mg->dscal, n * sizeof(double), cudaMemcpyDeviceToHost, s));
  F2D_CUDA(cudaStreamSynchronize(s));
  return F2D_OK;
}

// ---- device-side iteration control of Gmg.solve (hierarchy.py:154-192) -------------
// The loop `while nite < maxite and res0 > tol` runs inside the graph: a conditional WHILE
// node whose condition these two kernels set from the (all-reduced) norms, so a solve costs
// one graph launch and one host read instead of a host round trip per F-cycle.
__global__ void k_solve_init(const double *dscal, f2d_mg::SolveState *st, cudaGraphConditionalHandle h, double tol,
                             int maxite) {
  f2d::pdl_trigger();
  f2d::pdl_wait();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double normb = sqrt(dscal[0]);
  double res0 = 0.;
  bool go = false;
  if (normb > 0) {
    res0 = sqrt(dscal[1]) / normb;
    go = maxite > 0 && res0 > tol;
  }
  st->normb = normb; st->res0 = res0; st->res = res0;
  st->nite = 0; st->ndiv = 0; st->diverged = 0;
  cudaGraphSetConditional(h, go ? 1u : 0u);
}
__global__ void k_solve_step(const double *dscal, f2d_mg::SolveState *st, cudaGraphConditionalHandle h, double tol,
                             int maxite) {
  f2d::pdl_trigger();
  f2d::pdl_wait();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double res = sqrt(dscal[1]) / st->normb;
  const double conv = st->res0 / res;
  st->res0 = res; st->res = res;
  const int nite = st->nite + 1;
  st->nite = nite;
  int ndiv = st->ndiv;
  if (conv < 1) st->ndiv = ++ndiv;
  bool go = nite < maxite && res > tol;
  if (ndiv > 4) { st->diverged = 1; go = false; }
  cudaGraphSetConditional(h, go ? 1u : 0u);
}

int solve_pre(f2d_mg *mg, double *psi, const double *rhs, cudaStream_t s);
int solve_body(f2d_mg *mg, double *psi, const double *rhs, cudaStream_t s);

int build_solve_graph(f2d_mg *mg, double *psi, const double *rhs, double tol, int maxite, f2d_mg::SG &out) {
  cudaStream_t s = mg->cap;
  const long long before = g_launches;
  F2D_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  cudaGraph_t graph = nullptr;
  auto abort_capture = [&](int rc) {
    cudaGraph_t g = nullptr;
    cudaStreamEndCapture(s, &g);
    if (g) cudaGraphDestroy(g);
    g_launches = before;
    return rc;
  };
  int rc = solve_pre(mg, psi, rhs, s);
  if (rc != F2D_OK) return abort_capture(rc);
  cudaStreamCaptureStatus status;
  cudaGraph_t g = nullptr;
  const cudaGraphNode_t *deps = nullptr;
  size_t ndeps = 0;
  cudaError_t e = cudaStreamGetCaptureInfo_v2(s, &status, nullptr, &g, &deps, &ndeps);
  if (e != cudaSuccess) return abort_capture(cuda_fail(e, "cudaStreamGetCaptureInfo"));
  cudaGraphConditionalHandle h;
  e = cudaGraphConditionalHandleCreate(&h, g, 0, cudaGraphCondAssignDefault);
  if (e != cudaSuccess) return abort_capture(cuda_fail(e, "cudaGraphConditionalHandleCreate"));
  e = f2d::launch_pdl(k_solve_init, dim3(1), dim3(32), 0, s, mg->dscal, mg->dstate, h, tol, maxite);
  if (e != cudaSuccess) return abort_capture(cuda_fail(e, "k_solve_init"));
  ++g_launches;
  out.pre = g_launches - before;
  e = cudaStreamGetCaptureInfo_v2(s, &status, nullptr, &g, &deps, &ndeps);
  if (e != cudaSuccess) return abort_capture(cuda_fail(e, "cudaStreamGetCaptureInfo"));
  cudaGraphNodeParams np = {cudaGraphNodeTypeConditional};
  np.type = cudaGraphNodeTypeConditional;
  np.conditional.handle = h;
  np.conditional.type = cudaGraphCondTypeWhile;
  np.conditional.size = 1;
  cudaGraphNode_t wnode;
  e = cudaGraphAddNode(&wnode, g, deps, ndeps, &np);
  if (e != cudaSuccess) return abort_capture(cuda_fail(e, "cudaGraphAddNode(conditional)"));
  cudaGraph_t body = np.conditional.phGraph_out[0];
  e = cudaStreamUpdateCaptureDependencies(s, &wnode, 1, cudaStreamSetCaptureDependencies);
  if (e != cudaSuccess) return abort_capture(cuda_fail(e, "cudaStreamUpdateCaptureDependencies"));
  // ---- the loop body, captured into the conditional node's graph
  const long long b0 = g_launches;
  e = cudaStreamBeginCaptureToGraph(mg->cap2, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
  if (e != cudaSuccess) return abort_capture(cuda_fail(e, "cudaStreamBeginCaptureToGraph"));
  rc = solve_body(mg, psi, rhs, mg->cap2);
  if (rc == F2D_OK) {
    if (f2d::launch_pdl(k_solve_step, dim3(1), dim3(32), 0, mg->cap2, mg->dscal, mg->dstate, h, tol, maxite) != cudaSuccess)
      rc = F2D_ERR_CUDA;
    ++g_launches;
  }
  e = cudaStreamEndCapture(mg->cap2, nullptr);
  if (rc != F2D_OK) return abort_capture(rc);
  if (e != cudaSuccess) return abort_capture(cuda_fail(e, "cudaStreamEndCapture(body)"));
  out.body = g_launches - b0;
  // ---- result -> pinned host
  e = cudaMemcpyAsync(mg->hstate, mg->dstate, sizeof(f2d_mg::SolveState), cudaMemcpyDeviceToHost, s);
  if (e != cudaSuccess) return abort_capture(cuda_fail(e, "cudaMemcpyAsync(state)"));
  e = cudaStreamEndCapture(s, &graph);
  g_launches = before;
  if (e != cudaSuccess) return cuda_fail(e, "cudaStreamEndCapture(solve)");
  e = cudaGraphInstantiate(&out.exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGraphInstantiate(solve)");
  return F2D_OK;
}

// solve through the cached WHILE graph; one synchronising read of (nite, res)
int solve_graph(f2d_mg *mg, double *psi, const double *rhs, double tol, int maxite, int *nite_out, double *res_out,
                cudaStream_t s) {
  auto key = std::make_tuple((const void *)psi, (const void *)rhs, tol, maxite);
  auto it = mg->solve_cache.find(key);
  if (it == mg->solve_cache.end()) {
    f2d_mg::SG sg;
    TRY(build_solve_graph(mg, psi, rhs, tol, maxite, sg));
    it = mg->solve_cache.emplace(key, sg).first;
  }
  F2D_CUDA(cudaGraphLaunch(it->second.exec, s));
  F2D_CUDA(cudaStreamSynchronize(s));
  const f2d_mg::SolveState &st = *mg->hstate;
  g_launches += it->second.pre + (long long)st.nite * it->second.body;
  if (st.diverged) return fail(F2D_ERR_DIVERGE, "solver is not converging");
  if (nite_out) *nite_out = st.nite;
  if (res_out) *res_out = st.res;
  return F2D_OK;
}

void free_level(Level &l) {
  cudaFree(l.msk); cudaFree(l.A); cudaFree(l.r);
  if (!l.arena) { cudaFree(l.x); cudaFree(l.b); cudaFree(l.t); }   // arena memory belongs to the communicator
}

}  // namespace

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
// ---------------------------------------------------------------------------
// set-up
// ---------------------------------------------------------------------------
namespace {

#define MGC(call)                                                         \
  do {                                                                    \
    cudaError_t e__ = (call);                                             \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call);                 \
  } while (0)

Stencil9 finest_stencil(double dx, double dy, double hydroepsilon) {
  // level.py:261-302
  double bx = dy / dx * hydroepsilon, by = dx / dy, a = -2 * (bx + by);
  double st[9] = {0., by, 0., bx, a, bx, 0., by, 0.};
  if (dx == dy && hydroepsilon == 1.) {
    double aa = -6. / 2, bb = 1. / 2, cc = 0.5 / 2;
    double st9[9] = {cc, bb, cc, bb, aa, bb, cc, bb, cc};
    for (int k = 0; k < 9; k++) st[k] = st9[k];
  }
  double coef = 1. / (dx * dy);
  Stencil9 S9;
  for (int k = 0; k < 9; k++) S9.v[k] = st[k] * coef;
  return S9;
}

// device allocation of a level's arrays; x, b, t come from the symmetric heap when asked
int alloc_level(f2d_mg *mg, Level &l, bool arena, cudaStream_t s) {
  size_t nb = l.n() * sizeof(double);
  MGC(cudaMalloc(&l.msk, l.n()));
  MGC(cudaMalloc(&l.A, 6 * nb));   // 5 diagonals + omega / |centre| (finish_setup)
  MGC(cudaMalloc(&l.r, nb));
  l.arena = arena;
  if (arena) {
    l.x = (double *)comm_alloc(mg->comm, nb);
    l.b = (double *)comm_alloc(mg->comm, nb);
    l.t = (double *)comm_alloc(mg->comm, nb);
    if (!l.x || !l.b || !l.t) return fail(F2D_ERR_ARG, "mg: symmetric heap exhausted (raise the arena size)");
  } else {
    MGC(cudaMalloc(&l.x, nb));
    MGC(cudaMalloc(&l.b, nb));
    MGC(cudaMalloc(&l.t, nb));
  }
  MGC(cudaMemsetAsync(l.x, 0, nb, s));
  MGC(cudaMemsetAsync(l.b, 0, nb, s));
  MGC(cudaMemsetAsync(l.r, 0, nb, s));
  MGC(cudaMemsetAsync(l.t, 0, nb, s));
  return F2D_OK;
}

// fill of one field of a level: local periodic wrap, or x images + exchange on a slab
int level_fill(f2d_mg *mg, Level &l, double *x, cudaStream_t s) {
  if (l.ywrap) return f2d_fill_halo(x, NH, l.ny, l.nx, (f2d_stream_t)s);
  TRY(f2d_fill_halo_x(x, NH, l.ny, l.nx, (f2d_stream_t)s));
  double *arr[1] = {x};
  return comm_exchange(mg->comm, arr, 1, NH, l.ny, l.nx, s);
}

// coarse mask and Galerkin matrix of level c from level p (level.py:233-236,304-329);
// A9p / A9c: 9 planes with filled halos (A9c in the symmetric heap on slab levels)
int coarsen_level(f2d_mg *mg, Level &p, Level &c, const double *A9p, double *A9c, cudaStream_t s) {
  dim3 blk(32, 8);
  size_t pl = c.n();
  k_mask_to_double<<<nblocks1d(p.n()), 256, 0, s>>>(p.msk, p.t, p.n());
  k_fill_const<<<nblocks1d(pl), 256, 0, s>>>(c.t, 1., pl);
  k_restrict<false><<<grid2d(c.ny, c.nx, blk), blk, 0, s>>>(nullptr, p.t, c.t, c.ny, c.nx, p.nx, c.ywrap, f2d::Peer{});
  g_launches += 3;
  if (!c.ywrap) {
    double *arr[1] = {c.t};
    TRY(comm_exchange(mg->comm, arr, 1, NH, c.ny, c.nx, s));
  }
  k_threshold_mask<<<nblocks1d(pl), 256, 0, s>>>(c.t, c.msk, pl);
  k_coarsenmatrix<<<grid2d(c.ny, c.nx, blk), blk, 0, s>>>(A9p, A9c, p.msk, c.msk, c.ny, c.nx, p.ny, p.nx);
  g_launches += 2;
  for (int k = 0; k < 9; k++) TRY(level_fill(mg, c, A9c + k * pl, s));
  return F2D_OK;
}

// replicated hierarchy below L[0]: L[0].msk and A9 (9 planes of L[0], halos filled) given
int build_replicated(f2d_mg *mg, double *A9, bool a9_owned, cudaStream_t s) {
  double *A9prev = A9;
  bool prev_owned = a9_owned;
  MGC(cudaMemcpyAsync(mg->L[0].A, A9, 5 * mg->L[0].n() * sizeof(double), cudaMemcpyDeviceToDevice, s));
  for (size_t lev = 1; lev < mg->L.size(); lev++) {
    Level &l = mg->L[lev];
    double *A9c = nullptr;
    MGC(cudaMalloc(&A9c, 9 * l.n() * sizeof(double)));
    int rc = coarsen_level(mg, mg->L[lev - 1], l, A9prev, A9c, s);
    if (rc != F2D_OK) { cudaFree(A9c); return rc; }
    MGC(cudaMemcpyAsync(l.A, A9c, 5 * l.n() * sizeof(double), cudaMemcpyDeviceToDevice, s));
    MGC(cudaStreamSynchronize(s));
    if (prev_owned) cudaFree(A9prev);
    A9prev = A9c;
    prev_owned = true;
  }
  MGC(cudaStreamSynchronize(s));
  if (prev_owned) cudaFree(A9prev);
  return F2D_OK;
}

// coefficient class of one level: read the stencil at the first cell whose 3x3
// neighbourhood is fluid, then verify entry by entry that "constant stencil x mask
// products" reproduces the stored matrix (f2d_mg_fused.cuh); ranks must agree
int detect_mode(f2d_mg *mg, Level &l, int *dflag, unsigned long long *didx, cudaStream_t s) {
  dim3 blk(32, 8);
  l.mode = 0;
  unsigned long long none = ~0ull, idx = none;
  MGC(cudaMemcpyAsync(didx, &none, sizeof none, cudaMemcpyHostToDevice, s));
  fused::k_find_interior<<<grid2d(l.ny, l.nx, blk), blk, 0, s>>>(l.msk, l.ny, l.nx, didx);
  MGC(cudaMemcpyAsync(&idx, didx, sizeof idx, cudaMemcpyDeviceToHost, s));
  MGC(cudaStreamSynchronize(s));
  int flags[2] = {1, 1};
  if (idx == none) {
    flags[0] = 0;
  } else {
    for (int k = 0; k < 5; k++)
      MGC(cudaMemcpyAsync(&l.cst[k], l.A + k * l.n() + idx, sizeof(double), cudaMemcpyDeviceToHost, s));
    MGC(cudaMemcpyAsync(dflag, flags, sizeof flags, cudaMemcpyHostToDevice, s));
    MGC(cudaStreamSynchronize(s));
    fused::k_check_const<<<grid2d(l.ny, l.nx, blk), blk, 0, s>>>(level_k(mg, l), dflag);
    MGC(cudaMemcpyAsync(flags, dflag, sizeof flags, cudaMemcpyDeviceToHost, s));
    MGC(cudaStreamSynchronize(s));
    g_launches += 2;
  }
  if (!l.ywrap && comm_size(mg->comm) > 1) {
    // slab level: every rank must take the same class and the same constants
    double h[12];
    h[0] = flags[0]; h[1] = flags[1];
    for (int k = 0; k < 5; k++) { h[2 + k] = flags[0] ? l.cst[k] : -1e300; h[7 + k] = flags[0] ? -l.cst[k] : -1e300; }
    MGC(cudaMemcpyAsync(mg->dscal, h, sizeof h, cudaMemcpyHostToDevice, s));
    TRY(comm_allreduce(mg->comm, mg->dscal, 12, 0xffcu, s));   // flags: sums; constants: max of c and of -c
    MGC(cudaMemcpyAsync(h, mg->dscal, sizeof h, cudaMemcpyDeviceToHost, s));
    MGC(cudaStreamSynchronize(s));
    int G = comm_size(mg->comm);
    flags[0] = (h[0] == G);
    flags[1] = (h[1] == G);
    for (int k = 0; k < 5; k++)
      if (h[2 + k] != -h[7 + k]) flags[0] = 0;    // max(c) != min(c): not uniform across ranks
  }
  if (flags[0]) l.mode = flags[1] ? 1 : 2;
  return F2D_OK;
}

int finish_setup(f2d_mg *mg, double Rd, cudaStream_t s) {
  if (Rd > 0.) {
    for (auto *vec : {&mg->S, &mg->L})
      for (auto &l : *vec) {
        if (vec == &mg->S && &l == &mg->S.back()) continue;   // S[lg] shares L[0]'s matrix
        k_helmholtz<<<nblocks1d(l.n()), 256, 0, s>>>(l.A + 4 * l.n(), 1. / (Rd * Rd), l.n());
        ++g_launches;
      }
  }
  for (auto *vec : {&mg->S, &mg->L})
    for (auto &l : *vec) {
      k_inverse_diagonal<<<cdiv(l.n(), 256), 256, 0, s>>>(l.A + 4 * l.n(), l.A + 5 * l.n(), mg->omega, l.n());
      ++g_launches;
    }
  MGC(set_smem_all());
  {
    // per-block partial sums of k_resid_sumsq (finest level of the handle: L[0], or S[0] on slabs)
    size_t nb = 0;
    for (auto *lp : {mg->L.empty() ? nullptr : &mg->L[0], mg->S.empty() ? nullptr : &mg->S[0]})
      if (lp) {
        nb = std::max(nb, (size_t)cdiv(lp->nx - 2 * NH, RST) * (size_t)cdiv(lp->ny - 2 * NH, RSR));
        nb = std::max(nb, (size_t)cdiv(lp->nx - 2 * NH, QTX) * (size_t)cdiv(lp->ny - 2 * NH, QTY));
      }
    MGC(cudaMalloc(&mg->partials, nb * sizeof(double)));
    mg->npartials = nb;
  }
  {
    int *dflag = nullptr;
    unsigned long long *didx = nullptr;
    MGC(cudaMalloc(&dflag, 2 * sizeof(int)));
    MGC(cudaMalloc(&didx, sizeof(unsigned long long)));
    for (int g = 0; g < mg->lg; g++) TRY(detect_mode(mg, mg->S[g], dflag, didx, s));
    for (auto &l : mg->L) TRY(detect_mode(mg, l, dflag, didx, s));
    cudaFree(dflag);
    cudaFree(didx);
    // the mask-free kernels also skip the coarse-mask tests of the transfers they fuse:
    // a level stays in class 1 only if the next coarser level is all fluid as well
    for (size_t lev = mg->L.size() - 1; lev-- > 0;)
      if (mg->L[lev].mode == 1 && mg->L[lev + 1].mode != 1) mg->L[lev].mode = 2;
    for (int g = mg->lg - 1; g >= 0; g--) {
      int coarse_mode = (g + 1 < mg->lg) ? mg->S[g + 1].mode : mg->L[0].mode;
      if (mg->S[g].mode == 1 && coarse_mode != 1) mg->S[g].mode = 2;
    }
  }
  if (const char *force = getenv("F2D_MG_FORCE_STORED"))
    if (force[0] == '1') {
      for (auto &l : mg->L) l.mode = 0;
      for (auto &l : mg->S) l.mode = 0;
    }
  if (mg->lg >= 0 && !mg->S.empty()) {   // slab view of L[0]: same class, no local y wrap
    Level &v = mg->S.back();
    v.mode = mg->L[0].mode == 1 ? 1 : 0;
    for (int k = 0; k < 5; k++) v.cst[k] = mg->L[0].cst[k];
  }
  // shared-memory tail: the deepest levels whose interior is at most 64 wide
  {
    int t0 = (int)mg->L.size();
    size_t cells = 0;
    while (t0 > 0) {
      Level &l = mg->L[t0 - 1];
      if (l.ny - 2 * NH > tail::MAXN || l.nx - 2 * NH > tail::MAXN) break;
      if ((int)mg->L.size() - (t0 - 1) > tail::MAXL) break;
      if (3 * (cells + l.n()) * sizeof(double) > 220 * 1024) break;
      cells += l.n();
      t0--;
    }
    const char *notail = getenv("F2D_MG_NO_TAIL");
    if (t0 < (int)mg->L.size() && !(notail && notail[0] == '1')) {
      mg->tail0 = t0;
      if (const char *nt = getenv("F2D_TAIL_NT")) {
        int v = atoi(nt);
        if (v >= 32 && v <= tail::NT && v % 32 == 0) mg->tail_nt = v;
      }
      mg->tail_smem = 3 * cells * sizeof(double);
      mg->tail_const = true;
      for (size_t lev = t0; lev < mg->L.size(); lev++)
        if (mg->L[lev].mode != 1) mg->tail_const = false;
      MGC(cudaFuncSetAttribute(tail::k_mg_tail<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)mg->tail_smem));
      MGC(cudaFuncSetAttribute(tail::k_mg_tail<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)mg->tail_smem));
    }
  }
  // cluster tail (default): the deepest levels whose interior is at most F2D_CTAIL_MAXN (128)
  // wide on a cluster of F2D_CTAIL_NC (16, else 8) CTAs; F2D_MG_NO_CTAIL=1 keeps the one-CTA tail
  {
    const char *notail = getenv("F2D_MG_NO_TAIL"), *noct = getenv("F2D_MG_NO_CTAIL");
    const bool off = (notail && notail[0] == '1') || (noct && noct[0] == '1');
    int maxn = 128, nc_want = ctail::MAXNC;   // measured on B200 (tools/trace_ctail.py): 16 CTAs from 128^2 down is the fastest
    if (const char *e = getenv("F2D_CTAIL_MAXN")) maxn = atoi(e);
    if (const char *e = getenv("F2D_CTAIL_NC")) nc_want = atoi(e);
    if (nc_want != 16 && nc_want != 8 && nc_want != 4 && nc_want != 2 && nc_want != 1) nc_want = ctail::MAXNC;
    long long mincells = 2048;   // smaller levels are replicated: a cluster barrier costs more than their sweeps
    if (const char *e = getenv("F2D_CTAIL_MINCELLS")) mincells = atoll(e);
    const size_t budget = 220 * 1024;
    // level table for the levels [t0, last] on nc CTAs; returns the bytes of shared memory
    auto plan = [&](int t0, int nc, ctail::Params &P) -> size_t {
      P = ctail::Params{};
      P.nlev = (int)mg->L.size() - t0;
      int o = 0, tmax = 0;
      for (int k = 0; k < P.nlev; k++) {
        Level &l = mg->L[t0 + k];
        const int m = l.ny - 2 * NH, n = l.nx - 2 * NH;
        ctail::Lev &v = P.lv[k];
        v.k = level_k(mg, t0 + k);
        v.dist = (nc > 1 && m % nc == 0 && m / nc >= ctail::MINROWS && (long long)m * n >= mincells) ? 1 : 0;
        v.R = v.dist ? m / nc : m;
        v.off = o;
        v.lgn = 0;
        while ((1 << v.lgn) < n) v.lgn++;
        int cells = (v.R + 2 * NH) * l.nx;
        cells += cells & 1;
        o += cells;
        tmax = std::max(tmax, cells);
      }
      P.total = o;
      P.tsize = tmax;
      P.ndeepest = mg->ndeepest;
      return (2 * (size_t)o + tmax) * sizeof(double);
    };
    if (!off && mg->L.size() >= 1) {
      for (int nc = nc_want; nc >= 1 && !mg->ctail; nc = (nc == 16 ? 8 : (nc > 1 ? 1 : 0))) {
        if (nc > 8) {
          cudaError_t e1 = cudaFuncSetAttribute(ctail::k_mg_ctail<false, false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
          cudaError_t e2 = cudaFuncSetAttribute(ctail::k_mg_ctail<true, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
          if (e1 != cudaSuccess || e2 != cudaSuccess) { cudaGetLastError(); continue; }
        }
        int t0 = (int)mg->L.size();
        ctail::Params P;
        size_t smem = 0;
        while (t0 > 0) {
          Level &l = mg->L[t0 - 1];
          if (l.ny - 2 * NH > maxn || l.nx - 2 * NH > maxn) break;
          if ((int)mg->L.size() - (t0 - 1) > ctail::MAXL) break;
          ctail::Params Q;
          size_t need = plan(t0 - 1, nc, Q);
          if (need > budget) break;
          t0--;
        }
        if (t0 >= (int)mg->L.size()) continue;
        smem = plan(t0, nc, P);
        const int nc_launch = P.lv[0].dist ? nc : 1;
        if (nc > 1 && nc_launch == 1) continue;   // nothing to distribute with this cluster size: try a smaller one
        cudaError_t e1 = cudaFuncSetAttribute(ctail::k_mg_ctail<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaError_t e2 = cudaFuncSetAttribute(ctail::k_mg_ctail<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e1 != cudaSuccess || e2 != cudaSuccess) { cudaGetLastError(); continue; }
        if (nc_launch > 1) {
          // can the device co-schedule such a cluster at all?
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3(nc_launch);
          cfg.blockDim = dim3(ctail::NT);
          cfg.dynamicSmemBytes = smem;
          cudaLaunchAttribute at[1];
          at[0].id = cudaLaunchAttributeClusterDimension;
          at[0].val.clusterDim.x = nc_launch;
          at[0].val.clusterDim.y = 1;
          at[0].val.clusterDim.z = 1;
          cfg.attrs = at;
          cfg.numAttrs = 1;
          int nclusters = 0;
          cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, ctail::k_mg_ctail<false, false>, &cfg);
          if (e != cudaSuccess || nclusters < 1) { cudaGetLastError(); continue; }
        }
        bool all_const = true;
        for (size_t lev = t0; lev < mg->L.size(); lev++)
          if (mg->L[lev].mode != 1) all_const = false;
        mg->ctail = true;
        mg->ctail_nc = nc_launch;
        mg->ctp = P;
        mg->tail0 = t0;
        mg->tail_smem = smem;
        mg->tail_const = all_const;
      }
    }
  }
  // all-fluid doubly periodic SQUARE tail (every tail level in the constant-stencil class, sizes
  // 2^k ... 8, 4): the interior-only template kernel (f2d_mg_ptail.cuh), from F2D_PTAIL_MAXN (128)
  // down; F2D_MG_NO_PTAIL=1 keeps the general cluster tail
  {
    const char *notail = getenv("F2D_MG_NO_TAIL"), *noct = getenv("F2D_MG_NO_CTAIL"), *nopt = getenv("F2D_MG_NO_PTAIL");
    const bool off = (notail && notail[0] == '1') || (noct && noct[0] == '1') || (nopt && nopt[0] == '1');
    int maxn = 128, nc_want = 16;
    if (const char *e = getenv("F2D_PTAIL_MAXN")) maxn = atoi(e);
    if (const char *e = getenv("F2D_CTAIL_NC")) nc_want = atoi(e);
    const int last = (int)mg->L.size() - 1;
    if (!off && last >= 0 && mg->L[last].ny - 2 * NH == 4 && mg->L[last].nx - 2 * NH == 4) {
      // how far up do the square constant-stencil levels go?
      int t0 = last + 1, top = ptail::LGMIN - 1;
      while (t0 > 0) {
        Level &l = mg->L[t0 - 1];
        const int n = l.nx - 2 * NH, want = 4 << (last - (t0 - 1));
        if (l.mode != 1 || l.ywrap != 1 || n != want || l.ny - 2 * NH != want || want > maxn || want > 256) break;
        t0--;
        top++;
      }
      for (int nc = nc_want; nc >= 1 && !mg->ptail && top >= ptail::LGMIN; nc = (nc >= 16 ? 8 : (nc > 1 ? 1 : 0))) {
        int tp = top;
        if (nc == 1 && tp > 6) tp = 6;                 // one CTA holds up to 64^2
        if (nc == 8 && tp > 7) tp = 7;
        if (nc > 1 && tp < 6) continue;                // nothing to distribute below 64^2
        ptail::Params P;
        size_t smem = 0;
        if (ptail_launch(tp, nc, P, 0, s, true, &smem) != F2D_OK) continue;
        const int t0p = last - (tp - ptail::LGMIN);
        if (mg->tail0 >= 0 && mg->tail0 < t0p) continue;   // a general tail above it is not supported
        mg->ptail = true;
        mg->ptail_top = tp;
        mg->ctail_nc = nc;
        mg->tail0 = t0p;
        mg->tail_smem = smem;
        mg->tail_const = true;
      }
    }
  }
  MGC(cudaStreamSynchronize(s));
  MGC(cudaGetLastError());
  return F2D_OK;
}

// Gridinfo, single rank (level.py:62-93): halve (m, n) until n <= 4 or m <= 4
int level_sizes(int m, int n, std::vector<std::pair<int, int>> &out) {
  int lev = 0;
  while (true) {
    if (lev > 0) { n /= 2; m /= 2; }
    out.push_back({m, n});
    lev++;
    if (n <= 4 || m <= 4) break;
    if (lev > 20) return fail(F2D_ERR_ARG, "mg_create: too many levels");
  }
  return F2D_OK;
}

int common_init(f2d_mg *mg) {
  MGC(cudaStreamCreateWithFlags(&mg->cap, cudaStreamNonBlocking));
  MGC(cudaMalloc(&mg->scratch, f2d_reduce_scratch_len() * sizeof(double)));
  MGC(cudaMalloc(&mg->dscal, 16 * sizeof(double)));
  MGC(cudaMallocHost(&mg->hscal, 16 * sizeof(double)));
  MGC(cudaStreamCreateWithFlags(&mg->cap2, cudaStreamNonBlocking));
  MGC(cudaMalloc(&mg->dstate, sizeof(f2d_mg::SolveState)));
  MGC(cudaMallocHost(&mg->hstate, sizeof(f2d_mg::SolveState)));
  return F2D_OK;
}

}  // namespace

extern "C" int f2d_mg_destroy(f2d_mg_t *mg);
extern "C" int f2d_mg_set_relaxation(f2d_mg_t *mg, int mode);

extern "C" int f2d_mg_create(f2d_mg_t **out, const double *cornermask, int ny, int nx, double dx, double dy,
                             double omega, double hydroepsilon, double Rd, f2d_stream_t stream) {
  if (!out || !cornermask) return fail(F2D_ERR_ARG, "mg_create: null pointer");
  int m = ny - 2 * NH, n = nx - 2 * NH;
  if (m < 4 || n < 4) return fail(F2D_ERR_ARG, "mg_create: grid too small");
  if ((m & (m - 1)) || (n & (n - 1))) return fail(F2D_ERR_ARG, "mg_create: nx, ny must be powers of two");
  cudaStream_t s = S(stream);
  f2d_mg *mg = new f2d_mg();
  mg->omega = omega;
  if (const char *ng = getenv("F2D_MG_NO_GRAPHS")) mg->graphs = !(ng[0] == '1');
  if (const char *tm = getenv("F2D_MG_TMA")) mg->tma = !(tm[0] == '0');
  auto bail = [&](int rc) { f2d_mg_destroy(mg); return rc; };
  std::vector<std::pair<int, int>> sizes;
  int rc = level_sizes(m, n, sizes);
  if (rc != F2D_OK) return bail(rc);
  for (auto &sz : sizes) {
    Level l;
    l.ny = sz.first + 2 * NH;
    l.nx = sz.second + 2 * NH;
    mg->L.push_back(l);
  }
  if ((rc = common_init(mg)) != F2D_OK) return bail(rc);
  for (auto &l : mg->L)
    if ((rc = alloc_level(mg, l, false, s)) != F2D_OK) return bail(rc);
  Level &l0 = mg->L[0];
  double *A9 = nullptr;
  if (cudaMalloc(&A9, 9 * l0.n() * sizeof(double)) != cudaSuccess) return bail(fail(F2D_ERR_CUDA, "mg_create: out of memory"));
  dim3 blk(32, 8);
  k_mask_from_double<<<nblocks1d(l0.n()), 256, 0, s>>>(cornermask, l0.msk, l0.n());
  k_finest_matrix<<<grid2d(l0.ny, l0.nx, blk), blk, 0, s>>>(l0.msk, A9, finest_stencil(dx, dy, hydroepsilon), l0.ny, l0.nx);
  g_launches += 2;
  for (int k = 0; k < 9; k++)
    if ((rc = f2d_fill_halo(A9 + k * l0.n(), NH, l0.ny, l0.nx, stream)) != F2D_OK) { cudaFree(A9); return bail(rc); }
  if ((rc = build_replicated(mg, A9, true, s)) != F2D_OK) return bail(rc);
  if ((rc = finish_setup(mg, Rd, s)) != F2D_OK) return bail(rc);
  // level.py:153-163: flat cells take the line relaxation
  if (hydroepsilon * dy / dx <= 0.2 && (rc = f2d_mg_set_relaxation(mg, 1)) != F2D_OK) return bail(rc);
  *out = mg;
  return F2D_OK;
}

#include "f2d_mg_slab.cuh"

namespace {
// norms of the rhs and of the first residual -> dscal[0..1] (all-reduced on slabs)
int solve_pre(f2d_mg *mg, double *psi, const double *rhs, cudaStream_t s) {
  Level &l = mg->comm ? mg->S[0] : mg->L[0];
  if (mg->comm && !comm_owns(mg->comm, psi))
    return fail(F2D_ERR_ARG, "slab multigrid: psi must live in the symmetric heap (f2d_comm_alloc)");
  TRY(f2d_computenorm(l.msk, rhs, NH, l.ny, l.nx, mg->dscal, mg->scratch, (f2d_stream_t)s));
  TRY(op_resid_sumsq_L(mg, l, psi, rhs, l.b, mg->dscal + 1, s));
  if (mg->comm) {
    TRY(xch1(mg, l, l.b, s));
    TRY(comm_allreduce(mg->comm, mg->dscal, 2, 0u, s));
  }
  return F2D_OK;
}
// one iteration: psi += Fcycle(r); r = rhs - A psi; |r|^2 -> dscal[1]
int solve_body(f2d_mg *mg, double *psi, const double *rhs, cudaStream_t s) {
  Level &l = mg->comm ? mg->S[0] : mg->L[0];
  if (mg->comm) {
    TRY(slab_cycle_enqueue(mg, 1, 0, l.x, l.b, s, psi));
  } else if (mg->L.size() > 1) {
    TRY(fcycle_enqueue(mg, 0, l.x, l.b, s, psi));
  } else {
    TRY(fcycle_enqueue(mg, 0, l.x, l.b, s, nullptr));
    k_add_inplace<<<nblocks1d(l.n()), 256, 0, s>>>(psi, l.x, l.n());
    F2D_LAUNCHED();
  }
  TRY(op_resid_sumsq_L(mg, l, psi, rhs, l.b, mg->dscal + 1, s));
  if (mg->comm) {
    TRY(xch1(mg, l, l.b, s));
    TRY(comm_allreduce(mg->comm, mg->dscal + 1, 1, 0u, s));
  }
  return F2D_OK;
}
}  // namespace

extern "C" int f2d_mg_destroy(f2d_mg_t *mg) {
  if (!mg) return F2D_OK;
  for (auto &kv : mg->cache) cudaGraphExecDestroy(kv.second.exec);
  for (auto &kv : mg->solve_cache) cudaGraphExecDestroy(kv.second.exec);
  if (mg->cap2) cudaStreamDestroy(mg->cap2);
  cudaFree(mg->dstate);
  if (mg->hstate) cudaFreeHost(mg->hstate);
  for (auto &l : mg->L) free_level(l);
  for (auto &l : mg->S) free_level(l);
  cudaFree(mg->scratch);
  cudaFree(mg->partials);
  cudaFree(mg->dscal);
  if (mg->hscal) cudaFreeHost(mg->hscal);
  if (mg->cap) cudaStreamDestroy(mg->cap);
  delete mg;
  return F2D_OK;
}

// global level g: slab level (g < lg) or level g - lg of the replicated hierarchy
static Level *level_at(const f2d_mg_t *mg, int g) {
  f2d_mg *m = const_cast<f2d_mg *>(mg);
  if (!m || g < 0 || g >= m->lg + (int)m->L.size()) return nullptr;
  return g < m->lg ? &m->S[g] : &m->L[g - m->lg];
}
extern "C" int f2d_mg_nlevels(const f2d_mg_t *mg) { return mg ? mg->lg + (int)mg->L.size() : 0; }
extern "C" int f2d_mg_level_shape(const f2d_mg_t *mg, int lev, int *ny, int *nx) {
  Level *l = level_at(mg, lev);
  if (!l) return fail(F2D_ERR_ARG, "mg_level_shape: bad level");
  *ny = l->ny;
  *nx = l->nx;
  return F2D_OK;
}
extern "C" void *f2d_mg_level_ptr(f2d_mg_t *mg, int lev, int which) {
  Level *lp = level_at(mg, lev);
  if (!lp) return nullptr;
  Level &l = *lp;
  switch (which) {
    case 0: return l.msk;
    case 1: return l.A;
    case 2: return l.x;
    case 3: return l.b;
    case 4: return l.r;
  }
  return nullptr;
}
extern "C" int f2d_mg_level_matrix_mode(const f2d_mg_t *mg, int lev) {
  Level *l = level_at(mg, lev);
  return l ? l->mode : -1;
}
/* number of distributed (slab) levels of a handle made by f2d_mg_create_slab */
extern "C" int f2d_mg_slab_levels(const f2d_mg_t *mg) { return mg ? mg->lg : 0; }
extern "C" int f2d_mg_set_relaxation(f2d_mg_t *mg, int mode) {
  if (!mg || (mode != 0 && mode != 1)) return fail(F2D_ERR_ARG, "mg_set_relaxation: bad handle / mode");
  if (mode == mg->relax) return F2D_OK;
  if (mode == 1) {
    if (mg->comm) return fail(F2D_ERR_ARG, "mg_set_relaxation: the line relaxation needs npy = 1 (level.py:155-157)");
    size_t need = 0;
    for (auto &l : mg->L) need = std::max(need, 4 * (size_t)l.ny * sizeof(double));
    if (need > 200 * 1024) return fail(F2D_ERR_ARG, "mg_set_relaxation: columns longer than 6400 rows do not fit the line kernel");
    if (need > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(k_smooth_tridiag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need);
      if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_smooth_tridiag)");
    }
    // the shared-memory tail kernels and the constant-stencil classes belong to the Jacobi
    // path: every level goes through the per-operator kernels on the stored coefficients
    mg->tail0 = -1;
    mg->ctail = false;
    mg->ptail = false;
    for (auto &l : mg->L) l.mode = 0;
  } else {
    return fail(F2D_ERR_ARG, "mg_set_relaxation: a hierarchy switched to the line relaxation cannot be switched back");
  }
  mg->relax = mode;
  for (auto &kv : mg->cache) cudaGraphExecDestroy(kv.second.exec);   // graphs captured the Jacobi kernels
  mg->cache.clear();
  for (auto &kv : mg->solve_cache) cudaGraphExecDestroy(kv.second.exec);
  mg->solve_cache.clear();
  return F2D_OK;
}
extern "C" int f2d_mg_set_graphs(f2d_mg_t *mg, int enable) {
  if (!mg) return fail(F2D_ERR_ARG, "mg_set_graphs: null");
  mg->graphs = enable != 0;
  return F2D_OK;
}

#define CHECK_LEV(mg, lev, name)                                                       \
  if (!(mg) || (lev) < 0 || (lev) >= (int)(mg)->L.size()) return fail(F2D_ERR_ARG, name ": bad handle/level"); \
  if ((mg)->comm) return fail(F2D_ERR_ARG, name ": the per-level entry points are not available on a slab hierarchy")

extern "C" int f2d_mg_smooth(f2d_mg_t *mg, int lev, double *x, const double *b, int nite, f2d_stream_t s) {
  CHECK_LEV(mg, lev, "mg_smooth");
  if (!x || !b) return fail(F2D_ERR_ARG, "mg_smooth: null");
  return op_smooth(mg, lev, x, b, nite, S(s));
}
extern "C" int f2d_mg_residual(f2d_mg_t *mg, int lev, const double *x, const double *b, double *r, f2d_stream_t s) {
  CHECK_LEV(mg, lev, "mg_residual");
  if (!x || !b || !r) return fail(F2D_ERR_ARG, "mg_residual: null");
  return op_residual(mg, lev, x, b, r, S(s));
}
extern "C" int f2d_mg_restrict(f2d_mg_t *mg, int lev, const double *xf, double *xc, f2d_stream_t s) {
  CHECK_LEV(mg, lev + 1, "mg_restrict");
  if (lev < 0 || !xf || !xc) return fail(F2D_ERR_ARG, "mg_restrict: bad args");
  return op_restrict(mg, lev, xf, xc, S(s));
}
extern "C" int f2d_mg_interpolate(f2d_mg_t *mg, int lev, const double *xc, double *xf, int add, f2d_stream_t s) {
  CHECK_LEV(mg, lev + 1, "mg_interpolate");
  if (lev < 0 || !xf || !xc) return fail(F2D_ERR_ARG, "mg_interpolate: bad args");
  return op_interpolate(mg, lev, xc, xf, add, S(s));
}
extern "C" int f2d_mg_sumsq(f2d_mg_t *mg, int lev, const double *x, double *out, f2d_stream_t s) {
  CHECK_LEV(mg, lev, "mg_sumsq");
  Level &l = mg->L[lev];
  return f2d_computenorm(l.msk, x, NH, l.ny, l.nx, out, mg->scratch, s);
}
extern "C" int f2d_mg_vcycle(f2d_mg_t *mg, int lev1, f2d_stream_t s) {
  CHECK_LEV(mg, lev1, "mg_vcycle");
  return run_cycle(mg, 2, lev1, mg->L[lev1].x, mg->L[lev1].b, S(s));
}
extern "C" int f2d_mg_fcycle(f2d_mg_t *mg, int lev1, f2d_stream_t s) {
  CHECK_LEV(mg, lev1, "mg_fcycle");
  return run_cycle(mg, 1, lev1, mg->L[lev1].x, mg->L[lev1].b, S(s));
}
// hierarchy.py:207-218.  The residual computed before each V-cycle (:215) is dead (the
// V-cycle overwrites r[0] before reading it) and is skipped; psi/rhs are used in place
// of the copies x[0], b[0].
extern "C" int f2d_mg_two_vcycle(f2d_mg_t *mg, double *psi, const double *rhs, f2d_stream_t s) {
  if (!mg || !psi || !rhs) return fail(F2D_ERR_ARG, "mg_two_vcycle: null");
  return run_cycle(mg, 0, 0, psi, const_cast<double *>(rhs), S(s));
}
// hierarchy.py:154-192
extern "C" int f2d_mg_solve(f2d_mg_t *mg, double *psi, const double *rhs, double tol, int maxite, int *nite_out,
                            double *res_out, f2d_stream_t stream) {
  if (!mg || !psi || !rhs) return fail(F2D_ERR_ARG, "mg_solve: null");
  cudaStream_t s = S(stream);
  // (line relaxation: the host-driven loop below; its F-cycle is still one captured graph)
  if (mg->graphs && !(mg->comm && mg->lg == 0) && mg->relax == 0)
    return solve_graph(mg, psi, rhs, tol, maxite, nite_out, res_out, s);
  // host-driven loop (graphs disabled): same operations, one host round trip per F-cycle
  if (mg->comm) return slab_solve(mg, psi, rhs, tol, maxite, nite_out, res_out, s);
  Level &l = mg->L[0];
  TRY(f2d_computenorm(l.msk, rhs, NH, l.ny, l.nx, mg->dscal, mg->scratch, stream));
  TRY(op_resid_sumsq(mg, psi, rhs, l.b, mg->dscal + 1, s));
  TRY(read_scalars(mg, 2, s));
  double normb = sqrt(mg->hscal[0]);
  int nite = 0;
  double res = 0.;
  if (normb > 0) {
    double res0 = sqrt(mg->hscal[1]) / normb;
    res = res0;
    int ndiv = 0;
    const bool fuse_add = mg->L.size() > 1;   // `x += self.x[0]` done by the F-cycle's last kernel
    while (nite < maxite && res0 > tol) {
      TRY(run_cycle(mg, 1, 0, l.x, l.b, s, fuse_add ? psi : nullptr));
      if (!fuse_add) {
        k_add_inplace<<<nblocks1d(l.n()), 256, 0, s>>>(psi, l.x, l.n());
        F2D_LAUNCHED();
      }
      TRY(op_resid_sumsq(mg, psi, rhs, l.b, mg->dscal + 1, s));
      TRY(read_scalars(mg, 2, s));
      res = sqrt(mg->hscal[1]) / normb;
      double conv = res0 / res;
      res0 = res;
      nite++;
      if (conv < 1) ndiv++;
      if (ndiv > 4) return fail(F2D_ERR_DIVERGE, "solver is not converging");
    }
  }
  if (nite_out) *nite_out = nite;
  if (res_out) *res_out = res;
  return F2D_OK;
}

// bench.py: one operator of the cycles, `reps` launches back to back on the level's own arrays
// (the kernels the graphs replay, timed in isolation with CUDA events around the batch)
extern "C" int f2d_mg_bench_op(f2d_mg_t *mg, int kind, int lev, int reps, f2d_stream_t stream) {
  CHECK_LEV(mg, lev, "mg_bench_op");
  if (reps < 1) return fail(F2D_ERR_ARG, "mg_bench_op: reps");
  if (mg->relax != 0) return fail(F2D_ERR_ARG, "mg_bench_op: Jacobi hierarchies only");
  cudaStream_t s = S(stream);
  Level &l = mg->L[lev];
  const bool has_coarse = lev + 1 < (int)mg->L.size();
  for (int k = 0; k < reps; k++) {
    switch (kind) {
      case 0: TRY(smooth2(mg, lev, 0, l.x, l.b, l.t, nullptr, s)); break;
      case 1: TRY(smooth2(mg, lev, 1, l.x, l.b, l.t, nullptr, s)); break;
      case 2:
        if (!has_coarse) return fail(F2D_ERR_ARG, "mg_bench_op: no coarser level");
        TRY(smooth2(mg, lev, 2, l.x, l.b, l.x, mg->L[lev + 1].x, s));
        break;
      case 3:
        if (!has_coarse) return fail(F2D_ERR_ARG, "mg_bench_op: no coarser level");
        TRY(smooth2(mg, lev, 3, l.t, l.b, l.x, mg->L[lev + 1].x, s));
        break;
      case 4:
        if (!has_coarse) return fail(F2D_ERR_ARG, "mg_bench_op: no coarser level");
        TRY(op_resid_restrict(mg, lev, l.t, l.b, mg->L[lev + 1].b, s));
        break;
      case 5:
        if (!has_coarse) return fail(F2D_ERR_ARG, "mg_bench_op: no coarser level");
        TRY(op_restrict(mg, lev, l.b, mg->L[lev + 1].b, s));
        break;
      case 6:
        if (lev != 0) return fail(F2D_ERR_ARG, "mg_bench_op: the residual norm kernel belongs to level 0");
        TRY(op_resid_sumsq(mg, l.x, l.b, l.r, mg->dscal + 3, s));
        break;
      case 7:
        if (lev != mg->tail0) return fail(F2D_ERR_ARG, "mg_bench_op: not the first level of the tail");
        TRY(tail_launch(mg, 0, l.b, nullptr, l.x, s));
        break;
      case 8:
        if (lev != mg->tail0) return fail(F2D_ERR_ARG, "mg_bench_op: not the first level of the tail");
        TRY(tail_launch(mg, 2, l.b, nullptr, l.x, s));
        break;
      case 9:
        if (!has_coarse || !zrr_ok(mg, l, mg->L[lev + 1])) return fail(F2D_ERR_ARG, "mg_bench_op: no fused descent on this level");
        TRY(op_zsmooth_rr_L(mg, l, mg->L[lev + 1], l.b, l.t, mg->L[lev + 1].b, s));
        break;
      default: return fail(F2D_ERR_ARG, "mg_bench_op: kind 0..9");
    }
  }
  return F2D_OK;
}
/* first level handled by the shared-memory tail kernel (-1: none) */
extern "C" int f2d_mg_tail_level(const f2d_mg_t *mg) { return mg ? mg->tail0 : -1; }

// operators.py:421-498
extern "C" int f2d_mg_set_trace(f2d_mg_t *mg, long long *buf, int cap) {
  if (!mg) return fail(F2D_ERR_ARG, "mg_set_trace: null handle");
  mg->trace = cap > 1 ? buf : nullptr;
  mg->trace_cap = cap;
  for (auto &kv : mg->cache) cudaGraphExecDestroy(kv.second.exec);   // graphs captured the old pointer
  mg->cache.clear();
  for (auto &kv : mg->solve_cache) cudaGraphExecDestroy(kv.second.exec);
  mg->solve_cache.clear();
  return F2D_OK;
}

extern "C" int f2d_mg_set_uv_stage(f2d_mg_t *mg, const double *ub, const double *vb, const double *ue,
                                   const double *ve, double *uo, double *vo, double c) {
  if (!mg || !ub || !vb || !uo || !vo) return fail(F2D_ERR_ARG, "mg_set_uv_stage: null");
  if ((ue == nullptr) != (ve == nullptr)) return fail(F2D_ERR_ARG, "mg_set_uv_stage: ue and ve go together");
  mg->uvs.on = true;
  mg->uvs.ub = ub; mg->uvs.vb = vb; mg->uvs.ue = ue; mg->uvs.ve = ve;
  mg->uvs.uo = uo; mg->uvs.vo = vo; mg->uvs.c = c;
  return F2D_OK;
}

extern "C" int f2d_invert_vorticity(f2d_mg_t *mg, const int8_t *msk, const int8_t *mskp, const double *w,
                                    double *psi, double *u, double *v, double *work, const double *rhsp,
                                    const double *psi_island, int full, int perio, double area, double dx,
                                    double dy, int nh, int *nite, double *res, double *scratch,
                                    f2d_stream_t stream) {
  if (!mg || !w || !psi || !u || !v || !work) return fail(F2D_ERR_ARG, "invert_vorticity: null");
  if ((msk == nullptr) != (mskp == nullptr) || (!msk && psi_island))
    return fail(F2D_ERR_ARG, "invert_vorticity: msk and mskp may only be omitted together (all-fluid domain, no island)");
  if (nh != NH) return fail(F2D_ERR_NH, "invert_vorticity: nh must be 3");
  Level &l = mg->comm ? mg->S[0] : mg->L[0];
  size_t n = l.n();
  TRY(f2d_celltocorner(w, work, l.ny, l.nx, stream));
  if (rhsp) TRY(f2d_add_scaled(work, -1., rhsp, n, stream));
  if (full) {
    TRY(f2d_mg_solve(mg, psi, work, 1e-11, 4, nite, res, stream));
    if (perio) {
      if (!scratch) return fail(F2D_ERR_ARG, "invert_vorticity: scratch needed");
      TRY(f2d_domain_sum(psi, NH, l.ny, l.nx, mg->dscal + 2, scratch, stream));
      if (mg->comm) TRY(comm_allreduce(mg->comm, mg->dscal + 2, 1, 0u, S(stream)));   // area is the global one
      TRY(f2d_sub_devscalar(psi, mg->dscal + 2, area, n, stream));
    }
  } else {
    TRY(f2d_mg_two_vcycle(mg, psi, work, stream));
    if (nite) *nite = 1;
    if (res) *res = 0.;
  }
  if (mg->uvs.on) {
    const f2d_mg::UVStage R = mg->uvs;
    mg->uvs.on = false;
    if (!psi_island)
      return f2d_mask_orthogradient_stage(msk, mskp, psi, dx, dy, nh, u, v, R.ub, R.vb, R.ue, R.ve, R.uo, R.vo, R.c,
                                          l.ny, l.nx, stream);
    TRY(f2d_mul_mask(psi, mskp, n, stream));
    TRY(f2d_add_scaled(psi, 1., psi_island, n, stream));
    TRY(f2d_orthogradient(msk, psi, dx, dy, nh, u, v, l.ny, l.nx, stream));
    return f2d::uv_stage(u, v, R.ub, R.vb, R.ue, R.ve, R.uo, R.vo, R.c, n, stream);
  }
  if (!psi_island) return f2d_mask_orthogradient(msk, mskp, psi, dx, dy, nh, u, v, l.ny, l.nx, stream);
  TRY(f2d_mul_mask(psi, mskp, n, stream));
  TRY(f2d_add_scaled(psi, 1., psi_island, n, stream));
  return f2d_orthogradient(msk, psi, dx, dy, nh, u, v, l.ny, l.nx, stream);
}
