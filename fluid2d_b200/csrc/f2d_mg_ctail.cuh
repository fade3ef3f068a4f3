// Cluster coarse-tail multigrid kernel (included by f2d_multigrid.cu).
//
// Same job as f2d_mg_tail.cuh -- one launch runs a whole V-cycle (hierarchy.py:98-127) or
// F-cycle (hierarchy.py:131-151) of the sub-hierarchy of coarse levels with every array
// resident in shared memory -- but on a thread-block CLUSTER of NC CTAs, so that the
// sub-hierarchy can start at 128^2 instead of 64^2 and each operator application of the
// larger levels is shared by NC SMs:
//
//   * "distributed" levels (more than REPL_CELLS cells): x, b and the scratch t are split
//     in row bands, band c in the shared memory of CTA c.  A CTA computes the rows it owns;
//     the rows of its neighbours that the 9-point stencils, the restriction and the
//     interpolation read, and the periodic halo images it must store, are reached through
//     distributed shared memory (generic pointers from cluster.map_shared_rank kept in a
//     per-CTA table of row addresses).  One cluster barrier per operator application.
//   * "replicated" levels (the smallest ones, where latency rules): every CTA keeps the
//     whole level and computes all of it, redundantly and identically, with
//     __syncthreads() only -- no cluster traffic on the latency-critical bottom of the cycle.
//
// Arithmetic: the same expressions, in the same order, as the Fortran kernels (through
// fused::jacobi_val / resid_val / interp_w*), so results are bit-identical to the
// one-CTA tail and to the per-level kernels.
#pragma once
#include <cooperative_groups.h>

namespace ctail {

namespace cg = cooperative_groups;

constexpr int NH = 3;
constexpr int NT = 1024;
constexpr int MAXL = 10;
constexpr int NC = 8;            // CTAs per cluster (portable maximum)
constexpr int MAXN = 128;        // largest interior size handled
constexpr int REPL_CELLS = 600;  // levels with at most this many cells (halo included) are replicated

struct Params {
  int nlev;
  fused::LevelK lv[MAXL];   // geometry / matrix of each level (index 0 = finest of the tail)
  int dist[MAXL];           // 1: rows distributed over the CTAs; 0: replicated in every CTA
  int rows[MAXL];           // band height of a distributed level
  int off[MAXL];            // offset (doubles) of the level inside each local array
  int rp[MAXL];             // offset of the level inside each row-address table
  int total, rptotal;       // doubles per local array, entries per row-address table
  int ndeepest;
  const double *b_in;       // global rhs of the finest tail level
  const double *x_in;       // global first guess (program 1) or nullptr
  double *x_out;            // global result of the finest tail level
  double *acc;              // if set: acc += result instead of storing it (solve(), hierarchy.py:171)
  long long *trace;         // diagnostics (f2d_mg_set_trace): trace[0] = count, then one clock64() per barrier
  int trace_cap;
};

struct Ctx {
  double *A[3];             // this CTA's X, B, T arrays (all levels concatenated)
  double **RP[3];           // row-address tables of the distributed levels
  int rank;
  int *ntrace;              // shared-memory counter of the trace stamps
};

// one array of one level: row(j) is the address of element (j, 0), wherever it lives
struct View {
  double *base;
  double *const *rp;
  int nx, dist;
  __device__ __forceinline__ double *row(int j) const { return dist ? rp[j] : base + j * nx; }
};
enum { AX = 0, AB = 1, AT = 2 };
__device__ __forceinline__ View view(const Params &P, const Ctx &C, int a, int l) {
  View v;
  v.base = C.A[a] + P.off[l];
  v.rp = C.RP[a] + P.rp[l];
  v.nx = P.lv[l].nx;
  v.dist = P.dist[l];
  return v;
}

// barrier after an operator application that touched distributed arrays of the cluster
__device__ __forceinline__ void sync(const Params &P, const Ctx &C, bool cluster_wide) {
  if (cluster_wide) cg::this_cluster().sync();
  else __syncthreads();
  if (P.trace && C.rank == 0 && threadIdx.x == 0) {   // stores only: nothing waits on global memory
    const int k = ++*C.ntrace;
    if (k < P.trace_cap) { P.trace[k] = clock64(); P.trace[0] = k; }
  }
}

// cells (j, i) of [jlo, jhi] x [ilo, ihi] that this CTA computes: its own rows of a
// distributed level, every row of a replicated one
template <class F>
__device__ __forceinline__ void for_points(const Params &P, const Ctx &C, int l, int jlo, int jhi, int ilo, int ihi,
                                           F f) {
  if (P.dist[l]) {
    const int r0 = C.rank * P.rows[l];
    jlo = max(jlo, r0);
    jhi = min(jhi, r0 + P.rows[l] - 1);
  }
  const int w = ihi - ilo + 1, h = jhi - jlo + 1;
  if (w <= 0 || h <= 0) return;
  const int n = w * h;
  for (int p = threadIdx.x; p < n; p += NT) {
    const int dj = p / w;
    f(jlo + dj, ilo + (p - dj * w));
  }
}

template <bool MASKED, bool STORED>
__device__ __forceinline__ double jacobi_at(const fused::LevelK &L, const fused::Coefs<MASKED, STORED> &kc,
                                            const View &s, const View &b, int j, int i) {
  const int nx = L.nx;
  const size_t g = (size_t)j * nx + i;
  if (MASKED && L.msk[g] == 0) return 0.;
  fused::Coefs<MASKED, STORED> k;
  if (MASKED || STORED) k.load(L, g, MASKED ? L.msk + g : nullptr, nx); else k = kc;
  const double *lo = s.row(j - 1) + i, *mid = s.row(j) + i, *hi = s.row(j + 1) + i;
  return fused::jacobi_val<MASKED, STORED>(L, k, lo[-1], lo[0], lo[1], mid[-1], mid[0], mid[1], hi[-1], hi[0], hi[1],
                                           b.row(j)[i]);
}

// two damped-Jacobi sweeps + halo fill, x in place (scratch t)
template <bool MASKED, bool STORED>
__device__ void smooth2(const Params &P, const Ctx &C, int l) {
  const fused::LevelK &L = P.lv[l];
  const int ny = L.ny, nx = L.nx;
  const View x = view(P, C, AX, l), b = view(P, C, AB, l), t = view(P, C, AT, l);
  fused::Coefs<MASKED, STORED> kc;
  if (!MASKED && !STORED) kc.load(L, 0, nullptr, 0);
  for_points(P, C, l, 2, ny - 3, 2, nx - 3,
             [&](int j, int i) { t.row(j)[i] = jacobi_at<MASKED, STORED>(L, kc, x, b, j, i); });
  sync(P, C, P.dist[l]);
  for_points(P, C, l, NH, ny - 1 - NH, NH, nx - 1 - NH, [&](int j, int i) {
    const double val = jacobi_at<MASKED, STORED>(L, kc, t, b, j, i);
    x.row(j)[i] = val;
    f2d::for_each_halo_image(j, i, ny, nx, NH, [&](int jj, int ii) { x.row(jj)[ii] = val; });
  });
  sync(P, C, P.dist[l]);
}

// t = b - A x on the interior + halo fill
template <bool MASKED, bool STORED>
__device__ void residual(const Params &P, const Ctx &C, int l) {
  const fused::LevelK &L = P.lv[l];
  const int ny = L.ny, nx = L.nx;
  const View x = view(P, C, AX, l), b = view(P, C, AB, l), r = view(P, C, AT, l);
  fused::Coefs<MASKED, STORED> kc;
  if (!MASKED && !STORED) kc.load(L, 0, nullptr, 0);
  for_points(P, C, l, NH, ny - 1 - NH, NH, nx - 1 - NH, [&](int j, int i) {
    const size_t g = (size_t)j * nx + i;
    double val = 0.;
    if (!MASKED || L.msk[g] != 0) {
      fused::Coefs<MASKED, STORED> k;
      if (MASKED || STORED) k.load(L, g, MASKED ? L.msk + g : nullptr, nx); else k = kc;
      const double cdiag = STORED ? L.A[4 * (size_t)ny * nx + g] : L.c[4];
      const double *lo = x.row(j - 1) + i, *mid = x.row(j) + i, *hi = x.row(j + 1) + i;
      val = fused::resid_val<MASKED, STORED>(L, k, cdiag, lo[-1], lo[0], lo[1], mid[-1], mid[0], mid[1], hi[-1],
                                             hi[0], hi[1], b.row(j)[i]);
    }
    r.row(j)[i] = val;
    f2d::for_each_halo_image(j, i, ny, nx, NH, [&](int jj, int ii) { r.row(jj)[ii] = val; });
  });
  sync(P, C, P.dist[l]);
}

// full-weighting restriction of array `af` of level l into B of level l+1 (interior + halo fill)
template <bool MASKED>
__device__ void restrict_to(const Params &P, const Ctx &C, int l, int af) {
  const fused::LevelK &Lc = P.lv[l + 1];
  const int ny = Lc.ny, nx = Lc.nx;
  const View xf = view(P, C, af, l), xc = view(P, C, AB, l + 1);
  for_points(P, C, l + 1, NH, ny - 1 - NH, NH, nx - 1 - NH, [&](int j, int i) {
    double val = 0.;
    if (!MASKED || Lc.msk[j * nx + i] != 0) {
      const int fi = 2 * i - 2;
      const double *lo = xf.row(2 * j - 3) + fi, *mid = xf.row(2 * j - 2) + fi, *hi = xf.row(2 * j - 1) + fi;
      val = 0.25 * mid[0] + 0.125 * (((mid[-1] + mid[1]) + lo[0]) + hi[0]) +
            0.0625 * (((lo[-1] + lo[1]) + hi[-1]) + hi[1]);
    }
    xc.row(j)[i] = val;
    f2d::for_each_halo_image(j, i, ny, nx, NH, [&](int jj, int ii) { xc.row(jj)[ii] = val; });
  });
  sync(P, C, P.dist[l] || P.dist[l + 1]);
}

// X(l) = [X(l) +] I(X(l+1)) over the whole fine array
template <bool MASKED>
__device__ void interpolate(const Params &P, const Ctx &C, int l, bool add) {
  const fused::LevelK &Lf = P.lv[l], &Lc = P.lv[l + 1];
  const int ny = Lf.ny, nx = Lf.nx, nxc = Lc.nx;
  const View xc = view(P, C, AX, l + 1), xf = view(P, C, AX, l);
  for_points(P, C, l, 0, ny - 1, 0, nx - 1, [&](int j, int i) {
    double iv = 0.;
    if (!MASKED || Lf.msk[j * nx + i] > 0) {
      const int jc = (j >> 1) + 1, ic = (i >> 1) + 1;
      const int k = jc * nxc + ic;
      const int pj = j & 1, pi = i & 1;
      const int8_t *mc = Lc.msk;
      const double *c0 = xc.row(jc) + ic;
      if (!pj && !pi) {
        iv = c0[0];
      } else if (!pj) {
        const int s = MASKED ? mc[k] + mc[k + 1] : 2;
        iv = (c0[0] + c0[1]) * fused::interp_w2(s);
      } else if (!pi) {
        const int s = MASKED ? mc[k] + mc[k + nxc] : 2;
        iv = (c0[0] + xc.row(jc + 1)[ic]) * fused::interp_w2(s);
      } else {
        const int s = MASKED ? mc[k] + mc[k + 1] + mc[k + nxc] + mc[k + nxc + 1] : 4;
        const double *c1 = xc.row(jc + 1) + ic;
        iv = fused::interp_w4(s) * (((c0[0] + c0[1]) + c1[0]) + c1[1]);
      }
    }
    double *o = xf.row(j) + i;
    *o = add ? *o + iv : iv;
  });
  sync(P, C, P.dist[l]);
}

__device__ __forceinline__ void fill_zero(const Params &P, const Ctx &C, int l) {
  const View x = view(P, C, AX, l);
  for_points(P, C, l, 0, P.lv[l].ny - 1, 0, P.lv[l].nx - 1, [&](int j, int i) { x.row(j)[i] = 0.; });
  sync(P, C, P.dist[l]);
}

// deepest level: x = 0, then ndeepest double sweeps (hierarchy.py:114-116)
template <bool MASKED, bool STORED>
__device__ void coarsest(const Params &P, const Ctx &C) {
  const int last = P.nlev - 1;
  const fused::LevelK &L = P.lv[last];
  if (!MASKED && !STORED && !P.dist[last] && tail::coarsest_periodic_ok(L)) {
    // one warp, periodic indexing on the m x n unknowns (see f2d_mg_tail.cuh)
    tail::coarsest_periodic(L, C.A[AX] + P.off[last], C.A[AB] + P.off[last], C.A[AT] + P.off[last], P.ndeepest);
    return;
  }
  fill_zero(P, C, last);
  for (int k = 0; k < P.ndeepest; k++) smooth2<MASKED, STORED>(P, C, last);
}

// V-cycle of the levels [l1, nlev-1]; X of level l1 is whatever the arrays hold
template <bool MASKED, bool STORED>
__device__ void vcycle(const Params &P, const Ctx &C, int l1) {
  const int last = P.nlev - 1;
  for (int l = l1; l < last; l++) {
    if (l > l1) fill_zero(P, C, l);
    smooth2<MASKED, STORED>(P, C, l);
    residual<MASKED, STORED>(P, C, l);
    restrict_to<MASKED>(P, C, l, AT);
  }
  coarsest<MASKED, STORED>(P, C);
  for (int l = last - 1; l >= l1; l--) {
    interpolate<MASKED>(P, C, l, true);
    smooth2<MASKED, STORED>(P, C, l);
  }
}

// PROGRAM 0: V-cycle from the finest tail level, x = 0 initially
//         1: V-cycle, first guess read from x_in
//         2: F-cycle of the tail (restrict b down, coarsest solve, interpolate + V-cycle up)
template <bool MASKED, bool STORED>
__global__ void __launch_bounds__(NT, 1) k_mg_ctail(const __grid_constant__ Params P, int program) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ int ntrace;
  Ctx C;
  C.rank = (int)cluster.block_rank();
  C.ntrace = &ntrace;
  if (threadIdx.x == 0) ntrace = 0;
  double *base = reinterpret_cast<double *>(smem_raw);
  double **tab = reinterpret_cast<double **>(base + 3 * P.total);
  for (int a = 0; a < 3; a++) {
    C.A[a] = base + a * P.total;
    C.RP[a] = tab + a * P.rptotal;
  }
  // row-address tables: row j of a distributed level lives in CTA j / rows[l]
  for (int l = 0; l < P.nlev; l++) {
    if (!P.dist[l]) continue;
    const int ny = P.lv[l].ny, nx = P.lv[l].nx, R = P.rows[l];
    for (int q = threadIdx.x; q < 3 * ny; q += NT) {
      const int a = q / ny, j = q - a * ny;
      const int owner = j / R;
      double *local = C.A[a] + P.off[l] + (j - owner * R) * nx;
      C.RP[a][P.rp[l] + j] = cluster.map_shared_rank(local, owner);
    }
  }
  __syncthreads();
  {
    const View x = view(P, C, AX, 0), b = view(P, C, AB, 0);
    const int nx0 = P.lv[0].nx;
    for_points(P, C, 0, 0, P.lv[0].ny - 1, 0, nx0 - 1, [&](int j, int i) {
      b.row(j)[i] = P.b_in[j * nx0 + i];
      x.row(j)[i] = (program == 1) ? P.x_in[j * nx0 + i] : 0.;
    });
  }
  sync(P, C, true);   // every CTA of the cluster is running and has its band loaded
  if (program == 2) {
    const int last = P.nlev - 1;
    for (int l = 0; l < last; l++) restrict_to<MASKED>(P, C, l, AB);
    coarsest<MASKED, STORED>(P, C);
    for (int l = last - 1; l >= 0; l--) {
      interpolate<MASKED>(P, C, l, false);
      vcycle<MASKED, STORED>(P, C, l);
    }
  } else {
    vcycle<MASKED, STORED>(P, C, 0);
  }
  // the last operator ended with a cluster barrier: nobody touches this CTA's memory any more
  {
    const View x = view(P, C, AX, 0);
    const int nx0 = P.lv[0].nx;
    const bool writer = P.dist[0] || C.rank == 0;
    if (writer)
      for_points(P, C, 0, 0, P.lv[0].ny - 1, 0, nx0 - 1, [&](int j, int i) {
        const double v = x.row(j)[i];
        if (P.acc) P.acc[j * nx0 + i] = P.acc[j * nx0 + i] + v;
        else P.x_out[j * nx0 + i] = v;
      });
  }
}

}  // namespace ctail
