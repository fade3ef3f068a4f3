// Cluster coarse-tail multigrid kernel (included by f2d_multigrid.cu).
//
// Same job as f2d_mg_tail.cuh -- one launch runs a whole V-cycle (hierarchy.py:98-127) or
// F-cycle (hierarchy.py:131-151) of the sub-hierarchy of coarse levels with every array
// resident in shared memory -- but on a thread-block CLUSTER of NC CTAs (16, the sm_100
// maximum, or 8), so that the sub-hierarchy can start at 256^2 / 128^2 instead of 64^2 and
// each operator application of the larger levels is shared by NC SMs.
//
// Layout: the cluster is a one-dimensional y-slab decomposition in miniature.
//   * "distributed" levels (at least 4 interior rows per CTA): CTA c owns the interior rows
//     [c*R, (c+1)*R) and keeps them, for x and b, in a local (R + 6) x nx array whose 3 ghost
//     rows on either side mirror the neighbouring CTAs' edge rows (cyclically: the ghost rows
//     of the first / last CTA are the periodic halo rows of the level).  Every stencil read is
//     a LOCAL shared-memory read; the producer of a field stores its 3 top / bottom rows (and
//     their x images) into the neighbours' ghost rows through distributed shared memory
//     (st.shared::cluster) where the reference fills the halo, followed by one cluster
//     barrier.  In local coordinates a band is a small domain whose row jl is the global row
//     c*R + jl: restriction (2J-2) and interpolation ((j>>1)+1) keep their index maps.
//   * "replicated" levels (the smallest ones, where latency rules): every CTA keeps the
//     whole level and computes all of it, redundantly and identically, with
//     __syncthreads() only.  The restriction that leaves the last distributed level
//     broadcasts its rows to every CTA.
//   * the scratch array (sweep-1 result, residual) is one region shared by all levels.
//
// Arithmetic: the same expressions, in the same order, as the Fortran kernels (through
// fused::jacobi_val / resid_val / interp_w*), evaluated at the same cells: the first sweep on
// [2, n-3] in place (ring values included, with the coefficients stored AT the ring cell), the
// second sweep, the residual and the restriction on the interior followed by the periodic
// images -- so results are bit-identical to the one-CTA tail and to the per-level kernels.
#pragma once
#include <cooperative_groups.h>

namespace ctail {

namespace cg = cooperative_groups;

constexpr int NH = 3;
constexpr int NT = 1024;
constexpr int MAXL = 12;
constexpr int MAXNC = 16;        // CTAs per cluster (non-portable maximum of sm_100)
constexpr int MINROWS = 4;       // a level is distributed when every CTA owns at least this many rows

struct Lev {
  fused::LevelK k;          // geometry / matrix of the level (global shape)
  int dist;                 // 1: rows distributed over the CTAs; 0: replicated in every CTA
  int R;                    // interior rows a CTA owns (dist) / all interior rows (replicated)
  int off;                  // offset (doubles) of the level inside the local X and B arrays
  int lgn;                  // log2 of the interior width
};

struct Params {
  int nlev;
  Lev lv[MAXL];             // index 0 = finest of the tail
  int total;                // doubles per local X / B array
  int tsize;                // doubles of the shared scratch region
  int ndeepest;
  const double *b_in;       // global rhs of the finest tail level
  const double *x_in;       // global first guess (program 1) or nullptr
  double *x_out;            // global result of the finest tail level
  double *acc;              // if set: acc += result instead of storing it (solve(), hierarchy.py:171)
  long long *trace;         // diagnostics (f2d_mg_set_trace): trace[0] = count, then one clock64() per barrier
  int trace_cap;
};

struct Ctx {
  double *X, *B, *T;        // this CTA's arrays (X, B: all levels concatenated; T: one region)
  int rank, nc, south, north;
  int *ntrace;
};

__device__ __forceinline__ void sync(const Params &P, const Ctx &C, bool cluster_wide) {
  if (cluster_wide) cg::this_cluster().sync();
  else __syncthreads();
  if (P.trace && C.rank == 0 && threadIdx.x == 0) {   // stores only: nothing waits on global memory
    const int k = ++*C.ntrace;
    if (k < P.trace_cap) { P.trace[k] = clock64(); P.trace[0] = k; }
  }
}

// local geometry of a level in this CTA
struct Geo {
  int nyl, nx, R, base;     // local rows (R + 6 or ny), row pitch, own rows, global row of local row 0
};
__device__ __forceinline__ Geo geo(const Lev &lv, const Ctx &C) {
  Geo g;
  g.nx = lv.k.nx;
  g.R = lv.R;
  g.nyl = lv.R + 2 * NH;
  g.base = lv.dist ? C.rank * lv.R : 0;
  return g;
}

// store an interior value of a field together with the images the halo fill of the reference
// makes of it: x images locally; y images locally (replicated level: periodic wrap) or into the
// neighbouring CTAs' ghost rows (distributed level)
struct Put {
  double *a, *as, *an;      // local array, the same array in the south / north CTA
  int nx, n, R, dist;
  __device__ __forceinline__ void operator()(int jl, int i, double val) const {
    a[jl * nx + i] = val;
    const bool xlo = i < 2 * NH, xhi = i >= n;
    if (xlo) a[jl * nx + i + n] = val;
    if (xhi) a[jl * nx + i - n] = val;
    if (jl < 2 * NH) {        // my 3 lowest rows: top ghost rows of the south CTA / top halo rows
      double *q = (dist ? as : a) + (jl + R) * nx;
      q[i] = val;
      if (xlo) q[i + n] = val;
      if (xhi) q[i - n] = val;
    }
    if (jl >= R) {            // my 3 highest rows: bottom ghost rows of the north CTA / bottom halo rows
      double *q = (dist ? an : a) + (jl - R) * nx;
      q[i] = val;
      if (xlo) q[i + n] = val;
      if (xhi) q[i - n] = val;
    }
  }
};
__device__ __forceinline__ Put make_put(const Lev &lv, const Ctx &C, double *a) {
  Put p;
  p.a = a;
  p.nx = lv.k.nx;
  p.n = lv.k.nx - 2 * NH;
  p.R = lv.R;
  p.dist = lv.dist;
  p.as = p.an = a;
  if (lv.dist) {
    cg::cluster_group cl = cg::this_cluster();
    p.as = cl.map_shared_rank(a, C.south);
    p.an = cl.map_shared_rank(a, C.north);
  }
  return p;
}

template <bool MASKED, bool STORED, bool ZERO>
__device__ __forceinline__ double jacobi_at(const fused::LevelK &L, const fused::Coefs<MASKED, STORED> &kc,
                                            const double *__restrict__ s, const double *__restrict__ b, int jl, int i,
                                            int base) {
  const int nx = L.nx;
  const int c = jl * nx + i;
  const size_t g = (size_t)(base + jl) * nx + i;
  if (MASKED && __ldg(L.msk + g) == 0) return 0.;
  fused::Coefs<MASKED, STORED> k;
  if (MASKED || STORED) k.load(L, g, MASKED ? L.msk + g : nullptr, nx); else k = kc;
  if (ZERO) return fused::jacobi_val<MASKED, STORED>(L, k, 0., 0., 0., 0., 0., 0., 0., 0., 0., b[c]);
  const double *p = s + c;
  return fused::jacobi_val<MASKED, STORED>(L, k, p[-nx - 1], p[-nx], p[-nx + 1], p[-1], p[0], p[1], p[nx - 1], p[nx],
                                           p[nx + 1], b[c]);
}

// Grid.smooth: two damped-Jacobi sweeps + halo fill, x in place (scratch T).  ZERO: the input is
// identically zero (first visit of a level, hierarchy.py:101-102) and is not read.
template <bool MASKED, bool STORED, bool ZERO>
__device__ void smooth2(const Params &P, const Ctx &C, int l) {
  const Lev &lv = P.lv[l];
  const fused::LevelK &L = lv.k;
  const Geo G = geo(lv, C);
  double *x = C.X + lv.off, *t = C.T;
  const double *b = C.B + lv.off;
  fused::Coefs<MASKED, STORED> kc;
  if (!MASKED && !STORED) kc.load(L, 0, nullptr, 0);
  {
    // sweep 1 in place on local rows [2, R+3] x columns [2, nx-3]
    const int w = G.nx - 4, cnt = (G.R + 2) * w;
    for (int p = threadIdx.x; p < cnt; p += NT) {
      const int dj = p / w, jl = 2 + dj, i = 2 + (p - dj * w);
      t[jl * G.nx + i] = jacobi_at<MASKED, STORED, ZERO>(L, kc, x, b, jl, i, G.base);
    }
  }
  // the neighbours may still be reading the ghost rows of x that sweep 2 is about to overwrite
  sync(P, C, lv.dist);
  {
    const Put put = make_put(lv, C, x);
    const int cnt = G.R << lv.lgn, nm = (1 << lv.lgn) - 1;
    for (int p = threadIdx.x; p < cnt; p += NT) {
      const int jl = NH + (p >> lv.lgn), i = NH + (p & nm);
      put(jl, i, jacobi_at<MASKED, STORED, false>(L, kc, t, b, jl, i, G.base));
    }
  }
  sync(P, C, lv.dist);
}

// T = b - A x on the interior + halo fill
template <bool MASKED, bool STORED>
__device__ void residual(const Params &P, const Ctx &C, int l) {
  const Lev &lv = P.lv[l];
  const fused::LevelK &L = lv.k;
  const Geo G = geo(lv, C);
  const double *x = C.X + lv.off, *b = C.B + lv.off;
  fused::Coefs<MASKED, STORED> kc;
  if (!MASKED && !STORED) kc.load(L, 0, nullptr, 0);
  const Put put = make_put(lv, C, C.T);
  const int nx = G.nx, cnt = G.R << lv.lgn, nm = (1 << lv.lgn) - 1;
  for (int p = threadIdx.x; p < cnt; p += NT) {
    const int jl = NH + (p >> lv.lgn), i = NH + (p & nm);
    const int c = jl * nx + i;
    const size_t g = (size_t)(G.base + jl) * nx + i;
    double val = 0.;
    if (!MASKED || __ldg(L.msk + g) != 0) {
      fused::Coefs<MASKED, STORED> k;
      if (MASKED || STORED) k.load(L, g, MASKED ? L.msk + g : nullptr, nx); else k = kc;
      const double cdiag = STORED ? __ldg(L.A + 4 * (size_t)L.ny * nx + g) : L.c[4];
      const double *q = x + c;
      val = fused::resid_val<MASKED, STORED>(L, k, cdiag, q[-nx - 1], q[-nx], q[-nx + 1], q[-1], q[0], q[1], q[nx - 1],
                                             q[nx], q[nx + 1], b[c]);
    }
    put(jl, i, val);
  }
  sync(P, C, lv.dist);
}

// full-weighting restriction of the local fine array `xf` of level l into B of level l+1
// (coarse interior + halo fill)
template <bool MASKED>
__device__ void restrict_to(const Params &P, const Ctx &C, int l, const double *xf) {
  const Lev &lf = P.lv[l], &lc = P.lv[l + 1];
  const fused::LevelK &Lc = lc.k;
  const int nxf = lf.k.nx, nxc = Lc.nx;
  double *bc = C.B + lc.off;
  // rows this CTA produces: its own coarse rows (both distributed / both replicated), or the
  // R/2 coarse rows under its fine band when the coarse level is replicated
  const bool bcast = lf.dist && !lc.dist;
  const int rows = bcast ? lf.R / 2 : lc.R;
  const int jbase = bcast ? C.rank * rows : 0;     // first produced row (coarse local numbering, minus NH)
  const int gbase = lc.dist ? C.rank * lc.R : 0;   // global coarse row of coarse local row 0
  const int cnt = rows << lc.lgn, nm = (1 << lc.lgn) - 1;
  const Put put = make_put(lc, C, bc);
  for (int p = threadIdx.x; p < cnt; p += NT) {
    const int t = p >> lc.lgn, i = NH + (p & nm);
    const int jl = NH + jbase + t;                 // coarse row in the coarse array of this CTA
    double val = 0.;
    if (!MASKED || __ldg(Lc.msk + (size_t)(gbase + jl) * nxc + i) != 0) {
      // fine centre: local fine row 2*(NH + t) - 2 (band coordinates), column 2*i - 2
      const double *f = xf + (2 * (NH + t) - 2) * nxf + (2 * i - 2);
      val = 0.25 * f[0] + 0.125 * (((f[-1] + f[1]) + f[-nxf]) + f[nxf]) +
            0.0625 * (((f[-nxf - 1] + f[-nxf + 1]) + f[nxf - 1]) + f[nxf + 1]);
    }
    if (!bcast) {
      put(jl, i, val);
    } else {
      cg::cluster_group cl = cg::this_cluster();
      Put q = put;
      for (int r = 0; r < C.nc; r++) {
        q.a = cl.map_shared_rank(bc, r);
        q(jl, i, val);
      }
    }
  }
  sync(P, C, lf.dist || lc.dist);
}

// X(l) = [X(l) +] I(X(l+1)) over the whole local fine array (ghost rows included: the same
// expression the neighbour evaluates there)
template <bool MASKED>
__device__ void interpolate(const Params &P, const Ctx &C, int l, bool add) {
  const Lev &lf = P.lv[l], &lc = P.lv[l + 1];
  const fused::LevelK &Lf = lf.k, &Lc = lc.k;
  const Geo G = geo(lf, C);
  const int nx = G.nx, nxc = Lc.nx;
  const int gbc = lc.dist ? C.rank * lc.R : 0;
  const int shift = (G.base >> 1) - gbc;            // coarse local row = (jl >> 1) + 1 + shift
  const double *xc = C.X + lc.off;
  double *xf = C.X + lf.off;
  const int cnt = G.nyl * nx;
  for (int p = threadIdx.x; p < cnt; p += NT) {
    const int jl = p / nx, i = p - jl * nx;
    double iv = 0.;
    if (!MASKED || __ldg(Lf.msk + (size_t)(G.base + jl) * nx + i) > 0) {
      const int kl = ((jl >> 1) + 1 + shift) * nxc + (i >> 1) + 1;
      const int pj = jl & 1, pi = i & 1;
      const int8_t *mc = Lc.msk + (size_t)gbc * nxc + kl;   // mask of the same coarse cell (global array)
      const double *c0 = xc + kl;
      if (!pj && !pi) {
        iv = c0[0];
      } else if (!pj) {
        const int s = MASKED ? __ldg(mc) + __ldg(mc + 1) : 2;
        iv = (c0[0] + c0[1]) * fused::interp_w2(s);
      } else if (!pi) {
        const int s = MASKED ? __ldg(mc) + __ldg(mc + nxc) : 2;
        iv = (c0[0] + c0[nxc]) * fused::interp_w2(s);
      } else {
        const int s = MASKED ? __ldg(mc) + __ldg(mc + 1) + __ldg(mc + nxc) + __ldg(mc + nxc + 1) : 4;
        iv = fused::interp_w4(s) * (((c0[0] + c0[1]) + c0[nxc]) + c0[nxc + 1]);
      }
    }
    xf[p] = add ? xf[p] + iv : iv;
  }
  __syncthreads();   // local reads and writes only
  if (P.trace && C.rank == 0 && threadIdx.x == 0) {
    const int k = ++*C.ntrace;
    if (k < P.trace_cap) { P.trace[k] = clock64(); P.trace[0] = k; }
  }
}

// deepest level: x = 0, then ndeepest double sweeps (hierarchy.py:114-116)
template <bool MASKED, bool STORED>
__device__ void coarsest(const Params &P, const Ctx &C) {
  const int last = P.nlev - 1;
  const Lev &lv = P.lv[last];
  if (!MASKED && !STORED && !lv.dist && tail::coarsest_periodic_ok(lv.k)) {
    // one warp, periodic indexing on the m x n unknowns (see f2d_mg_tail.cuh)
    tail::coarsest_periodic(lv.k, C.X + lv.off, C.B + lv.off, C.T, P.ndeepest);
    return;
  }
  smooth2<MASKED, STORED, true>(P, C, last);
  for (int k = 1; k < P.ndeepest; k++) smooth2<MASKED, STORED, false>(P, C, last);
}

// V-cycle of the levels [l1, nlev-1]; xzero: X of level l1 is identically zero (not read)
template <bool MASKED, bool STORED>
__device__ void vcycle(const Params &P, const Ctx &C, int l1, bool xzero) {
  const int last = P.nlev - 1;
  if (l1 == last) {   // F-cycle only: "V-cycle" from the coarsest level
    coarsest<MASKED, STORED>(P, C);
    return;
  }
  for (int l = l1; l < last; l++) {
    if (l > l1 || xzero) smooth2<MASKED, STORED, true>(P, C, l);
    else smooth2<MASKED, STORED, false>(P, C, l);
    residual<MASKED, STORED>(P, C, l);
    restrict_to<MASKED>(P, C, l, C.T);
  }
  coarsest<MASKED, STORED>(P, C);
  for (int l = last - 1; l >= l1; l--) {
    interpolate<MASKED>(P, C, l, true);
    smooth2<MASKED, STORED, false>(P, C, l);
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
  C.nc = (int)cluster.num_blocks();
  C.south = (C.rank + C.nc - 1) % C.nc;
  C.north = (C.rank + 1) % C.nc;
  C.ntrace = &ntrace;
  if (threadIdx.x == 0) ntrace = 0;
  C.X = reinterpret_cast<double *>(smem_raw);
  C.B = C.X + P.total;
  C.T = C.B + P.total;
  const Lev &l0 = P.lv[0];
  const Geo G0 = geo(l0, C);
  f2d::pdl_trigger();
  f2d::pdl_wait();
  {
    // my band of the rhs (and of the first guess), ghost rows included: the global arrays
    // arrive halo-filled
    const size_t g0 = (size_t)G0.base * G0.nx;
    const int cnt = G0.nyl * G0.nx;
    double *b = C.B + l0.off, *x = C.X + l0.off;
    for (int p = threadIdx.x; p < cnt; p += NT) {
      b[p] = P.b_in[g0 + p];
      if (program == 1) x[p] = P.x_in[g0 + p];
    }
  }
  sync(P, C, true);   // every CTA of the cluster is running (its shared memory may be written)
  if (program == 2) {
    const int last = P.nlev - 1;
    for (int l = 0; l < last; l++) restrict_to<MASKED>(P, C, l, C.B + P.lv[l].off);
    coarsest<MASKED, STORED>(P, C);
    for (int l = last - 1; l >= 0; l--) {
      interpolate<MASKED>(P, C, l, false);
      vcycle<MASKED, STORED>(P, C, l, false);
    }
  } else {
    vcycle<MASKED, STORED>(P, C, 0, program == 0);
  }
  // the last operator ended with a barrier: nobody writes this CTA's arrays any more
  {
    // own rows; the first / last CTA also write the halo rows they hold as ghost rows
    const bool first = !l0.dist || C.rank == 0, lastc = !l0.dist || C.rank == C.nc - 1;
    if (l0.dist || C.rank == 0) {
      const int jlo = first ? 0 : NH, jhi = lastc ? G0.nyl : G0.R + NH;
      const double *x = C.X + l0.off;
      const size_t g0 = (size_t)G0.base * G0.nx;
      for (int p = jlo * G0.nx + threadIdx.x; p < jhi * G0.nx; p += NT) {
        const double v = x[p];
        if (P.acc) P.acc[g0 + p] = P.acc[g0 + p] + v;
        else P.x_out[g0 + p] = v;
      }
    }
  }
  // a CTA must not exit while its neighbours may still write its shared memory: the last
  // operator's barrier came after every remote store
}

}  // namespace ctail
