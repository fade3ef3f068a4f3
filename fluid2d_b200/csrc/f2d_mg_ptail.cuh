// Coarse-tail multigrid kernel for ALL-FLUID DOUBLY PERIODIC SQUARE hierarchies (included by
// f2d_multigrid.cu): every tail level is in the constant-stencil class (f2d_mg_fused.cuh, mode 1)
// and the levels are (2^TOP)^2, ..., 8^2, 4^2.
//
// Same job and same cluster decomposition as f2d_mg_ctail.cuh (one launch = one V- or F-cycle of
// the coarse sub-hierarchy, the larger levels split in row bands over the CTAs of a thread-block
// cluster, the smallest ones replicated), specialised for the case in which every halo value the
// reference computes or fills is the periodic image of an interior value -- bit for bit, because
// the stencil is the same constant at every cell (the argument of tail::coarsest_periodic):
//
//   * arrays hold INTERIOR cells only; the x neighbours and, on a replicated level, the y
//     neighbours are index wraps (i +- 1) & (n - 1): no halo columns, no halo images to store, no
//     ring cells to sweep;
//   * the whole cycle is a template recursion over the level size, so every shift, stride, trip
//     count, barrier size and shared-memory offset is an immediate: the coarse levels are bound by
//     the length of the dependent instruction chain of a phase (measured: a 9-point phase of a
//     16^2 level costs 350-450 cycles written this way, 800-1000 with run-time level tables);
//   * a band of a distributed level keeps 2 ghost rows on either side; the first sweep and the
//     residual are evaluated redundantly on the band +- 1 row, so the residual needs no exchange
//     and a smooth exchanges once: 4 cluster barriers per level visit instead of 6;
//   * a level is worked on by as many threads as it has cells per CTA (32 ... 1024), synchronised
//     by a named barrier of exactly that many threads (one warp: __syncwarp); the others wait at
//     the next barrier that includes them.
//
// Arithmetic: fused::jacobi_val / resid_val and the restriction / interpolation expressions of
// the Fortran, evaluated at interior cells only -- results identical to k_mg_ctail / k_mg_tail /
// the per-level kernels.
#pragma once
#include <cooperative_groups.h>

namespace ptail {

namespace cg = cooperative_groups;

constexpr int NH = 3;
constexpr int NT = 1024;
constexpr int MAXL = 8;          // levels 256^2 ... 4^2 at most (LG = 8 ... 2)
constexpr int G = 2;             // ghost rows of a band
constexpr int LGMIN = 2;         // the coarsest level is 4 x 4

struct Params {
  fused::LevelK k[MAXL];    // stencil constants, index = LG - LGMIN
  int ndeepest;
  const double *b_in;       // global (halo-filled) rhs of the finest tail level
  const double *x_in;       // global first guess (program 1) or nullptr
  double *x_out;            // global result
  double *acc;              // if set: acc += result (solve(), hierarchy.py:171)
  long long *trace;
  int trace_cap;
};

// compile-time geometry of level LG on a cluster of NC CTAs
template <int LG, int NC>
struct Geo {
  static constexpr int N = 1 << LG;
  static constexpr bool DIST = NC > 1 && N / NC >= 4 && N * N >= 2048;
  static constexpr int R = DIST ? N / NC : N;              // own rows
  static constexpr int ROWS = DIST ? R + 2 * G : N;        // local rows
  static constexpr int CELLS = ROWS * N;
  static constexpr int OWN = R * N;
  static constexpr int NA = OWN >= NT ? NT : (OWN <= 32 ? 32 : OWN);   // threads at work (OWN is a power of two)
};
// offset of level LG inside the X / B arrays of a tail that starts at TOP
template <int TOP, int LG, int NC>
struct Off { static constexpr int V = Off<TOP, LG + 1, NC>::V + Geo<LG + 1, NC>::CELLS; };
template <int TOP, int NC>
struct Off<TOP, TOP, NC> { static constexpr int V = 0; };
template <int TOP, int NC>
__host__ __device__ constexpr int total_cells() { return Off<TOP, LGMIN, NC>::V + Geo<LGMIN, NC>::CELLS; }
// the scratch region holds the largest local level array (a replicated level may be larger than
// the band of the distributed one above it)
template <int TOP, int LG, int NC>
struct MaxCells {
  static constexpr int A = Geo<LG, NC>::CELLS, B = MaxCells<TOP, LG - 1, NC>::V;
  static constexpr int V = A > B ? A : B;
};
template <int TOP, int NC>
struct MaxCells<TOP, LGMIN, NC> { static constexpr int V = Geo<LGMIN, NC>::CELLS > 64 ? Geo<LGMIN, NC>::CELLS : 64; };
template <int TOP, int NC>
__host__ __device__ constexpr size_t smem_bytes() {
  return (2 * (size_t)total_cells<TOP, NC>() + MaxCells<TOP, TOP, NC>::V) * sizeof(double);
}

// (passed by value: a few registers; the stencil constants of every level sit in shared memory)
struct Ctx {
  double *X, *B, *T;
  const double *K;          // shared: 8 doubles per level (SW, S, SE, W, C, 1-omega, omega/|C|, -)
  long long *trace;
  int *ntrace;
  int rank, south, north, ndeepest, trace_cap;
};

__device__ __forceinline__ void stamp(const Ctx &C) {
  if (C.trace && C.rank == 0 && threadIdx.x == 0) {
    const int k = ++*C.ntrace;
    if (k < C.trace_cap) { C.trace[k] = clock64(); C.trace[0] = k; }
  }
}
// stencil of level LG: the constants jacobi_val / resid_val read
template <int LG>
__device__ __forceinline__ void consts(const Ctx &C, fused::LevelK &L, fused::Coefs<false, false> &kc) {
  const double *k = C.K + 8 * (LG - LGMIN);
  L.c[0] = k[0]; L.c[1] = k[1]; L.c[2] = k[2]; L.c[3] = k[3]; L.c[4] = k[4];
  L.c2 = k[5]; L.c3 = k[6];
  kc.load(L, 0, nullptr, 0);
}
// barrier of the first NB threads of the CTA: one barrier id per size, so that the threads already
// waiting for a larger set never disturb a smaller one (called by every thread; the first NB take part)
template <int NB>
__device__ __forceinline__ void bar() {
  if (NB >= NT) { __syncthreads(); return; }
  if ((int)threadIdx.x >= NB) return;
  if (NB <= 32) { __syncwarp(); return; }
  constexpr int id = NB == 64 ? 1 : (NB == 128 ? 2 : (NB == 256 ? 3 : 4));
  asm volatile("bar.sync %0, %1;" ::"n"(id), "n"(NB) : "memory");
}
__device__ __forceinline__ void cluster_bar() { cg::this_cluster().sync(); }
__host__ __device__ constexpr int cmax(int a, int b) { return a > b ? a : b; }

// offsets (doubles) of the rows j-1, j, j+1 of own-relative row jrel inside a level's local array
template <int LG, int NC>
__device__ __forceinline__ void rows_of(int jrel, int &lo, int &mid, int &hi) {
  using Gm = Geo<LG, NC>;
  if (Gm::DIST) {
    mid = (jrel + G) << LG;
    lo = mid - Gm::N;
    hi = mid + Gm::N;
  } else {
    mid = jrel << LG;
    lo = ((jrel - 1) & (Gm::N - 1)) << LG;
    hi = ((jrel + 1) & (Gm::N - 1)) << LG;
  }
}

template <int LG, bool ZERO>
__device__ __forceinline__ double jac(const fused::LevelK &L, const fused::Coefs<false, false> &kc,
                                      const double *__restrict__ s, double bval, int lo, int mid, int hi, int i) {
  if (ZERO) return fused::jacobi_val<false, false>(L, kc, 0., 0., 0., 0., 0., 0., 0., 0., 0., bval);
  constexpr int NM = (1 << LG) - 1;
  const int il = (i - 1) & NM, ir = (i + 1) & NM;
  return fused::jacobi_val<false, false>(L, kc, s[lo + il], s[lo + i], s[lo + ir], s[mid + il], s[mid + i], s[mid + ir],
                                         s[hi + il], s[hi + i], s[hi + ir], bval);
}

// store own row jrel of a band and, when it is one of the DEPTH lowest / highest own rows, also
// the matching ghost row of the south / north CTA
template <int LG, int NC, int DEPTH>
__device__ __forceinline__ void put_band(double *a, double *as, double *an, int jrel, int i, double v) {
  using Gm = Geo<LG, NC>;
  const int c = ((jrel + G) << LG) + i;
  a[c] = v;
  if (jrel < DEPTH) as[c + Gm::R * Gm::N] = v;           // my row jrel = the south CTA's row R + jrel
  if (jrel >= Gm::R - DEPTH) an[c - Gm::R * Gm::N] = v;  // my row jrel = the north CTA's row jrel - R
}

// Grid.smooth on level LG; ZERO: x is identically zero and is not read.  NB_AFTER: threads of
// the phase that follows (the closing barrier includes them).
template <int TOP, int LG, int NC, bool ZERO, int NB_AFTER>
__device__ __forceinline__ void smooth2(const Ctx &C) {
  using Gm = Geo<LG, NC>;
  double *x = C.X + Off<TOP, LG, NC>::V, *t = C.T;
  const double *b = C.B + Off<TOP, LG, NC>::V;
  fused::LevelK L;
  fused::Coefs<false, false> kc;
  consts<LG>(C, L, kc);
  const int tid = threadIdx.x;
  if (tid < Gm::NA) {
    constexpr int J0 = Gm::DIST ? -1 : 0, CNT = (Gm::DIST ? Gm::R + 2 : Gm::R) * Gm::N;
#pragma unroll
    for (int p = tid; p < CNT; p += Gm::NA) {
      const int jrel = J0 + (p >> LG), i = p & (Gm::N - 1);
      int lo, mid, hi;
      rows_of<LG, NC>(jrel, lo, mid, hi);
      t[mid + i] = jac<LG, ZERO>(L, kc, x, b[mid + i], lo, mid, hi, i);
    }
  }
  // (band, x read: the neighbours may still be reading the ghost rows sweep 2 overwrites)
  if (Gm::DIST && !ZERO) cluster_bar(); else bar<Gm::NA>();
  stamp(C);
  if (tid < Gm::NA) {
    double *xs = x, *xn = x;
    if (Gm::DIST) {
      cg::cluster_group cl = cg::this_cluster();
      xs = cl.map_shared_rank(x, C.south);
      xn = cl.map_shared_rank(x, C.north);
    }
#pragma unroll
    for (int p = tid; p < Gm::OWN; p += Gm::NA) {
      const int jrel = p >> LG, i = p & (Gm::N - 1);
      int lo, mid, hi;
      rows_of<LG, NC>(jrel, lo, mid, hi);
      const double v = jac<LG, false>(L, kc, t, b[mid + i], lo, mid, hi, i);
      if (Gm::DIST) put_band<LG, NC, G>(x, xs, xn, jrel, i, v);
      else x[mid + i] = v;
    }
  }
  if (Gm::DIST) cluster_bar(); else bar<cmax(Gm::NA, NB_AFTER)>();
  stamp(C);
}

// full-weighting restriction of the fine array f (level LG layout) into B(LG-1)
template <int TOP, int LG, int NC>
__device__ __forceinline__ void restrict_from(const Ctx &C, const double *f) {
  using Gf = Geo<LG, NC>;
  using Gc = Geo<LG - 1, NC>;
  double *bc = C.B + Off<TOP, LG - 1, NC>::V;
  constexpr int ROWS = Gf::DIST ? Gf::R / 2 : Gc::N;       // coarse rows this CTA produces
  constexpr int CNT = ROWS * Gc::N;
  constexpr int NA = CNT >= Gf::NA ? Gf::NA : (CNT <= 32 ? 32 : CNT);
  const int tid = threadIdx.x;
  if (tid < NA) {
    cg::cluster_group cl = cg::this_cluster();
    double *bs = bc, *bn = bc;
    if (Gf::DIST && Gc::DIST) {
      bs = cl.map_shared_rank(bc, C.south);
      bn = cl.map_shared_rank(bc, C.north);
    }
#pragma unroll
    for (int p = tid; p < CNT; p += NA) {
      const int jc = p >> (LG - 1), ic = p & (Gc::N - 1);  // coarse row (relative to my first one), column
      int lo, mid, hi;
      rows_of<LG, NC>(2 * jc + 1, lo, mid, hi);             // fine centre: row 2 jc + 1, column 2 ic + 1
      const int i = 2 * ic + 1, il = 2 * ic, ir = (2 * ic + 2) & (Gf::N - 1);
      const double val = 0.25 * f[mid + i] + 0.125 * (((f[mid + il] + f[mid + ir]) + f[lo + i]) + f[hi + i]) +
                         0.0625 * (((f[lo + il] + f[lo + ir]) + f[hi + il]) + f[hi + ir]);
      if (!Gf::DIST) {
        bc[(jc << (LG - 1)) + ic] = val;
      } else if (Gc::DIST) {
        put_band<LG - 1, NC, 1>(bc, bs, bn, jc, ic, val);   // b is read on own rows +- 1
      } else {
        // the coarse level is replicated: my rows go to every CTA
        const int c = ((C.rank * ROWS + jc) << (LG - 1)) + ic;
#pragma unroll
        for (int q = 0; q < NC; q++) cl.map_shared_rank(bc, q)[c] = val;
      }
    }
  }
  if (Gf::DIST) cluster_bar(); else bar<cmax(NA, Gc::NA)>();
  stamp(C);
}

// B(LG-1) = R(b - A x): the residual on own rows +- 1 into T (no exchange), then the restriction
template <int TOP, int LG, int NC>
__device__ __forceinline__ void resid_restrict(const Ctx &C) {
  using Gm = Geo<LG, NC>;
  const double *x = C.X + Off<TOP, LG, NC>::V, *b = C.B + Off<TOP, LG, NC>::V;
  double *t = C.T;
  fused::LevelK L;
  fused::Coefs<false, false> kc;
  consts<LG>(C, L, kc);
  const int tid = threadIdx.x;
  if (tid < Gm::NA) {
    constexpr int J0 = Gm::DIST ? -1 : 0, CNT = (Gm::DIST ? Gm::R + 2 : Gm::R) * Gm::N, NM = Gm::N - 1;
#pragma unroll
    for (int p = tid; p < CNT; p += Gm::NA) {
      const int jrel = J0 + (p >> LG), i = p & NM;
      int lo, mid, hi;
      rows_of<LG, NC>(jrel, lo, mid, hi);
      const int il = (i - 1) & NM, ir = (i + 1) & NM;
      t[mid + i] = fused::resid_val<false, false>(L, kc, L.c[4], x[lo + il], x[lo + i], x[lo + ir], x[mid + il], x[mid + i],
                                                  x[mid + ir], x[hi + il], x[hi + i], x[hi + ir], b[mid + i]);
    }
  }
  bar<Gm::NA>();
  stamp(C);
  restrict_from<TOP, LG, NC>(C, t);
}

// X(LG) = [X(LG) +] I(X(LG-1)) on own rows +- 2 (band) / all rows
template <int TOP, int LG, int NC, bool ADD>
__device__ __forceinline__ void interpolate(const Ctx &C) {
  using Gf = Geo<LG, NC>;
  using Gc = Geo<LG - 1, NC>;
  const double *xc = C.X + Off<TOP, LG - 1, NC>::V;
  double *xf = C.X + Off<TOP, LG, NC>::V;
  constexpr int NMC = Gc::N - 1;
  const int tid = threadIdx.x;
  const int base = Gf::DIST ? C.rank * Gf::R : 0;          // first own fine row (interior numbering)
  const int basec = Gc::DIST ? C.rank * Gc::R : 0;
  if (tid < Gf::NA) {
    constexpr int J0 = Gf::DIST ? -G : 0, CNT = Gf::CELLS;
#pragma unroll
    for (int p = tid; p < CNT; p += Gf::NA) {
      const int jrel = J0 + (p >> LG), i = p & (Gf::N - 1);
      const int jf = base + jrel;                           // global interior row (-2 .. N+1 on a band)
      // halo-numbered parities: pj = (jf + 3) & 1
      const int pj = (jf + 1) & 1, pi = (i + 1) & 1;
      const int kc = pj ? (jf >> 1) - 1 : (jf - 1) >> 1;    // coarse row (global interior numbering; >> floors)
      const int ic = pi ? ((i >> 1) - 1) & NMC : (i - 1) >> 1;
      const int ic1 = (ic + 1) & NMC;
      int r0, r1;
      if (Gc::DIST) {
        r0 = (kc - basec + G) << (LG - 1);
        r1 = r0 + Gc::N;
      } else {
        r0 = (kc & NMC) << (LG - 1);
        r1 = ((kc + 1) & NMC) << (LG - 1);
      }
      double iv;
      if (!pj && !pi) iv = xc[r0 + ic];
      else if (!pj) iv = (xc[r0 + ic] + xc[r0 + ic1]) * fused::interp_w2(2);
      else if (!pi) iv = (xc[r0 + ic] + xc[r1 + ic]) * fused::interp_w2(2);
      else iv = fused::interp_w4(4) * (((xc[r0 + ic] + xc[r0 + ic1]) + xc[r1 + ic]) + xc[r1 + ic1]);
      xf[p] = ADD ? xf[p] + iv : iv;                        // p is the local index: rows J0.. in storage order
    }
  }
  bar<Gf::NA>();
  stamp(C);
}

// deepest level (4 x 4, replicated): x = 0, then ndeepest double sweeps, half a warp
template <int TOP, int NC, int NB_AFTER>
__device__ __forceinline__ void coarsest(const Ctx &C) {
  constexpr int LG = LGMIN, N = 1 << LG, U = N * N;
  double *x = C.X + Off<TOP, LG, NC>::V, *u = C.T;
  const double *b = C.B + Off<TOP, LG, NC>::V;
  if (threadIdx.x < 32) {
    fused::LevelK L;
    fused::Coefs<false, false> kc;
    consts<LG>(C, L, kc);
    const int p = threadIdx.x & (U - 1);     // lanes 16..31 repeat the work of lanes 0..15 (same values, same stores)
    int lo, mid, hi;
    rows_of<LG, NC>(p >> LG, lo, mid, hi);
    const int i = p & (N - 1);
    const double bq = b[p];
    const int nd2 = 2 * C.ndeepest;
    double *src = x, *dst = u;
    dst[p] = jac<LG, true>(L, kc, src, bq, lo, mid, hi, i);
    __syncwarp();
    for (int s = 1; s < nd2; s++) {
      double *tmp = src; src = dst; dst = tmp;
      dst[p] = jac<LG, false>(L, kc, src, bq, lo, mid, hi, i);
      __syncwarp();
    }
    // an even number of sweeps: the last one wrote x
  }
  bar<cmax(32, NB_AFTER)>();
  stamp(C);
}

// V-cycle from level LG down (hierarchy.py:98-127); ZERO: X(LG) is identically zero;
// NB_END: threads of the phase that follows the cycle
template <int TOP, int LG, int NC, bool ZERO, int NB_END>
struct Cycle {
  static __device__ __forceinline__ void run(const Ctx &C) {
    smooth2<TOP, LG, NC, ZERO, Geo<LG, NC>::NA>(C);
    resid_restrict<TOP, LG, NC>(C);
    Cycle<TOP, LG - 1, NC, true, Geo<LG, NC>::NA>::run(C);
    interpolate<TOP, LG, NC, true>(C);
    smooth2<TOP, LG, NC, false, NB_END>(C);
  }
};
template <int TOP, int NC, bool ZERO, int NB_END>
struct Cycle<TOP, LGMIN, NC, ZERO, NB_END> {
  static __device__ __forceinline__ void run(const Ctx &C) { coarsest<TOP, NC, NB_END>(C); }
};
// V-cycles of the F-cycle: not inlined into each other (one body per starting level)
template <int TOP, int LG, int NC>
__device__ __noinline__ void vcycle_from(const Ctx C) {
  Cycle<TOP, LG, NC, false, (LG < TOP ? Geo<(LG < TOP ? LG + 1 : LG), NC>::NA : NT)>::run(C);
}
// F-cycle (hierarchy.py:131-151): restrict b down, coarsest solve, then from each level upwards
// x = I(coarser x), V-cycle
template <int TOP, int LG, int NC>
struct FDown {
  static __device__ __forceinline__ void run(const Ctx &C) {
    restrict_from<TOP, LG, NC>(C, C.B + Off<TOP, LG, NC>::V);
    FDown<TOP, LG - 1, NC>::run(C);
  }
};
template <int TOP, int NC>
struct FDown<TOP, LGMIN, NC> {
  static __device__ __forceinline__ void run(const Ctx &) {}
};
template <int TOP, int LG, int NC>
struct FUp {   // levels LGMIN+1 .. LG
  static __device__ __forceinline__ void run(const Ctx &C) {
    FUp<TOP, LG - 1, NC>::run(C);
    interpolate<TOP, LG, NC, false>(C);
    vcycle_from<TOP, LG, NC>(C);
  }
};
template <int TOP, int NC>
struct FUp<TOP, LGMIN, NC> {
  static __device__ __forceinline__ void run(const Ctx &) {}
};

// PROGRAM 0: V-cycle, x = 0 initially; 1: V-cycle from x_in; 2: F-cycle of the tail
template <int TOP, int NC>
__global__ void __launch_bounds__(NT, 1) k_mg_ptail(const __grid_constant__ Params P, int program) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using G0 = Geo<TOP, NC>;
  __shared__ int ntrace;
  __shared__ double sk[8 * MAXL];
  Ctx C;
  C.rank = NC > 1 ? (int)cg::this_cluster().block_rank() : 0;
  C.south = (C.rank + NC - 1) % NC;
  C.north = (C.rank + 1) % NC;
  C.K = sk;
  C.ndeepest = P.ndeepest;
  C.trace = P.trace;
  C.trace_cap = P.trace_cap;
  C.ntrace = &ntrace;
  if (threadIdx.x == 0) ntrace = 0;
  if (threadIdx.x < 8 * (TOP - LGMIN + 1)) {
    const int lev = threadIdx.x >> 3, q = threadIdx.x & 7;
    const fused::LevelK &k = P.k[lev];
    sk[threadIdx.x] = q < 5 ? k.c[q] : (q == 5 ? k.c2 : (q == 6 ? k.c3 : 0.));
  }
  C.X = reinterpret_cast<double *>(smem_raw);
  C.B = C.X + total_cells<TOP, NC>();
  C.T = C.B + total_cells<TOP, NC>();
  constexpr int N0 = G0::N, NX0 = N0 + 2 * NH;
  const int base0 = G0::DIST ? C.rank * G0::R : 0;
  f2d::pdl_trigger();
  f2d::pdl_wait();
  {
    // b on own rows +- 1 (loaded +- 2), x (program 1) on own rows +- 2: the global arrays are halo
    // filled, so the rows outside the interior are read at their halo position (global row 3 + j)
    double *b = C.B, *x = C.X;
    constexpr int GH = G0::DIST ? G : 0;
    for (int p = threadIdx.x; p < G0::CELLS; p += NT) {
      const int jrel = (p >> TOP) - GH, i = p & (N0 - 1);
      const size_t g = (size_t)(NH + base0 + jrel) * NX0 + NH + i;
      b[p] = P.b_in[g];
      if (program == 1) x[p] = P.x_in[g];
    }
  }
  if (NC > 1) cluster_bar(); else __syncthreads();   // every CTA of the cluster is running
  stamp(C);
  if (program == 2) {
    FDown<TOP, TOP, NC>::run(C);
    coarsest<TOP, NC, (TOP > LGMIN ? Geo<(TOP > LGMIN ? LGMIN + 1 : LGMIN), NC>::NA : NT)>(C);
    FUp<TOP, TOP, NC>::run(C);
  } else if (program == 0) {
    Cycle<TOP, TOP, NC, true, NT>::run(C);
  } else {
    Cycle<TOP, TOP, NC, false, NT>::run(C);
  }
  if (!G0::DIST) __syncthreads();
  if (G0::DIST || C.rank == 0) {
    // own interior rows and the halo images they have in the global array
    const double *x = C.X;
    constexpr int GH = G0::DIST ? G : 0, NY0 = N0 + 2 * NH;
    for (int p = threadIdx.x; p < G0::OWN; p += NT) {
      const int jrel = p >> TOP, i = p & (N0 - 1);
      const double v = x[((jrel + GH) << TOP) + i];
      const int j = NH + base0 + jrel, ii = NH + i;
      if (P.acc) {
        P.acc[(size_t)j * NX0 + ii] = P.acc[(size_t)j * NX0 + ii] + v;
        f2d::for_each_halo_image(j, ii, NY0, NX0, NH, [&](int j2, int i2) {
          P.acc[(size_t)j2 * NX0 + i2] = P.acc[(size_t)j2 * NX0 + i2] + v;
        });
      } else {
        P.x_out[(size_t)j * NX0 + ii] = v;
        f2d::for_each_halo_image(j, ii, NY0, NX0, NH, [&](int j2, int i2) { P.x_out[(size_t)j2 * NX0 + i2] = v; });
      }
    }
  }
}

}  // namespace ptail
