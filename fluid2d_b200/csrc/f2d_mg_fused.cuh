// Fused, shared-memory-tiled multigrid kernels (included by f2d_multigrid.cu).
//
//   k_smooth2        : Grid.smooth = two damped-Jacobi sweeps (smoothtwicewithA,
//                      fortran_multigrid.f90:2-127) + halo fill, one pass over HBM.
//                      Variants: the input may be zero (first visit of a coarse level in
//                      a V-cycle, hierarchy.py:101-102) or xin + I(xcoarse) (interpolate
//                      :415-498 and the `x += r` of hierarchy.py:123-126 fused in).
//   k_resid_restrict : computeresidualwithA (:320-362) + restrict (:501-546) + halo
//                      fills; the fine residual never goes to HBM.
//   k_zsmooth_resid_restrict : both of the above for a level visited on the way down from a
//                      zero first guess (hierarchy.py:100-107), in one kernel that reads b once
//                      (all-fluid levels; see the comment in front of it).
//
// Tiles: 64 x 32 outputs per CTA of 256 threads (a thread marches a column strip with the 3x3
// window in registers); levels that would get fewer such tiles than the device has SMs use
// 64 x 8 tiles (template parameter TYP / RTYP), where a thread owns 2-3 rows and the stored
// coefficients of all of them are loaded before the first is used.
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
#include <cstddef>

namespace fused {

constexpr int NH = 3;
constexpr int TX = 64;   // outputs per tile in x
constexpr int TY = 32;   // outputs per tile in y (levels with fewer tiles than the device has SMs: TYS)
constexpr int TYS = 8;   // small-level tile height: four times as many CTAs, a quarter of the rows per thread
constexpr int NT = 256;  // threads per CTA

struct LevelK {
  int ny, nx;
  const int8_t *msk;
  const double *A;   // 5 planes SW,S,SE,W,C (STORED) + a 6th written at set-up: omega / |C|
  double c[5];       // SW,S,SE,W,C (constant classes)
  double c1, c2, c3; // omega, 1-omega, omega/|C|
  int ywrap;         // 1: y halo rows are local periodic images; 0: they belong to the neighbouring slabs
};

// coefficients of the 9-point operator at cell `g` (global index) / mask window
template <bool MASKED, bool STORED>
struct Coefs {
  double sw, s, se, w, e, nw, n, ne, c3;
  // m: pointer to the centre of a mask window with row stride ms (MASKED only)
  // STORED: the eight off-diagonal coefficients of cell g and, in c3, what the caller applies
  // with them: omega / |diagonal| for a sweep -- the 6th plane, divided once at set-up
  // (k_inverse_diagonal: the same IEEE division the sweep used to do per point, 25 dependent
  // instructions on a latency-bound path) -- or the diagonal itself for a residual (RESID).
  // Matrices and masks are written at set-up only: non-coherent loads (LDG.CONSTANT).
  template <bool RESID>
  __device__ __forceinline__ void load_stored(const LevelK &L, size_t g) {
    size_t pl = (size_t)L.ny * L.nx;
    const double *A1 = L.A, *A2 = L.A + pl, *A3 = L.A + 2 * pl, *A4 = L.A + 3 * pl;
    int nx = L.nx;
    sw = __ldg(A1 + g); s = __ldg(A2 + g); se = __ldg(A3 + g); w = __ldg(A4 + g);
    e = __ldg(A4 + g + 1); nw = __ldg(A3 + g + nx - 1); n = __ldg(A2 + g + nx); ne = __ldg(A1 + g + nx + 1);
    c3 = __ldg(L.A + (RESID ? 4 : 5) * pl + g);
  }
  __device__ __forceinline__ void load(const LevelK &L, size_t g, const int8_t *m, int ms) {
    if (STORED) {
      load_stored<false>(L, g);
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
  // constant stencil x mask products from a 3x3 mask window held in registers (rows lo / mid / hi)
  __device__ __forceinline__ void from_window(const LevelK &L, int l0, int l1, int l2, int m0, int m2, int h0, int h1,
                                              int h2) {
    sw = l0 ? L.c[0] : 0.; s = l1 ? L.c[1] : 0.; se = l2 ? L.c[2] : 0.;
    w = m0 ? L.c[3] : 0.;  e = m2 ? L.c[3] : 0.;
    nw = h0 ? L.c[2] : 0.; n = h1 ? L.c[1] : 0.; ne = h2 ? L.c[0] : 0.;
    c3 = L.c3;
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
// TMA boxes must start at an even column (16-byte aligned innermost coordinate): the x and
// coarse tiles start at odd columns (i0 - 2 = 1 + 64k, ci0 = 1 + 32k), so their boxes begin
// one column earlier and are two columns wider; the tile proper sits at column offset 1.
constexpr int XP = XW + 2, CP = CW + 2;

// b is read by the very thread that applies the operator at a point, and by nobody else: it
// goes from global memory straight into the registers of the column strip (one coalesced load
// per strip row, shared by both sweeps) instead of through a shared-memory tile -- 18 KB less
// shared memory per CTA and two shared loads less per point.
template <int TYP>
struct Smooth2SmemT {
  static constexpr int XH = TYP + 4, YH = TYP + 2, CH = XH / 2 + 2;
  alignas(128) double xs[XH][XP];   // TMA destinations: 128-byte aligned, dense boxes
  // the coarse tile is dead once the interpolation pass has folded it into xs (a block
  // barrier later sweep 1 starts writing y1): the two share their storage
  union {
    alignas(128) double y1[YH][YW];
    double cs[CH][CP];
  };
  alignas(8) uint64_t bar;          // mbarrier of the TMA loads
  // mask tiles last: the mask-free instantiations are launched with smooth2_smem_nomask<TYP>()
  // bytes only (38 KB at TYP = 32: the register file, not shared memory, then sets 4 CTAs per SM)
  int8_t ms[XH][XW];
  int8_t cm[CH][CW];
};
static_assert(sizeof(double[CH][CP]) <= sizeof(double[YH][YW]), "coarse tile must fit in the y1 tile");
static_assert(sizeof(double[TYS / 2 + 4][CP]) <= sizeof(double[TYS + 2][YW]), "coarse tile must fit in the y1 tile");
using Smooth2Smem = Smooth2SmemT<TY>;
template <int TYP>
constexpr size_t smooth2_smem_nomask() { return offsetof(Smooth2SmemT<TYP>, ms); }
template <int TYP>
constexpr int smooth2_xh() { return TYP + 4; }
template <int TYP>
constexpr int smooth2_ch() { return (TYP + 4) / 2 + 2; }
static_assert(XP % 2 == 0 && YW % 2 == 0 && CP % 2 == 0, "TMA boxes need an even number of doubles per row");

// ---- asynchronous tile loader -----------------------------------------------------
// Copies a ROWS x COLS tile whose origin is `src` (row pitch nx doubles) into `dst` (row
// pitch LD doubles): one tile row per warp and pass, 8 bytes per LDGSTS, every address a
// constant offset from two per-thread bases.  FULL: all elements are inside the array (no
// predicates); otherwise rows >= vrows / columns >= vcols are zero-filled.
template <int ROWS, int COLS, int LD, bool FULL>
__device__ __forceinline__ void load_tile(double *dst, const double *src, int nx, int vrows, int vcols) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NW = NT / 32;
  constexpr int NK = (ROWS + NW - 1) / NW, NC = (COLS + 31) / 32;
  const double *row = src + (size_t)warp * nx + lane;
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst + warp * LD + lane);
  const size_t rstep = (size_t)NW * nx;
#pragma unroll
  for (int k = 0; k < NK; k++) {
    if ((k + 1) * NW <= ROWS || warp + k * NW < ROWS) {
#pragma unroll
      for (int c = 0; c < NC; c++) {
        if ((c + 1) * 32 <= COLS || lane + c * 32 < COLS) {
          const unsigned da = d + (unsigned)((k * NW * LD + c * 32) * 8);
          if (FULL) {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(da), "l"(row + c * 32) : "memory");
          } else {
            bool in = (warp + k * NW < vrows) && (lane + c * 32 < vcols);
            int sz = in ? 8 : 0;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(da), "l"(in ? row + c * 32 : src), "r"(sz)
                         : "memory");
          }
        }
      }
    }
    row += rstep;
  }
}
// (fully unrolled, every load issued before the first store: one L2 round trip for the tile
// instead of one per pass -- the loop form cost the masked kernels 5 dependent round trips)
template <int ROWS, int COLS, int LD>
__device__ __forceinline__ void load_tile_i8(int8_t *dst, const int8_t *src, int nx, int vrows, int vcols) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NW = NT / 32, NK = (ROWS + NW - 1) / NW, NC = (COLS + 31) / 32;
  int8_t v[NK][NC];
#pragma unroll
  for (int k = 0; k < NK; k++) {
    const int r = warp + k * NW;
#pragma unroll
    for (int q = 0; q < NC; q++) {
      const int c = lane + q * 32;
      v[k][q] = (r < ROWS && c < COLS && r < vrows && c < vcols) ? src[(size_t)r * nx + c] : (int8_t)0;
    }
  }
#pragma unroll
  for (int k = 0; k < NK; k++) {
    const int r = warp + k * NW;
#pragma unroll
    for (int q = 0; q < NC; q++) {
      const int c = lane + q * 32;
      if (r < ROWS && c < COLS) dst[r * LD + c] = v[k][q];
    }
  }
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// one column strip of a Jacobi sweep: `sp` (pitch SLD) points at the window centre of the
// first point in the source tile, `bp` (pitch BLD) at its b, `mp` (pitch MLD) at its mask;
// the 3x3 window is carried in registers, so each point costs three shared loads.  GUARD:
// per-point range test [lo, n-1-lo] (rim tiles); the fast path has none.  out(k, val)
// consumes row k (called for in-range points only).
// bget(k): the right-hand side of row k (registers).
template <bool MASKED, bool STORED, bool ZERO, bool GUARD, int NR, int SLD, int MLD, class BGET, class OUT>
__device__ __forceinline__ void jacobi_strip(const LevelK &L, const Coefs<MASKED, STORED> &kc, const double *sp,
                                             BGET bget, const int8_t *mp, int nr, int j, int i, int lo,
                                             OUT out) {
  const int ny = L.ny, nx = L.nx;
  const size_t g = (size_t)j * nx + i;
  double a0 = 0., a1 = 0., a2 = 0., m0 = 0., m1 = 0., m2 = 0.;
  if (!ZERO) {
    a0 = sp[-SLD - 1]; a1 = sp[-SLD]; a2 = sp[-SLD + 1];
    m0 = sp[-1]; m1 = sp[0]; m2 = sp[1];
  }
  // mask-product levels: the 3x3 mask window marches in registers too (3 byte loads per point
  // instead of 9)
  constexpr bool MWIN = MASKED && !STORED;
  int wa0 = 0, wa1 = 0, wa2 = 0, wm0 = 0, wm1 = 0, wm2 = 0;
  if (MWIN) {
    wa0 = mp[-MLD - 1]; wa1 = mp[-MLD]; wa2 = mp[-MLD + 1];
    wm0 = mp[-1]; wm1 = mp[0]; wm2 = mp[1];
  }
  // stored coefficients: every row's loads are issued here, fluid cell or not (the row-by-row
  // form paid one L2 round trip per row: 13 us per kernel on a 256 x 64 level against 7 us for
  // the constant-stencil kernel)
  // (short strips only -- the small-level tiles: nine rows of coefficients in registers leave
  // one CTA per SM, which costs the large stored levels more than the round trips)
  constexpr bool PRE = STORED && NR <= 4;
  Coefs<MASKED, STORED> kpre[PRE ? NR : 1];
  if (PRE) {
#pragma unroll
    for (int k = 0; k < NR; k++) {
      const bool okk = k < nr && (!GUARD || (j + k >= lo && j + k <= ny - 1 - lo && i >= lo && i <= nx - 1 - lo));
      if (okk) kpre[k].template load_stored<false>(L, g + (size_t)k * nx);
    }
  }
#pragma unroll
  for (int k = 0; k < NR; k++) {
    if (k < nr) {
      double h0 = 0., h1 = 0., h2 = 0.;
      if (!ZERO) { h0 = sp[(k + 1) * SLD - 1]; h1 = sp[(k + 1) * SLD]; h2 = sp[(k + 1) * SLD + 1]; }
      int wh0 = 0, wh1 = 0, wh2 = 0;
      if (MWIN) { wh0 = mp[(k + 1) * MLD - 1]; wh1 = mp[(k + 1) * MLD]; wh2 = mp[(k + 1) * MLD + 1]; }
      const bool ok = !GUARD || (j + k >= lo && j + k <= ny - 1 - lo && i >= lo && i <= nx - 1 - lo);
      if (ok) {
        double val = 0.;
        // stored coefficients are fetched whether the cell is fluid or not: loads that do not hang
        // on the mask test can be issued ahead of the rows before them
        Coefs<MASKED, STORED> kk;
        if (PRE) kk = kpre[k];
        else if (STORED) kk.load(L, g + (size_t)k * nx, nullptr, MLD);
        if (!MASKED || (MWIN ? wm1 : (int)mp[k * MLD]) != 0) {
          if (MWIN) kk.from_window(L, wa0, wa1, wa2, wm0, wm2, wh0, wh1, wh2);
          else if (!STORED) kk = kc;
          val = jacobi_val<MASKED, STORED>(L, kk, a0, a1, a2, m0, m1, m2, h0, h1, h2, bget(k));
        }
        out(k, val);
      }
      a0 = m0; a1 = m1; a2 = m2;
      m0 = h0; m1 = h1; m2 = h2;
      if (MWIN) { wa0 = wm0; wa1 = wm1; wa2 = wm2; wm0 = wh0; wm1 = wh1; wm2 = wh2; }
    }
  }
}

// INPUT: 0 xin, 1 zero, 2 I(xc), 3 xin + I(xc)
// PEER (y-slab levels on several GPUs): the tiles of the first / last tile row wait for the
// neighbouring rank's epoch before reading halo rows, store their 3 outermost rows into the
// neighbour's halo rows as well (peer stores), and the last of them publishes the epoch --
// the halo exchange of the reference (halo.fill after every smooth) without a kernel of
// its own, overlapped with the interior tiles.
template <bool MASKED, bool STORED, int INPUT, bool PEER, int TYP>
__global__ void __launch_bounds__(NT)
k_smooth2(LevelK L, const double *__restrict__ xin, const double *__restrict__ b, double *__restrict__ xout,
          const double *__restrict__ xc, const int8_t *__restrict__ mskc, int nxc, int nyc, double *acc,
          f2d::Peer P, int use_tma, const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmc) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // tile height TYP (these shadow the namespace constants of the standard tile); a thread
  // group owns RG rows of the tile
  constexpr int TY = TYP, XH = TY + 4, YH = TY + 2, CH = XH / 2 + 2, RG = TY / 4;
  static_assert(TY % 4 == 0 && TY >= 4 && 2 * YH <= NT, "tile height");
  Smooth2SmemT<TYP> &S = *reinterpret_cast<Smooth2SmemT<TYP> *>(smem_raw);
  const int ny = L.ny, nx = L.nx;
  const int t = threadIdx.x;
  const int by = PEER ? f2d::peer_tile_row(blockIdx.y, gridDim.y) : blockIdx.y;
  const bool bsouth = PEER && by == 0, bnorth = PEER && by == (int)gridDim.y - 1;
  constexpr bool INTERP = INPUT >= 2;
  constexpr bool HAVE_X = (INPUT == 0 || INPUT == 3);
  // programmatic dependent launch: barrier and descriptors are set up while the kernel in
  // front of this one drains; nothing it wrote is touched before pdl_wait()
  f2d::pdl_trigger();
  if (use_tma && t == 0) {
    f2d::mbar_init(&S.bar, 1);
    if (HAVE_X) f2d::tma_prefetch_desc(&tmx);
    if (INTERP) f2d::tma_prefetch_desc(&tmc);
  }
  f2d::pdl_wait();
  if (PEER && (bsouth || bnorth)) f2d::peer_wait(P, bsouth, bnorth);
  const int i0 = NH + blockIdx.x * TX, j0 = NH + by * TY;
  constexpr bool ZERO = INPUT == 1;
  const int cj0 = ((j0 - 2) >> 1) + 1, ci0 = ((i0 - 2) >> 1) + 1;
  // inner tile: every point of the sweep-1 ring and of the output tile is a valid target
  const bool inner = (j0 + TY <= ny - 3) && (i0 + TX <= nx - 3);
  // ---- column strips.  Thread (tx, tg) owns column i0 + tx: in sweep 1 the rows
  // j0-1+r1 .. (9, 8, 8, 9 rows: the 34 rows of the tile + ring 1), in sweep 2 the 8 tile rows
  // among them -- so the b of its points is loaded once, into registers, for both sweeps.
  // The two ring columns i0-1 and i0+TX of sweep 1 are done point-wise by the first 68 threads.
  const int tx = t & (TX - 1), tg = t >> 6;  // TX == 64
  const int r1 = tg * RG + (tg > 0 ? 1 : 0);  // first sweep-1 row (y1 tile coordinates): 0, 9, 17, 25 (RG = 8)
  const int n1 = (tg == 0 || tg == 3) ? RG + 1 : RG;
  double bv[RG + 1], bx = 0.;
  auto load_b = [&]() {
    const int i = i0 + tx;
#pragma unroll
    for (int k = 0; k < RG + 1; k++) {
      const int j = j0 - 1 + r1 + k;
      bv[k] = (k < n1 && (inner || (j < ny && i < nx))) ? b[(size_t)j * nx + i] : 0.;
    }
    if (t < YH * 2) {
      const int j = j0 - 1 + (t >> 1), ii = (t & 1) ? i0 + TX : i0 - 1;
      if (j < ny && ii < nx) bx = b[(size_t)j * nx + ii];
    }
  };
  // ---- stage the tiles: TMA boxes (one instruction per tile, zero fill outside the
  // array) or, on levels smaller than a box, per-element asynchronous copies
  if (use_tma) {
    __syncthreads();
    if (t == 0) {
      // (PEER, boundary tiles: the only ones that read rows a neighbour stored) the neighbour's
      // halo rows were observed through the generic proxy.  Interior tiles skip the fence
      // (measured on 2 B200s, 16384 x 8192 per rank: 943 -> 901 us per launch in the ungraphed
      // step, k_resid_restrict 627 -> 564 us)
      if (PEER && (bsouth || bnorth)) asm volatile("fence.proxy.async;" ::: "memory");
      constexpr unsigned bytes = (HAVE_X ? XH * XP * 8 : 0) + (INTERP ? CH * CP * 8 : 0);
      if (bytes) f2d::mbar_expect_tx(&S.bar, bytes);
      if (HAVE_X) f2d::tma_load_2d(&S.xs[0][0], &tmx, &S.bar, i0 - 3, j0 - 2);
      if (INTERP) f2d::tma_load_2d(&S.cs[0][0], &tmc, &S.bar, ci0 - 1, cj0);
    }
    const int vr = ny - (j0 - 2), vc = nx - (i0 - 2);
    if (INTERP && MASKED) load_tile_i8<CH, CW, CW>(&S.cm[0][0], mskc + (size_t)cj0 * nxc + ci0, nxc, nyc - cj0, nxc - ci0);
    if (MASKED) load_tile_i8<XH, XW, XW>(&S.ms[0][0], L.msk + (size_t)(j0 - 2) * nx + (i0 - 2), nx, vr, vc);
    load_b();
    if (HAVE_X || INTERP) f2d::mbar_wait(&S.bar, 0);
  } else {
    const int vr = ny - (j0 - 2), vc = nx - (i0 - 2);
    if (HAVE_X) {
      const double *src = xin + (size_t)(j0 - 2) * nx + (i0 - 2);
      if (inner) load_tile<XH, XW, XP, true>(&S.xs[0][1], src, nx, 0, 0);
      else load_tile<XH, XW, XP, false>(&S.xs[0][1], src, nx, vr, vc);
    }
    if (INTERP) {
      load_tile<CH, CW, CP, false>(&S.cs[0][1], xc + (size_t)cj0 * nxc + ci0, nxc, nyc - cj0, nxc - ci0);
      if (MASKED) load_tile_i8<CH, CW, CW>(&S.cm[0][0], mskc + (size_t)cj0 * nxc + ci0, nxc, nyc - cj0, nxc - ci0);
    }
    if (MASKED) load_tile_i8<XH, XW, XW>(&S.ms[0][0], L.msk + (size_t)(j0 - 2) * nx + (i0 - 2), nx, vr, vc);
    load_b();
    cp_async_wait_all();
  }
  __syncthreads();
  // ---- fused interpolation: xs = [xin +] I(xc)  (fortran_multigrid.f90:415-498).
  // One thread per coarse cell of the tile produces the 2x2 fine block it anchors.
  if (INTERP) {
    // fine position (r,q) of the tile <-> coarse local (r+1)/2 ... ; the block anchored at
    // coarse local (lj,li) covers fine rows 2*lj-1, 2*lj (tile coordinates) when j0 is odd
    // (j0 = 3 + k*TY): fine tile row r has global parity (j0 - 2 + r) & 1 = (r + 1) & 1
    for (int p = t; p < (XH / 2) * (XW / 2); p += NT) {
      int bj = p / (XW / 2), bi = p % (XW / 2);   // 2x2 block index inside the x tile
      int r = 2 * bj, q = 2 * bi;                 // its first (odd-global) row / column
      // global (j,i) of (r,q) is odd/odd: coarse anchor lj = ((j>>1)+1-cj0) = bj, li = bi
      int lj = bj, li = bi;
      double c00 = S.cs[lj][li + 1], c01 = S.cs[lj][li + 2], c10 = S.cs[lj + 1][li + 1], c11 = S.cs[lj + 1][li + 2];
      int m00 = 1, m01 = 1, m10 = 1, m11 = 1;
      if (MASKED) { m00 = S.cm[lj][li]; m01 = S.cm[lj][li + 1]; m10 = S.cm[lj + 1][li]; m11 = S.cm[lj + 1][li + 1]; }
      // the tile origin is odd/odd in global coordinates, so inside the block
      //  (r  , q  ) odd row, odd col  : 4-point mean anchored at coarse (lj,li)
      //  (r  , q+1) odd row, even col : mean in y along coarse column li+1 -> (c01,c11)
      //  (r+1, q  ) even row, odd col : mean in x along coarse row lj+1    -> (c10,c11)
      //  (r+1, q+1) even row, even col: injection of c11
      double v_oo = interp_w4(m00 + m01 + m10 + m11) * (((c00 + c01) + c10) + c11);
      double v_oe = (c01 + c11) * interp_w2(m01 + m11);
      double v_eo = (c10 + c11) * interp_w2(m10 + m11);
      double v_ee = c11;
      auto put = [&](int rr, int qq, double iv) {
        int j = j0 - 2 + rr, i = i0 - 2 + qq;
        if (j < ny && i < nx) {
          if (MASKED && !(S.ms[rr][qq] > 0)) iv = 0.;
          S.xs[rr][qq + 1] = (INPUT == 3) ? S.xs[rr][qq + 1] + iv : iv;
        } else if (INPUT != 3) {
          S.xs[rr][qq + 1] = 0.;
        }
      };
      put(r, q, v_oo);
      put(r, q + 1, v_oe);
      put(r + 1, q, v_eo);
      put(r + 1, q + 1, v_ee);
    }
    __syncthreads();
  }
  // ---- sweep 1 on the tile + ring 1, restricted to [2, n-3] (all that sweep 2 reads).
  Coefs<MASKED, STORED> kc;
  if (!MASKED && !STORED) kc.load(L, 0, nullptr, 0);
  {
    const int j = j0 - 1 + r1, i = i0 + tx;
    const double *sp = &S.xs[r1 + 1][tx + 3];
    const int8_t *mp = &S.ms[r1 + 1][tx + 2];
    double *yp = &S.y1[r1][tx + 1];
    auto out = [&](int k, double val) { yp[k * YW] = val; };
    auto bget = [&](int k) { return bv[k]; };
    if (inner)
      jacobi_strip<MASKED, STORED, ZERO, false, RG + 1, XP, XW>(L, kc, sp, bget, mp, n1, j, i, 2, out);
    else
      jacobi_strip<MASKED, STORED, ZERO, true, RG + 1, XP, XW>(L, kc, sp, bget, mp, n1, j, i, 2, out);
    if (t < YH * 2) {   // ring columns 0 and YW-1 of the y1 tile
      const int r = t >> 1, q = (t & 1) ? YW - 1 : 0;
      double *y = &S.y1[r][q];
      auto out1 = [&](int, double val) { *y = val; };
      auto bget1 = [&](int) { return bx; };
      jacobi_strip<MASKED, STORED, ZERO, true, 1, XP, XW>(L, kc, &S.xs[r + 1][q + 2], bget1, &S.ms[r + 1][q + 1], 1,
                                                          j0 - 1 + r, i0 - 1 + q, 2, out1);
    }
  }
  __syncthreads();
  // ---- sweep 2 on the tile interior; tiles that touch the rim also store halo images
  {
    const bool rim = (j0 < 2 * NH) || (i0 < 2 * NH) || (j0 + TY > ny - 2 * NH) || (i0 + TX > nx - 2 * NH);
    const int r0 = tg * RG;
    const int j = j0 + r0, i = i0 + tx;
    double *base = acc ? acc : xout;
    double *dst = base + (size_t)j * nx + i;
    const bool accum = acc != nullptr;
    double *bsouth_p = PEER ? f2d::peer_addr(base, P.south_off) : nullptr;
    double *bnorth_p = PEER ? f2d::peer_addr(base, P.north_off) : nullptr;
    const int m2 = ny - 2 * NH;
    auto out = [&](int k, double val) {
      double *d = dst + (size_t)k * nx;
      // solve(): `x += self.x[0]` (hierarchy.py:171) fused into the last kernel of the
      // F-cycle: the correction is added to psi (acc) instead of being stored
      const double nv = accum ? *d + val : val;
      *d = nv;
      if (rim)
        f2d::for_each_halo_image(j + k, i, ny, nx, NH, [&](int j2, int i2) {
          double *e = base + (size_t)j2 * nx + i2;
          *e = accum ? *e + val : val;
        }, L.ywrap != 0);
      if (PEER) {
        // rows NH..2NH-1 are the south rank's top halo, rows m2..m2+NH-1 the north rank's
        // bottom halo (x images included, so the corners arrive too)
        const int jr = j + k;
        if (bsouth && jr < 2 * NH) {
          bsouth_p[(size_t)(jr + m2) * nx + i] = nv;
          f2d::for_each_halo_image(jr + m2, i, ny, nx, NH, [&](int j2, int i2) { bsouth_p[(size_t)j2 * nx + i2] = nv; }, false);
        }
        if (bnorth && jr >= m2) {
          bnorth_p[(size_t)(jr - m2) * nx + i] = nv;
          f2d::for_each_halo_image(jr - m2, i, ny, nx, NH, [&](int j2, int i2) { bnorth_p[(size_t)j2 * nx + i2] = nv; }, false);
        }
      }
    };
    const double *sp = &S.y1[r0 + 1][tx + 1];
    const int8_t *mp = &S.ms[r0 + 2][tx + 2];
    // tile row k of this strip is row k + 1 of the first group's sweep-1 strip (which starts on
    // the ring), row k of the others'
    auto bget = [&](int k) { return tg == 0 ? bv[k + 1] : bv[k]; };
    if (inner)
      jacobi_strip<MASKED, STORED, false, false, RG, YW, XW>(L, kc, sp, bget, mp, RG, j, i, NH, out);
    else
      jacobi_strip<MASKED, STORED, false, true, RG, YW, XW>(L, kc, sp, bget, mp, RG, j, i, NH, out);
  }
  if (PEER && (bsouth || bnorth)) f2d::peer_done(P, gridDim.x * (gridDim.y == 1 ? 1u : 2u));
}

// ---------------------------------------------------------------------------
// k_resid_restrict: coarse tile RTX x RTY, fine residual tile (2RTX+1) x (2RTY+1)
// ---------------------------------------------------------------------------
constexpr int RTX = 32, RTY = 16;
constexpr int RTYS = TYS / 2;   // small-level tiles
constexpr int RW = 2 * RTX + 1, RH = 2 * RTY + 1;  // residual tile
constexpr int RXW = RW + 2, RXH = RH + 2;          // x tile

// residual at an arbitrary fine cell straight from global memory (ring cells whose
// periodic source lies far from the tile)
template <bool MASKED, bool STORED>
__device__ __noinline__ double resid_global(const LevelK &L, const double *__restrict__ x, const double *__restrict__ b,
                                            int j, int i) {
  int nx = L.nx;
  size_t g = (size_t)j * nx + i;
  if (MASKED && L.msk[g] == 0) return 0.;
  Coefs<MASKED, STORED> k;
  k.load(L, g, MASKED ? L.msk + g : nullptr, nx);
  double cdiag = STORED ? __ldg(L.A + 4 * (size_t)L.ny * nx + g) : L.c[4];
  return resid_val<MASKED, STORED>(L, k, cdiag, x[g - nx - 1], x[g - nx], x[g - nx + 1], x[g - 1], x[g], x[g + 1],
                                   x[g + nx - 1], x[g + nx], x[g + nx + 1], b[g]);
}

constexpr int RXP = RXW + 1, RBP = RW + 1;         // row pitches of the x / b tiles (even: TMA boxes)
template <int RTYP>
struct ResidSmemT {
  static constexpr int RH = 2 * RTYP + 1, RXH = RH + 2;
  alignas(128) double xs[RXH][RXP];
  // b tile, overwritten IN PLACE by the residual: the thread that computes r(j,i) is the only
  // reader of b(j,i), so r(r,q) takes the place of b at bs[r][q + 1] (row pitch RBP)
  alignas(128) double bs[RH][RBP];
  alignas(8) uint64_t bar;
  int8_t ms[RXH][RXW];     // last: not allocated for the mask-free instantiations
};
using ResidSmem = ResidSmemT<RTY>;
template <int RTYP>
constexpr size_t resid_smem_nomask() { return offsetof(ResidSmemT<RTYP>, ms); }

// column strip of the residual tile (same register-window scheme as jacobi_strip)
template <bool MASKED, bool STORED, bool GUARD, int NR>
__device__ __forceinline__ void resid_strip(const LevelK &L, const Coefs<MASKED, STORED> &kc, const double *sp,
                                            const double *bp, const int8_t *mp, double *rp, int nr, int j, int i,
                                            const double *__restrict__ x, const double *__restrict__ b) {
  const int ny = L.ny, nx = L.nx;
  double a0 = sp[-RXP - 1], a1 = sp[-RXP], a2 = sp[-RXP + 1];
  double m0 = sp[-1], m1 = sp[0], m2 = sp[1];
  size_t g = (size_t)j * nx + i;
  constexpr bool MWIN = MASKED && !STORED;   // 3x3 mask window in registers (see jacobi_strip)
  int wa0 = 0, wa1 = 0, wa2 = 0, wm0 = 0, wm1 = 0, wm2 = 0;
  if (MWIN) {
    wa0 = mp[-RXW - 1]; wa1 = mp[-RXW]; wa2 = mp[-RXW + 1];
    wm0 = mp[-1]; wm1 = mp[0]; wm2 = mp[1];
  }
  // stored coefficients: all rows' loads first (see jacobi_strip); rows this strip evaluates in
  // place, i.e. inside [NH, n-NH) on GUARD tiles
  constexpr bool PRE = STORED && NR <= 4;
  Coefs<MASKED, STORED> kpre[PRE ? NR : 1];
  if (PRE) {
#pragma unroll
    for (int k = 0; k < NR; k++) {
      const bool okk = k < nr && (!GUARD || (j + k < ny - NH && i < nx - NH));
      if (okk) kpre[k].template load_stored<true>(L, g + (size_t)k * nx);
    }
  }
#pragma unroll
  for (int k = 0; k < NR; k++) {
    if (k < nr) {
      double h0 = sp[(k + 1) * RXP - 1], h1 = sp[(k + 1) * RXP], h2 = sp[(k + 1) * RXP + 1];
      int wh0 = 0, wh1 = 0, wh2 = 0;
      if (MWIN) { wh0 = mp[(k + 1) * RXW - 1]; wh1 = mp[(k + 1) * RXW]; wh2 = mp[(k + 1) * RXW + 1]; }
      double val = 0.;
      const int jj = j + k;
      if (GUARD && (jj > ny - NH || i > nx - NH)) {
        val = 0.;
      } else if (GUARD && (jj == ny - NH || i == nx - NH)) {
        // first halo ring on the high side: the reference reads the halo-filled residual
        // there, i.e. the residual of the periodic source cell
        // (on a y-slab the halo row already holds the neighbour's data: evaluate in place)
        val = resid_global<MASKED, STORED>(L, x, b, L.ywrap ? f2d::wrap_src(jj, ny, NH) : jj, f2d::wrap_src(i, nx, NH));
      } else if (!MASKED || (MWIN ? wm1 : (int)mp[k * RXW]) != 0) {
        Coefs<MASKED, STORED> kk;
        if (MWIN) kk.from_window(L, wa0, wa1, wa2, wm0, wm2, wh0, wh1, wh2);
        else if (PRE) kk = kpre[k];
        else if (STORED) kk.template load_stored<true>(L, g + (size_t)k * nx);
        else kk = kc;
        // (the residual uses the diagonal itself, not omega / |diagonal|: load_stored<true> leaves it in c3)
        double cdiag = STORED ? kk.c3 : L.c[4];
        val = resid_val<MASKED, STORED>(L, kk, cdiag, a0, a1, a2, m0, m1, m2, h0, h1, h2, bp[k * RBP]);
      }
      rp[k * RBP] = val;
      a0 = m0; a1 = m1; a2 = m2;
      m0 = h0; m1 = h1; m2 = h2;
      if (MWIN) { wa0 = wm0; wa1 = wm1; wa2 = wm2; wm0 = wh0; wm1 = wh1; wm2 = wh2; }
    }
  }
}

template <bool MASKED, bool STORED, bool PEER, int RTYP>
__global__ void __launch_bounds__(NT, (STORED && RTYP <= 4) ? 2 : (MASKED ? 4 : 5))   // short stored strips: coefficients in registers
k_resid_restrict(LevelK L, const double *__restrict__ x, const double *__restrict__ b, double *__restrict__ bc,
                 const int8_t *__restrict__ mskc, int nyc, int nxc, f2d::Peer P, int use_tma,
                 const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmb) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // coarse tile height RTYP (shadowing the constants of the standard tile); RG rows per thread group
  constexpr int RTY = RTYP, RH = 2 * RTY + 1, RXH = RH + 2, RG = RTY / 2;
  static_assert(RTY % 2 == 0 && RH <= NT, "tile height");
  ResidSmemT<RTYP> &S = *reinterpret_cast<ResidSmemT<RTYP> *>(smem_raw);
  const int ny = L.ny, nx = L.nx;
  const int t = threadIdx.x;
  const int by = PEER ? f2d::peer_tile_row(blockIdx.y, gridDim.y) : blockIdx.y;
  const bool bsouth = PEER && by == 0, bnorth = PEER && by == (int)gridDim.y - 1;
  f2d::pdl_trigger();
  if (use_tma && t == 0) {
    f2d::mbar_init(&S.bar, 1);
    f2d::tma_prefetch_desc(&tmx);
    f2d::tma_prefetch_desc(&tmb);
  }
  f2d::pdl_wait();
  if (PEER && (bsouth || bnorth)) f2d::peer_wait(P, bsouth, bnorth);
  const int ci0 = NH + blockIdx.x * RTX, cj0 = NH + by * RTY;  // first coarse output
  const int fi0 = 2 * ci0 - 3, fj0 = 2 * cj0 - 3;                      // first fine residual point
  // inner tile: the whole residual tile lies strictly inside the fine interior
  const bool inner = (fj0 + RH <= ny - NH) && (fi0 + RW <= nx - NH);
  if (use_tma) {
    __syncthreads();
    if (t == 0) {
      if (PEER && (bsouth || bnorth)) asm volatile("fence.proxy.async;" ::: "memory");
      f2d::mbar_expect_tx(&S.bar, (RXH * RXP + RH * RBP) * 8);
      f2d::tma_load_2d(&S.xs[0][0], &tmx, &S.bar, fi0 - 1, fj0 - 1);
      f2d::tma_load_2d(&S.bs[0][0], &tmb, &S.bar, fi0 - 1, fj0);   // even column: the tile sits at column offset 1
    }
    if (MASKED)
      load_tile_i8<RXH, RXW, RXW>(&S.ms[0][0], L.msk + (size_t)(fj0 - 1) * nx + (fi0 - 1), nx, ny - (fj0 - 1), nx - (fi0 - 1));
    f2d::mbar_wait(&S.bar, 0);
  } else {
    const double *xsrc = x + (size_t)(fj0 - 1) * nx + (fi0 - 1);
    const double *bsrc = b + (size_t)fj0 * nx + fi0;
    const int vr = ny - (fj0 - 1), vc = nx - (fi0 - 1);
    if (inner) {
      load_tile<RXH, RXW, RXP, true>(&S.xs[0][0], xsrc, nx, 0, 0);
      load_tile<RH, RW, RBP, true>(&S.bs[0][1], bsrc, nx, 0, 0);
    } else {
      load_tile<RXH, RXW, RXP, false>(&S.xs[0][0], xsrc, nx, vr, vc);
      load_tile<RH, RW, RBP, false>(&S.bs[0][1], bsrc, nx, vr - 1, vc - 1);
    }
    if (MASKED) load_tile_i8<RXH, RXW, RXW>(&S.ms[0][0], L.msk + (size_t)(fj0 - 1) * nx + (fi0 - 1), nx, vr, vc);
    cp_async_wait_all();
  }
  __syncthreads();
  // ---- fine residual on the (2RTX+1) x (2RTY+1) tile: column strips (rows 9,8,8,8) with
  // the 3x3 window in registers; the last column point-wise
  Coefs<MASKED, STORED> kc;
  if (!MASKED && !STORED) kc.load(L, 0, nullptr, 0);
  {
    const int tx = t & 63, tg = t >> 6;
    const int r0 = tg * RG + (tg < 1 ? 0 : 1), nr = tg < 1 ? RG + 1 : RG;
    if (inner)
      resid_strip<MASKED, STORED, false, RG + 1>(L, kc, &S.xs[r0 + 1][tx + 1], &S.bs[r0][tx + 1], &S.ms[r0 + 1][tx + 1],
                                                 &S.bs[r0][tx + 1], nr, fj0 + r0, fi0 + tx, x, b);
    else
      resid_strip<MASKED, STORED, true, RG + 1>(L, kc, &S.xs[r0 + 1][tx + 1], &S.bs[r0][tx + 1], &S.ms[r0 + 1][tx + 1],
                                                &S.bs[r0][tx + 1], nr, fj0 + r0, fi0 + tx, x, b);
    if (t < RH) {
      const int r = t, q = RW - 1;
      resid_strip<MASKED, STORED, true, 1>(L, kc, &S.xs[r + 1][q + 1], &S.bs[r][q + 1], &S.ms[r + 1][q + 1], &S.bs[r][q + 1], 1,
                                           fj0 + r, fi0 + q, x, b);
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
      const double *c = &S.bs[2 * r + 1][2 * q + 2];  // centre in the residual tile (column offset 1)
      val = 0.25 * c[0] + 0.125 * (((c[-1] + c[1]) + c[-RBP]) + c[RBP]) +
            0.0625 * (((c[-RBP - 1] + c[-RBP + 1]) + c[RBP - 1]) + c[RBP + 1]);
    }
    bc[g] = val;
    if (rim)
      f2d::for_each_halo_image(j, i, nyc, nxc, NH, [&](int jj, int ii) { bc[(size_t)jj * nxc + ii] = val; },
                               L.ywrap != 0);
    if (PEER) {
      const int m2c = nyc - 2 * NH;
      if (bsouth && j < 2 * NH) {
        double *q2 = f2d::peer_addr(bc, P.south_off);
        q2[(size_t)(j + m2c) * nxc + i] = val;
        f2d::for_each_halo_image(j + m2c, i, nyc, nxc, NH, [&](int jj, int ii) { q2[(size_t)jj * nxc + ii] = val; }, false);
      }
      if (bnorth && j >= m2c) {
        double *q2 = f2d::peer_addr(bc, P.north_off);
        q2[(size_t)(j - m2c) * nxc + i] = val;
        f2d::for_each_halo_image(j - m2c, i, nyc, nxc, NH, [&](int jj, int ii) { q2[(size_t)jj * nxc + ii] = val; }, false);
      }
    }
  }
  if (PEER && (bsouth || bnorth)) f2d::peer_done(P, gridDim.x * (gridDim.y == 1 ? 1u : 2u));
}

// ---------------------------------------------------------------------------
// k_zsmooth_resid_restrict: the whole visit of a level on the way down, below the level a
// cycle starts from (hierarchy.py:100-107: x = 0; smooth(x, b, npre = 1); residual; restrict)
//     t  = S2(0, b)          (double sweep from a zero first guess)
//     bc = R(b - A t)
// in ONE kernel that reads b once: the first sweep of a zero guess is pointwise
// (y = 0*c2 + c3*(0 - b) = -(c3*b), exactly), so sweep 2 forms its 3x3 window of y from the b tile
// on the fly; t stays in shared memory for the residual, the residual replaces b in place, the
// restriction reads it from there.  Replaces k_smooth2<INPUT = 1> + k_resid_restrict (two
// launches, t and b read back through TMA) on ALL-FLUID DOUBLY PERIODIC levels, where a halo
// cell evaluated in place from halo-filled inputs IS the periodic image of its source, bit for
// bit -- so no tile needs a guard: for every tile the b tile (residual tile + 2 rings) lies
// inside the array (rows / columns 1 .. n-1).
// Coarse tile RTX x RTY at (cj0, ci0); fine residual tile RH x RW from (fj0, fi0) = 2*(cj0, ci0) - 3;
// t tile = + 1 ring, b tile = + 2 rings.
// ---------------------------------------------------------------------------
constexpr int ZBW = RW + 4, ZBP = ZBW + 1;   // b tile 37 x 69 (standard tile); the TMA box starts one column early (even)
constexpr int ZTW = RW + 2, ZTP = ZTW + 1;   // t tile 35 x 67
static_assert(ZBP % 2 == 0, "TMA boxes need an even number of doubles per row");
template <int RTYP>
struct ZrrSmemT {
  static constexpr int RH = 2 * RTYP + 1, ZBH = RH + 4, ZTH = RH + 2;
  alignas(128) double bs[ZBH][ZBP];   // b tile (ring 2): tile column c sits at bs[.][c + 1]; later the residual in place
  alignas(128) double ts[ZTH][ZTP];   // t tile (ring 1)
  alignas(8) uint64_t bar;
};
template <int RTYP>
constexpr int zrr_bh() { return 2 * RTYP + 5; }
// t at t-tile point (rho, c) from the b tile: window of y = -(c3*b) at b-tile rows rho..rho+2,
// columns c..c+2, right-hand side b-tile (rho+1, c+1)
__device__ __forceinline__ double zsmooth_point(const LevelK &L, const Coefs<false, false> &kc, const double (*bs)[ZBP],
                                                int rho, int c) {
  const double nc3 = -kc.c3;
  const double *p = &bs[rho][c + 1];
  return jacobi_val<false, false>(L, kc, nc3 * p[0], nc3 * p[1], nc3 * p[2], nc3 * p[ZBP], nc3 * p[ZBP + 1],
                                  nc3 * p[ZBP + 2], nc3 * p[2 * ZBP], nc3 * p[2 * ZBP + 1], nc3 * p[2 * ZBP + 2],
                                  p[ZBP + 1]);
}
// PEER (y-slab levels): the y halo rows of b hold the neighbouring ranks' rows (pushed by the
// kernel that produced b), so the halo cells the tile evaluates in place are the neighbour's
// values bit for bit; the tiles of the first / last tile row wait for the neighbours' epoch
// first, push the 3 outermost rows of t and of bc into the neighbours' halo rows and publish
// the epoch -- ONE synchronisation point where the two-kernel form has two.
template <int RTYP, bool PEER>
__global__ void __launch_bounds__(NT, 4)
k_zsmooth_resid_restrict(LevelK L, double *__restrict__ tout, double *__restrict__ bc, int nyc, int nxc,
                         f2d::Peer P, const __grid_constant__ CUtensorMap tmb) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int RTY = RTYP, RH = 2 * RTY + 1, ZBH = RH + 4, ZTH = RH + 2;
  constexpr int TG = (ZTH + 3) / 4;    // t-tile rows per thread group (9, 9, 9, 8 of 35)
  constexpr int RG = RTY / 2;          // residual rows per thread group (9, 8, 8, 8 of 33)
  static_assert(3 * ZTH <= NT && RH <= NT, "tile height");
  ZrrSmemT<RTYP> &S = *reinterpret_cast<ZrrSmemT<RTYP> *>(smem_raw);
  const int ny = L.ny, nx = L.nx;
  const int t = threadIdx.x;
  const int by = PEER ? f2d::peer_tile_row(blockIdx.y, gridDim.y) : blockIdx.y;
  const bool bsouth = PEER && by == 0, bnorth = PEER && by == (int)gridDim.y - 1;
  const int ci0 = NH + blockIdx.x * RTX, cj0 = NH + by * RTY;
  const int fi0 = 2 * ci0 - 3, fj0 = 2 * cj0 - 3;
  f2d::pdl_trigger();
  if (t == 0) {
    f2d::mbar_init(&S.bar, 1);
    f2d::tma_prefetch_desc(&tmb);
  }
  f2d::pdl_wait();
  if (PEER && (bsouth || bnorth)) f2d::peer_wait(P, bsouth, bnorth);
  __syncthreads();
  if (t == 0) {
    if (PEER && (bsouth || bnorth)) asm volatile("fence.proxy.async;" ::: "memory");   // the neighbours' rows were observed through the generic proxy
    f2d::mbar_expect_tx(&S.bar, ZBH * ZBP * 8);
    f2d::tma_load_2d(&S.bs[0][0], &tmb, &S.bar, fi0 - 3, fj0 - 2);
  }
  Coefs<false, false> kc;
  kc.load(L, 0, nullptr, 0);
  const double nc3 = -kc.c3;
  const int tx = t & 63, tg = t >> 6;
  f2d::mbar_wait(&S.bar, 0);
  // ---- t = S2(0, b) on the t tile (35 x 67): column strips of 9, 9, 9, 8 rows with the 3x3
  // window of y in registers; columns 64..66 point-wise.  The tile's own 32 x 64 points
  // (t-tile rows 1..32, columns 1..64) also go to global memory, with their halo images.
  const bool rimf = (fj0 < 2 * NH) || (fi0 < 2 * NH) || (fj0 + 2 * RTY > ny - 2 * NH) || (fi0 + 2 * RTX > nx - 2 * NH);
  double *tsouth = PEER ? f2d::peer_addr(tout, P.south_off) : nullptr;
  double *tnorth = PEER ? f2d::peer_addr(tout, P.north_off) : nullptr;
  const int mrows = ny - 2 * NH;   // interior rows of the slab
  auto put_t = [&](int rho, int c, double val) {
    S.ts[rho][c] = val;
    if (rho >= 1 && rho <= 2 * RTY && c >= 1 && c <= 2 * RTX) {
      const int j = fj0 - 1 + rho, i = fi0 - 1 + c;
      tout[(size_t)j * nx + i] = val;
      if (rimf)
        f2d::for_each_halo_image(j, i, ny, nx, NH, [&](int j2, int i2) { tout[(size_t)j2 * nx + i2] = val; }, L.ywrap != 0);
      if (PEER) {
        if (bsouth && j < 2 * NH) {
          tsouth[(size_t)(j + mrows) * nx + i] = val;
          f2d::for_each_halo_image(j + mrows, i, ny, nx, NH, [&](int j2, int i2) { tsouth[(size_t)j2 * nx + i2] = val; }, false);
        }
        if (bnorth && j >= mrows) {
          tnorth[(size_t)(j - mrows) * nx + i] = val;
          f2d::for_each_halo_image(j - mrows, i, ny, nx, NH, [&](int j2, int i2) { tnorth[(size_t)j2 * nx + i2] = val; }, false);
        }
      }
    }
  };
  {
    const int r0 = tg * TG, nr = (ZTH - r0 < TG) ? ZTH - r0 : TG;
    const double *p = &S.bs[r0][tx + 1];
    double a0 = nc3 * p[0], a1 = nc3 * p[1], a2 = nc3 * p[2];
    double m0 = nc3 * p[ZBP], m1 = nc3 * p[ZBP + 1], m2 = nc3 * p[ZBP + 2];
    double bm = p[ZBP + 1];   // b at the window centre
#pragma unroll
    for (int k = 0; k < TG; k++) {
      if (k < nr) {
        const double *q = p + (k + 2) * ZBP;
        const double bh = q[1];
        const double h0 = nc3 * q[0], h1 = nc3 * bh, h2 = nc3 * q[2];
        put_t(r0 + k, tx, jacobi_val<false, false>(L, kc, a0, a1, a2, m0, m1, m2, h0, h1, h2, bm));
        a0 = m0; a1 = m1; a2 = m2;
        m0 = h0; m1 = h1; m2 = h2;
        bm = bh;
      }
    }
    if (t < 3 * ZTH) {
      const int rho = t / 3, c = 64 + t % 3;
      put_t(rho, c, zsmooth_point(L, kc, S.bs, rho, c));
    }
  }
  __syncthreads();
  // ---- residual on the 33 x 65 tile, in place over b: residual-tile (rho, c) = t-tile centre
  // (rho+1, c+1) = b-tile (rho+2, c+2)
  {
    const int r0 = tg * RG + (tg < 1 ? 0 : 1), nr = tg < 1 ? RG + 1 : RG;
    const double cdiag = L.c[4];
    auto strip = [&](int rho0, int c, int n) {
      const double *p = &S.ts[rho0][c];
      double *bp = &S.bs[rho0 + 2][c + 3];
      double a0 = p[0], a1 = p[1], a2 = p[2];
      double m0 = p[ZTP], m1 = p[ZTP + 1], m2 = p[ZTP + 2];
#pragma unroll
      for (int k = 0; k < RG + 1; k++) {
        if (k < n) {
          const double *q = p + (k + 2) * ZTP;
          const double h0 = q[0], h1 = q[1], h2 = q[2];
          bp[k * ZBP] = resid_val<false, false>(L, kc, cdiag, a0, a1, a2, m0, m1, m2, h0, h1, h2, bp[k * ZBP]);
          a0 = m0; a1 = m1; a2 = m2;
          m0 = h0; m1 = h1; m2 = h2;
        }
      }
    };
    strip(r0, tx, nr);
    if (t < RH) strip(t, RW - 1, 1);
  }
  __syncthreads();
  // ---- full-weighting restriction (as k_resid_restrict), halo images on rim tiles
  const bool rim = (cj0 < 2 * NH) || (ci0 < 2 * NH) || (cj0 + RTY > nyc - 2 * NH) || (ci0 + RTX > nxc - 2 * NH);
#pragma unroll
  for (int pidx = t; pidx < RTY * RTX; pidx += NT) {
    const int r = pidx / RTX, q = pidx % RTX;
    const int j = cj0 + r, i = ci0 + q;
    const double *c = &S.bs[2 * r + 3][2 * q + 4];   // residual-tile (2r+1, 2q+1)
    const double val = 0.25 * c[0] + 0.125 * (((c[-1] + c[1]) + c[-ZBP]) + c[ZBP]) +
                       0.0625 * (((c[-ZBP - 1] + c[-ZBP + 1]) + c[ZBP - 1]) + c[ZBP + 1]);
    bc[(size_t)j * nxc + i] = val;
    if (rim)
      f2d::for_each_halo_image(j, i, nyc, nxc, NH, [&](int jj, int ii) { bc[(size_t)jj * nxc + ii] = val; },
                               L.ywrap != 0);
    if (PEER) {
      const int m2c = nyc - 2 * NH;
      if (bsouth && j < 2 * NH) {
        double *q2 = f2d::peer_addr(bc, P.south_off);
        q2[(size_t)(j + m2c) * nxc + i] = val;
        f2d::for_each_halo_image(j + m2c, i, nyc, nxc, NH, [&](int jj, int ii) { q2[(size_t)jj * nxc + ii] = val; }, false);
      }
      if (bnorth && j >= m2c) {
        double *q2 = f2d::peer_addr(bc, P.north_off);
        q2[(size_t)(j - m2c) * nxc + i] = val;
        f2d::for_each_halo_image(j - m2c, i, nyc, nxc, NH, [&](int jj, int ii) { q2[(size_t)jj * nxc + ii] = val; }, false);
      }
    }
  }
  if (PEER && (bsouth || bnorth)) f2d::peer_done(P, gridDim.x * (gridDim.y == 1 ? 1u : 2u));
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
