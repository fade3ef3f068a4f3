// Multigrid on y-slabs (included by f2d_multigrid.cu): the distributed counterpart of
// gmg/hierarchy.py + gmg/level.py run under mpirun, with
//   halo.fill (halo.py:214-292)              -> x images by the producing kernel + comm_exchange
//   Subdomains.gather (subdomains.py:105-115) -> comm_gather at ONE level (lg): from there on
//                                               every rank holds the whole (small) grid and
//                                               computes it redundantly with the single-GPU
//                                               hierarchy mg->L (CUDA graphs, tail kernel)
//   Subdomains.split (:118-124)               -> k_scatter_rows
//   allreduce of norms (level.py:401)         -> comm_allreduce
// The arithmetic of every cell is the single-GPU one (Jacobi is decomposition independent),
// so fields agree with a single-GPU run bit for bit; sums agree to summation order.
#pragma once

namespace {

// slab rows [0, ny_loc) of rank `rank` copied out of a replicated full-height array
__global__ void k_scatter_rows(const double *__restrict__ full, double *__restrict__ slab, int ny_loc, int nx, int row0) {
  const size_t total = (size_t)ny_loc * nx;
  for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < total; p += (size_t)gridDim.x * blockDim.x)
    slab[p] = full[(size_t)row0 * nx + p];
}

inline int xch1(f2d_mg *mg, Level &l, double *x, cudaStream_t s) {
  double *arr[1] = {x};
  return comm_exchange(mg->comm, arr, 1, NH, l.ny, l.nx, s);
}

// slab array of level lg -> replicated L[0] array (+ local periodic fill of the full array)
int gather_to_full(f2d_mg *mg, const double *slab, double *full, cudaStream_t s) {
  Level &v = mg->S[mg->lg];
  Level &f = mg->L[0];
  // lg == 0: nothing but gathers between two uses of `full` -> explicit barrier first
  // lg >= 1: the slab array comes from a kernel that wrote its x halo columns; the gather also
  // pushes the rows that are the periodic y halo of the replicated array -- no fill kernel
  const bool yimages = mg->lg >= 1;
  TRY(comm_gather(mg->comm, slab, full, v.ny, v.nx, NH, s, mg->lg == 0, yimages));
  if (yimages) return F2D_OK;
  return f2d_fill_halo(full, NH, f.ny, f.nx, (f2d_stream_t)s);
}
// this rank's rows of the replicated level-lg array, halo rows included, as a slab array: the
// full array is halo filled, so rows [row0, row0 + ny_loc) ARE the slab with its y halo -- the
// consumers of x at level lg (interpolation fused into the smoother above it) read it in place
// instead of through a copy (Subdomains.split, subdomains.py:118-124)
double *full_as_slab(f2d_mg *mg, double *full) {
  Level &v = mg->S[mg->lg];
  return full + (size_t)comm_rank(mg->comm) * (v.ny - 2 * NH) * v.nx;
}
int scatter_from_full(f2d_mg *mg, const double *full, double *slab, cudaStream_t s) {
  Level &v = mg->S[mg->lg];
  int row0 = comm_rank(mg->comm) * (v.ny - 2 * NH);
  k_scatter_rows<<<nblocks1d(v.n()), 256, 0, s>>>(full, slab, v.ny, v.nx, row0);
  F2D_LAUNCHED();
  return F2D_OK;
}

// V-cycle from slab level lev1 <= lg (hierarchy.py:98-127); x0/b0 are slab arrays of lev1
// acc != nullptr (lg >= 1): the last smoother adds its result to acc (psi) instead of storing x0
int slab_vcycle(f2d_mg *mg, int lev1, double *x0, double *b0, cudaStream_t s, int first_input = 0,
                double *acc = nullptr) {
  const int lg = mg->lg;
  auto X = [&](int g) { return g == lev1 ? x0 : (g == lg ? full_as_slab(mg, mg->L[0].x) : mg->S[g].x); };
  auto B = [&](int g) { return g == lev1 ? b0 : mg->S[g].b; };
  for (int g = lev1; g < lg; g++) {
    Level &l = mg->S[g], &c = mg->S[g + 1];
    int input = g > lev1 ? 1 : first_input;
    // the smoother and the residual/restriction kernels fill the neighbours' halo rows of
    // their outputs themselves (fused::k_smooth2<..., PEER>): no exchange kernels here
    if (input == 1 && zrr_ok(mg, l, c)) {   // smooth from zero + residual + restriction: one kernel, one epoch
      TRY(op_zsmooth_rr_L(mg, l, c, B(g), l.t, B(g + 1), s));
      continue;
    }
    TRY(smooth2_L(mg, l, &c, input, X(g), B(g), l.t, input == 2 ? X(g + 1) : nullptr, s));
    TRY(op_resid_restrict_L(mg, l, c, l.t, B(g), B(g + 1), s));
  }
  // levels >= lg: gathered, then the replicated single-GPU cycle
  Level &f = mg->L[0];
  TRY(gather_to_full(mg, B(lg), f.b, s));
  if (lev1 == lg && first_input == 0) {
    TRY(gather_to_full(mg, X(lg), f.x, s));
    TRY(vcycle_enqueue(mg, 0, f.x, f.b, s, 0));
  } else if (lev1 == lg && first_input == 2) {
    return fail(F2D_ERR_ARG, "slab_vcycle: interpolation input at the gather level is handled by slab_fcycle");
  } else {
    TRY(vcycle_enqueue(mg, 0, f.x, f.b, s, 1));
  }
  if (lev1 == lg) TRY(scatter_from_full(mg, f.x, X(lg), s));   // (levels above read f.x in place)
  for (int g = lg - 1; g >= lev1; g--) {
    Level &l = mg->S[g], &c = mg->S[g + 1];
    TRY(smooth2_L(mg, l, &c, 3, l.t, B(g), X(g), X(g + 1), s, g == lev1 ? acc : nullptr));
  }
  return F2D_OK;
}

// F-cycle from slab level 0 (hierarchy.py:131-151)
int slab_fcycle(f2d_mg *mg, double *x0, double *b0, cudaStream_t s, double *acc = nullptr) {
  const int lg = mg->lg;
  auto X = [&](int g) { return g == 0 ? x0 : (g == lg ? full_as_slab(mg, mg->L[0].x) : mg->S[g].x); };
  auto B = [&](int g) { return g == 0 ? b0 : mg->S[g].b; };
  for (int g = 0; g < lg; g++) {
    TRY(op_restrict_L(mg, mg->S[g], mg->S[g + 1], B(g), B(g + 1), s));   // fills the neighbours' halo rows itself
  }
  Level &f = mg->L[0];
  TRY(gather_to_full(mg, B(lg), f.b, s));
  TRY(fcycle_enqueue(mg, 0, f.x, f.b, s));
  if (lg == 0) TRY(scatter_from_full(mg, f.x, X(lg), s));   // (levels above read f.x in place)
  for (int g = lg - 1; g >= 0; g--)
    for (int k = 0; k < mg->nvcyc; k++)
      TRY(slab_vcycle(mg, g, X(g), B(g), s, k == 0 ? 2 : 0, (g == 0 && k == mg->nvcyc - 1) ? acc : nullptr));
  return F2D_OK;
}

int slab_cycle_enqueue(f2d_mg *mg, int kind, int lev1, double *x0, double *b0, cudaStream_t s, double *acc) {
  if (lev1 != 0) return fail(F2D_ERR_ARG, "slab multigrid: cycles start from level 0");
  if (!comm_owns(mg->comm, x0))
    return fail(F2D_ERR_ARG, "slab multigrid: psi must live in the symmetric heap (f2d_comm_alloc)");
  if (kind == 0) {
    TRY(slab_vcycle(mg, 0, x0, b0, s));
    TRY(slab_vcycle(mg, 0, x0, b0, s));
  } else if (kind == 1) {
    if (acc && (mg->lg == 0 || !comm_owns(mg->comm, acc)))
      return fail(F2D_ERR_ARG, "slab multigrid: fused accumulation needs a slab level and psi in the symmetric heap");
    TRY(slab_fcycle(mg, x0, b0, s, acc));
  } else {
    TRY(slab_vcycle(mg, 0, x0, b0, s));
  }
  // the last smoother published its halo rows without waiting: kernels outside the
  // protocol (psi += x, the velocity operators) read them, so wait for the neighbours here
  return comm_drain(mg->comm, s);
}

// Gmg.solve on slabs (hierarchy.py:154-192): norms are all-reduced, so every rank takes
// the same number of F-cycles
int slab_solve(f2d_mg *mg, double *psi, const double *rhs, double tol, int maxite, int *nite_out, double *res_out,
               cudaStream_t s) {
  Level &l = mg->S[0];
  if (!comm_owns(mg->comm, psi))
    return fail(F2D_ERR_ARG, "slab multigrid: psi must live in the symmetric heap (f2d_comm_alloc)");
  f2d_stream_t stream = (f2d_stream_t)s;
  TRY(f2d_computenorm(l.msk, rhs, NH, l.ny, l.nx, mg->dscal, mg->scratch, stream));
  TRY(op_resid_sumsq_L(mg, l, psi, rhs, l.b, mg->dscal + 1, s));
  TRY(xch1(mg, l, l.b, s));
  TRY(comm_allreduce(mg->comm, mg->dscal, 2, 0u, s));
  TRY(read_scalars(mg, 2, s));
  double normb = sqrt(mg->hscal[0]);
  int nite = 0;
  double res = 0.;
  if (normb > 0) {
    double res0 = sqrt(mg->hscal[1]) / normb;
    res = res0;
    int ndiv = 0;
    const bool fuse_add = mg->lg >= 1;   // `x += self.x[0]` done by the F-cycle's last kernel
    while (nite < maxite && res0 > tol) {
      TRY(run_cycle(mg, 1, 0, l.x, l.b, s, fuse_add ? psi : nullptr));
      if (!fuse_add) {
        k_add_inplace<<<nblocks1d(l.n()), 256, 0, s>>>(psi, l.x, l.n());
        F2D_LAUNCHED();
      }
      TRY(op_resid_sumsq_L(mg, l, psi, rhs, l.b, mg->dscal + 1, s));
      TRY(xch1(mg, l, l.b, s));
      TRY(comm_allreduce(mg->comm, mg->dscal + 1, 1, 0u, s));
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

}  // namespace

extern "C" int f2d_mg_create_slab(f2d_mg_t **out, f2d_comm_t *comm, const double *cornermask, int ny_loc, int nx,
                                  double dx, double dy, double omega, double hydroepsilon, double Rd,
                                  f2d_stream_t stream) {
  if (!comm || comm_size(comm) == 1) return f2d_mg_create(out, cornermask, ny_loc, nx, dx, dy, omega, hydroepsilon, Rd, stream);
  if (!out || !cornermask) return fail(F2D_ERR_ARG, "mg_create_slab: null pointer");
  const int G = comm_size(comm), rank = comm_rank(comm);
  const int ml = ny_loc - 2 * NH, n = nx - 2 * NH, m = ml * G;
  if (ml < 8 || n < 4) return fail(F2D_ERR_ARG, "mg_create_slab: slab too small");
  if ((ml & (ml - 1)) || (n & (n - 1))) return fail(F2D_ERR_ARG, "mg_create_slab: local sizes must be powers of two");
  if (hydroepsilon * dy / dx <= 0.2)
    return fail(F2D_ERR_ARG, "mg_create_slab: small aspect ratio needs the tridiagonal relaxation (not built yet)");
  cudaStream_t s = S(stream);
  f2d_mg *mg = new f2d_mg();
  mg->omega = omega;
  if (const char *ng = getenv("F2D_MG_NO_GRAPHS")) mg->graphs = !(ng[0] == '1');
  if (const char *tm = getenv("F2D_MG_TMA")) mg->tma = !(tm[0] == '0');
  mg->comm = comm;
  auto bail = [&](int rc) { f2d_mg_destroy(mg); return rc; };
  std::vector<std::pair<int, int>> sizes;   // global (m, n) per level
  int rc = level_sizes(m, n, sizes);
  if (rc != F2D_OK) return bail(rc);
  // number of distributed levels: while the level has more than min_cells global cells and
  // the local slab keeps an even number (>= 8) of rows
  long long min_cells = 1LL << 20;
  if (const char *e = getenv("F2D_SLAB_MIN_CELLS")) min_cells = atoll(e);
  int lg = 0;
  while (lg < (int)sizes.size() - 1) {
    long long cells = (long long)sizes[lg].first * sizes[lg].second;
    int rows = sizes[lg].first / G;
    if (cells <= min_cells || rows < 8 || (rows & 1)) break;
    lg++;
  }
  mg->lg = lg;
  for (int g = 0; g <= lg; g++) {   // slab levels 0..lg-1 and the slab view of level lg
    Level l;
    l.ny = sizes[g].first / G + 2 * NH;
    l.nx = sizes[g].second + 2 * NH;
    l.ywrap = 0;
    mg->S.push_back(l);
  }
  for (size_t g = lg; g < sizes.size(); g++) {   // replicated hierarchy: global level lg and below
    Level l;
    l.ny = sizes[g].first + 2 * NH;
    l.nx = sizes[g].second + 2 * NH;
    mg->L.push_back(l);
  }
  if ((rc = common_init(mg)) != F2D_OK) return bail(rc);
  for (auto &l : mg->S)
    if ((rc = alloc_level(mg, l, true, s)) != F2D_OK) return bail(rc);
  for (size_t k = 0; k < mg->L.size(); k++)
    if ((rc = alloc_level(mg, mg->L[k], k == 0, s)) != F2D_OK) return bail(rc);   // L[0].x/b: gather targets
  dim3 blk(32, 8);
  // ---- slab chain: level 0 from the local corner mask, Galerkin coarsening down to lg
  Level &s0 = mg->S[0];
  double *A9 = (double *)comm_alloc(comm, 9 * s0.n() * sizeof(double));
  if (!A9) return bail(fail(F2D_ERR_ARG, "mg_create_slab: symmetric heap exhausted (raise the arena size)"));
  k_mask_from_double<<<nblocks1d(s0.n()), 256, 0, s>>>(cornermask, s0.msk, s0.n());
  k_finest_matrix<<<grid2d(s0.ny, s0.nx, blk), blk, 0, s>>>(s0.msk, A9, finest_stencil(dx, dy, hydroepsilon), s0.ny, s0.nx);
  g_launches += 2;
  for (int k = 0; k < 9; k++)
    if ((rc = level_fill(mg, s0, A9 + k * s0.n(), s)) != F2D_OK) return bail(rc);
  cudaMemcpyAsync(s0.A, A9, 5 * s0.n() * sizeof(double), cudaMemcpyDeviceToDevice, s);
  for (int g = 1; g <= lg; g++) {
    Level &p = mg->S[g - 1], &c = mg->S[g];
    double *A9c = (double *)comm_alloc(comm, 9 * c.n() * sizeof(double));
    if (!A9c) return bail(fail(F2D_ERR_ARG, "mg_create_slab: symmetric heap exhausted (raise the arena size)"));
    if ((rc = coarsen_level(mg, p, c, A9, A9c, s)) != F2D_OK) return bail(rc);
    cudaMemcpyAsync(c.A, A9c, 5 * c.n() * sizeof(double), cudaMemcpyDeviceToDevice, s);
    A9 = A9c;
  }
  // ---- gather level lg (mask + 9 matrix planes) onto every rank, then the replicated chain
  {
    Level &v = mg->S[lg], &f = mg->L[0];
    double *A9full = (double *)comm_alloc(comm, 9 * f.n() * sizeof(double));
    int8_t *mfull = (int8_t *)comm_alloc(comm, f.n());
    if (!A9full || !mfull) return bail(fail(F2D_ERR_ARG, "mg_create_slab: symmetric heap exhausted (raise the arena size)"));
    for (int k = 0; k < 9; k++) {
      if ((rc = comm_gather(comm, A9 + k * v.n(), A9full + k * f.n(), v.ny, v.nx, NH, s)) != F2D_OK) return bail(rc);
      if ((rc = f2d_fill_halo(A9full + k * f.n(), NH, f.ny, f.nx, stream)) != F2D_OK) return bail(rc);
    }
    if ((rc = comm_gather_i8(comm, v.msk, mfull, v.ny, v.nx, NH, s)) != F2D_OK) return bail(rc);
    // the corner mask is not halo filled in the reference (hierarchy.py:46): the first /
    // last rank contribute their outer halo rows, the x halo columns travel with the rows
    cudaMemcpyAsync(f.msk, mfull, f.n(), cudaMemcpyDeviceToDevice, s);
    (void)rank;
    if ((rc = build_replicated(mg, A9full, false, s)) != F2D_OK) return bail(rc);
  }
  if ((rc = finish_setup(mg, Rd, s)) != F2D_OK) return bail(rc);
  *out = mg;
  return F2D_OK;
}
