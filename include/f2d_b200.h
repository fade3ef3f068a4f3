/*
 * f2d_b200.h -- C ABI of libf2d_b200.so, the B200 (sm_100a) implementation of
 * Fluid2d's per-timestep hot path.
 *
 * This is the drop-in boundary: the entry points below are what the reference's
 * f2py binding for this path binds (the five modules built by build.py:12-39) plus
 * the Python-level operators that only make sense fused on a device
 * (gmg.Gmg.twoVcycle / solve, Operators.invert_vorticity, Timescheme combinations).
 * Each declaration cites the reference interface it replaces.
 *
 * Conventions
 *  - every field is a DEVICE pointer to a row-major [ny][nx] array of double, x
 *    contiguous, halo of width nh = 3 included (ny = nyl, nx = nxl of grid.py:29-30);
 *    this is the logical layout of the reference's numpy arrays.  Masks are int8
 *    (0 = solid, 1 = fluid).  nx must be even (rows stay 16-byte aligned).
 *  - Fortran x(j,i) (1-based) is x[(j-1)*nx + (i-1)].
 *  - every call is asynchronous on the cudaStream_t passed as `stream` unless it
 *    returns a scalar to the host (documented per call); nothing allocates or frees
 *    caller memory; nothing calls exit()/stop.
 *  - return value: F2D_OK or an error code; f2d_last_error() gives the message.
 *  - scalar results go to a caller-provided DEVICE slot (double *out).
 *  - reductions need a DEVICE scratch of f2d_reduce_scratch_len() doubles.
 *  - one host thread per GPU (one process per GPU); a handle is not thread-safe.
 */
#ifndef F2D_B200_H
#define F2D_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define F2D_ABI_VERSION 1

#define F2D_OK 0
#define F2D_ERR_NH 1    /* nh != 3 (fortran_advection.f90:30-34 prints and STOPs) */
#define F2D_ERR_ARG 2   /* bad shape / order / method / null pointer */
#define F2D_ERR_CUDA 3  /* CUDA runtime error (message in f2d_last_error) */
#define F2D_ERR_DIVERGE 4 /* solver diverged (hierarchy.py:185-188 exit(0)) */

typedef void *f2d_stream_t;     /* cudaStream_t */
typedef struct f2d_mg f2d_mg_t; /* multigrid hierarchy (gmg.hierarchy.Gmg) */

int f2d_abi_version(void);
const char *f2d_last_error(void);
/* number of kernels this library launched since load / since the last reset */
long long f2d_launch_count(void);
void f2d_launch_count_reset(void);
/* Per-kernel accounting for bench.py's roofline table (no reference counterpart; the
 * reference times whole phases with core/timers.py).  Between f2d_prof_begin(stream) and
 * f2d_prof_report every kernel launched on `stream` is followed by a CUDA event; the report is
 * one line "name<TAB>launches<TAB>microseconds" per kernel (instantiation and grid size in
 * the name).  Run with CUDA graphs disabled (f2d_mg_set_graphs(h, 0)). */
int f2d_mg_bench_op(f2d_mg_t *mg, int kind, int level, int reps, f2d_stream_t stream);
/* ^ measurement aid (bench.py): `reps` back-to-back launches of ONE operator of the cycles on the
 * level's own arrays.  kind: 0 Grid.smooth, 1 Grid.smooth from x = 0, 2 x = I(xc) + smooth,
 * 3 x += I(xc) + smooth, 4 residual + restriction, 5 restriction, 6 residual + norm (level 0),
 * 7 / 8 the coarse-tail kernel's V-cycle / F-cycle (first tail level, f2d_mg_tail_level),
 * 9 the fused descent of a level (smooth from x = 0 + residual + restriction in one kernel). */
int f2d_mg_tail_level(const f2d_mg_t *mg);
int f2d_prof_begin(f2d_stream_t stream);
int f2d_prof_report(char *buf, size_t cap);

/* plain device-to-device copy / fill of nbytes (numpy slice assignments of the
 * reference's Python, e.g. hierarchy.py:212-217, euler.py:115) */
int f2d_copy(void *dst, const void *src, size_t nbytes, f2d_stream_t stream);
int f2d_zero(void *dst, size_t nbytes, f2d_stream_t stream);

/* ---- gmg/fortran_multigrid.f90:365-412  fillhalo(x,nh)  (Halo.fill, halo.py:136-141)
 * doubly periodic halo fill, corners included */
int f2d_fill_halo(double *x, int nh, int ny, int nx, f2d_stream_t stream);
int f2d_fill_halo_i8(int8_t *x, int nh, int ny, int nx, f2d_stream_t stream);
/* x images only (y-slab decomposition: the y halo rows come from f2d_comm_exchange_y) */
int f2d_fill_halo_x(double *x, int nh, int ny, int nx, f2d_stream_t stream);

/* ---- core/fortran_advection.f90:2-165 adv_upwind(msk,x,y,u,v,cst,nh,method,order)
 *      core/fortran_fluxes.f90:2-170 when xflx,yflx != NULL (face fluxes stored too)
 * dq(interior) = -div(U q); cst5 is a HOST array {dx,dy,0.05,umax,aparab}
 * (operators.py:99-119).  method 0 minmax / 1 parabolic; order 1,3,5.
 * fill_halo = 1 also performs the periodic halo fill of dq that Operators.rhs_adv
 * always does next (operators.py:231); fill_halo = 2 stores the x images only (y-slab
 * decomposition: follow with f2d_comm_exchange_y).  Same convention for the other
 * operators that take fill_halo. */
int f2d_adv_upwind(const int8_t *msk, const double *q, double *dq, const double *u,
                   const double *v, double *xflx, double *yflx, const double *cst5,
                   int nh, int method, int order, int ny, int nx, int fill_halo,
                   f2d_stream_t stream);
/* ---- core/fortran_advection.f90:169-284 adv_centered (order 2,4,6) */
int f2d_adv_centered(const int8_t *msk, const double *q, double *dq, const double *u,
                     const double *v, double *xflx, double *yflx, const double *cst5,
                     int nh, int method, int order, int ny, int nx, int fill_halo,
                     f2d_stream_t stream);
/* Operators.rhs_adv (core/operators.py:214-236: the loop over the tracer list) in ONE launch: the
 * tracers of the model share the velocity tiles.  q, dq (and xbase, xout) are HOST arrays of
 * ntracers device pointers.  xbase / xout both NULL, or both given: the kernel then also writes
 * the Runge-Kutta stage state xout[t] = xbase[t] + coef*dq[t], halo included
 * (core/timescheme.py:172-176; the product is rounded before the sum, as numpy does). */
int f2d_adv_multi(const int8_t *msk, const double *const *q, double *const *dq, int ntracers,
                  const double *u, const double *v, const double *cst5, int nh, int upwind, int method,
                  int order, const double *const *xbase, double *const *xout, double coef, int ny, int nx,
                  int fill_halo, f2d_stream_t stream);

/* ---- core/fortran_operators.f90 */
/* :44-64 celltocorner(xr,xp) */
int f2d_celltocorner(const double *xr, double *xp, int ny, int nx, f2d_stream_t stream);
/* :102-122 cornertocell(xp,xr) */
int f2d_cornertocell(const double *xp, double *xr, int ny, int nx, f2d_stream_t stream);
/* :2-39 computeorthogradient(msk,psi,dx,dy,nh,u,v) */
int f2d_orthogradient(const int8_t *msk, const double *psi, double dx, double dy, int nh,
                      double *u, double *v, int ny, int nx, f2d_stream_t stream);
/* psi = psi*mskp followed by computeorthogradient in one pass (operators.py:481,493).
 * msk == mskp == NULL: all-fluid domain (cell mask 1 everywhere, corner mask 1 except on
 * the last row and column, operators.py:59-67) -- no mask byte is read; needs even nx. */
int f2d_mask_orthogradient(const int8_t *msk, const int8_t *mskp, double *psi, double dx,
                           double dy, int nh, double *u, double *v, int ny, int nx,
                           f2d_stream_t stream);
/* the same, followed by the Runge-Kutta stage update of the velocities it has just derived
 * (Timescheme.RK3_SSP, timescheme.py:172-180: self.x[u] = x[u] + dt*dx0[u], then
 * x[u] + dt/4*(dx0[u] + dx1[u]); v alike): uo = ub + c*u (ue == ve == NULL) or ub + c*(ue + u),
 * over the whole arrays, with numpy's rounding sequence -- in the same kernel while u, v are in
 * registers (even nx, 16-byte aligned fields), or by f2d_ts_xpay / f2d_ts_xpay2 behind it. */
int f2d_mask_orthogradient_stage(const int8_t *msk, const int8_t *mskp, double *psi, double dx,
                                 double dy, int nh, double *u, double *v, const double *ub,
                                 const double *vb, const double *ue, const double *ve, double *uo,
                                 double *vo, double c, int ny, int nx, f2d_stream_t stream);
/* :125-156 add_diffusion(msk,trac,dx,nh,Kdiff,dtrac) (+ optional halo fill,
 * operators.py:300) */
int f2d_add_diffusion(const int8_t *msk, const double *trac, double dx, int nh,
                      double Kdiff, double *dtrac, int ny, int nx, int fill_halo,
                      f2d_stream_t stream);
/* :330-381 add_torque(msk,buoy,dx,nh,gravity,domega); premask != 0 first does
 * domega *= msk (operators.py:311); fill_halo as above (operators.py:313) */
int f2d_add_torque(const int8_t *msk, const double *buoy, double dx, int nh,
                   double gravity, double *domega, int ny, int nx, int premask,
                   int fill_halo, f2d_stream_t stream);
/* :221-277 computenoslipsourceterm(msk,x,y,dx,dy,nh): y is overwritten on rows/cols
 * 1..m-nh+1 exactly as the Fortran scatter does (restated as a gather); rows/cols
 * beyond keep their input values.  The Fortran's scalar `total` is not returned
 * (its only caller discards it, operators.py:256-274). */
int f2d_noslip_source(const int8_t *msknoslip, const double *psi, double *y, double dx,
                      double dy, int nh, int ny, int nx, f2d_stream_t stream);

/* ---- core/fortran_diag.f90 (+ computenorm/computeinner of fortran_multigrid.f90)
 * interior masked reductions; results written to DEVICE out[]; `scratch` is a DEVICE
 * buffer of f2d_reduce_scratch_len() doubles.  Deterministic (fixed two-stage tree);
 * the summation ORDER differs from the Fortran's sequential loop. */
size_t f2d_reduce_scratch_len(void);
/* :3-30 computedotprod -> out[0] (also computeinner, fortran_multigrid.f90:841) */
int f2d_computedotprod(const int8_t *msk, const double *x, const double *y, int nh,
                       int ny, int nx, double *out, double *scratch, f2d_stream_t stream);
/* :33-60 computemax -> out[0] */
int f2d_computemax(const int8_t *msk, const double *x, int nh, int ny, int nx, double *out,
                   double *scratch, f2d_stream_t stream);
/* :63-90 computesum -> out[0] */
int f2d_computesum(const int8_t *msk, const double *x, int nh, int ny, int nx, double *out,
                   double *scratch, f2d_stream_t stream);
/* :93-120 computesumandnorm -> out[0]=sum, out[1]=sum of squares */
int f2d_computesumandnorm(const int8_t *msk, const double *x, int nh, int ny, int nx,
                          double *out, double *scratch, f2d_stream_t stream);
/* :123-152 computenormmaxu -> out[0]=sum x^2, out[1]=max|x| on east faces */
int f2d_computenormmaxu(const int8_t *msk, const double *x, int nh, int ny, int nx,
                        double *out, double *scratch, f2d_stream_t stream);
/* :155-195 computekemaxu -> out[0]=ke, out[1]=maxu */
int f2d_computekemaxu(const int8_t *msk, const double *u, const double *v, int nh, int ny,
                      int nx, double *out, double *scratch, f2d_stream_t stream);
/* :198-236 computekemaxuv -> out[0]=ke, out[1]=maxu, out[2]=maxv */
int f2d_computekemaxuv(const int8_t *msk, const double *u, const double *v, int nh, int ny,
                       int nx, double *out, double *scratch, f2d_stream_t stream);
/* :239-267 computekewithpsi -> out[0] */
int f2d_computekewithpsi(const int8_t *msk, const double *omega, const double *psi, int nh,
                         int ny, int nx, double *out, double *scratch,
                         f2d_stream_t stream);
/* fortran_multigrid.f90:813-839 computenorm -> out[0] = sum of squares */
int f2d_computenorm(const int8_t *msk, const double *x, int nh, int ny, int nx, double *out,
                    double *scratch, f2d_stream_t stream);
/* grid.py:132-140 domain_integration: unmasked interior sum -> out[0] */
int f2d_domain_sum(const double *x, int nh, int ny, int nx, double *out, double *scratch,
                   f2d_stream_t stream);
/* euler.py:185-223 Euler.diagnostics in one pass:
 * out[0..7] = maxu, ke, sum w, sum w^2, sum w*xr, sum w*yr, sum psi, sum w*source */
int f2d_diag_euler(const int8_t *msk, const double *u, const double *v, const double *w,
                   const double *psi, const double *source, const double *xr,
                   const double *yr, int nh, int ny, int nx, double *out, double *scratch,
                   f2d_stream_t stream);

/* ---- core/timescheme.py:78-201 whole-state combinations (n = number of doubles).
 * Evaluated with the rounding sequence of the numpy expression cited (no FMA).   */
/* y += c*a                                   :80 (EF), :87, :200 */
int f2d_ts_axpy(double *y, double c, const double *a, size_t n, f2d_stream_t stream);
/* out = x + c*a                              :145, :156, :172, :188 ... */
int f2d_ts_xpay(double *out, const double *x, double c, const double *a, size_t n,
                f2d_stream_t stream);
/* out = x + c*(a+b)                          :176 (RK3_SSP stage 2); out may be x (:149) */
int f2d_ts_xpay2(double *out, const double *x, double c, const double *a, const double *b,
                 size_t n, f2d_stream_t stream);
/* x += c*(a+b+4*d)                           :180 (RK3_SSP final) */
int f2d_ts_rk3ssp_final(double *x, double c, const double *a, const double *b,
                        const double *d, size_t n, f2d_stream_t stream);
/* x += c0*a - c1*b                           :90-91 (AB2), :100 */
int f2d_ts_ab2(double *x, double c0, const double *a, double c1, const double *b, size_t n,
               f2d_stream_t stream);
/* x += c0*a - c1*b + c2*d                    :102-103 (AB3) */
int f2d_ts_ab3(double *x, double c0, const double *a, double c1, const double *b, double c2,
               const double *d, size_t n, f2d_stream_t stream);
/* x = xb + c*a                               :116, :131 */
int f2d_ts_set_xpay(double *x, const double *xb, double c, const double *a, size_t n,
                    f2d_stream_t stream);
/* xs += c*(x + xb - 2*xs)                    :118 (Asselin filter) */
int f2d_ts_asselin(double *xs, double c, const double *x, const double *xb, size_t n,
                   f2d_stream_t stream);
/* x = (1/12)*(5*x + 8*xs - xb)               :133 (LFAM3) */
int f2d_ts_am3(double *x, const double *xs, const double *xb, size_t n,
               f2d_stream_t stream);
/* elementwise helpers for model glue (euler.py:111,134,141-145, boussinesq.py:100,
 * operators.py:276,290,444,477,481,486, quasigeostrophic.py:117-119):
 * y = y*a (a double or int8 field), y = alpha*y, y += alpha*a, y = a + alpha*b,
 * y -= (s/denom)[*mask] with s read from a DEVICE slot (operators.py:274-276,476-477:
 * a domain integral divided by an area). */
int f2d_mul_field(double *y, const double *a, size_t n, f2d_stream_t stream);
int f2d_mul_mask(double *y, const int8_t *a, size_t n, f2d_stream_t stream);
int f2d_scale(double *y, double alpha, size_t n, f2d_stream_t stream);
int f2d_add_scaled(double *y, double alpha, const double *a, size_t n, f2d_stream_t stream);
int f2d_add_scaled_mask(double *y, double alpha, const int8_t *a, size_t n,
                        f2d_stream_t stream);
int f2d_set_sum(double *y, const double *a, double alpha, const double *b, size_t n,
                f2d_stream_t stream);
/* y = y / d  (euler.py:112 source /= dt: a true division) */
int f2d_div_scalar(double *y, double d, size_t n, f2d_stream_t stream);
/* y -= (pa*a + pb*b)*mask  (operators.py:287, zero-momentum correction of the source) */
int f2d_sub_lin2_mask(double *y, double pa, const double *a, double pb, const double *b,
                      const int8_t *mask, size_t n, f2d_stream_t stream);
/* y -= (pa*a + pb*b)       (fluid2d.py:356 enforce_zero_momentum) */
int f2d_sub_lin2(double *y, double pa, const double *a, double pb, const double *b, size_t n,
                 f2d_stream_t stream);
int f2d_sub_devscalar(double *y, const double *dev_scalar, double denom, size_t n,
                      f2d_stream_t stream);
int f2d_sub_devscalar_mask(double *y, const double *dev_scalar, double denom,
                           const int8_t *a, size_t n, f2d_stream_t stream);

/* ---- thermal-wind model (core/thermalwind.py; operators.py:330-394).  numpy's rounding
 * sequence, no FMA.
 * :332-335 / :348-351 the in-place linear extrapolation diffx / diffz apply to the first halo
 * line on both sides before differencing: axis 0 -> x[:, -nh] = 2x[:, -nh-1] - x[:, -nh-2],
 * x[:, nh-1] = 2x[:, nh] - x[:, nh+1]; axis 1 -> the same on rows */
int f2d_extrapolate_bry(double *x, int nh, int ny, int nx, int axis, f2d_stream_t stream);
/* :374-382 y[1:-1,1:-1] = diffx(b)*gravity - diffz(V)*f0; y *= msk (b, V already
 * extrapolated; the halo fill of :383 is the caller's f2d_fill_halo) */
int f2d_tw_torque(const int8_t *msk, const double *b, const double *V, double dx, double dy,
                  double gravity, double f0, double *y, int ny, int nx, f2d_stream_t stream);
/* :389-391 y[:, 1:] = -0.5*f0*(u[:, :-1] + u[:, 1:]); y *= msk */
int f2d_tw_coriolis(const int8_t *msk, const double *u, double f0, double *y, int ny, int nx,
                    f2d_stream_t stream);
/* :354-355 + thermalwind.py:92-95  out = 0; out[1:-1,1:-1] = diffx(x)*diffz(y) -
 * diffz(x)*diffx(y); out *= msk  (x, y already extrapolated) */
int f2d_jacobian(const int8_t *msk, const double *x, const double *y, double dx, double dy,
                 double *out, int ny, int nx, f2d_stream_t stream);
/* thermalwind.py:136-137  out = x where x <= 0, else 0 */
int f2d_negative_part(double *out, const double *x, size_t n, f2d_stream_t stream);

/* ---- core/fluxes.py (diag_fluxes: reversible / irreversible advective fluxes)
 * :120-127 uc = 0.5*(u + roll(u,1,axis=1)), vc = 0.5*(v + roll(v,1,axis=0)) + fill_halo */
int f2d_flx_cellvel(const double *u, const double *v, double *uc, double *vc, int nh, int ny,
                    int nx, int fill_halo, f2d_stream_t stream);
/* :160-177 rev = cff*(fwd + sign*bwd), irr = cff*(fwd - sign*bwd)  (sign = +1 or -1) */
int f2d_flx_split(double *rev, double *irr, const double *fwd, const double *bwd, double cff,
                  double sign, size_t n, f2d_stream_t stream);

/* ---- core/output.py:90-95 (history files are float32): interior of a field cast to
 * float32 and packed [ny-2nh][nx-2nh] on the device, ready for one D2H copy */
int f2d_pack_interior_f32(const double *x, float *out, int nh, int ny, int nx,
                          f2d_stream_t stream);

/* ---- core/gmg: hierarchy.Gmg (hierarchy.py:21-218) + level.Grid (level.py:120-496)
 * + the kernels of gmg/fortran_multigrid.f90.
 * f2d_mg_create builds the whole hierarchy on the device: Gridinfo (level.py:24-117,
 * single rank), corner mask -> int8 (hierarchy.py:46), finest matrix
 * (level.py:261-302), mask coarsening (level.py:233-236), Galerkin coarse matrices
 * (coarsenmatrix, fortran_multigrid.f90:706-811 + halo fills level.py:323-327),
 * ninetofive (level.py:332), optional Helmholtz diagonal (hierarchy.py:80-84, Rd > 0).
 * cornermask: DEVICE [ny][nx] doubles holding 0/1 (operators.py:59-67).          */
int f2d_mg_create(f2d_mg_t **mg, const double *cornermask, int ny, int nx, double dx,
                  double dy, double omega, double hydroepsilon, double Rd,
                  f2d_stream_t stream);
int f2d_mg_destroy(f2d_mg_t *mg);
int f2d_mg_nlevels(const f2d_mg_t *mg);
int f2d_mg_level_shape(const f2d_mg_t *mg, int lev, int *ny, int *nx);
/* device pointers owned by the handle (tests, set-up inspection):
 * which = 0 msk(int8) / 1 A (5 planes [5][ny][nx]: SW,S,SE,W,C) / 2 x / 3 b / 4 r   */
void *f2d_mg_level_ptr(f2d_mg_t *mg, int lev, int which);
/* matrix class the kernels use at a level: 0 stored coefficients, 1 constant stencil
 * (all-fluid level), 2 constant stencil x mask products */
int f2d_mg_level_matrix_mode(const f2d_mg_t *mg, int lev);

/* Grid.smooth (level.py:340-368): nite x (smoothtwicewithA :2-127 + halo fill) */
int f2d_mg_smooth(f2d_mg_t *mg, int lev, double *x, const double *b, int nite,
                  f2d_stream_t stream);
/* Grid.residual (level.py:370-387): computeresidualwithA :320-362 + halo fill */
int f2d_mg_residual(f2d_mg_t *mg, int lev, const double *x, const double *b, double *r,
                    f2d_stream_t stream);
/* finetocoarse (level.py:473-496): restrict :501-546 + halo fill; lev = fine level */
int f2d_mg_restrict(f2d_mg_t *mg, int lev, const double *xfine, double *xcoarse,
                    f2d_stream_t stream);
/* coarsetofine (level.py:447-469): interpolate :415-498; add != 0 gives
 * xfine += I(xcoarse) (hierarchy.py:123-125); lev = fine level */
int f2d_mg_interpolate(f2d_mg_t *mg, int lev, const double *xcoarse, double *xfine, int add,
                       f2d_stream_t stream);
/* Grid.norm (level.py:389-416) without the sqrt: out[0] = sum of squares (DEVICE) */
int f2d_mg_sumsq(f2d_mg_t *mg, int lev, const double *x, double *out, f2d_stream_t stream);
/* Gmg.Vcycle / Gmg.Fcycle on the handle's own x,b,r (hierarchy.py:98-151) */
int f2d_mg_vcycle(f2d_mg_t *mg, int lev1, f2d_stream_t stream);
int f2d_mg_fcycle(f2d_mg_t *mg, int lev1, f2d_stream_t stream);
/* Gmg.twoVcycle(x,b) (hierarchy.py:207-218): psi is first guess and result */
int f2d_mg_two_vcycle(f2d_mg_t *mg, double *psi, const double *rhs, f2d_stream_t stream);
/* Gmg.solve(x,b,{maxite,tol}) (hierarchy.py:154-192).  SYNCHRONISES the stream (the
 * iteration count depends on residual norms read back by the host).  *nite, *res are
 * HOST outputs.  Returns F2D_ERR_DIVERGE where the reference would exit(0). */
int f2d_mg_solve(f2d_mg_t *mg, double *psi, const double *rhs, double tol, int maxite,
                 int *nite, double *res, f2d_stream_t stream);
/* relaxation of Grid.smooth (level.py:153-163, 340-349): 0 = two damped-Jacobi sweeps per
 * application (smoothtwicewithA, the default), 1 = line relaxation -- THREE applications of
 * smoothtridiag (fortran_multigrid.f90:215-317: per column a tridiagonal solve in y, columns
 * swept west to east in place) + halo fill per requested iteration.  f2d_mg_create selects 1
 * by itself when hydroepsilon*dy/dx <= 0.2; param.relaxation = 'tridiagonal' asks for it.
 * Single-GPU hierarchies only (the reference also refuses it with npy > 1, level.py:155-157). */
int f2d_mg_set_relaxation(f2d_mg_t *mg, int mode);
/* use CUDA graphs for the cycles (default 1) */
int f2d_mg_set_graphs(f2d_mg_t *mg, int enable);
/* diagnostics: the cluster tail kernel appends one SM clock stamp per barrier of its rank-0
 * CTA to the DEVICE buffer buf (buf[0] = count, zero it first; cap entries); NULL turns it off */
int f2d_mg_set_trace(f2d_mg_t *mg, long long *buf, int cap);

/* ---- Operators.invert_vorticity (operators.py:421-498) as one call:
 * work = celltocorner(w) [- rhsp]; full ? solve(psi, work, 4, 1e-11)[, psi -= mean if
 * perio] : twoVcycle(psi, work); psi *= mskp [+ psi_island]; (u,v) = orthogradient(psi).
 * w, psi, u, v, work: DEVICE fields; mskp int8 corner mask; rhsp / psi_island may be
 * NULL (island.py:21-43); msk and mskp may both be NULL on an all-fluid domain without
 * island (see f2d_mask_orthogradient).  full != 0 synchronises (see f2d_mg_solve);
 * nite/res HOST outputs (may be NULL). `scratch` as for the reductions. */
int f2d_invert_vorticity(f2d_mg_t *mg, const int8_t *msk, const int8_t *mskp,
                         const double *w, double *psi, double *u, double *v, double *work,
                         const double *rhsp, const double *psi_island, int full, int perio,
                         double area, double dx, double dy, int nh, int *nite, double *res,
                         double *scratch, f2d_stream_t stream);
/* One shot: the NEXT f2d_invert_vorticity on this handle also writes the Runge-Kutta stage state
 * of the velocities, uo = ub + c*u (ue == ve == NULL) or ub + c*(ue + u), v alike (see
 * f2d_mask_orthogradient_stage; with an island the update runs as kernels of its own). */
int f2d_mg_set_uv_stage(f2d_mg_t *mg, const double *ub, const double *vb, const double *ue,
                        const double *ve, double *uo, double *vo, double c);

/* ---- multi-GPU: y-slab decomposition (npx = 1, npy = nranks), one process per GPU.
 * Replaces the mpi4py layer: gmg/halo.py (8 persistent Send/Recv per fill),
 * gmg/subdomains.py (Allgatherv gluing of coarse levels), level.py:401 (allreduce of
 * norms), mpitools.py (allgather of diagnostics).
 * Every rank creates a communicator with an arena of the same size; the 64-byte CUDA IPC
 * handles are exchanged by the host (torch.distributed) and passed, rank-ordered, to
 * f2d_comm_connect.  Buffers whose halos are exchanged must come from f2d_comm_alloc,
 * called in the same order with the same sizes on every rank (symmetric heap). */
typedef struct f2d_comm f2d_comm_t;
int f2d_comm_create(f2d_comm_t **comm, int rank, int nranks, size_t arena_bytes,
                    void *ipc_handle_out /* 64 bytes */);
int f2d_comm_connect(f2d_comm_t *comm, const void *all_handles /* nranks x 64 bytes */);
void *f2d_comm_alloc(f2d_comm_t *comm, size_t nbytes);
int f2d_comm_rank(const f2d_comm_t *comm);
int f2d_comm_size(const f2d_comm_t *comm);
int f2d_comm_destroy(f2d_comm_t *comm);
/* lock-step synchronisations completed by this rank (debug / test aid; synchronises) */
long long f2d_comm_epoch(f2d_comm_t *comm);
/* device-side barrier of all ranks (no host synchronisation) */
int f2d_comm_barrier(f2d_comm_t *comm, f2d_stream_t stream);
/* Halo.fill of the y direction (halo.py:214-292): push the nh top / bottom interior rows
 * (full width) into the neighbours' halo rows over NVLink, then lock-step with them.
 * With one rank it is the plain periodic fill. */
int f2d_comm_exchange_y(f2d_comm_t *comm, double *x, int nh, int ny, int nx,
                        f2d_stream_t stream);
/* in-place all-reduce of n <= 32 DEVICE doubles (bit k of maxmask: max instead of sum),
 * folded in rank order (deterministic); mpitools.py:16-40, level.py:401 */
int f2d_comm_allreduce(f2d_comm_t *comm, double *vals, int n, unsigned int maxmask,
                       f2d_stream_t stream);
/* multigrid on slabs: the local corner-mask slab [ny_loc][nx]; levels with more than
 * F2D_SLAB_MIN_CELLS (default 2^20) global cells stay distributed (halo rows exchanged
 * after every operator), coarser levels are gathered onto every rank and computed
 * redundantly (the reference's "peak" levels, level.py:71-86, taken to their end state).
 * psi / rhs passed to the cycles must live in the symmetric heap. */
int f2d_mg_slab_levels(const f2d_mg_t *mg);
int f2d_mg_create_slab(f2d_mg_t **mg, f2d_comm_t *comm, const double *cornermask, int ny_loc,
                       int nx, double dx, double dy, double omega, double hydroepsilon,
                       double Rd, f2d_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* F2D_B200_H */
