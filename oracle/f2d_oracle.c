/*
 * f2d_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * A plain-C restatement of the five Fortran-90 kernels files of pvthinker/Fluid2d
 * (core/fortran_advection.f90, core/fortran_fluxes.f90, core/fortran_operators.f90,
 * core/fortran_diag.f90, core/gmg/fortran_multigrid.f90).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this library; the product (fluid2d_b200/) never does.
 *
 * PARITY STATUS: the reference's Fortran cannot be compiled in the build container
 * (no gfortran / meson / mpi4py) and the reference ships no golden vectors, so the
 * KERNEL arithmetic below is "parity unpinned" against the gfortran binary (restated by
 * reading the source).  What checks the restatement: oracle/fortran_source.py executes the
 * reference's Fortran SOURCE TEXT through a statement-by-statement translation (language
 * semantics modelled, not compiled) and tests/test_oracle_vs_fortran_source.py requires
 * every routine below to agree with it bit for bit.
 * The ORCHESTRATION above these kernels is pinned: tests/golden/make_golden.py runs
 * the reference's own, unmodified Python (operators.py, gmg/level.py, gmg/hierarchy.py, euler.py,
 * timescheme.py, fluid2d.py ...) on top of this library and freezes its output.
 *
 * Conventions
 *  - arrays are numpy C-order [m][n]: m = rows (y, index j), n = columns (x, index i);
 *    the Fortran x(j,i) (1-based) is x[(j-1)*n + (i-1)].  The macros below keep the
 *    1-based notation so every loop bound can be read against the .f90 line cited.
 *  - the matrix A is numpy C-order [m][n][nd] (nd = 5 or 9); Fortran A(j,i,k).
 *  - gfortran default-real literals (1./30., 0.333333...) are float32 constants
 *    promoted to double: they are written here as (double)(float) expressions.
 *  - compile with -ffp-contract=off: every expression is evaluated in the order the
 *    Fortran source writes it, without FMA contraction, so that the library gives
 *    the same bits on every host.
 *  - loops whose iterations are independent are OpenMP-parallel over rows; running
 *    mask sums of adv_upwind/adv_centered are restated as window sums (same
 *    integers).  Reductions are sequential in Fortran order unless
 *    f2d_oracle_set_reduce_mode(1) selects per-row partial sums (CPU baseline only).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef int8_t i8;

#define IDX(j, i) ((size_t)((j)-1) * (size_t)n + (size_t)((i)-1))

static int g_reduce_mode = 0; /* 0: sequential (Fortran order); 1: row partials */

void f2d_oracle_set_reduce_mode(int mode) { g_reduce_mode = mode; }
int f2d_oracle_get_reduce_mode(void) { return g_reduce_mode; }

int f2d_oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void f2d_oracle_set_num_threads(int nt) {
#ifdef _OPENMP
  omp_set_num_threads(nt);
#else
  (void)nt;
#endif
}

/* ------------------------------------------------------------------------- */
/* fortran_advection.f90 / fortran_fluxes.f90                                */
/* ------------------------------------------------------------------------- */

typedef struct {
  double d1, d2, d3, d4, d5, c1, c2, c3;
  double zdx, zdy, a, KK, u1, aa, bb;
  int method, order;
} upw_cst;

/* flux-splitting speed, fortran_advection.f90:76-83 (and :116-121) */
static inline double split_speed(const upw_cst *k, double vel) {
  double UU = 0.;
  if (k->method == 0) UU = fabs(vel);
  if (k->method == 1) {
    UU = fabs(vel);
    if (UU < k->u1) UU = k->aa * (vel * vel) + k->bb;
  }
  if (k->method == 2) UU = 2 * k->KK + log(cosh(vel * k->a)) / k->a;
  return UU;
}

/* east-face flux at (j,i), fortran_advection.f90:73-111 */
static inline double upw_fx(const upw_cst *k, const i8 *msk, const double *x,
                            const double *u, int n, int j, int i) {
  int mx2 = msk[IDX(j, i)] + msk[IDX(j, i + 1)];
  if (mx2 != 2) return 0.;
  /* window sums: before the running update mx5 covers i-2..i+2, mx3 i-1..i+1;
     after it (:96-97) they cover i-1..i+3 and i..i+2 */
  int mx5p = msk[IDX(j, i - 2)] + msk[IDX(j, i - 1)] + msk[IDX(j, i)] +
             msk[IDX(j, i + 1)] + msk[IDX(j, i + 2)];
  int mx3p = msk[IDX(j, i - 1)] + msk[IDX(j, i)] + msk[IDX(j, i + 1)];
  int mx5m = mx5p - msk[IDX(j, i - 2)] + msk[IDX(j, i + 3)];
  int mx3m = mx3p - msk[IDX(j, i - 1)] + msk[IDX(j, i + 2)];
  double uu = u[IDX(j, i)];
  double UU = split_speed(k, uu);
  double up = 0.5 * (uu + UU);
  double um = 0.5 * (uu - UU);
  double qp, qm;
  if (mx5p == 5 && k->order == 5)
    qp = k->d1 * x[IDX(j, i - 2)] + k->d2 * x[IDX(j, i - 1)] + k->d3 * x[IDX(j, i)] +
         k->d4 * x[IDX(j, i + 1)] + k->d5 * x[IDX(j, i + 2)];
  else if (mx3p == 3 && k->order >= 3)
    qp = k->c1 * x[IDX(j, i - 1)] + k->c2 * x[IDX(j, i)] + k->c3 * x[IDX(j, i + 1)];
  else
    qp = x[IDX(j, i)];
  if (mx5m == 5 && k->order == 5)
    qm = k->d5 * x[IDX(j, i - 1)] + k->d4 * x[IDX(j, i)] + k->d3 * x[IDX(j, i + 1)] +
         k->d2 * x[IDX(j, i + 2)] + k->d1 * x[IDX(j, i + 3)];
  else if (mx3m == 3 && k->order >= 3)
    qm = k->c3 * x[IDX(j, i)] + k->c2 * x[IDX(j, i + 1)] + k->c1 * x[IDX(j, i + 2)];
  else
    qm = x[IDX(j, i + 1)];
  return up * qp + um * qm;
}

/* north-face flux at (j,i), fortran_advection.f90:113-146 */
static inline double upw_fy(const upw_cst *k, const i8 *msk, const double *x,
                            const double *v, int n, int j, int i) {
  int my2 = msk[IDX(j, i)] + msk[IDX(j + 1, i)];
  if (my2 != 2) return 0.;
  int my5p = msk[IDX(j - 2, i)] + msk[IDX(j - 1, i)] + msk[IDX(j, i)] +
             msk[IDX(j + 1, i)] + msk[IDX(j + 2, i)];
  int my3p = msk[IDX(j - 1, i)] + msk[IDX(j, i)] + msk[IDX(j + 1, i)];
  int my5m = my5p - msk[IDX(j - 2, i)] + msk[IDX(j + 3, i)];
  int my3m = my3p - msk[IDX(j - 1, i)] + msk[IDX(j + 2, i)];
  double vv = v[IDX(j, i)];
  double UU = split_speed(k, vv);
  double up = 0.5 * (vv + UU);
  double um = 0.5 * (vv - UU);
  double qp, qm;
  if (my5p == 5 && k->order == 5)
    qp = k->d1 * x[IDX(j - 2, i)] + k->d2 * x[IDX(j - 1, i)] + k->d3 * x[IDX(j, i)] +
         k->d4 * x[IDX(j + 1, i)] + k->d5 * x[IDX(j + 2, i)];
  else if (my3p == 3 && k->order >= 3)
    qp = k->c1 * x[IDX(j - 1, i)] + k->c2 * x[IDX(j, i)] + k->c3 * x[IDX(j + 1, i)];
  else
    qp = x[IDX(j, i)];
  if (my5m == 5 && k->order == 5)
    qm = k->d5 * x[IDX(j - 1, i)] + k->d4 * x[IDX(j, i)] + k->d3 * x[IDX(j + 1, i)] +
         k->d2 * x[IDX(j + 2, i)] + k->d1 * x[IDX(j + 3, i)];
  else if (my3m == 3 && k->order >= 3)
    qm = k->c3 * x[IDX(j, i)] + k->c2 * x[IDX(j + 1, i)] + k->c1 * x[IDX(j + 2, i)];
  else
    qm = x[IDX(j + 1, i)];
  return up * qp + um * qm;
}

/*
 * adv_upwind: fortran_advection.f90:2-165; with xflx/yflx != NULL it is the
 * fortran_fluxes.f90:2-170 variant (same arithmetic + the two flux stores :150-162).
 * Returns 1 when nh != 3 (the Fortran prints "NHALO = 3 is compulsory" and STOPs).
 */
int f2d_oracle_adv_upwind(const i8 *msk, const double *x, double *y, const double *u,
                          const double *v, double *xflx, double *yflx,
                          const double *cst, int nh, int method, int order, int m,
                          int n) {
  if (nh != 3) return 1;
  upw_cst k;
  /* :36-44 default-real literals */
  k.d1 = (double)(1.f / 30.f);
  k.d2 = (double)(-13.f / 60.f);
  k.d3 = (double)(47.f / 60.f);
  k.d4 = (double)(9.f / 20.f);
  k.d5 = (double)(-1.f / 20.f);
  k.c1 = (double)(-1.f / 6.f);
  k.c2 = (double)(5.f / 6.f);
  k.c3 = (double)(2.f / 6.f);
  double dx = cst[0], dy = cst[1], logcosh = cst[2], umax = cst[3], aparab = cst[4];
  k.zdx = 1. / dx;
  k.zdy = 1. / dy;
  k.a = logcosh / umax;                               /* :55 */
  k.KK = (umax - log(cosh(umax * k.a)) / k.a) * 0.5;  /* :56 */
  k.u1 = aparab * umax;                               /* :58 */
  k.aa = 1. / (2. * k.u1);                            /* :59 */
  k.bb = k.u1 * 0.5;                                  /* :60 */
  k.method = method;
  k.order = order;

#pragma omp parallel
  {
    double *fx = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    double *fy = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    double *fym = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    int nt = 1, tid = 0;
#ifdef _OPENMP
    nt = omp_get_num_threads();
    tid = omp_get_thread_num();
#endif
    /* rows nh+1..m-nh are written; split them in contiguous blocks */
    int nrows = m - 2 * nh;
    int r0 = (int)((long)nrows * tid / nt), r1 = (int)((long)nrows * (tid + 1) / nt);
    int jstart = nh + 1 + r0, jend = nh + 1 + r1; /* [jstart, jend) */
    if (jstart < jend) {
      /* fym of the first row of the block = fy of the row below (:157-160) */
      for (int i = nh; i <= n - nh; i++) fym[i] = upw_fy(&k, msk, x, v, n, jstart - 1, i);
      if (yflx && jstart - 1 == nh)
        for (int i = nh + 1; i <= n - nh; i++) yflx[IDX(nh, i)] = fym[i];
      for (int j = jstart; j < jend; j++) {
        for (int i = nh; i <= n - nh; i++) {
          fx[i] = upw_fx(&k, msk, x, u, n, j, i);
          fy[i] = upw_fy(&k, msk, x, v, n, j, i);
        }
        if (xflx) xflx[IDX(j, nh)] = fx[nh];
        for (int i = nh + 1; i <= n - nh; i++) {
          y[IDX(j, i)] = -k.zdx * (fx[i] - fx[i - 1]) - k.zdy * (fy[i] - fym[i]); /* :152 */
          if (xflx) xflx[IDX(j, i)] = fx[i];
          if (yflx) yflx[IDX(j, i)] = fy[i];
          fym[i] = fy[i];
        }
      }
    }
    free(fx);
    free(fy);
    free(fym);
  }
  return 0;
}

typedef struct {
  double e1, e2, e3, d1, d2, c1, zdx, zdy;
  int order;
} cen_cst;

/* fortran_advection.f90:229-247 */
static inline double cen_fx(const cen_cst *k, const i8 *msk, const double *x,
                            const double *u, int n, int j, int i) {
  int mx6 = msk[IDX(j, i - 2)] + msk[IDX(j, i - 1)] + msk[IDX(j, i)] + msk[IDX(j, i + 1)] +
            msk[IDX(j, i + 2)];
  mx6 = mx6 + msk[IDX(j, i + 3)];
  int mx4 = msk[IDX(j, i - 1)] + msk[IDX(j, i)] + msk[IDX(j, i + 1)] + msk[IDX(j, i + 2)];
  int mx2 = msk[IDX(j, i)] + msk[IDX(j, i + 1)];
  if (mx2 != 2) return 0.;
  double qp = 0.;
  if (mx6 == 6 && k->order == 6) {
    qp = k->e1 * (x[IDX(j, i - 2)] + x[IDX(j, i + 3)]) +
         k->e2 * (x[IDX(j, i - 1)] + x[IDX(j, i + 2)]);
    qp = qp + k->e3 * (x[IDX(j, i)] + x[IDX(j, i + 1)]);
  } else if (mx4 == 4 && k->order == 4) {
    qp = k->d1 * (x[IDX(j, i - 1)] + x[IDX(j, i + 2)]) +
         k->d2 * (x[IDX(j, i)] + x[IDX(j, i + 1)]);
  } else if (mx2 == 2 && k->order >= 2) {
    qp = k->c1 * (x[IDX(j, i)] + x[IDX(j, i + 1)]);
  }
  return u[IDX(j, i)] * qp;
}

/* fortran_advection.f90:249-265; my6 window j-2..j+3, my4 j-1..j+2, my2 j..j+1 */
static inline double cen_fy(const cen_cst *k, const i8 *msk, const double *x,
                            const double *v, int n, int j, int i) {
  int my2 = msk[IDX(j, i)] + msk[IDX(j + 1, i)];
  if (my2 != 2) return 0.;
  int my4 = my2 + msk[IDX(j - 1, i)] + msk[IDX(j + 2, i)];
  int my6 = my4 + msk[IDX(j - 2, i)] + msk[IDX(j + 3, i)];
  double qp = 0.;
  if (my6 == 6 && k->order == 6) {
    qp = k->e1 * (x[IDX(j - 2, i)] + x[IDX(j + 3, i)]);
    qp = qp + k->e2 * (x[IDX(j - 1, i)] + x[IDX(j + 2, i)]) +
         k->e3 * (x[IDX(j, i)] + x[IDX(j + 1, i)]);
  } else if (my4 == 4 && k->order == 4) {
    qp = k->d1 * (x[IDX(j - 1, i)] + x[IDX(j + 2, i)]) +
         k->d2 * (x[IDX(j, i)] + x[IDX(j + 1, i)]);
  } else if (my2 == 2 && k->order >= 2) {
    qp = k->c1 * (x[IDX(j, i)] + x[IDX(j + 1, i)]);
  }
  return v[IDX(j, i)] * qp;
}

/* adv_centered: fortran_advection.f90:169-284 (fortran_fluxes.f90:174-279 with fluxes) */
int f2d_oracle_adv_centered(const i8 *msk, const double *x, double *y, const double *u,
                            const double *v, double *xflx, double *yflx,
                            const double *cst, int nh, int method, int order, int m,
                            int n) {
  (void)method;
  if (nh != 3) return 1;
  /* core/fortran_fluxes.f90's adv_centered (the routine that also stores the face fluxes) has
   * no 6th-order branch: with order = 6 its tests `order.eq.4` fail and it lands on the
   * `order.ge.2` two-point mean (found by running the Fortran source, oracle/fortran_source.py) */
  if (xflx != NULL && order == 6) order = 2;
  cen_cst k;
  k.e1 = (double)(1.f / 60.f);
  k.e2 = (double)(-2.f / 15.f);
  k.e3 = (double)(37.f / 60.f);
  k.d1 = (double)(-1.f / 12.f);
  k.d2 = (double)(7.f / 12.f);
  k.c1 = (double)(1.f / 2.f);
  k.zdx = 1. / cst[0];
  k.zdy = 1. / cst[1];
  k.order = order;
#pragma omp parallel
  {
    double *fx = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    double *fy = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    double *fym = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    int nt = 1, tid = 0;
#ifdef _OPENMP
    nt = omp_get_num_threads();
    tid = omp_get_thread_num();
#endif
    int nrows = m - 2 * nh;
    int r0 = (int)((long)nrows * tid / nt), r1 = (int)((long)nrows * (tid + 1) / nt);
    int jstart = nh + 1 + r0, jend = nh + 1 + r1;
    if (jstart < jend) {
      for (int i = nh; i <= n - nh; i++) fym[i] = cen_fy(&k, msk, x, v, n, jstart - 1, i);
      if (yflx && jstart - 1 == nh)
        for (int i = nh + 1; i <= n - nh; i++) yflx[IDX(nh, i)] = fym[i];
      for (int j = jstart; j < jend; j++) {
        for (int i = nh; i <= n - nh; i++) {
          fx[i] = cen_fx(&k, msk, x, u, n, j, i);
          fy[i] = cen_fy(&k, msk, x, v, n, j, i);
        }
        if (xflx) xflx[IDX(j, nh)] = fx[nh];
        for (int i = nh + 1; i <= n - nh; i++) {
          y[IDX(j, i)] = -k.zdx * (fx[i] - fx[i - 1]) - k.zdy * (fy[i] - fym[i]);
          if (xflx) xflx[IDX(j, i)] = fx[i];
          if (yflx) yflx[IDX(j, i)] = fy[i];
          fym[i] = fy[i];
        }
      }
    }
    free(fx);
    free(fy);
    free(fym);
  }
  return 0;
}

/* ------------------------------------------------------------------------- */
/* fortran_operators.f90                                                     */
/* ------------------------------------------------------------------------- */

/* computeorthogradient: fortran_operators.f90:2-39 */
void f2d_oracle_computeorthogradient(const i8 *msk, const double *psi, double dx,
                                     double dy, int nh, double *u, double *v, int m,
                                     int n) {
  (void)nh;
  double zdx = 1. / dx, zdy = 1. / dy;
#pragma omp parallel for schedule(static)
  for (int j = 2; j <= m - 1; j++) {
    for (int i = 2; i <= n - 1; i++) {
      int mm = msk[IDX(j, i)] + msk[IDX(j, i + 1)];
      if (mm == 2)
        u[IDX(j, i)] = zdy * (psi[IDX(j - 1, i)] - psi[IDX(j, i)]);
      else
        u[IDX(j, i)] = 0.;
      mm = msk[IDX(j, i)] + msk[IDX(j + 1, i)];
      if (mm == 2)
        v[IDX(j, i)] = zdx * (psi[IDX(j, i)] - psi[IDX(j, i - 1)]);
      else
        v[IDX(j, i)] = 0.;
    }
  }
}

/* celltocorner: fortran_operators.f90:44-64 */
void f2d_oracle_celltocorner(const double *xr, double *xp, int m, int n) {
#pragma omp parallel for schedule(static)
  for (int j = 1; j <= m - 1; j++)
    for (int i = 1; i <= n - 1; i++)
      xp[IDX(j, i)] = 0.25 * (xr[IDX(j, i)] + xr[IDX(j, i + 1)] + xr[IDX(j + 1, i)] +
                              xr[IDX(j + 1, i + 1)]);
}

/* cornertocell: fortran_operators.f90:102-122 */
void f2d_oracle_cornertocell(const double *xp, double *xr, int m, int n) {
#pragma omp parallel for schedule(static)
  for (int j = 2; j <= m; j++)
    for (int i = 2; i <= n; i++)
      xr[IDX(j, i)] = 0.25 * (xp[IDX(j, i)] + xp[IDX(j, i - 1)] + xp[IDX(j - 1, i)] +
                              xp[IDX(j - 1, i - 1)]);
}

/* add_diffusion: fortran_operators.f90:125-156 */
void f2d_oracle_add_diffusion(const i8 *msk, const double *trac, double dx, int nh,
                              double Kdiff, double *dtrac, int m, int n) {
  (void)nh;
  double coef = Kdiff / (dx * dx);
#pragma omp parallel for schedule(static)
  for (int j = 2; j <= m - 1; j++) {
    for (int i = 2; i <= n - 1; i++) {
      if (msk[IDX(j, i)] == 1) {
        double c = trac[IDX(j, i)];
        dtrac[IDX(j, i)] =
            dtrac[IDX(j, i)] +
            coef * (+msk[IDX(j, i - 1)] * (trac[IDX(j, i - 1)] - c) +
                    msk[IDX(j, i + 1)] * (trac[IDX(j, i + 1)] - c) +
                    msk[IDX(j - 1, i)] * (trac[IDX(j - 1, i)] - c) +
                    msk[IDX(j + 1, i)] * (trac[IDX(j + 1, i)] - c));
      }
    }
  }
}

/* computenoslipsourceterm: fortran_operators.f90:221-277 (sequential scatter) */
double f2d_oracle_computenoslipsourceterm(const i8 *msk, const double *x, double *y,
                                          double dx, double dy, int nh, int m, int n) {
  double cff = 1. / (2 * dx * dy);
  double total = 0.;
  for (int j = 1; j <= nh; j++)
    for (int i = 1; i <= n; i++) y[IDX(j, i)] = 0.;
  for (int j = nh + 1; j <= m - nh + 1; j++) {
    for (int i = 1; i <= nh; i++) y[IDX(j, i)] = 0.;
    for (int i = nh + 1; i <= n - nh + 1; i++) {
      y[IDX(j, i)] = 0.;
      int msku = msk[IDX(j, i - 1)] + msk[IDX(j, i)];
      if (msku == 1) {
        double v = (x[IDX(j, i)] + x[IDX(j - 1, i)] - x[IDX(j, i - 2)] - x[IDX(j - 1, i - 2)]) * cff;
        if (msk[IDX(j, i)] != 0) {
          y[IDX(j, i)] = y[IDX(j, i)] - v;
          total = total - v;
        } else {
          y[IDX(j, i - 1)] = y[IDX(j, i - 1)] + v;
          total = total + v;
        }
      }
      int mskv = msk[IDX(j - 1, i)] + msk[IDX(j, i)];
      if (mskv == 1) {
        double u = -(x[IDX(j, i)] + x[IDX(j, i - 1)] - x[IDX(j - 2, i)] - x[IDX(j - 2, i - 1)]) * cff;
        if (msk[IDX(j, i)] != 0) {
          y[IDX(j, i)] = y[IDX(j, i)] + u;
          total = total + u;
        } else {
          y[IDX(j - 1, i)] = y[IDX(j - 1, i)] - u;
          total = total - u;
        }
      }
    }
  }
  return total;
}

/* add_torque: fortran_operators.f90:330-381 (ml/mr are max(1, pair sums)) */
void f2d_oracle_add_torque(const i8 *msk, const double *buoy, double dx, int nh,
                           double gravity, double *domega, int m, int n) {
  double coef = 0.5 * gravity / (dx);
#pragma omp parallel for schedule(static)
  for (int j = 1 + nh; j <= m - nh; j++) {
    int i = 1 + nh;
    int ml = msk[IDX(j, i)] + msk[IDX(j, i - 1)];
    if (ml < 1) ml = 1;
    for (i = 1 + nh; i <= n - nh; i++) {
      int mr = msk[IDX(j, i + 1)] + msk[IDX(j, i)];
      if (mr < 1) mr = 1;
      int mm = ml + mr;
      if (mm == 4)
        domega[IDX(j, i)] = domega[IDX(j, i)] +
                            (buoy[IDX(j, i + 1)] - buoy[IDX(j, i - 1)]) * (coef)*msk[IDX(j, i)];
      ml = mr;
    }
  }
}

/* ------------------------------------------------------------------------- */
/* fortran_diag.f90 (+ computenorm / computeinner of fortran_multigrid.f90)  */
/* ------------------------------------------------------------------------- */

/* generic interior masked sum of f(j,i); sequential or per-row partials */
#define REDUCE_SUM(EXPR_COND, EXPR_VAL, RESULT)                         \
  do {                                                                  \
    double acc__ = 0.;                                                  \
    if (g_reduce_mode == 0) {                                           \
      for (int j = nh + 1; j <= m - nh; j++)                            \
        for (int i = nh + 1; i <= n - nh; i++)                          \
          if (EXPR_COND) acc__ = acc__ + (EXPR_VAL);                    \
    } else {                                                            \
      double *rows__ = (double *)malloc(sizeof(double) * (size_t)(m + 1)); \
      _Pragma("omp parallel for schedule(static)")                      \
      for (int j = nh + 1; j <= m - nh; j++) {                          \
        double a__ = 0.;                                                \
        for (int i = nh + 1; i <= n - nh; i++)                          \
          if (EXPR_COND) a__ = a__ + (EXPR_VAL);                        \
        rows__[j] = a__;                                                \
      }                                                                 \
      for (int j = nh + 1; j <= m - nh; j++) acc__ = acc__ + rows__[j]; \
      free(rows__);                                                     \
    }                                                                   \
    RESULT = acc__;                                                     \
  } while (0)

/* computedotprod: fortran_diag.f90:3-30 */
double f2d_oracle_computedotprod(const i8 *msk, const double *x, const double *y, int nh,
                                 int m, int n) {
  double z;
  REDUCE_SUM(msk[IDX(j, i)] != 0, x[IDX(j, i)] * y[IDX(j, i)], z);
  return z;
}

/* computemax: fortran_diag.f90:33-60 */
double f2d_oracle_computemax(const i8 *msk, const double *x, int nh, int m, int n) {
  double y = 0.;
  for (int j = nh + 1; j <= m - nh; j++)
    for (int i = nh + 1; i <= n - nh; i++)
      if (msk[IDX(j, i)] != 0) y = fmax(y, fabs(x[IDX(j, i)]));
  return y;
}

/* computesum: fortran_diag.f90:63-90 */
double f2d_oracle_computesum(const i8 *msk, const double *x, int nh, int m, int n) {
  double y;
  REDUCE_SUM(msk[IDX(j, i)] != 0, x[IDX(j, i)], y);
  return y;
}

/* computesumandnorm: fortran_diag.f90:93-120 (tests msk == 1) */
void f2d_oracle_computesumandnorm(const i8 *msk, const double *x, int nh, int m, int n,
                                  double *y, double *y2) {
  double s, s2;
  REDUCE_SUM(msk[IDX(j, i)] == 1, x[IDX(j, i)], s);
  REDUCE_SUM(msk[IDX(j, i)] == 1, x[IDX(j, i)] * x[IDX(j, i)], s2);
  *y = s;
  *y2 = s2;
}

/* computenormmaxu: fortran_diag.f90:123-152 */
void f2d_oracle_computenormmaxu(const i8 *msk, const double *x, int nh, int m, int n,
                                double *y, double *ymax) {
  double s = 0., mx = 0.;
  for (int j = nh + 1; j <= m - nh; j++)
    for (int i = nh + 1; i <= n - nh; i++)
      if (msk[IDX(j, i)] + msk[IDX(j, i + 1)] == 2) {
        s = s + x[IDX(j, i)] * x[IDX(j, i)];
        mx = fmax(mx, fabs(x[IDX(j, i)]));
      }
  *y = s;
  *ymax = mx;
}

/* computekemaxu: fortran_diag.f90:155-195 */
void f2d_oracle_computekemaxu(const i8 *msk, const double *u, const double *v, int nh,
                              int m, int n, double *ke_out, double *maxu_out) {
  double ke = 0., maxu = 0.;
  if (g_reduce_mode == 0) {
    for (int j = nh + 1; j <= m - nh; j++)
      for (int i = nh + 1; i <= n - nh; i++)
        if (msk[IDX(j, i)] == 1) {
          double zu = u[IDX(j, i)] * u[IDX(j, i)] + u[IDX(j, i - 1)] * u[IDX(j, i - 1)];
          double zv = v[IDX(j, i)] * v[IDX(j, i)] + v[IDX(j - 1, i)] * v[IDX(j - 1, i)];
          double um = fabs(u[IDX(j, i)] + u[IDX(j, i - 1)]);
          double vm = fabs(v[IDX(j, i)] + v[IDX(j - 1, i)]);
          ke = ke + zu + zv;
          maxu = fmax(maxu, um + vm);
        }
  } else {
    double *rk = (double *)malloc(sizeof(double) * (size_t)(m + 1));
    double *rm = (double *)malloc(sizeof(double) * (size_t)(m + 1));
#pragma omp parallel for schedule(static)
    for (int j = nh + 1; j <= m - nh; j++) {
      double k_ = 0., m_ = 0.;
      for (int i = nh + 1; i <= n - nh; i++)
        if (msk[IDX(j, i)] == 1) {
          double zu = u[IDX(j, i)] * u[IDX(j, i)] + u[IDX(j, i - 1)] * u[IDX(j, i - 1)];
          double zv = v[IDX(j, i)] * v[IDX(j, i)] + v[IDX(j - 1, i)] * v[IDX(j - 1, i)];
          double um = fabs(u[IDX(j, i)] + u[IDX(j, i - 1)]);
          double vm = fabs(v[IDX(j, i)] + v[IDX(j - 1, i)]);
          k_ = k_ + zu + zv;
          m_ = fmax(m_, um + vm);
        }
      rk[j] = k_;
      rm[j] = m_;
    }
    for (int j = nh + 1; j <= m - nh; j++) {
      ke = ke + rk[j];
      maxu = fmax(maxu, rm[j]);
    }
    free(rk);
    free(rm);
  }
  *ke_out = ke * 0.25;
  *maxu_out = maxu * 0.5;
}

/* computekemaxuv: fortran_diag.f90:198-236 */
void f2d_oracle_computekemaxuv(const i8 *msk, const double *u, const double *v, int nh,
                               int m, int n, double *ke_out, double *maxu_out,
                               double *maxv_out) {
  double ke = 0., maxu = 0., maxv = 0.;
  for (int j = nh + 1; j <= m - nh; j++)
    for (int i = nh + 1; i <= n - nh; i++)
      if (msk[IDX(j, i)] == 1) {
        double zu = u[IDX(j, i)] * u[IDX(j, i)] + u[IDX(j, i - 1)] * u[IDX(j, i - 1)];
        double zv = v[IDX(j, i)] * v[IDX(j, i)] + v[IDX(j - 1, i)] * v[IDX(j - 1, i)];
        ke = ke + zu + zv;
        maxu = fmax(maxu, zu);
        maxv = fmax(maxv, zv);
      }
  *ke_out = ke * 0.25;
  *maxu_out = sqrt(maxu / 2.);
  *maxv_out = sqrt(maxv / 2.);
}

/* computekewithpsi: fortran_diag.f90:239-267 */
double f2d_oracle_computekewithpsi(const i8 *msk, const double *omega, const double *psi,
                                   int nh, int m, int n) {
  double ke = 0.;
  for (int j = nh + 1; j <= m - nh; j++)
    for (int i = nh + 1; i <= n - nh; i++)
      if (msk[IDX(j, i)] == 1)
        ke = ke - 0.125 *
                      (psi[IDX(j, i)] + psi[IDX(j - 1, i)] + psi[IDX(j - 1, i - 1)] +
                       psi[IDX(j, i - 1)]) *
                      omega[IDX(j, i)];
  return ke;
}

/* computenorm: fortran_multigrid.f90:813-839  (returns sum of squares) */
double f2d_oracle_computenorm(const i8 *msk, const double *x, int nh, int m, int n) {
  double y;
  REDUCE_SUM(msk[IDX(j, i)] != 0, x[IDX(j, i)] * x[IDX(j, i)], y);
  return y;
}

/* computeinner: fortran_multigrid.f90:841-867 */
double f2d_oracle_computeinner(const i8 *msk, const double *x, const double *y, int nh,
                               int m, int n) {
  double z;
  REDUCE_SUM(msk[IDX(j, i)] != 0, x[IDX(j, i)] * y[IDX(j, i)], z);
  return z;
}

/* ------------------------------------------------------------------------- */
/* gmg/fortran_multigrid.f90                                                 */
/* ------------------------------------------------------------------------- */

#define AIDX(j, i, k) ((IDX(j, i)) * (size_t)nd + (size_t)((k)-1))

/* one damped-Jacobi value at (j,i) from field `s`; fortran_multigrid.f90:66-85 */
static inline double jacobi_pt(const i8 *msk, const double *A, int nd, const double *s,
                               const double *b, double c1, double c2, int n, int j,
                               int i) {
  if (msk[IDX(j, i)] != 0) {
    int ip = i + 1, im = i - 1, jp = j + 1;
    double c3 = c1 / fabs(A[AIDX(j, i, 5)]);
    return s[IDX(j, i)] * c2 +
           c3 * (+A[AIDX(j, i, 1)] * s[IDX(j - 1, i - 1)] + A[AIDX(j, i, 2)] * s[IDX(j - 1, i)] +
                 A[AIDX(j, i, 3)] * s[IDX(j - 1, i + 1)] + A[AIDX(j, i, 4)] * s[IDX(j, i - 1)] +
                 A[AIDX(j, ip, 4)] * s[IDX(j, ip)] + A[AIDX(jp, im, 3)] * s[IDX(jp, im)] +
                 A[AIDX(jp, i, 2)] * s[IDX(jp, i)] + A[AIDX(jp, ip, 1)] * s[IDX(jp, ip)] -
                 b[IDX(j, i)]);
  }
  return 0.;
}

/*
 * smoothtwicewithA: fortran_multigrid.f90:2-127.  The Fortran pipelines the two
 * sweeps through a 3-row buffer `yo`; sweep 1 (rows/cols 2..m-1 / 2..n-1) only ever
 * reads the ORIGINAL x (rows already overwritten are behind the read front), sweep 2
 * is written on rows 3..m-2, cols 3..n-2.  This restatement keeps sweep 1 in a full
 * scratch array `yo` (m*n doubles) so both sweeps are row-parallel.  Entries of x
 * outside 3..m-2 x 3..n-2 (which the Fortran leaves in a buffer-dependent state:
 * lines :93-95 copy three sweep-1 values into columns 2 and n-1) are NOT specified
 * here -- the reference always follows this call with a halo fill (level.py:365)
 * that overwrites all of them; this restatement leaves them untouched.
 */
void f2d_oracle_smoothtwicewitha(const i8 *msk, const double *A, int nd, double *x,
                                 const double *b, double coef, int m, int n,
                                 double *yo) {
  double c1 = coef, c2 = 1. - c1;
#pragma omp parallel for schedule(static)
  for (int j = 2; j <= m - 1; j++)
    for (int i = 2; i <= n - 1; i++)
      yo[IDX(j, i)] = jacobi_pt(msk, A, nd, x, b, c1, c2, n, j, i);
#pragma omp parallel for schedule(static)
  for (int j = 3; j <= m - 2; j++)
    for (int i = 3; i <= n - 2; i++)
      x[IDX(j, i)] = jacobi_pt(msk, A, nd, yo, b, c1, c2, n, j, i);
}

/* computeresidualwithA: fortran_multigrid.f90:320-362 */
void f2d_oracle_computeresidualwitha(const i8 *msk, const double *A, int nd,
                                     const double *x, const double *b, double *y, int m,
                                     int n) {
#pragma omp parallel for schedule(static)
  for (int j = 2; j <= m - 1; j++) {
    for (int i = 2; i <= n - 1; i++) {
      if (msk[IDX(j, i)] != 0) {
        y[IDX(j, i)] = b[IDX(j, i)] - A[AIDX(j, i, 1)] * x[IDX(j - 1, i - 1)] -
                       A[AIDX(j, i, 2)] * x[IDX(j - 1, i)] -
                       A[AIDX(j, i, 3)] * x[IDX(j - 1, i + 1)] -
                       A[AIDX(j, i, 4)] * x[IDX(j, i - 1)] - A[AIDX(j, i, 5)] * x[IDX(j, i)] -
                       A[AIDX(j, i + 1, 4)] * x[IDX(j, i + 1)] -
                       A[AIDX(j + 1, i - 1, 3)] * x[IDX(j + 1, i - 1)] -
                       A[AIDX(j + 1, i, 2)] * x[IDX(j + 1, i)] -
                       A[AIDX(j + 1, i + 1, 1)] * x[IDX(j + 1, i + 1)];
      } else {
        y[IDX(j, i)] = 0.;
      }
    }
  }
}

/* fillhalo: fortran_multigrid.f90:365-412 (doubly periodic, corners included) */
void f2d_oracle_fillhalo(double *x, int nh, int m, int n) {
  int n2 = n - 2 * nh, m2 = m - 2 * nh;
  for (int j = 1; j <= nh; j++) {
    for (int i = 1; i <= nh; i++) x[IDX(j, i)] = x[IDX(m2 + j, n2 + i)];
    for (int i = 1; i <= n2; i++) x[IDX(j, i + nh)] = x[IDX(m2 + j, i + nh)];
    for (int i = 1; i <= nh; i++) x[IDX(j, i + n - nh)] = x[IDX(m2 + j, nh + i)];
  }
#pragma omp parallel for schedule(static)
  for (int j = 1; j <= m2; j++) {
    int jj = j + nh;
    for (int i = 1; i <= nh; i++) {
      x[IDX(jj, i)] = x[IDX(jj, n2 + i)];
      x[IDX(jj, i + n - nh)] = x[IDX(jj, i + nh)];
    }
  }
  for (int j = 1; j <= nh; j++) {
    int jj = j + m2 + nh;
    for (int i = 1; i <= nh; i++) x[IDX(jj, i)] = x[IDX(j + nh, n2 + i)];
    for (int i = 1; i <= n - 2 * nh; i++) x[IDX(jj, i + nh)] = x[IDX(j + nh, i + nh)];
    for (int i = 1; i <= nh; i++) x[IDX(jj, i + n - nh)] = x[IDX(j + nh, i + nh)];
  }
}

#define IDX1(j, i) ((size_t)((j)-1) * (size_t)n1 + (size_t)((i)-1))
#define IDX2(j, i) ((size_t)((j)-1) * (size_t)n2 + (size_t)((i)-1))

/* interpolate: fortran_multigrid.f90:415-498 (x1 fine <- x2 coarse) */
void f2d_oracle_interpolate(const i8 *msk1, const i8 *msk2, const double *x2, int nh,
                            double *x1, int m2, int n2, int m1, int n1) {
  (void)m1;
  const double c2[3] = {0., 1., 0.5};
  /* data c3/0.,1.,0.5,0.3333333333333333333333333333,0.25/ : default-real literal */
  const double c3[5] = {0., 1., 0.5, (double)0.3333333333333333333333333333f, 0.25};
  int j1s = 1, i1s = 1;
  if (nh == 2) { j1s = 2; i1s = 2; }
#pragma omp parallel for schedule(static)
  for (int j2 = 2; j2 <= m2 - 2; j2++) {
    int j1 = j1s + 2 * (j2 - 2);
    int i1 = i1s;
    for (int i2 = 2; i2 <= n2 - 2; i2++) {
      if (msk1[IDX1(j1, i1)] > 0)
        x1[IDX1(j1, i1)] = x2[IDX2(j2, i2)];
      else
        x1[IDX1(j1, i1)] = 0.;
      if (msk1[IDX1(j1, i1 + 1)] > 0) {
        int s = msk2[IDX2(j2, i2)] + msk2[IDX2(j2, i2 + 1)];
        x1[IDX1(j1, i1 + 1)] = (x2[IDX2(j2, i2)] + x2[IDX2(j2, i2 + 1)]) * c2[s];
      } else
        x1[IDX1(j1, i1 + 1)] = 0.;
      if (msk1[IDX1(j1 + 1, i1)] > 0) {
        int s = msk2[IDX2(j2, i2)] + msk2[IDX2(j2 + 1, i2)];
        x1[IDX1(j1 + 1, i1)] = (x2[IDX2(j2, i2)] + x2[IDX2(j2 + 1, i2)]) * c2[s];
      } else
        x1[IDX1(j1 + 1, i1)] = 0.;
      if (msk1[IDX1(j1 + 1, i1 + 1)] > 0) {
        int s = msk2[IDX2(j2, i2)] + msk2[IDX2(j2, i2 + 1)] + msk2[IDX2(j2 + 1, i2)] +
                msk2[IDX2(j2 + 1, i2 + 1)];
        x1[IDX1(j1 + 1, i1 + 1)] =
            c3[s] * (x2[IDX2(j2, i2)] + x2[IDX2(j2, i2 + 1)] + x2[IDX2(j2 + 1, i2)] +
                     x2[IDX2(j2 + 1, i2 + 1)]);
      } else
        x1[IDX1(j1 + 1, i1 + 1)] = 0.;
      i1 = i1 + 2;
    }
  }
}

/* restrict: fortran_multigrid.f90:501-546 (x2 coarse <- x1 fine) */
void f2d_oracle_restrict(const i8 *msk2, const double *x1, int nh, double *x2, int m2,
                         int n2, int m1, int n1) {
  (void)m1;
#pragma omp parallel for schedule(static)
  for (int j2 = nh; j2 <= m2 - nh; j2++) {
    int j1 = nh + 2 * (j2 - nh);
    int i1 = nh;
    for (int i2 = nh; i2 <= n2 - nh; i2++) {
      if (msk2[IDX2(j2, i2)] != 0)
        x2[IDX2(j2, i2)] =
            0.25 * x1[IDX1(j1, i1)] +
            0.125 * (x1[IDX1(j1, i1 - 1)] + x1[IDX1(j1, i1 + 1)] + x1[IDX1(j1 - 1, i1)] +
                     x1[IDX1(j1 + 1, i1)]) +
            0.0625 * (x1[IDX1(j1 - 1, i1 - 1)] + x1[IDX1(j1 - 1, i1 + 1)] +
                      x1[IDX1(j1 + 1, i1 - 1)] + x1[IDX1(j1 + 1, i1 + 1)]);
      else
        x2[IDX2(j2, i2)] = 0.;
      i1 = i1 + 2;
    }
  }
}

/*
 * coarsenmatrix: fortran_multigrid.f90:706-811.  Afine [m1][n1][9], Acoarse
 * [m2][n2][9] (intent(out): entries outside nh..m2-nh are left as the caller
 * allocated them -- the caller halo-fills every diagonal, level.py:323-327).
 * coef(ki,kj) = data coef/0.125,0.25,0.125,0.25,0.5,0.25,0.125,0.25,0.125/
 * (symmetric, so Fortran column-major order is immaterial).  The `m`/`loc`
 * computation at :753-772 is dead code and is omitted.
 */
void f2d_oracle_coarsenmatrix(const double *Afine, double *Acoarse, const i8 *msk1,
                              const i8 *msk2, int nh, int m1, int n1, int m2, int n2) {
  (void)m1;
  static const double coefv[3][3] = {
      {0.125, 0.25, 0.125}, {0.25, 0.5, 0.25}, {0.125, 0.25, 0.125}};
#define COEF(ki, kj) coefv[(kj) + 1][(ki) + 1]
#pragma omp parallel for schedule(static)
  for (int j2 = nh; j2 <= m2 - nh; j2++) {
    for (int i2 = nh; i2 <= n2 - nh; i2++) {
      double *out = Acoarse + IDX2(j2, i2) * 9;
      if (msk2[IDX2(j2, i2)] == 1) {
        for (int l = 1; l <= 9; l++) {
          int i1 = 2 * (i2 - nh) + nh;
          int j1 = 2 * (j2 - nh) + nh;
          int di2 = (l - 1) % 3 - 1;
          int dj2 = (l - 1) / 3 - 1;
          double z5[5][5]; /* z5[j+2][i+2] */
          double z3[3][3]; /* z3[jj+1][ii+1] */
          memset(z5, 0, sizeof z5);
          for (int kj = -1; kj <= 1; kj++)
            for (int ki = -1; ki <= 1; ki++) {
              int i = 2 * di2 + ki;
              int j = 2 * dj2 + kj;
              if (abs(i) <= 2 && abs(j) <= 2) {
                if (msk1[IDX1(j1 + j, i1 + i)] == 1) z5[j + 2][i + 2] = 2. * COEF(ki, kj);
              }
            }
          memset(z3, 0, sizeof z3);
          for (int jj = -1; jj <= 1; jj++)
            for (int ii = -1; ii <= 1; ii++)
              for (int kj = -1; kj <= 1; kj++)
                for (int ki = -1; ki <= 1; ki++) {
                  int k = 1 + (ki + 1) + (kj + 1) * 3;
                  if (msk1[IDX1(j1 + jj, i1 + ii)] == 1)
                    z3[jj + 1][ii + 1] =
                        z3[jj + 1][ii + 1] +
                        Afine[IDX1(j1 + jj, i1 + ii) * 9 + (size_t)(k - 1)] *
                            z5[jj + kj + 2][ii + ki + 2];
                }
          double w = 0.;
          if (msk2[IDX2(j2, i2)] != 0)
            for (int jj = -1; jj <= 1; jj++)
              for (int ii = -1; ii <= 1; ii++)
                w = w + 0.5 * COEF(ii, jj) * z3[jj + 1][ii + 1];
          out[l - 1] = w;
        }
      } else {
        for (int l = 0; l < 9; l++) out[l] = 0.;
      }
    }
  }
#undef COEF
}

/* halotobuffer: fortran_multigrid.f90:549-616.  b0..b7 are C-order (intent c). */
void f2d_oracle_halotobuffer(const double *x, double *b0, double *b1, double *b2,
                             double *b3, double *b4, double *b5, double *b6, double *b7,
                             int nh, int m, int n) {
  int n2 = n - 2 * nh, m2 = m - 2 * nh;
  for (int j = 1; j <= nh; j++)
    for (int i = 1; i <= nh; i++) b0[(j - 1) * nh + (i - 1)] = x[IDX(j + nh, i + nh)];
  for (int j = 1; j <= nh; j++)
    for (int i = 1; i <= n2; i++) b1[(size_t)(j - 1) * n2 + (i - 1)] = x[IDX(j + nh, i + nh)];
  for (int j = 1; j <= nh; j++)
    for (int i = 1; i <= nh; i++) b2[(j - 1) * nh + (i - 1)] = x[IDX(j + nh, i + n2)];
  for (int j = 1; j <= m2; j++)
    for (int i = 1; i <= nh; i++) b3[(size_t)(j - 1) * nh + (i - 1)] = x[IDX(j + nh, i + nh)];
  for (int j = 1; j <= m2; j++)
    for (int i = 1; i <= nh; i++) b4[(size_t)(j - 1) * nh + (i - 1)] = x[IDX(j + nh, i + n2)];
  for (int j = 1; j <= nh; j++)
    for (int i = 1; i <= nh; i++) b5[(j - 1) * nh + (i - 1)] = x[IDX(j + m2, i + nh)];
  for (int j = 1; j <= nh; j++)
    for (int i = 1; i <= n2; i++) b6[(size_t)(j - 1) * n2 + (i - 1)] = x[IDX(j + m2, i + nh)];
  for (int j = 1; j <= nh; j++)
    for (int i = 1; i <= nh; i++) b7[(j - 1) * nh + (i - 1)] = x[IDX(j + m2, i + n - 2 * nh)];
}

/* buffertohalo: fortran_multigrid.f90:619-670 */
void f2d_oracle_buffertohalo(double *x, const double *b0, const double *b1,
                             const double *b2, const double *b3, const double *b4,
                             const double *b5, const double *b6, const double *b7, int nh,
                             int m, int n) {
  int n2 = n - 2 * nh, m2 = m - 2 * nh;
  for (int j = 1; j <= nh; j++) {
    for (int i = 1; i <= nh; i++) x[IDX(j, i)] = b7[(j - 1) * nh + (i - 1)];
    for (int i = 1; i <= n2; i++) x[IDX(j, i + nh)] = b6[(size_t)(j - 1) * n2 + (i - 1)];
    for (int i = 1; i <= nh; i++) x[IDX(j, i + n - nh)] = b5[(j - 1) * nh + (i - 1)];
  }
  for (int j = 1; j <= m2; j++) {
    int jj = j + nh;
    for (int i = 1; i <= nh; i++) {
      x[IDX(jj, i)] = b4[(size_t)(j - 1) * nh + (i - 1)];
      x[IDX(jj, i + n - nh)] = b3[(size_t)(j - 1) * nh + (i - 1)];
    }
  }
  for (int j = 1; j <= nh; j++) {
    int jj = j + m - nh;
    for (int i = 1; i <= nh; i++) x[IDX(jj, i)] = b2[(j - 1) * nh + (i - 1)];
    for (int i = 1; i <= n2; i++) x[IDX(jj, i + nh)] = b1[(size_t)(j - 1) * n2 + (i - 1)];
    for (int i = 1; i <= nh; i++) x[IDX(jj, i + n - nh)] = b0[(j - 1) * nh + (i - 1)];
  }
}

/* buffertodomain: fortran_multigrid.f90:673-703.  b is [mp][np][m][n] (C order),
   x is [m1][n1]; tiles overlap by their halos, later tiles overwrite earlier ones. */
void f2d_oracle_buffertodomain(const double *b, double *x, int nh, int m1, int n1, int m,
                               int n, int mp, int np) {
  (void)m1;
  for (int l = 1; l <= mp; l++) {
    int jj = 1 + (l - 1) * (m - 2 * nh);
    for (int j = 1; j <= m; j++) {
      for (int k = 1; k <= np; k++) {
        int ii = 1 + (k - 1) * (n - 2 * nh);
        for (int i = 1; i <= n; i++) {
          x[IDX1(jj, ii)] =
              b[(((size_t)(l - 1) * np + (k - 1)) * m + (j - 1)) * (size_t)n + (i - 1)];
          ii = ii + 1;
        }
      }
      jj = jj + 1;
    }
  }
}

/* tridiag: fortran_multigrid.f90:290-317 (1-based arrays of length l) */
static void tridiag(const double *d, const double *dd, const double *b, double *xc,
                    double *gam, int l) {
  int k0 = 4;
  double bet = 1. / d[k0];
  xc[k0] = b[k0] * bet;
  for (int k = k0 + 1; k <= l - 4; k++) {
    gam[k] = dd[k - 1] * bet;
    bet = 1. / (d[k] - dd[k - 1] * gam[k]);
    xc[k] = (b[k] - dd[k - 1] * xc[k - 1]) * bet;
  }
  for (int k = l - 5; k >= k0; k--) xc[k] = xc[k] - gam[k + 1] * xc[k + 1];
}

/* smoothtridiag: fortran_multigrid.f90:215-288 (columns processed in order, in place) */
void f2d_oracle_smoothtridiag(const i8 *msk, const double *A, int nd, double *x,
                              const double *b, int m, int n) {
  double *y = (double *)calloc((size_t)m + 2, sizeof(double));
  double *d = (double *)calloc((size_t)m + 2, sizeof(double));
  double *ud = (double *)calloc((size_t)m + 2, sizeof(double));
  double *rhs = (double *)calloc((size_t)m + 2, sizeof(double));
  double *gam = (double *)calloc((size_t)m + 2, sizeof(double));
  for (int i = 2; i <= n - 1; i++) {
    int ip = i + 1, im = i - 1;
    if (msk[IDX(4, i)] != 0) {
      for (int j = 0; j <= m; j++) y[j] = 0.;
      for (int j = 4; j <= m - 4; j++) {
        int jm = j - 1, jp = j + 1;
        rhs[j] = b[IDX(j, i)] - A[AIDX(j, i, 1)] * x[IDX(jm, im)] -
                 A[AIDX(j, i, 3)] * x[IDX(jm, ip)] - A[AIDX(j, i, 4)] * x[IDX(j, im)] -
                 A[AIDX(j, ip, 4)] * x[IDX(j, ip)] - A[AIDX(jp, im, 3)] * x[IDX(jp, im)] -
                 A[AIDX(jp, ip, 1)] * x[IDX(jp, ip)];
        d[j] = A[AIDX(j, i, 5)];
        ud[j] = A[AIDX(jp, i, 2)];
      }
      tridiag(d, ud, rhs, y, gam, m);
      for (int j = 4; j <= m - 4; j++) x[IDX(j, i)] = y[j];
    }
  }
  free(y);
  free(d);
  free(ud);
  free(rhs);
  free(gam);
}

/* ------------------------------------------------------------------------- */
/* numpy whole-state combinations of timescheme.py, restated for the CPU      */
/* baseline (same operation order as the numpy expressions, no FMA)           */
/* ------------------------------------------------------------------------- */

/* out = x + c*d0          (timescheme.py:172  self.x = x + dt * self.dx0) */
void f2d_oracle_axpy1(double *out, const double *x, double c, const double *d0,
                      size_t len) {
#pragma omp parallel for schedule(static)
  for (size_t k = 0; k < len; k++) out[k] = x[k] + c * d0[k];
}
/* out = x + c*(d0+d1)     (timescheme.py:176) */
void f2d_oracle_axpy2(double *out, const double *x, double c, const double *d0,
                      const double *d1, size_t len) {
#pragma omp parallel for schedule(static)
  for (size_t k = 0; k < len; k++) out[k] = x[k] + c * (d0[k] + d1[k]);
}
/* x += c*(d0+d1+4*d2)     (timescheme.py:180) */
void f2d_oracle_axpy3(double *x, double c, const double *d0, const double *d1,
                      const double *d2, size_t len) {
#pragma omp parallel for schedule(static)
  for (size_t k = 0; k < len; k++) x[k] = x[k] + c * (d0[k] + d1[k] + 4 * d2[k]);
}
