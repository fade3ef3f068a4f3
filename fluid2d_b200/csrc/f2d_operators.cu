// Stencil operators of core/fortran_operators.f90, the periodic halo fill of
// gmg/fortran_multigrid.f90:365-412, and the whole-state combinations of
// core/timescheme.py -- memory-bound elementwise / 5-point kernels.
#include <stdarg.h>
#include <map>
#include <string>
#include <vector>

#include "f2d_common.cuh"

namespace f2d {
char g_err[512] = "";
long long g_launches = 0;
bool g_prof = false;
int g_pdl = -1;
bool pdl_enabled() {
  if (g_pdl < 0) {
    const char *e = getenv("F2D_PDL");
    g_pdl = (e && e[0] == '1') ? 1 : 0;   // off by default: measured neutral inside the graphs
  }
  // per-kernel accounting puts an event behind every launch: keep the kernels apart there
  return g_pdl == 1 && !g_prof;
}
namespace {
cudaStream_t g_prof_stream = nullptr;
struct ProfMark { std::string name; cudaEvent_t start, end; };   // start == nullptr: the previous mark's end
std::vector<ProfMark> g_prof_marks;
char g_prof_tag[160] = "";
cudaEvent_t g_prof_start = nullptr;
bool prof_capturing() {
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  return cudaStreamIsCapturing(g_prof_stream, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone;
}
}  // namespace
// names the next launch and records its start: an event recorded in front of a kernel carries
// the time the stream reached it, so a host that issues launches slower than the device runs
// them does not inflate the kernel's figure
void prof_tag(const char *fmt, ...) {
  if (!g_prof || prof_capturing()) return;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_prof_tag, sizeof g_prof_tag, fmt, ap);
  va_end(ap);
  if (g_prof_start) cudaEventDestroy(g_prof_start);
  g_prof_start = nullptr;
  if (cudaEventCreate(&g_prof_start) == cudaSuccess) cudaEventRecord(g_prof_start, g_prof_stream);
}
void prof_mark(const char *fallback) {
  if (prof_capturing()) {
    g_prof_tag[0] = 0;
    return;
  }
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, g_prof_stream);
  g_prof_marks.push_back(ProfMark{g_prof_tag[0] ? g_prof_tag : fallback, g_prof_tag[0] ? g_prof_start : nullptr, e});
  if (g_prof_tag[0]) g_prof_start = nullptr;
  g_prof_tag[0] = 0;
}
}  // namespace f2d

using namespace f2d;

extern "C" int f2d_abi_version(void) { return F2D_ABI_VERSION; }
extern "C" const char *f2d_last_error(void) { return f2d::g_err; }
extern "C" long long f2d_launch_count(void) { return f2d::g_launches; }
extern "C" void f2d_launch_count_reset(void) { f2d::g_launches = 0; }

extern "C" int f2d_prof_begin(f2d_stream_t s) {
  for (auto &m : f2d::g_prof_marks) {
    cudaEventDestroy(m.end);
    if (m.start) cudaEventDestroy(m.start);
  }
  f2d::g_prof_marks.clear();
  f2d::g_prof_stream = S(s);
  f2d::g_prof = true;
  f2d::prof_mark("(begin)");
  return F2D_OK;
}
extern "C" int f2d_prof_report(char *buf, size_t cap) {
  if (!buf || cap < 2) return fail(F2D_ERR_ARG, "prof_report: no buffer");
  f2d::g_prof = false;
  F2D_CUDA(cudaStreamSynchronize(f2d::g_prof_stream));
  std::map<std::string, std::pair<long long, double>> agg;   // name -> (launches, microseconds)
  std::vector<std::string> order;
  for (size_t k = 1; k < f2d::g_prof_marks.size(); k++) {
    float ms = 0.f;
    const auto &mk = f2d::g_prof_marks[k];
    cudaEventElapsedTime(&ms, mk.start ? mk.start : f2d::g_prof_marks[k - 1].end, mk.end);
    auto it = agg.find(mk.name);
    if (it == agg.end()) {
      order.push_back(mk.name);
      it = agg.emplace(mk.name, std::make_pair(0LL, 0.)).first;
    }
    it->second.first++;
    it->second.second += 1e3 * ms;
  }
  for (auto &m : f2d::g_prof_marks) {
    cudaEventDestroy(m.end);
    if (m.start) cudaEventDestroy(m.start);
  }
  f2d::g_prof_marks.clear();
  size_t used = 0;
  buf[0] = 0;
  for (auto &name : order) {
    auto &v = agg[name];
    int n = snprintf(buf + used, cap - used, "%s\t%lld\t%.3f\n", name.c_str(), v.first, v.second);
    if (n < 0 || (size_t)n >= cap - used) { buf[used] = 0; break; }
    used += (size_t)n;
  }
  return F2D_OK;
}

extern "C" int f2d_copy(void *dst, const void *src, size_t nbytes, f2d_stream_t s) {
  if (!dst || !src) return fail(F2D_ERR_ARG, "copy: null");
  F2D_CUDA(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToDevice, S(s)));
  return F2D_OK;
}
extern "C" int f2d_zero(void *dst, size_t nbytes, f2d_stream_t s) {
  if (!dst) return fail(F2D_ERR_ARG, "zero: null");
  F2D_CUDA(cudaMemsetAsync(dst, 0, nbytes, S(s)));
  return F2D_OK;
}

// ---------------------------------------------------------------------------
// halo fill: one thread per halo cell, pulls from the periodic interior source
// ---------------------------------------------------------------------------
template <typename T>
__global__ void k_fill_halo(T *__restrict__ x, int ny, int nx, int nh, int ywrap) {
  // halo cells: 2*nh full rows + (ny-2nh) rows x 2*nh columns; without ywrap (y-slab
  // decomposition) only the side columns of the interior rows are local
  long long nrowcells = 2LL * nh * nx;
  long long total = nrowcells + 2LL * nh * (ny - 2 * nh);
  for (long long t = (ywrap ? 0 : nrowcells) + blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    int j, i;
    if (t < nrowcells) {
      int r = (int)(t / nx);
      i = (int)(t % nx);
      j = r < nh ? r : ny - 2 * nh + r;  // r in [nh,2nh) -> top rows
    } else {
      long long q = t - nrowcells;
      int r = (int)(q / (2 * nh));
      int c = (int)(q % (2 * nh));
      j = nh + r;
      i = c < nh ? c : nx - 2 * nh + c;
    }
    x[(size_t)j * nx + i] = x[(size_t)wrap_src(j, ny, nh) * nx + wrap_src(i, nx, nh)];
  }
}

template <typename T>
static int fill_halo_t(T *x, int nh, int ny, int nx, cudaStream_t s, int ywrap = 1) {
  if (!x || ny <= 2 * nh || nx <= 2 * nh || nh < 1) return fail(F2D_ERR_ARG, "fill_halo: bad shape");
  if (ny - 2 * nh < nh || nx - 2 * nh < nh) return fail(F2D_ERR_ARG, "fill_halo: interior narrower than halo");
  long long total = 2LL * nh * nx + 2LL * nh * (ny - 2 * nh);
  int threads = 256;
  int blocks = cdiv(total, threads);
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_fill_halo<T><<<blocks, threads, 0, s>>>(x, ny, nx, nh, ywrap);
  F2D_LAUNCHED();
  return F2D_OK;
}

extern "C" int f2d_fill_halo(double *x, int nh, int ny, int nx, f2d_stream_t s) {
  return fill_halo_t<double>(x, nh, ny, nx, S(s));
}
extern "C" int f2d_fill_halo_i8(int8_t *x, int nh, int ny, int nx, f2d_stream_t s) {
  return fill_halo_t<int8_t>(x, nh, ny, nx, S(s));
}
extern "C" int f2d_fill_halo_x(double *x, int nh, int ny, int nx, f2d_stream_t s) {
  return fill_halo_t<double>(x, nh, ny, nx, S(s), 0);
}

// ---------------------------------------------------------------------------
// 2-D launch helper: thread (i,j) over a [ny][nx] array, 32x8 blocks
// ---------------------------------------------------------------------------
static inline dim3 grid2d(int ny, int nx, dim3 b) { return dim3(cdiv(nx, b.x), cdiv(ny, b.y)); }
#define IJ()                                       \
  int i = blockIdx.x * blockDim.x + threadIdx.x;   \
  int j = blockIdx.y * blockDim.y + threadIdx.y;   \
  if (i >= nx || j >= ny) return;                  \
  size_t c = (size_t)j * nx + i

// celltocorner: fortran_operators.f90:44-64, xp(1..m-1,1..n-1)
__global__ void k_celltocorner(const double *__restrict__ xr, double *__restrict__ xp, int ny, int nx) {
  IJ();
  if (j > ny - 2 || i > nx - 2) return;
  xp[c] = 0.25 * (((xr[c] + xr[c + 1]) + xr[c + nx]) + xr[c + nx + 1]);
}
// Strip version (nx even): a thread owns two adjacent columns and marches C2C_ROWS rows
// north; each row is read once as a 16-byte vector (+ the east neighbour from the next
// lane), the pair sums (xr(j,i)+xr(j,i+1)) of the row are reused as the south pair of the
// next row.  Same association as the Fortran: ((a + b) + c) + d.
constexpr int C2C_ROWS = 16;
__global__ void __launch_bounds__(128) k_celltocorner_strip(const double *__restrict__ xr, double *__restrict__ xp,
                                                            int ny, int nx) {
  const int i = (blockIdx.x * 128 + threadIdx.x) * 2;
  const int j0 = blockIdx.y * C2C_ROWS;
  const bool active = i < nx;
  const unsigned lane = threadIdx.x & 31;
  auto row = [&](int j, double2 &a, double &e) {   // xr(j,i), xr(j,i+1), xr(j,i+2)
    a = make_double2(0., 0.);
    if (active) a = *reinterpret_cast<const double2 *>(xr + (size_t)j * nx + i);
    e = __shfl_down_sync(0xffffffffu, a.x, 1);
    if (active && lane == 31) e = (i + 2 < nx) ? xr[(size_t)j * nx + i + 2] : 0.;
  };
  double2 a, an;
  double e, en;
  row(j0, a, e);
  for (int r = 0; r < C2C_ROWS; r++) {
    const int j = j0 + r;
    if (j > ny - 2) break;
    row(j + 1, an, en);
    if (active) {
      const size_t c = (size_t)j * nx + i;
      double o0 = 0.25 * (((a.x + a.y) + an.x) + an.y);
      double o1 = 0.25 * (((a.y + e) + an.y) + en);
      if (i + 1 <= nx - 2) *reinterpret_cast<double2 *>(xp + c) = make_double2(o0, o1);
      else xp[c] = o0;   // i + 1 is the last column, which celltocorner leaves untouched
    }
    a = an;
    e = en;
  }
}
extern "C" int f2d_celltocorner(const double *xr, double *xp, int ny, int nx, f2d_stream_t s) {
  if (!xr || !xp || ny < 2 || nx < 2) return fail(F2D_ERR_ARG, "celltocorner: bad args");
  if (nx % 2 == 0 && ((reinterpret_cast<uintptr_t>(xr) | reinterpret_cast<uintptr_t>(xp)) & 15) == 0) {
    dim3 g(cdiv(nx, 256), cdiv(ny - 1, C2C_ROWS));
    k_celltocorner_strip<<<g, 128, 0, S(s)>>>(xr, xp, ny, nx);
    F2D_LAUNCHED();
    return F2D_OK;
  }
  dim3 b(32, 8);
  k_celltocorner<<<grid2d(ny, nx, b), b, 0, S(s)>>>(xr, xp, ny, nx);
  F2D_LAUNCHED();
  return F2D_OK;
}

// cornertocell: fortran_operators.f90:102-122, xr(2..m,2..n)
__global__ void k_cornertocell(const double *__restrict__ xp, double *__restrict__ xr, int ny, int nx) {
  IJ();
  if (j < 1 || i < 1) return;
  xr[c] = 0.25 * (((xp[c] + xp[c - 1]) + xp[c - nx]) + xp[c - nx - 1]);
}
extern "C" int f2d_cornertocell(const double *xp, double *xr, int ny, int nx, f2d_stream_t s) {
  if (!xr || !xp || ny < 2 || nx < 2) return fail(F2D_ERR_ARG, "cornertocell: bad args");
  dim3 b(32, 8);
  k_cornertocell<<<grid2d(ny, nx, b), b, 0, S(s)>>>(xp, xr, ny, nx);
  F2D_LAUNCHED();
  return F2D_OK;
}

// computeorthogradient: fortran_operators.f90:2-39, rows/cols 2..m-1 / 2..n-1
__global__ void k_orthogradient(const int8_t *__restrict__ msk, const double *__restrict__ psi,
                                double zdx, double zdy, double *__restrict__ u,
                                double *__restrict__ v, int ny, int nx) {
  IJ();
  if (j < 1 || j > ny - 2 || i < 1 || i > nx - 2) return;
  int m0 = msk[c];
  double p = psi[c];
  u[c] = (m0 + msk[c + 1] == 2) ? zdy * (psi[c - nx] - p) : 0.;
  v[c] = (m0 + msk[c + nx] == 2) ? zdx * (p - psi[c - 1]) : 0.;
}
extern "C" int f2d_orthogradient(const int8_t *msk, const double *psi, double dx, double dy, int nh,
                                 double *u, double *v, int ny, int nx, f2d_stream_t s) {
  (void)nh;
  if (!msk || !psi || !u || !v || ny < 3 || nx < 3) return fail(F2D_ERR_ARG, "orthogradient: bad args");
  dim3 b(32, 8);
  k_orthogradient<<<grid2d(ny, nx, b), b, 0, S(s)>>>(msk, psi, 1. / dx, 1. / dy, u, v, ny, nx);
  F2D_LAUNCHED();
  return F2D_OK;
}

// psi *= mskp, then computeorthogradient, in one pass (operators.py:481,493 without
// islands).  A thread needs the masked psi of its south and west neighbours; it forms
// them itself from psi and mskp.  psi is updated in place: a neighbour may already have
// been masked when it is read, which is harmless because masking is idempotent
// (x*1 = x, x*0 = +-0) -- this is why the island case (psi += psi_island) is not fused.
__global__ void k_mask_orthogradient(const int8_t *__restrict__ msk, const int8_t *__restrict__ mskp,
                                     double *psi, double zdx, double zdy, double *__restrict__ u,
                                     double *__restrict__ v, int ny, int nx) {
  IJ();
  double p = __dmul_rn(psi[c], (double)mskp[c]);
  psi[c] = p;
  if (j < 1 || j > ny - 2 || i < 1 || i > nx - 2) return;
  int m0 = msk[c];
  double ps = __dmul_rn(psi[c - nx], (double)mskp[c - nx]);
  double pw = __dmul_rn(psi[c - 1], (double)mskp[c - 1]);
  u[c] = (m0 + msk[c + 1] == 2) ? zdy * (ps - p) : 0.;
  v[c] = (m0 + msk[c + nx] == 2) ? zdx * (p - pw) : 0.;
}
// Strip version (nx even, so that every row start is 16-byte aligned): a thread owns two
// adjacent columns and marches OG_ROWS rows north, keeping the masked psi of the row below
// in registers; psi / u / v move as 16-byte vectors, the west neighbour of the pair comes
// from the lane to the left.  ALLFLUID (msk == NULL): the cell mask is 1 everywhere and the
// corner mask is 1 except on the last row and column, so no mask byte is read at all.
constexpr int OG_ROWS = 16;
// RK (Timescheme.RK3_SSP, timescheme.py:172-180): the velocities (u, v) this kernel derives are the
// tendencies du, dv of a stage, and the stage state uo = ub + c*du (ue == NULL) or
// ub + c*(ue + du) -- v alike -- is written while they are in registers, for every cell of the
// array (the outermost ring, which computeorthogradient leaves alone, combines the stored du),
// with numpy's rounding sequence: what f2d_ts_xpay / f2d_ts_xpay2 would compute from the stored
// fields, without reading them back.
struct UVStage {
  const double *ub, *vb, *ue, *ve;
  double *uo, *vo;
  double c;
};
template <bool ALLFLUID, bool RK>
__global__ void __launch_bounds__(128) k_mask_orthogradient_strip(const int8_t *__restrict__ msk,
                                                                  const int8_t *__restrict__ mskp, double *psi,
                                                                  double zdx, double zdy, double *__restrict__ u,
                                                                  double *__restrict__ v, int ny, int nx, UVStage R) {
  const int i = (blockIdx.x * 128 + threadIdx.x) * 2;
  const int j0 = blockIdx.y * OG_ROWS;
  const bool active = i < nx;
  const unsigned lane = threadIdx.x & 31;
  auto masked2 = [&](int j) -> double2 {
    const size_t c = (size_t)j * nx + i;
    double2 p = *reinterpret_cast<const double2 *>(psi + c);
    double m0, m1;
    if (ALLFLUID) {
      m0 = (j < ny - 1) ? 1. : 0.;
      m1 = (j < ny - 1 && i + 1 < nx - 1) ? 1. : 0.;
    } else {
      char2 m = *reinterpret_cast<const char2 *>(mskp + c);
      m0 = (double)m.x;
      m1 = (double)m.y;
    }
    return make_double2(__dmul_rn(p.x, m0), __dmul_rn(p.y, m1));
  };
  auto masked1 = [&](int j, int ii) -> double {
    const size_t c = (size_t)j * nx + ii;
    double m = ALLFLUID ? ((j < ny - 1 && ii < nx - 1) ? 1. : 0.) : (double)mskp[c];
    return __dmul_rn(psi[c], m);
  };
  double2 ps = make_double2(0., 0.);
  if (active && j0 > 0) ps = masked2(j0 - 1);
  for (int r = 0; r < OG_ROWS; r++) {
    const int j = j0 + r;
    if (j >= ny) break;
    double2 p = make_double2(0., 0.);
    if (active) p = masked2(j);
    double pw = __shfl_up_sync(0xffffffffu, p.y, 1);
    if (active) {
      const size_t c = (size_t)j * nx + i;
      *reinterpret_cast<double2 *>(psi + c) = p;
      const bool wr = j >= 1 && j <= ny - 2;
      double2 du = make_double2(0., 0.), dv = make_double2(0., 0.);
      if (RK) {
        // cells this kernel does not write keep their stored tendency
        if (!wr || i < 1) { du.x = u[c]; dv.x = v[c]; }
        if (!wr || i + 1 > nx - 2) { du.y = u[c + 1]; dv.y = v[c + 1]; }
      }
      if (wr) {
        if (lane == 0) pw = i > 0 ? masked1(j, i - 1) : 0.;
        bool ue0 = true, ue1 = true, vn0 = true, vn1 = true;
        if (!ALLFLUID) {
          char2 m = *reinterpret_cast<const char2 *>(msk + c);
          char2 mn = *reinterpret_cast<const char2 *>(msk + c + nx);
          int me = (i + 2 < nx) ? msk[c + 2] : 0;
          ue0 = m.x + m.y == 2;
          ue1 = m.y + me == 2;
          vn0 = m.x + mn.x == 2;
          vn1 = m.y + mn.y == 2;
        }
        double2 uu, vv;
        uu.x = ue0 ? zdy * (ps.x - p.x) : 0.;
        uu.y = ue1 ? zdy * (ps.y - p.y) : 0.;
        vv.x = vn0 ? zdx * (p.x - pw) : 0.;
        vv.y = vn1 ? zdx * (p.y - p.x) : 0.;
        if (i >= 1 && i + 1 <= nx - 2) {
          *reinterpret_cast<double2 *>(u + c) = uu;
          *reinterpret_cast<double2 *>(v + c) = vv;
        } else {
          if (i >= 1) { u[c] = uu.x; v[c] = vv.x; }
          if (i + 1 <= nx - 2) { u[c + 1] = uu.y; v[c + 1] = vv.y; }
        }
        if (RK) {
          if (i >= 1) { du.x = uu.x; dv.x = vv.x; }
          if (i + 1 <= nx - 2) { du.y = uu.y; dv.y = vv.y; }
        }
      }
      if (RK) {
        const double2 ub = *reinterpret_cast<const double2 *>(R.ub + c);
        const double2 vb = *reinterpret_cast<const double2 *>(R.vb + c);
        if (R.ue) {
          const double2 ue = *reinterpret_cast<const double2 *>(R.ue + c);
          const double2 ve = *reinterpret_cast<const double2 *>(R.ve + c);
          du.x = add_rn(ue.x, du.x); du.y = add_rn(ue.y, du.y);
          dv.x = add_rn(ve.x, dv.x); dv.y = add_rn(ve.y, dv.y);
        }
        *reinterpret_cast<double2 *>(R.uo + c) = make_double2(add_rn(ub.x, mul_rn(R.c, du.x)), add_rn(ub.y, mul_rn(R.c, du.y)));
        *reinterpret_cast<double2 *>(R.vo + c) = make_double2(add_rn(vb.x, mul_rn(R.c, dv.x)), add_rn(vb.y, mul_rn(R.c, dv.y)));
      }
    }
    ps = p;
  }
}
/* msk == NULL: all-fluid domain (cell mask 1 everywhere, corner mask 1 except on the last
 * row and column) */
extern "C" int f2d_mask_orthogradient(const int8_t *msk, const int8_t *mskp, double *psi, double dx, double dy,
                                      int nh, double *u, double *v, int ny, int nx, f2d_stream_t s) {
  (void)nh;
  if (!psi || !u || !v || ny < 3 || nx < 3) return fail(F2D_ERR_ARG, "mask_orthogradient: bad args");
  if ((msk == nullptr) != (mskp == nullptr)) return fail(F2D_ERR_ARG, "mask_orthogradient: msk and mskp go together");
  const bool vec_ok = nx % 2 == 0 && ((reinterpret_cast<uintptr_t>(psi) | reinterpret_cast<uintptr_t>(u) |
                                       reinterpret_cast<uintptr_t>(v)) & 15) == 0 &&
                      (!msk || ((reinterpret_cast<uintptr_t>(msk) | reinterpret_cast<uintptr_t>(mskp)) & 1) == 0);
  if (vec_ok) {
    dim3 g(cdiv(nx, 256), cdiv(ny, OG_ROWS));
    UVStage R = {};
    if (msk) k_mask_orthogradient_strip<false, false><<<g, 128, 0, S(s)>>>(msk, mskp, psi, 1. / dx, 1. / dy, u, v, ny, nx, R);
    else k_mask_orthogradient_strip<true, false><<<g, 128, 0, S(s)>>>(msk, mskp, psi, 1. / dx, 1. / dy, u, v, ny, nx, R);
    F2D_LAUNCHED();
    return F2D_OK;
  }
  if (!msk) return fail(F2D_ERR_ARG, "mask_orthogradient: the mask-free form needs even nx and 16-byte aligned fields");
  dim3 b(32, 8);
  k_mask_orthogradient<<<grid2d(ny, nx, b), b, 0, S(s)>>>(msk, mskp, psi, 1. / dx, 1. / dy, u, v, ny, nx);
  F2D_LAUNCHED();
  return F2D_OK;
}

namespace f2d {
// the stage update of the velocities as kernels of its own (f2d_ts_xpay / f2d_ts_xpay2 on u and v)
int uv_stage(const double *u, const double *v, const double *ub, const double *vb, const double *ue,
             const double *ve, double *uo, double *vo, double c, size_t n, f2d_stream_t s) {
  int rc;
  if (ue) {
    rc = f2d_ts_xpay2(uo, ub, c, ue, u, n, s);
    if (rc == F2D_OK) rc = f2d_ts_xpay2(vo, vb, c, ve, v, n, s);
  } else {
    rc = f2d_ts_xpay(uo, ub, c, u, n, s);
    if (rc == F2D_OK) rc = f2d_ts_xpay(vo, vb, c, v, n, s);
  }
  return rc;
}
}  // namespace f2d

/* f2d_mask_orthogradient followed by the Runge-Kutta stage update of the velocities,
 * uo = ub + c*(u) or ub + c*(ue + u) (ue, ve both NULL or both given), v alike, over the whole
 * arrays: one kernel when the strip kernel applies (even nx, 16-byte aligned fields), otherwise the
 * orthogradient followed by f2d_ts_xpay / f2d_ts_xpay2 -- the same numbers either way. */
extern "C" int f2d_mask_orthogradient_stage(const int8_t *msk, const int8_t *mskp, double *psi, double dx, double dy,
                                            int nh, double *u, double *v, const double *ub, const double *vb,
                                            const double *ue, const double *ve, double *uo, double *vo, double c,
                                            int ny, int nx, f2d_stream_t s) {
  if (!psi || !u || !v || !ub || !vb || !uo || !vo || ny < 3 || nx < 3)
    return fail(F2D_ERR_ARG, "mask_orthogradient_stage: bad args");
  if ((msk == nullptr) != (mskp == nullptr)) return fail(F2D_ERR_ARG, "mask_orthogradient_stage: msk and mskp go together");
  if ((ue == nullptr) != (ve == nullptr)) return fail(F2D_ERR_ARG, "mask_orthogradient_stage: ue and ve go together");
  uintptr_t al = reinterpret_cast<uintptr_t>(psi) | reinterpret_cast<uintptr_t>(u) | reinterpret_cast<uintptr_t>(v) |
                 reinterpret_cast<uintptr_t>(ub) | reinterpret_cast<uintptr_t>(vb) | reinterpret_cast<uintptr_t>(uo) |
                 reinterpret_cast<uintptr_t>(vo) | reinterpret_cast<uintptr_t>(ue) | reinterpret_cast<uintptr_t>(ve);
  const bool vec_ok = nx % 2 == 0 && (al & 15) == 0 &&
                      (!msk || ((reinterpret_cast<uintptr_t>(msk) | reinterpret_cast<uintptr_t>(mskp)) & 1) == 0);
  // the stage state must not be one of the arrays the kernel still reads at other cells
  const bool distinct = uo != u && uo != v && vo != u && vo != v && (void *)uo != (void *)psi && (void *)vo != (void *)psi;
  if (vec_ok && distinct) {
    dim3 g(cdiv(nx, 256), cdiv(ny, OG_ROWS));
    UVStage R = {ub, vb, ue, ve, uo, vo, c};
    prof_tag("f2d_mask_orthogradient_stage<%d extra>", ue ? 1 : 0);
    if (msk) k_mask_orthogradient_strip<false, true><<<g, 128, 0, S(s)>>>(msk, mskp, psi, 1. / dx, 1. / dy, u, v, ny, nx, R);
    else k_mask_orthogradient_strip<true, true><<<g, 128, 0, S(s)>>>(msk, mskp, psi, 1. / dx, 1. / dy, u, v, ny, nx, R);
    F2D_LAUNCHED();
    return F2D_OK;
  }
  int rc = f2d_mask_orthogradient(msk, mskp, psi, dx, dy, nh, u, v, ny, nx, s);
  if (rc != F2D_OK) return rc;
  return f2d::uv_stage(u, v, ub, vb, ue, ve, uo, vo, c, (size_t)ny * nx, s);
}

// add_diffusion: fortran_operators.f90:125-156, rows/cols 2..m-1 where msk==1
// jlo..jhi: rows written.  On a y-slab (fill mode 2) only the interior rows: the y halo rows of
// the tendency belong to the neighbours' exchange, which may already be storing into them while
// this kernel runs (a read-modify-write there would corrupt what the neighbour pushed).
__global__ void k_add_diffusion(const int8_t *__restrict__ msk, const double *__restrict__ t,
                                double coef, double *__restrict__ d, int ny, int nx, int jlo, int jhi) {
  IJ();
  if (j < jlo || j > jhi || i < 1 || i > nx - 2) return;
  if (msk[c] != 1) return;
  double tc = t[c];
  double acc = msk[c - 1] * (t[c - 1] - tc);
  acc = acc + msk[c + 1] * (t[c + 1] - tc);
  acc = acc + msk[c - nx] * (t[c - nx] - tc);
  acc = acc + msk[c + nx] * (t[c + nx] - tc);
  d[c] = d[c] + coef * acc;
}
extern "C" int f2d_add_diffusion(const int8_t *msk, const double *trac, double dx, int nh, double Kdiff,
                                 double *dtrac, int ny, int nx, int fill, f2d_stream_t s) {
  if (!msk || !trac || !dtrac || ny < 3 || nx < 3) return fail(F2D_ERR_ARG, "add_diffusion: bad args");
  dim3 b(32, 8);
  const int jlo = fill == 2 ? nh : 1, jhi = fill == 2 ? ny - 1 - nh : ny - 2;
  k_add_diffusion<<<grid2d(ny, nx, b), b, 0, S(s)>>>(msk, trac, Kdiff / (dx * dx), dtrac, ny, nx, jlo, jhi);
  F2D_LAUNCHED();
  if (fill == 2) return f2d_fill_halo_x(dtrac, nh, ny, nx, s);
  if (fill) return f2d_fill_halo(dtrac, nh, ny, nx, s);
  return F2D_OK;
}

// add_torque: fortran_operators.f90:330-381.  ml/mr = max(1, pair sums); the update is
// applied where ml+mr == 4, i.e. msk(i-1)=msk(i)=msk(i+1)=1.  premask: y *= msk first
// on the WHOLE array (operators.py:311).
// slab != 0 (fill mode 2): the y halo rows are left to the neighbours' exchange (see k_add_diffusion)
__global__ void k_add_torque(const int8_t *__restrict__ msk, const double *__restrict__ b, double coef,
                             double *__restrict__ d, int ny, int nx, int nh, int premask, int slab) {
  IJ();
  if (slab && (j < nh || j >= ny - nh)) return;
  double y = d[c];
  int m0 = msk[c];
  bool touched = false;
  if (premask) { y = y * (double)m0; touched = true; }
  if (j >= nh && j < ny - nh && i >= nh && i < nx - nh) {
    int ml = m0 + msk[c - 1];
    if (ml < 1) ml = 1;
    int mr = msk[c + 1] + m0;
    if (mr < 1) mr = 1;
    if (ml + mr == 4) { y = y + ((b[c + 1] - b[c - 1]) * coef) * (double)m0; touched = true; }
  }
  if (touched) d[c] = y;
}
extern "C" int f2d_add_torque(const int8_t *msk, const double *buoy, double dx, int nh, double gravity,
                              double *domega, int ny, int nx, int premask, int fill, f2d_stream_t s) {
  if (!msk || !buoy || !domega || ny <= 2 * nh || nx <= 2 * nh) return fail(F2D_ERR_ARG, "add_torque: bad args");
  dim3 b(32, 8);
  k_add_torque<<<grid2d(ny, nx, b), b, 0, S(s)>>>(msk, buoy, 0.5 * gravity / dx, domega, ny, nx, nh, premask, fill == 2);
  F2D_LAUNCHED();
  if (fill == 2) return f2d_fill_halo_x(domega, nh, ny, nx, s);
  if (fill) return f2d_fill_halo(domega, nh, ny, nx, s);
  return F2D_OK;
}

// computenoslipsourceterm: fortran_operators.f90:221-277 as a gather.
// The Fortran visits (j,i), j=nh+1..m-nh+1, i=nh+1..n-nh+1 in row-major order; at each
// visit it sets y(j,i)=0, then adds face terms either to y(j,i) (cell is fluid) or to
// the already-visited y(j,i-1) / y(j-1,i) (cell is solid).  Gathered per target cell
// T=(j,i), in the order the Fortran performs the additions:
//   1. own west face  (msk(T)!=0, msk(j,i-1)+msk(T)==1)      : y -= vW(j,i)
//   2. own south face (msk(T)!=0, msk(j-1,i)+msk(T)==1)      : y += uS(j,i)
//   3. east neighbour E=(j,i+1) solid, visited, msk(T)+msk(E)==1 : y += vW(j,i+1)
//   4. north neighbour N=(j+1,i) solid, visited, msk(T)+msk(N)==1: y -= uS(j+1,i)
// Cells with row<=nh or (visited row and col<=nh) are zeroed and then only receive 3./4.
// vW(j,i) = (x(j,i)+x(j-1,i)-x(j,i-2)-x(j-1,i-2))*cff ; uS(j,i) = -(x(j,i)+x(j,i-1)-x(j-2,i)-x(j-2,i-1))*cff
__global__ void k_noslip_source(const int8_t *__restrict__ msk, const double *__restrict__ x,
                                double *__restrict__ y, double cff, int ny, int nx, int nh) {
  IJ();
  // 1-based coordinates of the Fortran
  int J = j + 1, I = i + 1;
  int jlo = nh + 1, jhi = ny - nh + 1, ilo = nh + 1, ihi = nx - nh + 1;
  bool visited = (J >= jlo && J <= jhi && I >= ilo && I <= ihi);
  bool zeroed = visited || (J <= nh) || (J >= jlo && J <= jhi && I <= nh);
  // scatter targets can be (j,i-1) with i-1 = nh (zeroed column) and (j-1,i) with j-1 = nh
  if (!zeroed) return;
  double acc = 0.;
  int mT = msk[c];
  if (visited && mT != 0) {
    if (msk[c - 1] + mT == 1) {
      double vW = (((x[c] + x[c - nx]) - x[c - 2]) - x[c - nx - 2]) * cff;
      acc = acc - vW;
    }
    if (msk[c - nx] + mT == 1) {
      double uS = -((((x[c] + x[c - 1]) - x[c - 2 * nx]) - x[c - 2 * nx - 1]) * cff);
      acc = acc + uS;
    }
  }
  // east neighbour visited?
  if (J >= jlo && J <= jhi && (I + 1) >= ilo && (I + 1) <= ihi) {
    int mE = msk[c + 1];
    if (mE == 0 && mT + mE == 1) {
      size_t e = c + 1;
      double vW = (((x[e] + x[e - nx]) - x[e - 2]) - x[e - nx - 2]) * cff;
      acc = acc + vW;
    }
  }
  if ((J + 1) >= jlo && (J + 1) <= jhi && I >= ilo && I <= ihi) {
    int mN = msk[c + nx];
    if (mN == 0 && mT + mN == 1) {
      size_t n_ = c + nx;
      double uS = -((((x[n_] + x[n_ - 1]) - x[n_ - 2 * nx]) - x[n_ - 2 * nx - 1]) * cff);
      acc = acc - uS;
    }
  }
  y[c] = acc;
}
extern "C" int f2d_noslip_source(const int8_t *msk, const double *psi, double *y, double dx, double dy,
                                 int nh, int ny, int nx, f2d_stream_t s) {
  if (!msk || !psi || !y || ny <= 2 * nh || nx <= 2 * nh) return fail(F2D_ERR_ARG, "noslip_source: bad args");
  dim3 b(32, 8);
  k_noslip_source<<<grid2d(ny, nx, b), b, 0, S(s)>>>(msk, psi, y, 1. / (2 * dx * dy), ny, nx, nh);
  F2D_LAUNCHED();
  return F2D_OK;
}

// ---------------------------------------------------------------------------
// elementwise: time-scheme combinations and model glue.  __dmul_rn/__dadd_rn keep
// numpy's rounding sequence (a product is rounded before it is added).
// ---------------------------------------------------------------------------
template <class F>
__global__ void k_elementwise(size_t n, F f) {
  for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) f(k);
}
template <class F>
static int elementwise(size_t n, cudaStream_t s, F f) {
  if (n == 0) return F2D_OK;
  int threads = 256;
  long long blocks = (long long)((n + threads - 1) / threads);
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  prof_tag("k_elementwise %zu doubles", n);
  k_elementwise<<<(int)blocks, threads, 0, s>>>(n, f);
  F2D_LAUNCHED();
  return F2D_OK;
}

// Streaming map out[k] = f(in0[k], ..., in{N-1}[k]) for the whole-state combinations of
// the time schemes: 16-byte accesses, two vectors per thread and iteration with every load
// issued before the first store (out may alias an input, so the compiler cannot hoist them
// itself) -- 4..10 independent 16 B loads in flight per thread keep HBM3e busy.
template <int NIN>
struct VecIn { const double2 *p[NIN]; };
template <int NIN, class F>
__device__ __forceinline__ double2 vec_apply(const double2 (&a)[NIN], F f) {
  double lo[NIN], hi[NIN];
#pragma unroll
  for (int q = 0; q < NIN; q++) { lo[q] = a[q].x; hi[q] = a[q].y; }
  return make_double2(f(lo), f(hi));
}
template <int NIN, class F>
__global__ void __launch_bounds__(256) k_map_vec(size_t nvec, double2 *out, VecIn<NIN> in, F f) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  for (; k + stride < nvec; k += 2 * stride) {
    double2 a[NIN], b[NIN];
#pragma unroll
    for (int q = 0; q < NIN; q++) a[q] = in.p[q][k];
#pragma unroll
    for (int q = 0; q < NIN; q++) b[q] = in.p[q][k + stride];
    out[k] = vec_apply<NIN>(a, f);
    out[k + stride] = vec_apply<NIN>(b, f);
  }
  if (k < nvec) {
    double2 a[NIN];
#pragma unroll
    for (int q = 0; q < NIN; q++) a[q] = in.p[q][k];
    out[k] = vec_apply<NIN>(a, f);
  }
}
template <int NIN, class F>
__global__ void k_map_scalar(size_t k0, size_t n, double *out, VecIn<NIN> in, F f) {
  for (size_t k = k0 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
    double a[NIN];
#pragma unroll
    for (int q = 0; q < NIN; q++) a[q] = reinterpret_cast<const double *>(in.p[q])[k];
    out[k] = f(a);
  }
}
template <int NIN, class F>
static int map_fields(size_t n, cudaStream_t s, double *out, const double *const (&inp)[NIN], F f) {
  if (n == 0) return F2D_OK;
  VecIn<NIN> in;
  bool aligned = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  for (int q = 0; q < NIN; q++) {
    in.p[q] = reinterpret_cast<const double2 *>(inp[q]);
    aligned = aligned && (reinterpret_cast<uintptr_t>(inp[q]) & 15) == 0;
  }
  size_t nvec = aligned ? n / 2 : 0;
  if (nvec) {
    long long blocks = (long long)((nvec + 511) / 512);
    if (blocks > 148LL * 8) blocks = 148LL * 8;
    prof_tag("k_map_vec<%d in> %zu doubles", NIN, n);
    k_map_vec<NIN><<<(int)blocks, 256, 0, s>>>(nvec, reinterpret_cast<double2 *>(out), in, f);
    F2D_LAUNCHED();
  }
  if (2 * nvec < n) {   // unaligned fields, or the odd last element
    size_t rest = n - 2 * nvec;
    long long blocks = (long long)((rest + 255) / 256);
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    k_map_scalar<NIN><<<(int)blocks, 256, 0, s>>>(2 * nvec, n, out, in, f);
    F2D_LAUNCHED();
  }
  return F2D_OK;
}

extern "C" int f2d_ts_axpy(double *y, double c, const double *a, size_t n, f2d_stream_t s) {
  const double *const in[2] = {y, a};
  return map_fields<2>(n, S(s), y, in, [=] __device__(const double (&v)[2]) { return add_rn(v[0], mul_rn(c, v[1])); });
}
extern "C" int f2d_ts_xpay(double *out, const double *x, double c, const double *a, size_t n, f2d_stream_t s) {
  const double *const in[2] = {x, a};
  return map_fields<2>(n, S(s), out, in, [=] __device__(const double (&v)[2]) { return add_rn(v[0], mul_rn(c, v[1])); });
}
extern "C" int f2d_ts_xpay2(double *out, const double *x, double c, const double *a, const double *b, size_t n,
                            f2d_stream_t s) {
  const double *const in[3] = {x, a, b};
  return map_fields<3>(n, S(s), out, in,
                       [=] __device__(const double (&v)[3]) { return add_rn(v[0], mul_rn(c, add_rn(v[1], v[2]))); });
}
extern "C" int f2d_ts_rk3ssp_final(double *x, double c, const double *a, const double *b, const double *d,
                                   size_t n, f2d_stream_t s) {
  const double *const in[4] = {x, a, b, d};
  return map_fields<4>(n, S(s), x, in, [=] __device__(const double (&v)[4]) {
    return add_rn(v[0], mul_rn(c, add_rn(add_rn(v[1], v[2]), mul_rn(4., v[3]))));
  });
}
extern "C" int f2d_ts_ab2(double *x, double c0, const double *a, double c1, const double *b, size_t n,
                          f2d_stream_t s) {
  const double *const in[3] = {x, a, b};
  return map_fields<3>(n, S(s), x, in, [=] __device__(const double (&v)[3]) {
    return add_rn(v[0], add_rn(mul_rn(c0, v[1]), -mul_rn(c1, v[2])));
  });
}
extern "C" int f2d_ts_ab3(double *x, double c0, const double *a, double c1, const double *b, double c2,
                          const double *d, size_t n, f2d_stream_t s) {
  const double *const in[4] = {x, a, b, d};
  return map_fields<4>(n, S(s), x, in, [=] __device__(const double (&v)[4]) {
    return add_rn(v[0], add_rn(add_rn(mul_rn(c0, v[1]), -mul_rn(c1, v[2])), mul_rn(c2, v[3])));
  });
}
extern "C" int f2d_ts_set_xpay(double *x, const double *xb, double c, const double *a, size_t n, f2d_stream_t s) {
  const double *const in[2] = {xb, a};
  return map_fields<2>(n, S(s), x, in, [=] __device__(const double (&v)[2]) { return add_rn(v[0], mul_rn(c, v[1])); });
}
extern "C" int f2d_ts_asselin(double *xs, double c, const double *x, const double *xb, size_t n, f2d_stream_t s) {
  const double *const in[3] = {xs, x, xb};
  return map_fields<3>(n, S(s), xs, in, [=] __device__(const double (&v)[3]) {
    return add_rn(v[0], mul_rn(c, add_rn(add_rn(v[1], v[2]), -mul_rn(2., v[0]))));
  });
}
extern "C" int f2d_ts_am3(double *x, const double *xs, const double *xb, size_t n, f2d_stream_t s) {
  const double w = 1. / 12.;
  const double *const in[3] = {x, xs, xb};
  return map_fields<3>(n, S(s), x, in, [=] __device__(const double (&v)[3]) {
    return mul_rn(w, add_rn(add_rn(mul_rn(5., v[0]), mul_rn(8., v[1])), -v[2]));
  });
}
extern "C" int f2d_mul_field(double *y, const double *a, size_t n, f2d_stream_t s) {
  return elementwise(n, S(s), [=] __device__(size_t k) { y[k] = mul_rn(y[k], a[k]); });
}
extern "C" int f2d_mul_mask(double *y, const int8_t *a, size_t n, f2d_stream_t s) {
  return elementwise(n, S(s), [=] __device__(size_t k) { y[k] = mul_rn(y[k], (double)a[k]); });
}
extern "C" int f2d_scale(double *y, double alpha, size_t n, f2d_stream_t s) {
  return elementwise(n, S(s), [=] __device__(size_t k) { y[k] = mul_rn(y[k], alpha); });
}
extern "C" int f2d_add_scaled(double *y, double alpha, const double *a, size_t n, f2d_stream_t s) {
  return elementwise(n, S(s), [=] __device__(size_t k) { y[k] = add_rn(y[k], mul_rn(alpha, a[k])); });
}
extern "C" int f2d_add_scaled_mask(double *y, double alpha, const int8_t *a, size_t n, f2d_stream_t s) {
  return elementwise(n, S(s), [=] __device__(size_t k) { y[k] = add_rn(y[k], mul_rn(alpha, (double)a[k])); });
}
extern "C" int f2d_set_sum(double *y, const double *a, double alpha, const double *b, size_t n, f2d_stream_t s) {
  return elementwise(n, S(s), [=] __device__(size_t k) { y[k] = add_rn(a[k], mul_rn(alpha, b[k])); });
}
// core/fluxes.py:120-127: cell-centred velocities uc = 0.5*(u + roll(u,1,axis=1)),
// vc = 0.5*(v + roll(v,1,axis=0)) on the interior, periodic images stored by the
// owning cell (the fill_halo that follows each in the reference)
__global__ void k_flx_cellvel(const double *__restrict__ u, const double *__restrict__ v, double *__restrict__ uc,
                              double *__restrict__ vc, int nh, int ny, int nx, int fill) {
  int i = blockIdx.x * blockDim.x + threadIdx.x + nh;
  int j = blockIdx.y * blockDim.y + threadIdx.y + nh;
  if (i >= nx - nh || j >= ny - nh) return;
  size_t c = (size_t)j * nx + i;
  double a = mul_rn(0.5, add_rn(u[c], u[c - 1]));
  double b = mul_rn(0.5, add_rn(v[c], v[c - nx]));
  uc[c] = a;
  vc[c] = b;
  if (fill)
    for_each_halo_image(j, i, ny, nx, nh, [&](int jj, int ii) {
      uc[(size_t)jj * nx + ii] = a;
      vc[(size_t)jj * nx + ii] = b;
    }, fill == 1);
}
extern "C" int f2d_flx_cellvel(const double *u, const double *v, double *uc, double *vc, int nh, int ny, int nx,
                               int fill_halo, f2d_stream_t s) {
  if (nh < 1 || ny <= 2 * nh || nx <= 2 * nh) return fail(F2D_ERR_ARG, "flx_cellvel: bad shape");
  dim3 blk(32, 8), grd(cdiv(nx - 2 * nh, 32), cdiv(ny - 2 * nh, 8));
  k_flx_cellvel<<<grd, blk, 0, S(s)>>>(u, v, uc, vc, nh, ny, nx, fill_halo);
  F2D_LAUNCHED();
  return F2D_OK;
}
// core/fluxes.py:160-177: rev = cff*(fwd + sign*bwd), irr = cff*(fwd - sign*bwd), sign = +-1
extern "C" int f2d_flx_split(double *rev, double *irr, const double *fwd, const double *bwd, double cff, double sign,
                             size_t n, f2d_stream_t s) {
  return elementwise(n, S(s), [=] __device__(size_t k) {
    double sb = mul_rn(sign, bwd[k]);
    rev[k] = mul_rn(cff, add_rn(fwd[k], sb));
    irr[k] = mul_rn(cff, add_rn(fwd[k], -sb));
  });
}
// history snapshot (output.py:90-95, NcfileIO.write): the interior of a field, cast to
// float32 on the device, packed [ny-2nh][nx-2nh] -- half the bytes cross PCIe
__global__ void k_pack_interior_f32(const double *__restrict__ x, float *__restrict__ out, int nh, int ny, int nx) {
  const int mi = nx - 2 * nh, mj = ny - 2 * nh;
  const size_t n = (size_t)mi * mj;
  for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(p / mi), i = (int)(p - (size_t)j * mi);
    out[p] = (float)x[(size_t)(j + nh) * nx + i + nh];
  }
}
extern "C" int f2d_pack_interior_f32(const double *x, float *out, int nh, int ny, int nx, f2d_stream_t s) {
  if (!x || !out || nh < 0 || ny <= 2 * nh || nx <= 2 * nh) return fail(F2D_ERR_ARG, "pack_interior_f32: bad args");
  const size_t n = (size_t)(ny - 2 * nh) * (nx - 2 * nh);
  long long blocks = (long long)((n + 255) / 256);
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  k_pack_interior_f32<<<(int)blocks, 256, 0, S(s)>>>(x, out, nh, ny, nx);
  F2D_LAUNCHED();
  return F2D_OK;
}
extern "C" int f2d_div_scalar(double *y, double d, size_t n, f2d_stream_t s) {
  return elementwise(n, S(s), [=] __device__(size_t k) { y[k] = __ddiv_rn(y[k], d); });
}
extern "C" int f2d_sub_lin2_mask(double *y, double pa, const double *a, double pb, const double *b,
                                 const int8_t *mask, size_t n, f2d_stream_t s) {
  return elementwise(n, S(s), [=] __device__(size_t k) {
    y[k] = add_rn(y[k], -mul_rn(add_rn(mul_rn(pa, a[k]), mul_rn(pb, b[k])), (double)mask[k]));
  });
}
extern "C" int f2d_sub_lin2(double *y, double pa, const double *a, double pb, const double *b, size_t n,
                            f2d_stream_t s) {
  return elementwise(n, S(s), [=] __device__(size_t k) {
    y[k] = add_rn(y[k], -add_rn(mul_rn(a[k], pa), mul_rn(b[k], pb)));
  });
}
extern "C" int f2d_sub_devscalar(double *y, const double *dev_scalar, double denom, size_t n, f2d_stream_t s) {
  return elementwise(n, S(s), [=] __device__(size_t k) { y[k] = add_rn(y[k], -__ddiv_rn(dev_scalar[0], denom)); });
}
extern "C" int f2d_sub_devscalar_mask(double *y, const double *dev_scalar, double denom, const int8_t *a, size_t n,
                                      f2d_stream_t s) {
  return elementwise(n, S(s), [=] __device__(size_t k) {
    y[k] = add_rn(y[k], -mul_rn(__ddiv_rn(dev_scalar[0], denom), (double)a[k]));
  });
}

// ---------------------------------------------------------------------------
// thermal-wind model (core/thermalwind.py, operators.py:330-394): centred differences with
// the reference's in-place linear extrapolation of the first halo line, the two right-hand-
// side terms and the Jacobian of the Ertel PV.  numpy rounding sequence (no FMA).
// ---------------------------------------------------------------------------
// operators.py:332-335 (axis 0: diffx / diff1x) and :348-351 (axis 1: diffz):
//   x[:, -nh] = 2*x[:, -nh-1] - x[:, -nh-2];  x[:, nh-1] = 2*x[:, nh] - x[:, nh+1]
__global__ void k_extrapolate_bry(double *__restrict__ x, int nh, int ny, int nx, int axis) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (axis == 0) {
    if (p >= ny) return;
    double *r = x + (size_t)p * nx;
    r[nx - nh] = add_rn(mul_rn(2., r[nx - nh - 1]), -r[nx - nh - 2]);
    r[nh - 1] = add_rn(mul_rn(2., r[nh]), -r[nh + 1]);
  } else {
    if (p >= nx) return;
    x[(size_t)(ny - nh) * nx + p] = add_rn(mul_rn(2., x[(size_t)(ny - nh - 1) * nx + p]), -x[(size_t)(ny - nh - 2) * nx + p]);
    x[(size_t)(nh - 1) * nx + p] = add_rn(mul_rn(2., x[(size_t)nh * nx + p]), -x[(size_t)(nh + 1) * nx + p]);
  }
}
extern "C" int f2d_extrapolate_bry(double *x, int nh, int ny, int nx, int axis, f2d_stream_t s) {
  if (!x || nh < 1 || ny < 2 * nh + 2 || nx < 2 * nh + 2 || (axis != 0 && axis != 1))
    return fail(F2D_ERR_ARG, "extrapolate_bry: bad args");
  int n = axis == 0 ? ny : nx;
  k_extrapolate_bry<<<cdiv(n, 128), 128, 0, S(s)>>>(x, nh, ny, nx, axis);
  F2D_LAUNCHED();
  return F2D_OK;
}
__device__ __forceinline__ double tw_diffx(const double *__restrict__ a, size_t c, double dx) {
  return __ddiv_rn(mul_rn(0.5, add_rn(a[c + 1], -a[c - 1])), dx);
}
__device__ __forceinline__ double tw_diffz(const double *__restrict__ a, size_t c, int nx, double dy) {
  return __ddiv_rn(mul_rn(0.5, add_rn(a[c + nx], -a[c - nx])), dy);
}
// operators.py:374-382: y[1:-1,1:-1] = diffx(b)*gravity - diffz(V)*f0 ; y *= msk
// (the outer ring of y is left to the halo fill that follows)
__global__ void k_tw_torque(const int8_t *__restrict__ msk, const double *__restrict__ b, const double *__restrict__ V,
                            double dx, double dy, double gravity, double f0, double *__restrict__ y, int ny, int nx) {
  int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  if (i >= nx - 1 || j >= ny - 1) return;
  size_t c = (size_t)j * nx + i;
  double t = mul_rn(tw_diffx(b, c, dx), gravity);
  t = add_rn(t, -mul_rn(tw_diffz(V, c, nx, dy), f0));
  y[c] = mul_rn(t, (double)msk[c]);
}
extern "C" int f2d_tw_torque(const int8_t *msk, const double *b, const double *V, double dx, double dy,
                             double gravity, double f0, double *y, int ny, int nx, f2d_stream_t s) {
  if (!msk || !b || !V || !y || ny < 3 || nx < 3) return fail(F2D_ERR_ARG, "tw_torque: bad args");
  dim3 blk(32, 8), grd(cdiv(nx - 2, 32), cdiv(ny - 2, 8));
  k_tw_torque<<<grd, blk, 0, S(s)>>>(msk, b, V, dx, dy, gravity, f0, y, ny, nx);
  F2D_LAUNCHED();
  return F2D_OK;
}
// operators.py:389-391: y[:, 1:] = -0.5*f0*(u[:, :-1] + u[:, 1:]) ; y *= msk
__global__ void k_tw_coriolis(const int8_t *__restrict__ msk, const double *__restrict__ u, double f0,
                              double *__restrict__ y, int ny, int nx) {
  int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= nx || j >= ny) return;
  size_t c = (size_t)j * nx + i;
  double t = mul_rn(mul_rn(-0.5, f0), add_rn(u[c - 1], u[c]));
  y[c] = mul_rn(t, (double)msk[c]);
}
extern "C" int f2d_tw_coriolis(const int8_t *msk, const double *u, double f0, double *y, int ny, int nx,
                               f2d_stream_t s) {
  if (!msk || !u || !y || ny < 1 || nx < 2) return fail(F2D_ERR_ARG, "tw_coriolis: bad args");
  dim3 blk(32, 8), grd(cdiv(nx - 1, 32), cdiv(ny, 8));
  k_tw_coriolis<<<grd, blk, 0, S(s)>>>(msk, u, f0, y, ny, nx);
  F2D_LAUNCHED();
  return F2D_OK;
}
// operators.py:354-355 + thermalwind.py:92-95: out = 0 ; out[1:-1,1:-1] =
// diffx(x)*diffz(y) - diffz(x)*diffx(y) ; out *= msk   (the boundary lines of x and y must
// have been extrapolated by f2d_extrapolate_bry, as diffx / diffz do in place)
__global__ void k_jacobian(const int8_t *__restrict__ msk, const double *__restrict__ x, const double *__restrict__ y,
                           double dx, double dy, double *__restrict__ out, int ny, int nx) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= nx || j >= ny) return;
  size_t c = (size_t)j * nx + i;
  double t = 0.;
  if (i >= 1 && i < nx - 1 && j >= 1 && j < ny - 1) {
    t = add_rn(mul_rn(tw_diffx(x, c, dx), tw_diffz(y, c, nx, dy)), -mul_rn(tw_diffz(x, c, nx, dy), tw_diffx(y, c, dx)));
  }
  out[c] = mul_rn(t, (double)msk[c]);
}
extern "C" int f2d_jacobian(const int8_t *msk, const double *x, const double *y, double dx, double dy, double *out,
                            int ny, int nx, f2d_stream_t s) {
  if (!msk || !x || !y || !out || ny < 3 || nx < 3) return fail(F2D_ERR_ARG, "jacobian: bad args");
  dim3 blk(32, 8), grd(cdiv(nx, 32), cdiv(ny, 8));
  k_jacobian<<<grd, blk, 0, S(s)>>>(msk, x, y, dx, dy, out, ny, nx);
  F2D_LAUNCHED();
  return F2D_OK;
}
// thermalwind.py:136-137: qEneg = qE.copy(); qEneg[qE > 0] = 0
extern "C" int f2d_negative_part(double *out, const double *x, size_t n, f2d_stream_t s) {
  if (!out || !x) return fail(F2D_ERR_ARG, "negative_part: null");
  return elementwise(n, S(s), [=] __device__(size_t k) { out[k] = x[k] > 0. ? 0. : x[k]; });
}
