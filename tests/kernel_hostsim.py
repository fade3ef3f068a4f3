"""Host-side functional simulation of selected CUDA kernels.  TEST INFRASTRUCTURE.

The source text of a kernel is cut out of fluid2d_b200/csrc/*.cu and compiled by g++ behind a
few shims (`__global__` etc. empty, threadIdx / blockIdx / blockDim as globals, `__syncthreads`
a no-op, the round-to-nearest intrinsics as plain operators under -ffp-contract=off) and driven
with ONE thread per block over the whole grid.  With a single thread the kernel body runs as
sequential code, so this checks what a kernel computes -- index arithmetic, operation order,
loop bounds -- against the oracle without a GPU.  It cannot see races, launch geometry or
anything else that only exists on the device: the -m gpu tests remain the parity tests.
Used for the kernels that were written when no GPU was available.
"""
import ctypes
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), "fluid2d_b200", "csrc")
BUILD = os.path.join(HERE, "_hostsim")

SHIM = r"""
#include <cstddef>
#include <cstdint>
#include <cmath>
struct dim3_ { int x, y, z; };
static dim3_ threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, blockDim = {1, 1, 1}, gridDim = {1, 1, 1};
#define __global__
#define __device__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__
static inline void __syncthreads() {}
static inline double mul_rn(double a, double b) { return a * b; }
static inline double add_rn(double a, double b) { return a + b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
constexpr int NH = 3;
static double tri_smem_storage[1 << 16];
#define HOSTSIM_EXTERN_SHARED(name) double *name = tri_smem_storage
"""


def cut(path, signature_regex):
    """source text of the function whose header matches signature_regex (brace matching)"""
    src = open(path).read()
    m = re.search(signature_regex, src)
    if not m:
        raise LookupError(signature_regex)
    start = m.start()
    k = src.index("{", m.end()-1)
    depth = 0
    for p in range(k, len(src)):
        if src[p] == "{":
            depth += 1
        elif src[p] == "}":
            depth -= 1
            if depth == 0:
                return src[start:p+1]
    raise LookupError("unbalanced braces after " + signature_regex)


DRIVERS = r"""
extern "C" void sim_smooth_tridiag(const int8_t *msk, const double *A, double *x, const double *b, int ny, int nx) {
  k_smooth_tridiag(msk, A, x, b, ny, nx);
}
template <class F> static void grid2(int gx, int gy, F f) {
  for (int by = 0; by < gy; by++) for (int bx = 0; bx < gx; bx++) { blockIdx.x = bx; blockIdx.y = by; f(); }
  blockIdx.x = blockIdx.y = 0;
}
extern "C" void sim_extrapolate_bry(double *x, int nh, int ny, int nx, int axis) {
  grid2(axis == 0 ? ny : nx, 1, [&] { k_extrapolate_bry(x, nh, ny, nx, axis); });
}
extern "C" void sim_tw_torque(const int8_t *msk, const double *b, const double *V, double dx, double dy, double g,
                              double f0, double *y, int ny, int nx) {
  grid2(nx, ny, [&] { k_tw_torque(msk, b, V, dx, dy, g, f0, y, ny, nx); });
}
extern "C" void sim_tw_coriolis(const int8_t *msk, const double *u, double f0, double *y, int ny, int nx) {
  grid2(nx, ny, [&] { k_tw_coriolis(msk, u, f0, y, ny, nx); });
}
extern "C" void sim_jacobian(const int8_t *msk, const double *x, const double *y, double dx, double dy, double *out,
                             int ny, int nx) {
  grid2(nx, ny, [&] { k_jacobian(msk, x, y, dx, dy, out, ny, nx); });
}
"""


def build():
    mg = os.path.join(CSRC, "f2d_multigrid.cu")
    op = os.path.join(CSRC, "f2d_operators.cu")
    parts = [SHIM]
    tri = cut(mg, r"__global__ void __launch_bounds__\(256\)\s*k_smooth_tridiag\(")
    tri = tri.replace("extern __shared__ double tri_smem[];", "HOSTSIM_EXTERN_SHARED(tri_smem);")
    parts.append(tri)
    parts.append(cut(op, r"__global__ void k_extrapolate_bry\("))
    parts.append(cut(op, r"__device__ __forceinline__ double tw_diffx\("))
    parts.append(cut(op, r"__device__ __forceinline__ double tw_diffz\("))
    parts.append(cut(op, r"__global__ void k_tw_torque\("))
    parts.append(cut(op, r"__global__ void k_tw_coriolis\("))
    parts.append(cut(op, r"__global__ void k_jacobian\("))
    parts.append(DRIVERS)
    os.makedirs(BUILD, exist_ok=True)
    cpp = os.path.join(BUILD, "hostsim.cpp")
    so = os.path.join(BUILD, "libhostsim.so")
    text = "\n".join(parts)
    if not (os.path.exists(cpp) and open(cpp).read() == text and os.path.exists(so)):
        open(cpp, "w").write(text)
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-shared", "-fPIC", cpp, "-o", so])
    return ctypes.CDLL(so)
