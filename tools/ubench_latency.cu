// Micro-benchmarks behind the latency-bound coarse-level design (B200, sm_100a): dependent fp64
// chain latency, fp64 issue rate of one SM, shared-memory load latency, __syncthreads and cluster
// barrier cost, integer division.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ub tools/ubench_latency.cu && /tmp/ub
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__global__ void k_dfma_chain(double *out, long long *cyc, double a, double b) {
  double x = threadIdx.x * 1e-3;
  long long t0 = clock64();
#pragma unroll 1
  for (int k = 0; k < 64; k++) {
#pragma unroll
    for (int q = 0; q < 16; q++) x = fma(x, a, b);
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_dadd_chain(double *out, long long *cyc, double a) {
  double x = threadIdx.x * 1e-3;
  long long t0 = clock64();
#pragma unroll 1
  for (int k = 0; k < 64; k++) {
#pragma unroll
    for (int q = 0; q < 16; q++) x = __dadd_rn(x, a);
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// 8 independent chains per thread: issue rate
__global__ void k_dfma_tput(double *out, long long *cyc, double a, double b) {
  double x[8];
  for (int q = 0; q < 8; q++) x[q] = threadIdx.x * 1e-3 + q;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int k = 0; k < 64; k++) {
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int q = 0; q < 8; q++) x[q] = fma(x[q], a, b);
  }
  __syncthreads();
  long long t1 = clock64();
  double s = 0;
  for (int q = 0; q < 8; q++) s += x[q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_lds_chain(double *out, long long *cyc) {
  __shared__ int idx[1024];
  for (int k = threadIdx.x; k < 1024; k += blockDim.x) idx[k] = (k + 33) & 1023;
  __syncthreads();
  int p = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int k = 0; k < 256; k++) p = idx[p];
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = p;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_syncthreads(double *out, long long *cyc) {
  long long t0 = clock64();
#pragma unroll 1
  for (int k = 0; k < 256; k++) __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_clustersync(double *out, long long *cyc) {
  cg::cluster_group cl = cg::this_cluster();
  cl.sync();
  long long t0 = clock64();
#pragma unroll 1
  for (int k = 0; k < 256; k++) cl.sync();
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// barrier.cluster.arrive.release + wait.acquire without the L1 invalidation? (same thing spelled out)
__global__ void k_clusterbar_relaxed(double *out, long long *cyc) {
  cg::cluster_group cl = cg::this_cluster();
  cl.sync();
  long long t0 = clock64();
#pragma unroll 1
  for (int k = 0; k < 256; k++) {
    asm volatile("barrier.cluster.arrive.relaxed.aligned;\nbarrier.cluster.wait.aligned;\n" ::: "memory");
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_idiv(int *out, long long *cyc, int w) {
  int p = threadIdx.x + 12345;
  long long t0 = clock64();
#pragma unroll 1
  for (int k = 0; k < 256; k++) p = p / w + 100000 + k;
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = p;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// a 9-point damped-Jacobi phase as the tail does it: 9 LDS + b, 8 FMA chain, store, barrier
__global__ void k_jacobi_phase(double *out, long long *cyc, int n, int reps, double c0, double c1, double c2) {
  extern __shared__ double sm[];
  const int nx = n + 6;
  double *x = sm, *t = sm + nx * nx, *b = t + nx * nx;
  for (int k = threadIdx.x; k < nx * nx; k += blockDim.x) { x[k] = k * 1e-3; t[k] = 0; b[k] = 1.; }
  __syncthreads();
  const int lg = 31 - __clz(n);
  long long t0 = clock64();
#pragma unroll 1
  for (int r = 0; r < reps; r++) {
    for (int p = threadIdx.x; p < n * n; p += blockDim.x) {
      const int j = 3 + (p >> lg), i = 3 + (p & (n - 1));
      const double *q = x + j * nx + i;
      double acc = c0 * q[-nx - 1];
      acc = acc + c1 * q[-nx];
      acc = acc + c0 * q[-nx + 1];
      acc = acc + c1 * q[-1];
      acc = acc + c1 * q[1];
      acc = acc + c0 * q[nx - 1];
      acc = acc + c1 * q[nx];
      acc = acc + c0 * q[nx + 1];
      t[j * nx + i] = q[0] * c2 + c2 * (acc - b[j * nx + i]);
    }
    __syncthreads();
    double *tmp = x; x = t; t = tmp;
  }
  long long t1 = clock64();
  out[threadIdx.x] = x[threadIdx.x % (nx * nx)];
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// the same phase on an interior-only periodic array (wrap by index masks, no halo images)
__global__ void k_jacobi_phase_periodic(double *out, long long *cyc, int n, int reps, double c0, double c1, double c2) {
  extern __shared__ double sm[];
  double *x = sm, *t = sm + n * n, *b = t + n * n;
  for (int k = threadIdx.x; k < n * n; k += blockDim.x) { x[k] = k * 1e-3; t[k] = 0; b[k] = 1.; }
  __syncthreads();
  const int lg = 31 - __clz(n), nm = n - 1;
  long long t0 = clock64();
#pragma unroll 1
  for (int r = 0; r < reps; r++) {
    for (int p = threadIdx.x; p < n * n; p += blockDim.x) {
      const int j = p >> lg, i = p & nm;
      const int lo = ((j - 1) & nm) << lg, mid = j << lg, hi = ((j + 1) & nm) << lg, il = (i - 1) & nm, ir = (i + 1) & nm;
      double acc = c0 * x[lo + il];
      acc = acc + c1 * x[lo + i];
      acc = acc + c0 * x[lo + ir];
      acc = acc + c1 * x[mid + il];
      acc = acc + c1 * x[mid + ir];
      acc = acc + c0 * x[hi + il];
      acc = acc + c1 * x[hi + i];
      acc = acc + c0 * x[hi + ir];
      t[p] = x[p] * c2 + c2 * (acc - b[p]);
    }
    __syncthreads();
    double *tmp = x; x = t; t = tmp;
  }
  long long t1 = clock64();
  out[threadIdx.x] = x[threadIdx.x % (n * n)];
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 4096);
  long long h[64];
  auto rd = [&](const char *name, double per) {
    cudaDeviceSynchronize();
    cudaError_t e = cudaGetLastError();
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    printf("%-44s %8.1f cycles per item (%s)\n", name, h[0] / per, cudaGetErrorString(e));
  };
  for (int rep = 0; rep < 1; rep++) {
    k_dfma_chain<<<1, 32>>>(out, cyc, 1.0000001, 1e-9); rd("DFMA dependent chain, 1 warp", 1024);
    k_dadd_chain<<<1, 32>>>(out, cyc, 1e-9); rd("DADD dependent chain, 1 warp", 1024);
    k_dfma_tput<<<1, 1024>>>(out, cyc, 1.0000001, 1e-9); rd("DFMA 32 warps x 8 chains: cycles per warp-instr/SM", 64.*32*32);
    k_dfma_tput<<<1, 128>>>(out, cyc, 1.0000001, 1e-9); rd("DFMA 4 warps x 8 chains: cycles per warp-instr/SM", 64.*32*4);
    k_lds_chain<<<1, 32>>>(out, cyc); rd("LDS dependent chain", 256);
    k_syncthreads<<<1, 1024>>>(out, cyc); rd("__syncthreads, 1024 threads", 256);
    k_syncthreads<<<1, 256>>>(out, cyc); rd("__syncthreads, 256 threads", 256);
    k_idiv<<<1, 32>>>((int *)out, cyc, 130); rd("int division (runtime divisor) chain", 256);
    for (int nc : {2, 8, 16}) {
      for (int nt : {1024, 256}) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(nc); cfg.blockDim = dim3(nt);
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = nc; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      cudaFuncSetAttribute(k_clustersync, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      cudaFuncSetAttribute(k_clusterbar_relaxed, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      cudaLaunchKernelEx(&cfg, k_clustersync, out, cyc);
      char nm[96]; snprintf(nm, sizeof nm, "cluster.sync, %d CTAs x %d threads", nc, nt); rd(nm, 256);
      cudaLaunchKernelEx(&cfg, k_clusterbar_relaxed, out, cyc);
      snprintf(nm, sizeof nm, "barrier.cluster relaxed, %d CTAs x %d thr", nc, nt); rd(nm, 256);
      }
    }
    cudaFuncSetAttribute(k_jacobi_phase, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_jacobi_phase_periodic, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int n : {4, 8, 16, 32, 64}) {
      for (int nt : {256, 1024}) {
        k_jacobi_phase<<<1, nt, 3 * (n + 6) * (n + 6) * 8>>>(out, cyc, n, 64, 0.25, 0.5, 0.1);
        char nm[96]; snprintf(nm, sizeof nm, "jacobi phase %d^2, %d threads", n, nt); rd(nm, 64);
        k_jacobi_phase_periodic<<<1, nt, 3 * n * n * 8>>>(out, cyc, n, 64, 0.25, 0.5, 0.1);
        snprintf(nm, sizeof nm, "jacobi phase periodic %d^2, %d threads", n, nt); rd(nm, 64);
      }
    }
  }
  return 0;
}
