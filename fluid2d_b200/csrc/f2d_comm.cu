// Multi-GPU plumbing for the y-slab decomposition: a symmetric heap shared through CUDA
// IPC, halo-row exchange by direct peer stores over NVLink, and device-side lock-step
// synchronisation -- the replacement of the reference's mpi4py layer (gmg/halo.py:
// Send_init/Recv_init + Startall/Waitall per fill; gmg/subdomains.py: Allgatherv;
// level.py:401 allreduce; mpitools.py allgather).
//
// Model: one process per GPU, rank r owns rows [r*ny/G, (r+1)*ny/G) of the global grid
// (npx = 1, npy = G), x is periodic inside the slab.  Every rank allocates the SAME
// sequence of buffers from its arena, so a local address translates to any peer's by an
// offset (symmetric heap).  A halo fill is: the producing kernel stores its own x-images;
// k_exchange_y pushes the 3 top / bottom interior rows (full width, corners included)
// into the neighbours' halo rows; its last block then bumps this rank's epoch counter,
// publishes it in the neighbours' control blocks and spins until both neighbours have
// published the same epoch.  All ranks run the same kernel sequence (SPMD; the
// data-dependent iteration count of solve() is derived from all-reduced norms), so a
// rank is never more than one exchange ahead of its neighbours and no buffer's halo is
// overwritten while its previous contents are still being read.
#include <vector>

#include "f2d_common.cuh"

using namespace f2d;

struct f2d_comm {
  int rank = 0, nranks = 1;
  size_t arena_bytes = 0, used = 0;
  char *base = nullptr;                    // my arena
  char *peer[MAXRANKS] = {nullptr};        // everybody's arena (peer[rank] == base)
  Ctrl *ctrl[MAXRANKS] = {nullptr};        // control blocks (at the arena start)
  Ctrl **d_ctrl = nullptr;                 // device copy of ctrl[]
  bool connected = false;
};

namespace {

__device__ __forceinline__ void publish_and_wait(Ctrl *me, Ctrl *const *peers, int rank, int nranks, int all) {
  __threadfence_system();
  unsigned long long D = me->done + 1;
  me->done = D;
  int north = (rank + 1) % nranks, south = (rank + nranks - 1) % nranks;
  // always published to every rank: slot[r] tracks rank r's progress everywhere
  for (int r = 0; r < nranks; r++)
    if (r != rank) *((volatile unsigned long long *)&peers[r]->slot[rank]) = D;
  if (all) {
    for (int r = 0; r < nranks; r++)
      if (r != rank)
        while (ld_acquire_sys(&me->slot[r]) < D) {}
  } else {
    while (ld_acquire_sys(&me->slot[north]) < D) {}
    while (ld_acquire_sys(&me->slot[south]) < D) {}
  }
}

struct XchArgs {
  double *self[4], *north[4], *south[4];
  int narr, ny, nx, nh;
};

// push my top interior rows into north's bottom halo and my bottom interior rows into
// south's top halo (all arrays of `a`), then lock-step with both neighbours
__global__ void k_exchange_y(XchArgs a, Ctrl *me, Ctrl *const *peers, int rank, int nranks) {
  // the fused kernels publish without waiting: before storing into a neighbour make sure
  // it has completed every synchronising kernel this rank has (no reader of the old halo)
  if (threadIdx.x == 0) {
    const unsigned long long d = *((volatile unsigned long long *)&me->done);
    const int rn = (rank + 1) % nranks, rs = (rank + nranks - 1) % nranks;
    while (ld_acquire_sys(&me->slot[rs]) < d) {}
    while (ld_acquire_sys(&me->slot[rn]) < d) {}
  }
  __syncthreads();
  const size_t rowlen = (size_t)a.nx;
  const size_t per = (size_t)a.nh * rowlen;          // elements of one 3-row strip
  const size_t total = 2 * per * a.narr;
  for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < total; p += (size_t)gridDim.x * blockDim.x) {
    int arr = (int)(p / (2 * per));
    size_t q = p % (2 * per);
    bool up = q < per;                                // to the north neighbour
    size_t e = up ? q : q - per;
    if (up)
      a.north[arr][e] = a.self[arr][(size_t)(a.ny - 2 * a.nh) * rowlen + e];        // -> rows 0..nh-1
    else
      a.south[arr][(size_t)(a.ny - a.nh) * rowlen + e] = a.self[arr][(size_t)a.nh * rowlen + e];  // -> rows ny-nh..
  }
  __syncthreads();   // the block's stores happen-before thread 0's system fence (cumulativity)
  if (threadIdx.x == 0) {
    __threadfence_system();
    unsigned int t = atomicAdd(&me->blocks_done, 1u);
    if (t == gridDim.x - 1) {
      me->blocks_done = 0;
      publish_and_wait(me, peers, rank, nranks, 0);
    }
  }
}

// wait (without publishing) until both neighbours have caught up: placed before kernels
// that read halo rows but do not take part in the protocol themselves
__global__ void k_drain(Ctrl *me, int rank, int nranks) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const unsigned long long d = *((volatile unsigned long long *)&me->done);
    const int rn = (rank + 1) % nranks, rs = (rank + nranks - 1) % nranks;
    while (ld_acquire_sys(&me->slot[rs]) < d) {}
    while (ld_acquire_sys(&me->slot[rn]) < d) {}
  }
}

__global__ void k_barrier(Ctrl *me, Ctrl *const *peers, int rank, int nranks, int all) {
  if (threadIdx.x == 0 && blockIdx.x == 0) publish_and_wait(me, peers, rank, nranks, all);
}

// in-place all-reduce of n (<= RED_SLOTS) doubles: every rank deposits its values in every
// rank's staging row, all-rank barrier, then folds the rows in rank order (deterministic)
__global__ void k_allreduce(double *vals, int n, unsigned maxmask, Ctrl *me, Ctrl *const *peers, int rank, int nranks) {
  __shared__ int buf;
  if (threadIdx.x == 0) {
    buf = (int)(me->red_epoch & 1ull);
    me->red_epoch++;
  }
  __syncthreads();
  int bsel = buf;
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    double v = vals[k];
    for (int r = 0; r < nranks; r++) peers[r]->red[bsel][rank][k] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) publish_and_wait(me, peers, rank, nranks, 1);
  __syncthreads();
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    double acc = ((volatile double *)me->red[bsel][0])[k];
    for (int r = 1; r < nranks; r++) {
      double o = ((volatile double *)me->red[bsel][r])[k];
      acc = ((maxmask >> k) & 1u) ? fmax(acc, o) : acc + o;
    }
    vals[k] = acc;
  }
}

// gather: copy my interior rows of a slab-shaped array into every rank's replicated
// full-height array (rows offset by my slab position), then all-rank barrier
__global__ void k_gather_push(const double *slab, int ny_loc, int nx, int nh, size_t full_off /*bytes from arena base*/,
                              int row0 /*first global interior row of my slab*/, char *const *arena, Ctrl *me,
                              Ctrl *const *peers, int rank, int nranks, int yimages) {
  // no rank may still be reading the previous contents of `full`: every rank must have
  // completed the synchronising kernels this rank has (see comm_gather)
  if (threadIdx.x == 0) {
    const unsigned long long d = *((volatile unsigned long long *)&me->done);
    for (int r = 0; r < nranks; r++)
      if (r != rank)
        while (ld_acquire_sys(&me->slot[r]) < d) {}
  }
  __syncthreads();
  const int nrows = ny_loc - 2 * nh;
  const int M = nranks * nrows;   // interior rows of the replicated array
  const size_t total = (size_t)nrows * nx;
  for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < total; p += (size_t)gridDim.x * blockDim.x) {
    int r = (int)(p / nx), c = (int)(p % nx);
    double v = slab[(size_t)(nh + r) * nx + c];
    const int gi = row0 + r;      // global interior row
    size_t dst = (size_t)(nh + gi) * nx + c;
    for (int k = 0; k < nranks; k++) reinterpret_cast<double *>(arena[k] + full_off)[dst] = v;
    // yimages: the periodic y halo rows of the replicated array (the x halo columns travel with
    // the rows), so that no halo-fill kernel has to follow the gather
    if (yimages) {
      if (gi < nh) {
        const size_t d2 = (size_t)(nh + M + gi) * nx + c;
        for (int k = 0; k < nranks; k++) reinterpret_cast<double *>(arena[k] + full_off)[d2] = v;
      }
      if (gi >= M - nh) {
        const size_t d2 = (size_t)(gi - M + nh) * nx + c;
        for (int k = 0; k < nranks; k++) reinterpret_cast<double *>(arena[k] + full_off)[d2] = v;
      }
    }
  }
  __syncthreads();   // the block's stores happen-before thread 0's system fence (cumulativity)
  if (threadIdx.x == 0) {
    __threadfence_system();
    unsigned int t = atomicAdd(&me->blocks_done, 1u);
    if (t == gridDim.x - 1) {
      me->blocks_done = 0;
      publish_and_wait(me, peers, rank, nranks, 1);
    }
  }
}
// int8 variant for masks; the masks' halo rows are not periodic images (hierarchy.py:46),
// so the first / last rank also contribute their outer halo rows
__global__ void k_gather_push_i8(const int8_t *slab, int ny_loc, int nx, int nh, size_t full_off, int row0,
                                 char *const *arena, Ctrl *me, Ctrl *const *peers, int rank, int nranks) {
  // no rank may still be reading the previous contents of `full`: every rank must have
  // completed the synchronising kernels this rank has (see comm_gather)
  if (threadIdx.x == 0) {
    const unsigned long long d = *((volatile unsigned long long *)&me->done);
    for (int r = 0; r < nranks; r++)
      if (r != rank)
        while (ld_acquire_sys(&me->slot[r]) < d) {}
  }
  __syncthreads();
  const int nrows = ny_loc - 2 * nh;
  const int rlo = rank == 0 ? -nh : 0, rhi = rank == nranks - 1 ? nrows + nh : nrows;
  const size_t total = (size_t)(rhi - rlo) * nx;
  for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < total; p += (size_t)gridDim.x * blockDim.x) {
    int r = rlo + (int)(p / nx), c = (int)(p % nx);
    int8_t v = slab[(size_t)(nh + r) * nx + c];
    size_t dst = (size_t)(nh + row0 + r) * nx + c;
    for (int k = 0; k < nranks; k++) reinterpret_cast<int8_t *>(arena[k] + full_off)[dst] = v;
  }
  __syncthreads();   // the block's stores happen-before thread 0's system fence (cumulativity)
  if (threadIdx.x == 0) {
    __threadfence_system();
    unsigned int t = atomicAdd(&me->blocks_done, 1u);
    if (t == gridDim.x - 1) {
      me->blocks_done = 0;
      publish_and_wait(me, peers, rank, nranks, 1);
    }
  }
}

}  // namespace

namespace f2d {
// internal API used by the multigrid
int comm_rank(const f2d_comm *c) { return c ? c->rank : 0; }
int comm_size(const f2d_comm *c) { return c ? c->nranks : 1; }
void *comm_alloc(f2d_comm *c, size_t nbytes) {
  size_t a = (c->used + 255) & ~(size_t)255;
  if (a + nbytes > c->arena_bytes) return nullptr;
  c->used = a + nbytes;
  return c->base + a;
}
template <typename T>
static T *peer_ptr(const f2d_comm *c, int r, T *p) {
  return reinterpret_cast<T *>(c->peer[r] + (reinterpret_cast<char *>(p) - c->base));
}
int comm_exchange(f2d_comm *c, double *const *arrs, int narr, int nh, int ny, int nx, cudaStream_t s) {
  if (!c || !c->connected) return fail(F2D_ERR_ARG, "exchange: communicator not connected");
  if (narr < 1 || narr > 4) return fail(F2D_ERR_ARG, "exchange: 1..4 arrays");
  XchArgs a;
  a.narr = narr; a.ny = ny; a.nx = nx; a.nh = nh;
  int north = (c->rank + 1) % c->nranks, south = (c->rank + c->nranks - 1) % c->nranks;
  for (int k = 0; k < narr; k++) {
    char *p = reinterpret_cast<char *>(arrs[k]);
    if (p < c->base || p >= c->base + c->arena_bytes) return fail(F2D_ERR_ARG, "exchange: array not in the symmetric heap");
    a.self[k] = arrs[k];
    a.north[k] = peer_ptr(c, north, arrs[k]);
    a.south[k] = peer_ptr(c, south, arrs[k]);
  }
  size_t total = 2 * (size_t)nh * nx * narr;
  int blocks = cdiv(total, 256);
  if (blocks > 64) blocks = 64;
  k_exchange_y<<<blocks, 256, 0, s>>>(a, c->ctrl[c->rank], c->d_ctrl, c->rank, c->nranks);
  F2D_LAUNCHED();
  return F2D_OK;
}
int comm_allreduce(f2d_comm *c, double *vals, int n, unsigned maxmask, cudaStream_t s) {
  if (!c || c->nranks == 1) return F2D_OK;
  if (n > RED_SLOTS) return fail(F2D_ERR_ARG, "allreduce: too many values");
  k_allreduce<<<1, 32, 0, s>>>(vals, n, maxmask, c->ctrl[c->rank], c->d_ctrl, c->rank, c->nranks);
  F2D_LAUNCHED();
  return F2D_OK;
}
int comm_barrier(f2d_comm *c, int all, cudaStream_t s) {
  if (!c || c->nranks == 1) return F2D_OK;
  k_barrier<<<1, 32, 0, s>>>(c->ctrl[c->rank], c->d_ctrl, c->rank, c->nranks, all);
  F2D_LAUNCHED();
  return F2D_OK;
}
// slab (interior rows) -> the same rows of every rank's replicated array `full` (symmetric address)
int comm_gather(f2d_comm *c, const double *slab, double *full, int ny_loc, int nx, int nh, cudaStream_t s,
                bool barrier_first, bool yimages) {
  // The replicated levels run without any synchronisation, so a fast rank could push the
  // next cycle's data into `full` while a slow rank still reads the previous contents.
  // The push kernel first waits until every rank has completed the synchronising kernels
  // this rank has; that is enough when at least one such kernel separates the readers of
  // `full` from this gather (slab levels above the gather level).  Otherwise
  // (barrier_first) all ranks first agree that everything enqueued so far has completed.
  if (barrier_first) {
    int rc = comm_barrier(c, 1, s);
    if (rc != F2D_OK) return rc;
  }
  char **d_arena = reinterpret_cast<char **>(c->d_ctrl + MAXRANKS);
  size_t off = reinterpret_cast<char *>(full) - c->base;
  int row0 = c->rank * (ny_loc - 2 * nh);
  size_t total = (size_t)(ny_loc - 2 * nh) * nx;
  int blocks = cdiv(total, 256);
  if (blocks > 128) blocks = 128;
  k_gather_push<<<blocks, 256, 0, s>>>(slab, ny_loc, nx, nh, off, row0, d_arena, c->ctrl[c->rank], c->d_ctrl, c->rank,
                                       c->nranks, yimages ? 1 : 0);
  F2D_LAUNCHED();
  return F2D_OK;
}
int comm_gather_i8(f2d_comm *c, const int8_t *slab, int8_t *full, int ny_loc, int nx, int nh, cudaStream_t s) {
  char **d_arena = reinterpret_cast<char **>(c->d_ctrl + MAXRANKS);
  size_t off = reinterpret_cast<char *>(full) - c->base;
  int row0 = c->rank * (ny_loc - 2 * nh);
  size_t total = (size_t)(ny_loc - 2 * nh) * nx;
  int blocks = cdiv(total, 256);
  if (blocks > 128) blocks = 128;
  k_gather_push_i8<<<blocks, 256, 0, s>>>(slab, ny_loc, nx, nh, off, row0, d_arena, c->ctrl[c->rank], c->d_ctrl,
                                          c->rank, c->nranks);
  F2D_LAUNCHED();
  return F2D_OK;
}
Peer comm_peer(const f2d_comm *c) {
  Peer P;
  P.me = nullptr; P.peers = nullptr; P.north_off = P.south_off = 0; P.rank = 0; P.nranks = 1;
  if (!c || c->nranks == 1) return P;
  int north = (c->rank + 1) % c->nranks, south = (c->rank + c->nranks - 1) % c->nranks;
  P.me = c->ctrl[c->rank];
  P.peers = c->d_ctrl;
  P.north_off = (long long)((intptr_t)c->peer[north] - (intptr_t)c->base);
  P.south_off = (long long)((intptr_t)c->peer[south] - (intptr_t)c->base);
  P.rank = c->rank;
  P.nranks = c->nranks;
  return P;
}
int comm_drain(f2d_comm *c, cudaStream_t s) {
  if (!c || c->nranks == 1) return F2D_OK;
  k_drain<<<1, 32, 0, s>>>(c->ctrl[c->rank], c->rank, c->nranks);
  F2D_LAUNCHED();
  return F2D_OK;
}
bool comm_owns(const f2d_comm *c, const void *p) {
  const char *q = reinterpret_cast<const char *>(p);
  return c && q >= c->base && q < c->base + c->arena_bytes;
}
}  // namespace f2d

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" int f2d_comm_create(f2d_comm_t **out, int rank, int nranks, size_t arena_bytes, void *ipc_handle_out) {
  if (!out || !ipc_handle_out || nranks < 1 || nranks > MAXRANKS || rank < 0 || rank >= nranks)
    return fail(F2D_ERR_ARG, "comm_create: bad arguments");
  if (nranks & (nranks - 1)) return fail(F2D_ERR_ARG, "comm_create: the number of ranks must be a power of two");
  f2d_comm *c = new f2d_comm();
  c->rank = rank;
  c->nranks = nranks;
  c->arena_bytes = arena_bytes;
  cudaError_t e = cudaMalloc(&c->base, arena_bytes);
  if (e != cudaSuccess) { delete c; return cuda_fail(e, "cudaMalloc(arena)"); }
  e = cudaMemset(c->base, 0, arena_bytes);
  if (e != cudaSuccess) { cudaFree(c->base); delete c; return cuda_fail(e, "cudaMemset(arena)"); }
  c->used = (sizeof(Ctrl) + 255) & ~(size_t)255;
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, c->base);
  if (e != cudaSuccess) { cudaFree(c->base); delete c; return cuda_fail(e, "cudaIpcGetMemHandle"); }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
  memcpy(ipc_handle_out, &h, 64);
  *out = c;
  return F2D_OK;
}

extern "C" int f2d_comm_connect(f2d_comm_t *c, const void *all_handles) {
  if (!c || !all_handles) return fail(F2D_ERR_ARG, "comm_connect: null");
  const char *hs = reinterpret_cast<const char *>(all_handles);
  for (int r = 0; r < c->nranks; r++) {
    if (r == c->rank) {
      c->peer[r] = c->base;
    } else {
      cudaIpcMemHandle_t h;
      memcpy(&h, hs + 64 * r, 64);
      void *p = nullptr;
      F2D_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
      c->peer[r] = reinterpret_cast<char *>(p);
    }
    c->ctrl[r] = reinterpret_cast<Ctrl *>(c->peer[r]);
  }
  // device table: MAXRANKS control-block pointers followed by MAXRANKS arena bases
  void *tab[2 * MAXRANKS] = {nullptr};
  for (int r = 0; r < c->nranks; r++) {
    tab[r] = c->ctrl[r];
    tab[MAXRANKS + r] = c->peer[r];
  }
  F2D_CUDA(cudaMalloc(&c->d_ctrl, sizeof tab));
  F2D_CUDA(cudaMemcpy(c->d_ctrl, tab, sizeof tab, cudaMemcpyHostToDevice));
  c->connected = true;
  return F2D_OK;
}

extern "C" void *f2d_comm_alloc(f2d_comm_t *c, size_t nbytes) { return c ? comm_alloc(c, nbytes) : nullptr; }
extern "C" int f2d_comm_rank(const f2d_comm_t *c) { return comm_rank(c); }
extern "C" int f2d_comm_size(const f2d_comm_t *c) { return comm_size(c); }

extern "C" int f2d_comm_destroy(f2d_comm_t *c) {
  if (!c) return F2D_OK;
  for (int r = 0; r < c->nranks; r++)
    if (r != c->rank && c->peer[r]) cudaIpcCloseMemHandle(c->peer[r]);
  cudaFree(c->d_ctrl);
  cudaFree(c->base);
  delete c;
  return F2D_OK;
}

extern "C" int f2d_comm_barrier(f2d_comm_t *c, f2d_stream_t s) { return comm_barrier(c, 1, S(s)); }

/* number of lock-step synchronisations this rank has completed (synchronises the device);
 * equal on every rank at matching program points -- a debugging / test aid */
extern "C" long long f2d_comm_epoch(f2d_comm_t *c) {
  if (!c) return 0;
  unsigned long long d = 0;
  cudaDeviceSynchronize();
  cudaMemcpy(&d, &c->ctrl[c->rank]->done, sizeof d, cudaMemcpyDeviceToHost);
  return (long long)d;
}

extern "C" int f2d_comm_exchange_y(f2d_comm_t *c, double *x, int nh, int ny, int nx, f2d_stream_t s) {
  if (!c || c->nranks == 1) return f2d_fill_halo(x, nh, ny, nx, s);
  double *arrs[1] = {x};
  return comm_exchange(c, arrs, 1, nh, ny, nx, S(s));
}

extern "C" int f2d_comm_allreduce(f2d_comm_t *c, double *vals, int n, unsigned int maxmask, f2d_stream_t s) {
  return comm_allreduce(c, vals, n, maxmask, S(s));
}
