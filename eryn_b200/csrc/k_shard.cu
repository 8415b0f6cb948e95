// Part of eryn_b200 (kernel overview in common.cuh). Built with --fmad=false.
//
// Multi-GPU plumbing of the temperature-sharded run: IPC-shareable device memory and the
// "publish" kernel = all-gather of logl written as NVLink peer stores + per-rank flag words.
#include "common.cuh"

namespace eb {

struct PublishArgs {
  int rank, world, W;
  int t_lo, nrows;                         // this rank's rows [t_lo, t_lo + nrows) of the full [T][W] matrix
  const double* logl_local;
  double* logl_all_peer[EB_MAX_RANKS];
  unsigned long long* flags_peer[EB_MAX_RANKS];
  eb_ctrl* ctrl;
};

__host__ __device__ inline int publish_grid(size_t ndoubles) {
  size_t g = (ndoubles / 2 + 255) / 256;
  return g < 1 ? 1 : g > 148 ? 148 : (int)g;
}

// K5: every CTA copies a contiguous slice of the local logl rows into the logl_all buffer of every rank (16-byte peer
// stores, coalesced); after a block barrier its thread 0 issues one system-scope release fence and ADDS 1 to this
// rank's flag word on every rank (NVLink atomics).  A flag word therefore counts publishing CTAs: a reader of iteration
// `it` waits for (it + 1) * publish_grid(rows of that rank).  No ticket, no last-block election, no host.
__global__ void __launch_bounds__(256) publish_logl_kernel(const __grid_constant__ PublishArgs p) {
  const size_t n = (size_t)p.nrows * p.W;                   // doubles to publish
  const size_t off = (size_t)p.t_lo * p.W;
  const size_t n2 = n >> 1;
  const bool vec = ((off & 1) == 0);                        // 16-byte alignment of the destination slice
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec) {
    const double2* src = reinterpret_cast<const double2*>(p.logl_local);
    for (size_t i = i0; i < n2; i += stride) {
      const double2 v = src[i];
      for (int g = 0; g < p.world; ++g) reinterpret_cast<double2*>(p.logl_all_peer[g] + off)[i] = v;
    }
    if ((n & 1) && i0 == 0)
      for (int g = 0; g < p.world; ++g) p.logl_all_peer[g][off + n - 1] = p.logl_local[n - 1];
  } else {
    for (size_t i = i0; i < n; i += stride) {
      const double v = p.logl_local[i];
      for (int g = 0; g < p.world; ++g) p.logl_all_peer[g][off + i] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x < p.world) {
    asm volatile("fence.acq_rel.sys;" ::: "memory");   // cumulative over the block's stores (ordered by the barrier)
    atomicAdd_system(p.flags_peer[threadIdx.x] + p.rank, 1ull);
  }
}

}  // namespace eb

using namespace eb;

extern "C" {

int eb_publish_logl(const eb_publish* pub, eb_ctrl* ctrl, void* stream) {
  if (!pub || !ctrl) return fail(EB_ERR_INVALID, "publish description/ctrl is NULL");
  if (pub->world < 1 || pub->world > EB_MAX_RANKS || pub->rank < 0 || pub->rank >= pub->world)
    return fail(EB_ERR_INVALID, "bad rank/world %d/%d", pub->rank, pub->world);
  if (pub->nwalkers < 1 || pub->ntemps_total < 1) return fail(EB_ERR_INVALID, "bad shape");
  if (pub->temp_begin[0] != 0 || pub->temp_begin[pub->world] != pub->ntemps_total)
    return fail(EB_ERR_INVALID, "temp_begin must run from 0 to ntemps_total");
  PublishArgs a;
  memset(&a, 0, sizeof(a));
  a.rank = pub->rank; a.world = pub->world; a.W = pub->nwalkers;
  a.t_lo = pub->temp_begin[pub->rank];
  a.nrows = pub->temp_begin[pub->rank + 1] - a.t_lo;
  if (a.nrows < 1) return fail(EB_ERR_INVALID, "rank %d owns no temperature", pub->rank);
  if (!pub->logl_local) return fail(EB_ERR_INVALID, "logl_local is NULL");
  for (int g = 0; g < pub->world; ++g) {
    if (!pub->logl_all_peer[g] || !pub->flags_peer[g]) return fail(EB_ERR_INVALID, "peer pointers of rank %d are NULL", g);
    a.logl_all_peer[g] = pub->logl_all_peer[g];
    a.flags_peer[g] = (unsigned long long*)pub->flags_peer[g];
  }
  a.logl_local = pub->logl_local;
  a.ctrl = ctrl;
  if (a.world > 256) return fail(EB_ERR_INVALID, "world too large");
  const int grid = publish_grid((size_t)a.nrows * a.W);
  publish_logl_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  return check_launch("publish_logl");
}

int eb_dev_malloc(size_t bytes, void** out) {
  if (!out || bytes == 0) return fail(EB_ERR_INVALID, "eb_dev_malloc: bad argument");
  EB_CUDA(cudaMalloc(out, bytes));
  EB_CUDA(cudaMemset(*out, 0, bytes));
  return EB_OK;
}

int eb_dev_free(void* p) {
  if (p) EB_CUDA(cudaFree(p));
  return EB_OK;
}

int eb_ipc_export(const void* dev_ptr, uint8_t* handle64) {
  if (!dev_ptr || !handle64) return fail(EB_ERR_INVALID, "eb_ipc_export: NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == EB_IPC_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  EB_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr)));
  memcpy(handle64, &h, sizeof(h));
  return EB_OK;
}

int eb_ipc_open(const uint8_t* handle64, void** out) {
  if (!handle64 || !out) return fail(EB_ERR_INVALID, "eb_ipc_open: NULL argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  EB_CUDA(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
  return EB_OK;
}

int eb_ipc_close(void* p) {
  if (p) EB_CUDA(cudaIpcCloseMemHandle(p));
  return EB_OK;
}

}  // extern "C"
