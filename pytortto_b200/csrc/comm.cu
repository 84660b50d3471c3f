// Small-message all-reduce over NVLink peer memory for the SyncBN statistics (one process per GPU).
//
// A data-parallel step needs 2 x (#BatchNorm layers) all-reduces of 2C doubles (C <= 4096): far too small for
// bandwidth to matter and, through a general collective library, ~70 latency-bound calls per step on the compute
// stream.  Here every rank owns a communication buffer (cudaMalloc + CUDA IPC, mapped into every peer) made of
// fixed "slots" {flag, arrive counter, data[]}; a call site (one BatchNorm layer, forward or backward) always uses
// the same slot, so nothing but device memory changes between steps (CUDA-graph safe).
//   publish : sum this rank's per-chunk partials (fixed order) into its own slot, then flag = flag + 1
//             (last-arriving block, release at system scope)
//   gather  : wait until every peer's flag for that slot has reached this rank's own flag, then add the peers' data
//             IN RANK ORDER (bit-identical result on every rank) straight over NVLink (peer loads bypass L1) into a
//             local [n] buffer that the ordinary bn finalize kernels consume.
// Why reuse of a slot is safe: a rank reaches the same call site again only after every peer has passed all the
// call sites in between, each of which needed this rank's later publishes, which are stream-ordered after its gather.
#include <string.h>

#include "common.cuh"

namespace ttb {

struct SlotHeader {
  unsigned long long flag;    // number of publishes completed on this slot (monotonic)
  unsigned int arrive;        // block arrival counter of the running publish
  unsigned int pad;
};
constexpr size_t kSlotHeaderBytes = 16;

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_volatile_f64(const double* p) {
  double v;
  asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// grid = ceil(n / 32) blocks of (32 values x 32 chunk lanes)
__global__ void __launch_bounds__(1024)
comm_publish_kernel(const double* __restrict__ partials, int num_chunks, int n, char* slot) {
  __shared__ double sm[32][33];
  SlotHeader* hdr = reinterpret_cast<SlotHeader*>(slot);
  double* data = reinterpret_cast<double*>(slot + kSlotHeaderBytes);
  const int i = blockIdx.x * 32 + threadIdx.x;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;  // four loads in flight per thread (the loop is latency-bound)
  if (i < n) {
    int k = threadIdx.y;
    for (; k + 96 < num_chunks; k += 128) {
      s0 += partials[(int64_t)k * n + i];
      s1 += partials[(int64_t)(k + 32) * n + i];
      s2 += partials[(int64_t)(k + 64) * n + i];
      s3 += partials[(int64_t)(k + 96) * n + i];
    }
    for (; k < num_chunks; k += 32) s0 += partials[(int64_t)k * n + i];
  }
  sm[threadIdx.y][threadIdx.x] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (threadIdx.y == 0 && i < n) {
    double t = 0.0;
    for (int j = 0; j < 32; ++j) t += sm[j][threadIdx.x];
    data[i] = t;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    unsigned int ticket = atomicAdd(&hdr->arrive, 1u);
    if (ticket == gridDim.x - 1) {  // every block's data is written and fenced: publish
      hdr->arrive = 0;
      __threadfence_system();
      st_release_sys(&hdr->flag, hdr->flag + 1ull);
    }
  }
}

// peers[r] = base of rank r's communication buffer as mapped in THIS process (peers[rank] = own buffer)
__global__ void __launch_bounds__(256)
comm_gather_kernel(char* const* __restrict__ peers, int world, int rank, size_t slot_offset, int n,
                   double* __restrict__ out, unsigned long long spin_limit) {
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  // one thread per peer polls that peer's flag (all NVLink round trips in flight at once)
  for (int r = threadIdx.x; r < world; r += blockDim.x) {
    if (r == rank) continue;
    const unsigned long long want = ld_acquire_sys(&reinterpret_cast<const SlotHeader*>(peers[rank] + slot_offset)->flag);
    const unsigned long long* f = &reinterpret_cast<const SlotHeader*>(peers[r] + slot_offset)->flag;
    unsigned long long spins = 0;
    while (ld_acquire_sys(f) < want) {
      if (++spins > spin_limit) {  // a peer never arrived (ranks diverged): fail loudly instead of hanging the GPU
        atomicExch(&bad, 1);
        break;
      }
      __nanosleep(32);
    }
  }
  __syncthreads();
  if (bad) __trap();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double t = 0.0;
  for (int r = 0; r < world; ++r)  // fixed rank order: the same bits on every rank
    t += ld_volatile_f64(reinterpret_cast<const double*>(peers[r] + slot_offset + kSlotHeaderBytes) + i);
  out[i] = t;
}

}  // namespace ttb

using namespace ttb;

extern "C" {

int ttb_comm_alloc(size_t bytes, void** dev_ptr, unsigned char* handle_out /*64 bytes*/) {
  TTB_REQUIRE(dev_ptr && handle_out && bytes > 0, "comm_alloc: bad arguments");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {
    set_error("comm_alloc: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return 1;
  }
  e = cudaMemset(p, 0, bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    set_error("comm_alloc: %s", cudaGetErrorString(e));
    cudaFree(p);
    return 1;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  memcpy(handle_out, &h, 64);
  *dev_ptr = p;
  return 0;
}

int ttb_comm_open(const unsigned char* handle /*64 bytes*/, void** peer_ptr) {
  TTB_REQUIRE(handle && peer_ptr, "comm_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  cudaError_t e = cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    set_error("comm_open: cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

int ttb_comm_close(void* peer_ptr) {
  if (peer_ptr && cudaIpcCloseMemHandle(peer_ptr) != cudaSuccess) {
    set_error("comm_close failed");
    return 1;
  }
  return 0;
}

int ttb_comm_free(void* dev_ptr) {
  if (dev_ptr && cudaFree(dev_ptr) != cudaSuccess) {
    set_error("comm_free failed");
    return 1;
  }
  return 0;
}

size_t ttb_comm_slot_bytes(int max_values) { return kSlotHeaderBytes + (size_t)max_values * sizeof(double); }

int ttb_comm_publish(const double* partials, int num_chunks, int n, void* my_buf, size_t slot_offset, void* stream) {
  TTB_REQUIRE(partials && my_buf && n > 0 && num_chunks > 0, "comm_publish: bad arguments");
  comm_publish_kernel<<<(n + 31) / 32, dim3(32, 32), 0, as_stream(stream)>>>(partials, num_chunks, n,
                                                                             reinterpret_cast<char*>(my_buf) + slot_offset);
  return check_launch("comm_publish");
}

int ttb_comm_gather(void* const* peers_dev, int world, int rank, size_t slot_offset, int n, double* out, void* stream) {
  TTB_REQUIRE(peers_dev && out && n > 0 && world > 0 && rank >= 0 && rank < world, "comm_gather: bad arguments");
  // >= 100 ns per probe: give a missing peer about a minute before trapping
  comm_gather_kernel<<<(n + 255) / 256, 256, 0, as_stream(stream)>>>(reinterpret_cast<char* const*>(peers_dev), world, rank,
                                                                     slot_offset, n, out, 600000000ull);
  return check_launch("comm_gather");
}

}  // extern "C"
