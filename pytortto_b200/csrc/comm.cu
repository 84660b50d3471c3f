// Small-message all-reduce over NVLink peer memory for the SyncBN statistics (one process per GPU).
//
// A data-parallel step needs 2 x (#BatchNorm layers) all-reduces of 2C doubles (C <= 4096): far too small for
// bandwidth to matter and, through a general collective library, ~70 latency-bound calls per step on the compute
// stream.  Here every rank owns a communication buffer (cudaMalloc + CUDA IPC, mapped into every peer) made of
// fixed "slots" {flag, arrive counter, data[]}; a call site (one BatchNorm layer, forward or backward) always uses
// the same slot, so nothing but device memory changes between steps (CUDA-graph safe).
// One kernel per exchange (comm_allreduce_kernel below): sum this rank's per-chunk partials, PUSH the result into the
// slot of every peer (remote stores over NVLink into the area reserved for this rank) as self-validating 8-byte words
// {sequence | 32 data bits}, poll the areas of the own slot that the peers fill - LOCAL memory, so a value is seen one
// one-way NVLink latency after it was sent instead of after 1-2 round trips of remote polling - and add the values IN RANK
// ORDER (bit-identical result on every rank) into a local [n] buffer that the ordinary bn finalize kernels consume.
// Slot layout: header | parity 0: [world][kCommMaxValues] word pairs | parity 1: the same.
// Why reuse of a slot is safe: the data area is DOUBLE-BUFFERED by the parity of the exchange's sequence number.  Rank A
// can start exchange s+1 (other buffer) while a slow peer B is still reading A's words of exchange s, but A cannot finish
// s+1 - and so cannot reach s+2, which overwrites the buffer of s - before B has published its s+1 value, which B does
// only after its own exchange s (all reads of A's buffer s) has completed.  This holds even when one call site is the only
// synchronisation point of the program (one BatchNorm layer, forward only).
#include <string.h>

#include "bn_finalize.cuh"
#include "common.cuh"

namespace ttb {

struct SlotHeader {
  unsigned int seq;           // number of exchanges completed on this slot (owner-written)
  unsigned int arrive;        // block arrival counter of the running exchange
  unsigned int pad[2];
};
constexpr size_t kSlotHeaderBytes = 16;
constexpr int kCommMaxValues = 4096;                          // values per exchange (2C doubles, C <= 2048)
constexpr size_t kRankStride = (size_t)kCommMaxValues * 16;   // bytes of the area one source rank writes
// area of a slot that source rank `src` fills in the exchange with sequence number `seq` (two buffers, by parity)
__device__ __forceinline__ size_t slot_area_offset(unsigned int seq, int world, int src) {
  return kSlotHeaderBytes + ((size_t)(seq & 1u) * world + src) * kRankStride;
}

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_volatile_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

constexpr int kMaxWorld = 8;  // one NVSwitch box (CUDA IPC does not leave the node)

// ONE kernel per exchange, one NVLink round trip ("LL" style: every 8-byte word carries its own sequence number, so
// the data is its own flag and no fence / separate flag write is needed; aligned 8-byte accesses are atomic).
//   value i of this rank  = sum over chunks of partials[chunk][i]            (fixed order)
//   own slot word pair i  = {seq | low 32 bits}, {seq | high 32 bits}
//   out[i]                = sum over ranks r = 0..world-1 of rank r's value i (rank order => same bits everywhere)
// grid = ceil(n / 32) blocks of (32 values x 32 chunk lanes); peers[r] = rank r's buffer mapped in this process.
__global__ void __launch_bounds__(1024)
comm_allreduce_kernel(const double* __restrict__ partials, int num_chunks, int n, char* const* __restrict__ peers,
                      int world, int rank, size_t slot_offset, double* __restrict__ out, unsigned long long spin_limit) {
  pdl_entry();
  __shared__ double sm[32][33];
  char* own = peers[rank] + slot_offset;
  SlotHeader* hdr = reinterpret_cast<SlotHeader*>(own);
  const unsigned int seq = ld_volatile_u32(&hdr->seq) + 1u;  // every block reads it before any block can finish last
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
  __shared__ double mine_sm[32];
  if (threadIdx.y == 0 && i < n) {
    double mine = 0.0;
    for (int j = 0; j < 32; ++j) mine += sm[j][threadIdx.x];
    mine_sm[threadIdx.x] = mine;
  }
  __syncthreads();  // sm[][] is reused below: row r = the value of rank r
  // thread (x = value, y = peer rank) serves ONE peer: it pushes this rank's value into that peer's slot and then polls
  // the area of the own slot that the peer fills, so the transfers of all peers overlap; rank order is restored by the
  // sum below
  bool failed = false;
  const int r = threadIdx.y;
  if (r < world && i < n) {
    double v;
    if (r == rank) {
      v = mine_sm[threadIdx.x];
    } else {
      const unsigned long long bits = (unsigned long long)__double_as_longlong(mine_sm[threadIdx.x]);
      const unsigned long long tag = (unsigned long long)seq << 32;
      unsigned long long* w =
          reinterpret_cast<unsigned long long*>(peers[r] + slot_offset + slot_area_offset(seq, world, rank)) + 2 * (size_t)i;
      st_volatile_u64(w, tag | (bits & 0xffffffffull));
      st_volatile_u64(w + 1, tag | (bits >> 32));
      const unsigned long long* pw =
          reinterpret_cast<const unsigned long long*>(own + slot_area_offset(seq, world, r)) + 2 * (size_t)i;
      unsigned long long a, b, spins = 0;
      for (;;) {
        a = ld_volatile_u64(pw);
        b = ld_volatile_u64(pw + 1);
        if ((unsigned int)(a >> 32) == seq && (unsigned int)(b >> 32) == seq) break;
        if (++spins > spin_limit) {  // a peer never arrived (ranks diverged): fail loudly, do not hang
          failed = true;
          break;
        }
      }
      v = __longlong_as_double((long long)((a & 0xffffffffull) | (b << 32)));
    }
    sm[r][threadIdx.x] = v;
  }
  if (failed) __trap();
  __syncthreads();
  if (threadIdx.y == 0 && i < n) {
    double acc = 0.0;
    for (int q = 0; q < world; ++q) acc += sm[q][threadIdx.x];  // rank order => the same bits on every rank
    out[i] = acc;
  }
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    __threadfence();
    const unsigned int ticket = atomicAdd(&hdr->arrive, 1u);
    if (ticket == gridDim.x - 1) {  // last block of this exchange: the slot's sequence number advances
      hdr->arrive = 0;
      __threadfence();
      hdr->seq = seq;
    }
  }
}

// The same exchange fused with the BatchNorm finalisation that follows it (compute + collective + compute in ONE
// kernel): block = 32 channels x 32 lanes; channel i owns values i (first sum) and c + i (second sum) of the [2][c]
// statistics vector.  Chunk partials are summed, both values are published in this rank's slot and polled from every
// peer (thread (x, y) polls peer y), summed in rank order, and `fin(i, total0, total1)` writes the per-channel
// coefficients - what would otherwise be bn_finalize_kernel / bn_bwd_finalize_kernel as a second launch.
template <class Finalize>
__global__ void __launch_bounds__(1024)
comm_bn_finalize_kernel(const double* __restrict__ partials, int num_chunks, int c, char* const* __restrict__ peers, int world,
                        int rank, size_t slot_offset, unsigned long long spin_limit, Finalize fin) {
  pdl_entry();
  __shared__ double sm0[32][33], sm1[32][33];
  __shared__ double mine_sm[2][32];
  char* own = peers[rank] + slot_offset;
  SlotHeader* hdr = reinterpret_cast<SlotHeader*>(own);
  const unsigned int seq = ld_volatile_u32(&hdr->seq) + 1u;
  const int i = blockIdx.x * 32 + threadIdx.x;
  const int n = 2 * c;
  double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;  // two loads per value in flight per thread
  if (i < c) {
    int k = threadIdx.y;
    for (; k + 32 < num_chunks; k += 64) {
      a0 += partials[(int64_t)k * n + i];
      b0 += partials[(int64_t)k * n + c + i];
      a1 += partials[(int64_t)(k + 32) * n + i];
      b1 += partials[(int64_t)(k + 32) * n + c + i];
    }
    for (; k < num_chunks; k += 32) {
      a0 += partials[(int64_t)k * n + i];
      b0 += partials[(int64_t)k * n + c + i];
    }
  }
  sm0[threadIdx.y][threadIdx.x] = a0 + a1;
  sm1[threadIdx.y][threadIdx.x] = b0 + b1;
  __syncthreads();
  if (threadIdx.y < 2 && i < c) {  // y = 0 sums value i, y = 1 value c + i
    double (*sm)[33] = threadIdx.y == 0 ? sm0 : sm1;
    double mine = 0.0;
    for (int j = 0; j < 32; ++j) mine += sm[j][threadIdx.x];
    mine_sm[threadIdx.y][threadIdx.x] = mine;
  }
  __syncthreads();  // sm0 / sm1 are reused: row r = the two values of rank r
  bool failed = false;
  const int r = threadIdx.y;
  if (r < world && i < c) {
    double v0, v1;
    if (r == rank) {
      v0 = mine_sm[0][threadIdx.x];
      v1 = mine_sm[1][threadIdx.x];
    } else {
      // push both values of channel i into peer r's slot (the area reserved for this rank), then poll what peer r pushes here
      const unsigned long long tag = (unsigned long long)seq << 32;
      const unsigned long long b0 = (unsigned long long)__double_as_longlong(mine_sm[0][threadIdx.x]);
      const unsigned long long b1 = (unsigned long long)__double_as_longlong(mine_sm[1][threadIdx.x]);
      unsigned long long* wbase =
          reinterpret_cast<unsigned long long*>(peers[r] + slot_offset + slot_area_offset(seq, world, rank));
      st_volatile_u64(wbase + 2 * (size_t)i, tag | (b0 & 0xffffffffull));
      st_volatile_u64(wbase + 2 * (size_t)i + 1, tag | (b0 >> 32));
      st_volatile_u64(wbase + 2 * ((size_t)c + i), tag | (b1 & 0xffffffffull));
      st_volatile_u64(wbase + 2 * ((size_t)c + i) + 1, tag | (b1 >> 32));
      const unsigned long long* base = reinterpret_cast<const unsigned long long*>(own + slot_area_offset(seq, world, r));
      const unsigned long long* p0 = base + 2 * (size_t)i;
      const unsigned long long* p1 = base + 2 * ((size_t)c + i);
      unsigned long long w0, w1, w2, w3, spins = 0;
      for (;;) {
        w0 = ld_volatile_u64(p0);
        w1 = ld_volatile_u64(p0 + 1);
        w2 = ld_volatile_u64(p1);
        w3 = ld_volatile_u64(p1 + 1);
        if ((unsigned int)(w0 >> 32) == seq && (unsigned int)(w1 >> 32) == seq && (unsigned int)(w2 >> 32) == seq &&
            (unsigned int)(w3 >> 32) == seq)
          break;
        if (++spins > spin_limit) {  // a peer never arrived (ranks diverged): fail loudly, do not hang
          failed = true;
          break;
        }
      }
      v0 = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
      v1 = __longlong_as_double((long long)((w2 & 0xffffffffull) | (w3 << 32)));
    }
    sm0[r][threadIdx.x] = v0;
    sm1[r][threadIdx.x] = v1;
  }
  if (failed) __trap();
  __syncthreads();
  if (threadIdx.y == 0 && i < c) {
    double t0 = 0.0, t1 = 0.0;
    for (int q = 0; q < world; ++q) {  // rank order => the same bits on every rank
      t0 += sm0[q][threadIdx.x];
      t1 += sm1[q][threadIdx.x];
    }
    fin(i, t0, t1);
  }
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    __threadfence();
    const unsigned int ticket = atomicAdd(&hdr->arrive, 1u);
    if (ticket == gridDim.x - 1) {
      hdr->arrive = 0;
      __threadfence();
      hdr->seq = seq;
    }
  }
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

size_t ttb_comm_slot_bytes(int max_values) {  // header + 2 parities x kMaxWorld source areas; max_values <= 4096
  if (max_values < 0 || max_values > kCommMaxValues) return 0;
  return kSlotHeaderBytes + 2 * (size_t)kMaxWorld * kRankStride;
}

int ttb_comm_allreduce(const double* partials, int num_chunks, int n, void* const* peers_dev, int world, int rank,
                       size_t slot_offset, double* out, void* stream) {
  TTB_REQUIRE(partials && peers_dev && out && n > 0 && n <= kCommMaxValues && num_chunks > 0, "comm_allreduce: bad arguments");
  TTB_REQUIRE(world > 0 && world <= kMaxWorld && rank >= 0 && rank < world, "comm_allreduce: world size %d not in 1..%d", world,
              kMaxWorld);
  // about a minute of polling before a missing peer is declared lost
  launch_k(comm_allreduce_kernel, (n + 31) / 32, dim3(32, 32), 0, as_stream(stream), 
      partials, num_chunks, n, reinterpret_cast<char* const*>(peers_dev), world, rank, slot_offset, out, 200000000ull);
  return check_launch("comm_allreduce");
}

static int comm_check(const void* partials, const void* peers_dev, int c, int num_chunks, int world, int rank, const char* what) {
  TTB_REQUIRE(partials && peers_dev && c > 0 && 2 * c <= kCommMaxValues && num_chunks > 0, "%s: bad arguments", what);
  TTB_REQUIRE(world > 0 && world <= kMaxWorld && rank >= 0 && rank < world, "%s: world size %d not in 1..%d", what, world, kMaxWorld);
  return 0;
}

/* ttb_comm_allreduce of the [2][C] forward statistics + ttb_bn_finalize in one kernel; `count` = GLOBAL element count */
int ttb_comm_bn_finalize(const double* partials, int num_chunks, void* const* peers_dev, int world, int rank, size_t slot_offset,
                         int64_t count, int c, float eps, float momentum, const float* gamma, const float* beta,
                         float* running_mean, float* running_var, float* mean, float* var_eps, float* sd, float* scale,
                         float* shift, void* stream) {
  if (comm_check(partials, peers_dev, c, num_chunks, world, rank, "comm_bn_finalize")) return 1;
  TTB_REQUIRE(count > 0, "comm_bn_finalize: count must be positive");
  BnFwdFinalize fin;
  fin.count = (double)count;
  fin.eps = eps;
  fin.momentum = momentum;
  bn_fwd_host_factors(count, momentum, &fin.unbias, &fin.one_minus_momentum);
  fin.gamma = gamma; fin.beta = beta; fin.running_mean = running_mean; fin.running_var = running_var;
  fin.mean = mean; fin.var_eps = var_eps; fin.sd = sd; fin.scale = scale; fin.shift = shift;
  launch_k(comm_bn_finalize_kernel<BnFwdFinalize>, (c + 31) / 32, dim3(32, 32), 0, as_stream(stream), 
      partials, num_chunks, c, reinterpret_cast<char* const*>(peers_dev), world, rank, slot_offset, 200000000ull, fin);
  return check_launch("comm_bn_finalize");
}

/* ttb_comm_allreduce of the [2][C] backward sums + ttb_bn_bwd_finalize in one kernel; `count` = GLOBAL element count */
int ttb_comm_bn_bwd_finalize(const double* partials, int num_chunks, void* const* peers_dev, int world, int rank,
                             size_t slot_offset, int64_t count, int c, const float* gamma, const float* var_eps,
                             const float* sd, float* dgamma, float* dbeta, float* coef, void* stream) {
  if (comm_check(partials, peers_dev, c, num_chunks, world, rank, "comm_bn_bwd_finalize")) return 1;
  BnBwdFinalize fin;
  fin.count = (double)count;
  fin.c = c;
  fin.gamma = gamma; fin.var_eps = var_eps; fin.sd = sd; fin.dgamma = dgamma; fin.dbeta = dbeta; fin.coef = coef;
  launch_k(comm_bn_finalize_kernel<BnBwdFinalize>, (c + 31) / 32, dim3(32, 32), 0, as_stream(stream), 
      partials, num_chunks, c, reinterpret_cast<char* const*>(peers_dev), world, rank, slot_offset, 200000000ull, fin);
  return check_launch("comm_bn_bwd_finalize");
}

}  // extern "C"
