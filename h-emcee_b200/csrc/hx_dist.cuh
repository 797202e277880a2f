// hx_dist.cuh — the exchange step of the row-sharded ensemble over NVLink peer memory.
//
// Every rank keeps a full copy of both halves [2n, dim] (+ log-probs [2n]) in a cudaMalloc'ed region that the other
// ranks map through CUDA IPC.  A half-move updates the rank's own rows of one half in place; k_dist_push then stores
// those rows straight into every peer's copy (NVLink writes, no staging, no NCCL) and, once all of its stores are
// fenced, bumps this rank's slot of every peer's flag array to the half-move's epoch.  The next half-move starts with
// k_dist_wait, which spins until all slots of the LOCAL flag array reached the epoch, i.e. until every peer's rows of
// the complement have landed.  One barrier per half-move is enough: a rank can only be one half-move ahead of a peer,
// and consecutive half-moves write different halves.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace hx {

constexpr int kMaxRanks = 8;

struct PeerTable {
  void* base[kMaxRanks];     // base[r] = rank r's region mapped into this process (base[rank] = own region)
};

// what a finishing kernel needs to push its rows itself (fused exchange): byte offsets inside the regions
struct DistPush {
  PeerTable pt;
  int rank, world;
  size_t x_off, lp_off, flag_off;   // this rank's rows of the updated half, their log-probs, the flag array
  uint32_t* counter;                // zero-initialised CTA counter in the own region
  uint32_t epoch;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Copy `count16` 16-byte words starting at byte offset x_off (the rank's rows of the updated half) and `lp_count16`
// words at lp_off (their log-probs) from the own region into every peer's region at the same offsets; the last CTA to
// finish publishes the epoch.  `counter` is a zero-initialised uint32 in the own region.
static __global__ void __launch_bounds__(256) k_dist_push(PeerTable pt, int rank, int world, size_t x_off, size_t count16, size_t lp_off,
                                                   size_t lp_count16, size_t flag_off, uint32_t* counter, uint32_t epoch) {
  const uint4* src = reinterpret_cast<const uint4*>(static_cast<const char*>(pt.base[rank]) + x_off);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count16; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = src[i];
#pragma unroll 1
    for (int p = 0; p < world; ++p)
      if (p != rank) reinterpret_cast<uint4*>(static_cast<char*>(pt.base[p]) + x_off)[i] = v;
  }
  const uint4* lsrc = reinterpret_cast<const uint4*>(static_cast<const char*>(pt.base[rank]) + lp_off);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < lp_count16; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = lsrc[i];
#pragma unroll 1
    for (int p = 0; p < world; ++p)
      if (p != rank) reinterpret_cast<uint4*>(static_cast<char*>(pt.base[p]) + lp_off)[i] = v;
  }
  // all stores of this CTA are performed system-wide before it is counted as done
  __threadfence_system();
  __syncthreads();
  __shared__ unsigned last;
  if (threadIdx.x == 0) last = (atomicAdd(counter, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (last) {
    __threadfence_system();
    if (threadIdx.x < world) {
      uint32_t* flags = reinterpret_cast<uint32_t*>(static_cast<char*>(pt.base[threadIdx.x]) + flag_off);
      st_release_sys(flags + rank, epoch);
    }
    if (threadIdx.x == 0) *counter = 0u;
  }
}

// tail of a kernel that stored rows into the peers: fence, count the CTA, and let the last CTA publish the epoch.
// Must be reached by ALL threads of ALL CTAs of the grid.
__device__ __forceinline__ void dist_publish(const DistPush& dp) {
  __threadfence_system();
  __syncthreads();
  __shared__ unsigned last;
  if (threadIdx.x == 0) last = (atomicAdd(dp.counter, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (last) {
    __threadfence_system();
    if ((int)threadIdx.x < dp.world) {
      uint32_t* flags = reinterpret_cast<uint32_t*>(static_cast<char*>(dp.pt.base[threadIdx.x]) + dp.flag_off);
      st_release_sys(flags + dp.rank, dp.epoch);
    }
    if (threadIdx.x == 0) *dp.counter = 0u;
  }
}

// spin until every rank's slot of the local flag array reached `epoch`; gives up after ~timeout_cycles (sets bit 1 of *err)
static __global__ void k_dist_wait(const uint32_t* flags, int world, uint32_t epoch, long long timeout_cycles, int32_t* err) {
  if (threadIdx.x < world) {
    const long long t0 = clock64();
    // epochs are compared modulo 2^32 (signed difference)
    while ((int32_t)(ld_acquire_sys(flags + threadIdx.x) - epoch) < 0) {
      __nanosleep(200);
      if (clock64() - t0 > timeout_cycles) {
        if (err) atomicOr(err, 2);
        break;
      }
    }
  }
  __threadfence_system();
}

}  // namespace hx
