// One-shot all-reduce over NVLink peer memory, fused with the compute that follows it.
//
// A column-sharded step exchanges two small activations per step ([rows, H] pre-activations and
// [rows, 4 + H] row statistics + dL/dh: 0.5-4 MB): a latency problem, not a bandwidth one. Every
// rank owns an exchange region (flags + two slots) that all peers map through CUDA IPC. The
// producing kernel writes its partial result straight into the rank's slot; then ONE kernel per
// rank
//   1. tells every peer "my slot of epoch e is complete" (a release store into the peer's flag
//      array, per CTA) and waits for the same word from every peer,
//   2. reads the slot of every rank over NVLink, adds them in rank order (the same order on
//      every rank: all ranks get bit-identical sums, so replicated parameters stay identical),
//   3. runs the epilogue on the sum in registers: bias + activation + dropout for the encoder
//      exchange, a plain store for the statistics / dL/dh exchange.
// Slots alternate with the epoch's parity. A slot is rewritten for epoch e+2 only after the rank
// has left the handshake of epoch e+1, which every peer enters after it finished reading epoch e,
// so one handshake per exchange is enough. Flags only grow (no reset, no second barrier).
// Waits are bounded: a protocol error traps instead of hanging the GPUs.
#pragma once

#include "ocf_kernels.cuh"

namespace ocf {
namespace peer {

constexpr int MAX_PEERS = 8;
constexpr int AR_CTAS = 64;
constexpr int AR_THREADS = 256;
constexpr size_t FLAG_BYTES = 16384;      // >= AR_CTAS * MAX_PEERS * 4, keeps the slots 16 KB aligned

struct PeerDev {
  const float4* slot[MAX_PEERS];    // this epoch's slot on every rank (own rank included)
  uint32_t* flags[MAX_PEERS];       // flag array of every rank: [AR_CTAS][MAX_PEERS] epochs
  int rank, world;
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
__device__ __forceinline__ float4 ld_peer(const float4* p) {
  float4 v;
  asm volatile("ld.global.cv.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// CTA c of this rank <-> CTA c of every peer. The slot was completed by an earlier kernel on this
// stream, so announcing it needs no grid-wide synchronisation.
__device__ __forceinline__ void handshake(const PeerDev& pd) {
  if ((int)threadIdx.x < pd.world) {
    __threadfence_system();
    st_release_sys(pd.flags[threadIdx.x] + blockIdx.x * MAX_PEERS + pd.rank, pd.epoch);
    const uint32_t* mine = pd.flags[pd.rank] + blockIdx.x * MAX_PEERS + threadIdx.x;
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(mine) - pd.epoch) < 0) {
      if (clock64() - t0 > 6000000000LL) asm volatile("trap;");
    }
  }
  __syncthreads();
}

__device__ __forceinline__ float4 sum_ranks(const PeerDev& pd, int idx) {
  float4 s = ld_peer(pd.slot[0] + idx);
  for (int p = 1; p < pd.world; ++p) {
    const float4 v = ld_peer(pd.slot[p] + idx);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  return s;
}

// encoder exchange: z = sum over ranks of the partial pre-activations, then bias + activation
// (+ dropout) of the first hidden layer (model.py:66-73) on the sum.
__global__ void __launch_bounds__(AR_THREADS)
k_allreduce_bias_act(PeerDev pd, float4* __restrict__ zsum, ActArgs g) {
  handshake(pd);
  const int count4 = g.B * g.hp4;
  for (int idx = blockIdx.x * AR_THREADS + threadIdx.x; idx < count4; idx += gridDim.x * AR_THREADS) {
    const float4 z = sum_ranks(pd, idx);
    zsum[idx] = z;
    bias_act_elem(g, idx, z);
  }
}

// decoder exchange: [rows, 4] row statistics followed by [rows, hp] dL/dh.
__global__ void __launch_bounds__(AR_THREADS)
k_allreduce_store(PeerDev pd, int n0_4, float4* __restrict__ dst0, int n1_4, float4* __restrict__ dst1) {
  handshake(pd);
  const int count4 = n0_4 + n1_4;
  for (int idx = blockIdx.x * AR_THREADS + threadIdx.x; idx < count4; idx += gridDim.x * AR_THREADS) {
    const float4 v = sum_ranks(pd, idx);
    if (idx < n0_4) dst0[idx] = v; else dst1[idx - n0_4] = v;
  }
}

}  // namespace peer
}  // namespace ocf
