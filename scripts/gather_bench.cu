// Microbenchmark behind the design of the row-gather kernels (K2 / K3 / K4b all read random 4*HP-byte rows of a
// weight matrix far larger than L2): how many bytes must be in flight per SM, and through which path, to reach
// the HBM roofline?   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_bench gather_bench.cu
//   ldg<R>   R rows in flight per warp through registers (LDG.128, what the kernels did in round 1)
//   bulk<D>  a per-warp ring of D rows in shared memory filled by cp.async.bulk (UBLKCP) + mbarrier, summed from smem
// Every warp walks `per_warp` random rows of a [n_rows, HP] fp32 table and accumulates them (so the loads are live).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int NV, int R>
__global__ void __launch_bounds__(128) k_ldg(const float* __restrict__ W, const int* __restrict__ idx, int per_warp, float4* __restrict__ out) {
  constexpr int HP = NV * 128;
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int* my = idx + (size_t)gw * per_warp;
  float4 acc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) acc[v] = make_float4(0, 0, 0, 0);
  for (int i = 0; i < per_warp; i += R) {
    float4 w[R][NV];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float* row = W + (size_t)my[i + r] * HP + lane * 4;
#pragma unroll
      for (int v = 0; v < NV; ++v) w[r][v] = __ldg(reinterpret_cast<const float4*>(row + v * 128));
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int v = 0; v < NV; ++v) { acc[v].x += w[r][v].x; acc[v].y += w[r][v].y; acc[v].z += w[r][v].z; acc[v].w += w[r][v].w; }
  }
#pragma unroll
  for (int v = 0; v < NV; ++v) out[((size_t)gw * NV + v) * 32 + lane] = acc[v];
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT;\nDONE:\n}" ::"r"(smem_u32(b)),
      "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(b)) : "memory");
}

template <int NV, int D>
__global__ void __launch_bounds__(128) k_bulk(const float* __restrict__ W, const int* __restrict__ idx, int per_warp, float4* __restrict__ out) {
  constexpr int HP = NV * 128;
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* ring = reinterpret_cast<float*>(smem) + (size_t)warp * D * HP;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)4 * D * HP * 4) + warp * D;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int* my = idx + (size_t)gw * per_warp;
  if (lane == 0) {
    for (int d = 0; d < D; ++d) mbar_init(&bars[d], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (lane == 0)
    for (int d = 0; d < D && d < per_warp; ++d) { mbar_expect(&bars[d], HP * 4); bulk_g2s(ring + d * HP, W + (size_t)my[d] * HP, HP * 4, &bars[d]); }
  float4 acc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) acc[v] = make_float4(0, 0, 0, 0);
  for (int i = 0; i < per_warp; ++i) {
    const int d = i % D;
    mbar_wait(&bars[d], (i / D) & 1);
    const float4* row = reinterpret_cast<const float4*>(ring + d * HP) + lane;
#pragma unroll
    for (int v = 0; v < NV; ++v) { const float4 w = row[v * 32]; acc[v].x += w.x; acc[v].y += w.y; acc[v].z += w.z; acc[v].w += w.w; }
    __syncwarp();
    if (lane == 0 && i + D < per_warp) { mbar_expect(&bars[d], HP * 4); bulk_g2s(ring + d * HP, W + (size_t)my[i + D] * HP, HP * 4, &bars[d]); }
  }
#pragma unroll
  for (int v = 0; v < NV; ++v) out[((size_t)gw * NV + v) * 32 + lane] = acc[v];
}

template <typename F>
static float time_ms(F&& launch, int reps) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  launch(); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int r = 0; r < reps; ++r) launch();
  CK(cudaEventRecord(b)); CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, a, b));
  return ms / reps;
}

template <int NV>
static void run(size_t n_rows, int total_rows_to_read) {
  constexpr int HP = NV * 128;
  float* W; CK(cudaMalloc(&W, n_rows * HP * 4)); CK(cudaMemset(W, 0, n_rows * HP * 4));
  std::vector<int> h(total_rows_to_read);
  uint64_t s = 88172645463325252ull;
  for (auto& x : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; x = (int)(s % n_rows); }
  int* idx; CK(cudaMalloc(&idx, h.size() * 4)); CK(cudaMemcpy(idx, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  float4* out; CK(cudaMalloc(&out, (size_t)total_rows_to_read * HP * 4 / 8 + (1 << 20)));
  const double bytes = (double)total_rows_to_read * HP * 4;
  printf("HP=%d table %.0f MB, %d rows read (%.0f MB) per launch\n", HP, n_rows * HP * 4 / 1e6, total_rows_to_read, bytes / 1e6);
  for (int per_warp : {16, 32, 64, 128}) {
    const int warps = total_rows_to_read / per_warp, ctas = warps / 4;
    auto report = [&](const char* name, float ms) { printf("  per_warp %3d ctas %5d (%.1f/SM)  %-10s %7.1f us  %6.0f GB/s\n", per_warp, ctas, ctas / 148.0, name, ms * 1e3, bytes / ms / 1e6); };
    report("ldg<2>", time_ms([&] { k_ldg<NV, 2><<<ctas, 128>>>(W, idx, per_warp, out); }, 20));
    report("ldg<4>", time_ms([&] { k_ldg<NV, 4><<<ctas, 128>>>(W, idx, per_warp, out); }, 20));
    if (NV <= 4) report("ldg<8>", time_ms([&] { k_ldg<NV, 8><<<ctas, 128>>>(W, idx, per_warp, out); }, 20));
    {
      constexpr int D = 4; const size_t sm = (size_t)4 * D * HP * 4 + 4 * D * 8;
      CK(cudaFuncSetAttribute(k_bulk<NV, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      report("bulk<4>", time_ms([&] { k_bulk<NV, D><<<ctas, 128, sm>>>(W, idx, per_warp, out); }, 20));
    }
    {
      constexpr int D = 8; const size_t sm = (size_t)4 * D * HP * 4 + 4 * D * 8;
      CK(cudaFuncSetAttribute(k_bulk<NV, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      report("bulk<8>", time_ms([&] { k_bulk<NV, D><<<ctas, 128, sm>>>(W, idx, per_warp, out); }, 20));
    }
    if (NV <= 4) {
      constexpr int D = 16; const size_t sm = (size_t)4 * D * HP * 4 + 4 * D * 8;
      CK(cudaFuncSetAttribute(k_bulk<NV, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      report("bulk<16>", time_ms([&] { k_bulk<NV, D><<<ctas, 128, sm>>>(W, idx, per_warp, out); }, 20));
    }
  }
  CK(cudaFree(W)); CK(cudaFree(idx)); CK(cudaFree(out));
}

int main() {
  run<4>(600000, 98304);       // 1.2 GB table, 201 MB read per launch: the shape of K2 on the ML-10M workload
  run<4>(600000, 393216);      // 805 MB per launch: the shape of K4b's reads
  run<8>(480189, 131072);      // Netflix rows (4 KB), 537 MB per launch
  return 0;
}
