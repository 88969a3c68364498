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
#include <algorithm>
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

// ---- variants with software prefetch into L2 (cp.async.bulk.prefetch.L2: one instruction per row, no registers,
// no shared memory: the bytes in flight are bounded by L2, not by the register file) ----
__device__ __forceinline__ void prefetch_row(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

template <int NV, int P>
__global__ void __launch_bounds__(128) k_ldg_pf(const float* __restrict__ W, const int* __restrict__ idx, int per_warp, float4* __restrict__ out) {
  constexpr int HP = NV * 128;
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int* my = idx + (size_t)gw * per_warp;
  float4 acc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) acc[v] = make_float4(0, 0, 0, 0);
  if (lane < P && lane < per_warp) prefetch_row(W + (size_t)my[lane] * HP, HP * 4);
  for (int i = 0; i < per_warp; i += 2) {
    if (lane < 2 && i + P + lane < per_warp) prefetch_row(W + (size_t)my[i + P + lane] * HP, HP * 4);
    float4 w[2][NV];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float* row = W + (size_t)my[i + r] * HP + lane * 4;
#pragma unroll
      for (int v = 0; v < NV; ++v) w[r][v] = __ldg(reinterpret_cast<const float4*>(row + v * 128));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int v = 0; v < NV; ++v) { acc[v].x += w[r][v].x; acc[v].y += w[r][v].y; acc[v].z += w[r][v].z; acc[v].w += w[r][v].w; }
  }
#pragma unroll
  for (int v = 0; v < NV; ++v) out[((size_t)gw * NV + v) * 32 + lane] = acc[v];
}

// read-modify-write of (weight row, state row) pairs, the traffic of the fused update kernel: tasks dealt to
// warps grid-stride; WPB warps per CTA; P = rows prefetched ahead into L2 (0: none)
template <int NV, int P, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_rmw(float* __restrict__ W, float* __restrict__ S, const int* __restrict__ idx, int n_tasks) {
  constexpr int HP = NV * 128;
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  if (P > 0 && lane < P && gw + lane * nw < n_tasks) {
    prefetch_row(W + (size_t)idx[gw + lane * nw] * HP, HP * 4);
    prefetch_row(S + (size_t)idx[gw + lane * nw] * HP, HP * 4);
  }
  for (int t = gw; t < n_tasks; t += nw) {
    if (P > 0 && lane == 0 && t + P * nw < n_tasks) {
      prefetch_row(W + (size_t)idx[t + P * nw] * HP, HP * 4);
      prefetch_row(S + (size_t)idx[t + P * nw] * HP, HP * 4);
    }
    float* wr = W + (size_t)idx[t] * HP + lane * 4;
    float* sr = S + (size_t)idx[t] * HP + lane * 4;
    float4 w[NV], s[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) { w[v] = *reinterpret_cast<const float4*>(wr + v * 128); s[v] = *reinterpret_cast<const float4*>(sr + v * 128); }
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      s[v].x = fmaf(w[v].x, w[v].x, s[v].x); s[v].y = fmaf(w[v].y, w[v].y, s[v].y); s[v].z = fmaf(w[v].z, w[v].z, s[v].z); s[v].w = fmaf(w[v].w, w[v].w, s[v].w);
      w[v].x -= 1e-3f * s[v].x; w[v].y -= 1e-3f * s[v].y; w[v].z -= 1e-3f * s[v].z; w[v].w -= 1e-3f * s[v].w;
      *reinterpret_cast<float4*>(wr + v * 128) = w[v];
      *reinterpret_cast<float4*>(sr + v * 128) = s[v];
    }
  }
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
  for (int per_warp : {16, 32, 64}) {
    const int warps = total_rows_to_read / per_warp, ctas = warps / 4;
    auto report = [&](const char* name, float ms) { printf("  per_warp %3d ctas %5d (%.1f/SM)  %-10s %7.1f us  %6.0f GB/s\n", per_warp, ctas, ctas / 148.0, name, ms * 1e3, bytes / ms / 1e6); };
    report("ldg2+pf4", time_ms([&] { k_ldg_pf<NV, 4><<<ctas, 128>>>(W, idx, per_warp, out); }, 20));
    report("ldg2+pf8", time_ms([&] { k_ldg_pf<NV, 8><<<ctas, 128>>>(W, idx, per_warp, out); }, 20));
    report("ldg2+pf16", time_ms([&] { k_ldg_pf<NV, 16><<<ctas, 128>>>(W, idx, per_warp, out); }, 20));
  }
  CK(cudaFree(W)); CK(cudaFree(idx)); CK(cudaFree(out));
}

// distinct random rows (a permutation prefix): every row is read and written once, like the update kernel's tasks
template <int NV>
static void run_rmw(size_t n_rows, int n_tasks, bool sorted = false) {
  constexpr int HP = NV * 128;
  float *W, *S; CK(cudaMalloc(&W, n_rows * HP * 4)); CK(cudaMalloc(&S, n_rows * HP * 4));
  CK(cudaMemset(W, 0, n_rows * HP * 4)); CK(cudaMemset(S, 0, n_rows * HP * 4));
  std::vector<int> perm(n_rows);
  for (size_t i = 0; i < n_rows; ++i) perm[i] = (int)i;
  uint64_t s = 1234567ull;
  for (size_t i = n_rows - 1; i > 0; --i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; std::swap(perm[i], perm[s % (i + 1)]); }
  if (sorted) std::sort(perm.begin(), perm.begin() + n_tasks);      // the same random subset of rows, visited in address order
  int* idx; CK(cudaMalloc(&idx, (size_t)n_tasks * 4)); CK(cudaMemcpy(idx, perm.data(), (size_t)n_tasks * 4, cudaMemcpyHostToDevice));
  const double bytes = (double)n_tasks * HP * 4 * 4;      // W + S, read + write
  printf("%s ", sorted ? "SORTED" : "RANDOM");
  printf("RMW HP=%d tables 2 x %.0f MB, %d tasks (%.0f MB of traffic) per launch\n", HP, n_rows * HP * 4 / 1e6, n_tasks, bytes / 1e6);
  auto report = [&](const char* name, float ms) { printf("  %-34s %7.1f us  %6.0f GB/s\n", name, ms * 1e3, bytes / ms / 1e6); };
  report("256 thr x 3/SM (as K4b), no pf", time_ms([&] { k_rmw<NV, 0, 256, 3><<<148 * 3, 256>>>(W, S, idx, n_tasks); }, 10));
  report("256 thr x 6 CTAs/SM grid, no pf", time_ms([&] { k_rmw<NV, 0, 256, 3><<<148 * 6, 256>>>(W, S, idx, n_tasks); }, 10));
  report("256 thr x 3/SM, pf 1", time_ms([&] { k_rmw<NV, 1, 256, 3><<<148 * 3, 256>>>(W, S, idx, n_tasks); }, 10));
  report("256 thr x 3/SM, pf 2", time_ms([&] { k_rmw<NV, 2, 256, 3><<<148 * 3, 256>>>(W, S, idx, n_tasks); }, 10));
  report("256 thr x 3/SM, pf 4", time_ms([&] { k_rmw<NV, 4, 256, 3><<<148 * 3, 256>>>(W, S, idx, n_tasks); }, 10));
  report("128 thr x 12/SM, no pf", time_ms([&] { k_rmw<NV, 0, 128, 12><<<148 * 12, 128>>>(W, S, idx, n_tasks); }, 10));
  report("128 thr x 16/SM, no pf", time_ms([&] { k_rmw<NV, 0, 128, 16><<<148 * 16, 128>>>(W, S, idx, n_tasks); }, 10));
  report("128 thr x 12/SM, pf 2", time_ms([&] { k_rmw<NV, 2, 128, 12><<<148 * 12, 128>>>(W, S, idx, n_tasks); }, 10));
  report("128 thr x 8/SM, pf 4", time_ms([&] { k_rmw<NV, 4, 128, 8><<<148 * 8, 128>>>(W, S, idx, n_tasks); }, 10));
  CK(cudaFree(W)); CK(cudaFree(S)); CK(cudaFree(idx));
}

int main(int argc, char** argv) {
  const bool all = argc > 1;
  if (!all) {
    run_rmw<4>(300000, 97000, false); run_rmw<4>(300000, 97000, true);
    run_rmw<8>(480189, 200000, true);
    return 0;
  }
  run<4>(600000, 98304);       // 1.2 GB table, 201 MB read per launch: the shape of K2 on the ML-10M workload
  if (all) run<4>(600000, 393216);      // 805 MB per launch
  if (all) run<8>(480189, 131072);      // Netflix rows (4 KB), 537 MB per launch
  run_rmw<4>(300000, 97000);   // the update kernel on the ML-10M workload: 97 k (weight row, state row) pairs of 2 KB, 795 MB
  run_rmw<8>(480189, 200000);  // Netflix rows
  return 0;
}
