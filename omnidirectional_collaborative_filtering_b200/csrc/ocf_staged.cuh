// Row gathers staged through shared memory by the copy engine (cp.async.bulk + mbarrier).
//
// K3 and K4b read one contiguous weight row (4 * HP bytes) per rating / per task, at random places of a catalogue
// far larger than L2. With plain loads the bytes a warp keeps in flight are the registers it can spare (two rows in
// K3), and every dependent round costs a full DRAM round trip: K3 ran at 0.47 of the HBM roofline on the ML-10M
// shape. Here every warp owns a small ring of row slots in shared memory; one elected lane asks the copy engine for
// the next rows (`cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes`, one instruction per row) while
// the warp works on the rows that have landed. In-flight bytes are set by the ring (SLOTS rows per warp), not by
// the register file, and the load latency leaves the warp's dependency chain.
#pragma once

#include "ocf_kernels.cuh"
#include "ocf_score_tc.cuh"

namespace ocf {
namespace st {

using tc::mbar_expect_tx;
using tc::mbar_init;
using tc::mbar_wait;
using tc::smem_u32;

// one row (bytes: a multiple of 16, both addresses 16-byte aligned) global -> shared, completion counted on `bar`
__device__ __forceinline__ void bulk_row(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int NV> struct DecCfg { static constexpr int SLOTS = NV <= 4 ? 4 : 2; };

// ============================================================================================
// K3, staged: same contract and same arguments as k_dec_fwd (ocf_kernels.cuh), for work items of
// at most a few hundred ratings (the catalogues whose rows are short: there a CTA's life is its
// chain of dependent row loads). Target j of a 128-entry tile goes to warp j % 4; a warp keeps
// SLOTS rows in flight and finishes two rows per round (two interleaved shuffle reductions).
// ============================================================================================
template <int NV, bool TRAIN>
__global__ void __launch_bounds__(128)
k_dec_fwd_st(BatchDev bt, const float* __restrict__ WdecT, const float* __restrict__ bdec,
             const float* __restrict__ h, float gscale, int loss_kind,
             float* __restrict__ dy, float4* __restrict__ P2, float* __restrict__ itemstats,
             float* __restrict__ dense_out, int n_cols, TailBuf tb, float4* __restrict__ dh_out, float* __restrict__ rowstats) {
  pdl_trigger();
  pdl_wait();
  constexpr int HP = NV * 128;
  constexpr int SLOTS = DecCfg<NV>::SLOTS;
  constexpr uint32_t ROW_BYTES = HP * 4;
  __shared__ __align__(128) float4 ring[4][SLOTS][HP / 4];      // per warp: SLOTS rows; reused for the warps' dh rows at the end
  __shared__ __align__(8) uint64_t bars[4][SLOTS];
  __shared__ float sred[4][3];
  __shared__ int2 s_list[128];              // (column, entry) of the target entries of 128 entries
  __shared__ float s_t[128];
  __shared__ float s_bias[128];
  __shared__ int s_cnt[4];
  if ((int)blockIdx.x >= bt.hdr->n_items) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < 4 * SLOTS) mbar_init(&bars[0][0] + threadIdx.x, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  const float aux_val = bt.hdr->aux_value;
  const int4 it = bt.items[blockIdx.x];
  const int b = it.x, len = it.z;
  const int p0 = bt.ent_off[b] + it.y;
  float4 hreg[NV], dh[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    hreg[v] = ldg4(h + (size_t)b * HP + v * 128 + lane * 4);
    dh[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float sse = 0.f, sae = 0.f, cnt = 0.f;
  uint32_t phase = 0;                        // parity of every slot's next completion (warp-uniform)

  for (int tile = 0; tile < len; tile += 128) {
    const int i = tile + (int)threadIdx.x;
    int c = 0, code = 0; float t = 0.f;
    if (i < len) { c = bt.ent_col[p0 + i]; t = bt.ent_val[p0 + i]; code = bt.codes[p0 + i]; }
    const bool tgt = (code & CODE_TGT) != 0;
    if (TRAIN && i < len && !tgt) dy[p0 + i] = 0.f;
    const float bias = tgt ? __ldg(bdec + c) : 0.f;
    const unsigned bal = __ballot_sync(FULL, tgt);
    if (lane == 0) s_cnt[warp] = __popc(bal);
    __syncthreads();                         // (first tile: also publishes the barrier inits)
    int mine = 0, total = 0;
#pragma unroll
    for (int w2 = 0; w2 < 4; ++w2) { const int n = s_cnt[w2]; if (w2 < warp) mine += n; total += n; }
    if (tgt) { const int pos = mine + __popc(bal & ((1u << lane) - 1u)); s_list[pos] = make_int2(c, i); s_t[pos] = t; s_bias[pos] = bias; }
    __syncthreads();
    const int n_w = total > warp ? (total - warp + 3) >> 2 : 0;          // this warp's targets: warp, warp + 4, ...
    auto issue = [&](int k) {                // lane 0: row of the warp's k-th target -> slot k % SLOTS
      const int slot = k % SLOTS;
      mbar_expect_tx(&bars[warp][slot], ROW_BYTES);
      bulk_row(&ring[warp][slot][0], WdecT + (size_t)s_list[warp + 4 * k].x * HP, ROW_BYTES, &bars[warp][slot]);
    };
    if (lane == 0)
      for (int k = 0; k < min(SLOTS, n_w); ++k) issue(k);
    // one landed row: the loss terms of its target and (training) its share of dL/dh
    auto finish = [&](int j, float d, const float4 (&w)[NV]) {
      const int2 e0 = s_list[j];
      const float t0 = s_t[j];
      const float y = aux_val * (d + s_bias[j]), e = y - t0;
      sse = fmaf(e, e, sse); sae += fabsf(e); cnt += ((t0 + y) != 0.f) ? 1.f : 0.f;
      if (dense_out != nullptr && lane == 0) dense_out[(size_t)b * n_cols + e0.x] = y;
      if (TRAIN) {
        const float ge = loss_kind == OCF_LOSS_MSE ? gscale * e : gscale * (float)((e > 0.f) - (e < 0.f));
        const float dyv = aux_val * ge;
        if (lane == 0) dy[p0 + e0.y] = dyv;
#pragma unroll
        for (int v = 0; v < NV; ++v) fma4(dh[v], dyv, w[v]);
      }
    };
    for (int k = 0; k < n_w; k += 2) {
      const bool two = k + 1 < n_w;
      const int s0 = k % SLOTS, s1 = (k + 1) % SLOTS;
      float4 w0[NV], w1[NV];
      mbar_wait(&bars[warp][s0], (phase >> s0) & 1u); phase ^= 1u << s0;
#pragma unroll
      for (int v = 0; v < NV; ++v) w0[v] = ring[warp][s0][v * 32 + lane];
      if (two) {
        mbar_wait(&bars[warp][s1], (phase >> s1) & 1u); phase ^= 1u << s1;
#pragma unroll
        for (int v = 0; v < NV; ++v) w1[v] = ring[warp][s1][v * 32 + lane];
      } else {
#pragma unroll
        for (int v = 0; v < NV; ++v) w1[v] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      __syncwarp();                          // every lane has its part of both rows: the slots may be refilled
      if (lane == 0) {
        if (k + SLOTS < n_w) issue(k + SLOTS);
        if (k + 1 + SLOTS < n_w) issue(k + 1 + SLOTS);
      }
      float d0 = 0.f, d1 = 0.f;
#pragma unroll
      for (int v = 0; v < NV; ++v) { d0 = dot4(w0[v], hreg[v], d0); d1 = dot4(w1[v], hreg[v], d1); }
      d0 = warp_sum(d0); d1 = warp_sum(d1);
      finish(warp + 4 * k, d0, w0);
      if (two) finish(warp + 4 * (k + 1), d1, w1);
    }
    if (tile + 128 < len) __syncthreads();   // the list is rebuilt for the next 128 entries
  }
  float4 (*red)[HP / 4] = reinterpret_cast<float4 (*)[HP / 4]>(&ring[0][0][0]);       // red[w] = ring[w][0]: the warp's own slot, drained
  if (lane == 0) { sred[warp][0] = sse; sred[warp][1] = sae; sred[warp][2] = cnt; }
  if (TRAIN) {
#pragma unroll
    for (int v = 0; v < NV; ++v) ring[warp][0][v * 32 + lane] = dh[v];
  }
  __syncthreads();
  if (TRAIN) {
    for (int u = threadIdx.x; u < HP / 4; u += 128) {
      float4 s = ring[0][0][u];
      const float4 a = ring[1][0][u], b2 = ring[2][0][u], c2 = ring[3][0][u];
      s.x = ((s.x + a.x) + b2.x) + c2.x; s.y = ((s.y + a.y) + b2.y) + c2.y;
      s.z = ((s.z + a.z) + b2.z) + c2.z; s.w = ((s.w + a.w) + b2.w) + c2.w;
      P2[(size_t)blockIdx.x * (HP / 4) + u] = s;
    }
  }
  (void)red;
  if (threadIdx.x < 3)
    itemstats[(size_t)blockIdx.x * ROWSTAT_W + threadIdx.x] =
        ((sred[0][threadIdx.x] + sred[1][threadIdx.x]) + sred[2][threadIdx.x]) + sred[3][threadIdx.x];
  if (threadIdx.x == 3) itemstats[(size_t)blockIdx.x * ROWSTAT_W + 3] = 0.f;
  row_tail(bt, (int)blockIdx.x, b, P2, TRAIN ? HP / 4 : 0, itemstats, tb,
           [&](int u, const float4 s) { dh_out[(size_t)b * (HP / 4) + u] = s; },
           [&](int k, float s) { rowstats[(size_t)b * ROWSTAT_W + k] = s; });
}

// ============================================================================================
// K4b, staged: the lean variant of k_row_update (catalogues far beyond L2: one warp per task, no
// heavy / wide walk) with the task's weight row and optimizer-state rows brought in by the copy
// engine. A warp keeps RING tasks' rows in flight (RING x (1 + states) x 4 HP bytes) while it walks
// the matches of the oldest one (L2 reads of activation rows), so the DRAM latency of a task is
// hidden behind the previous task's walk and update instead of heading every task's chain, and
// the bytes in flight per SM no longer depend on the register file.
// Task descriptors are fetched 32 at a time (lane l holds the warp's (32 q + l)-th task).
// Same RowArgs, same arithmetic and summation order as k_row_update: results are bit-identical.
// ============================================================================================
template <int NV, int KIND> struct UpdCfg {
  static constexpr int NROWS = KIND == OCF_OPT_SGD ? 1 : (KIND == OCF_OPT_ADAM ? 3 : 2);   // W + optimizer states
  static constexpr int RING = 2;
  static constexpr int SLOT_BYTES = NROWS * NV * 128 * 4;
  static constexpr int WARPS = SLOT_BYTES <= 4096 ? 8 : 4;
  static constexpr int SMEM = WARPS * RING * SLOT_BYTES;
  static constexpr int CTAS_PER_SM = (200 * 1024) / SMEM < 1 ? 1 : ((200 * 1024) / SMEM > 4 ? 4 : (200 * 1024) / SMEM);
};

template <int NV, int KIND>
__global__ void __launch_bounds__(UpdCfg<NV, KIND>::WARPS * 32, UpdCfg<NV, KIND>::CTAS_PER_SM)
k_row_update_st(RowArgs a) {
  pdl_trigger();
  pdl_wait();
  using C = UpdCfg<NV, KIND>;
  constexpr int HP = NV * 128;
  constexpr uint32_t ROW_BYTES = HP * 4;
  extern __shared__ __align__(128) uint8_t dsm[];
  __shared__ __align__(8) uint64_t bars[C::WARPS][C::RING];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4* ring = reinterpret_cast<float4*>(dsm) + (size_t)warp * C::RING * C::NROWS * (HP / 4);
  if (threadIdx.x < C::WARPS * C::RING) mbar_init(&bars[0][0] + threadIdx.x, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const long long n_tasks = a.dense ? (long long)a.n_cols * a.n_arr : (long long)a.counters[1];
  const int n_mine = gwarp < n_tasks ? (int)((n_tasks - gwarp + nwarps - 1) / nwarps) : 0;   // tasks gwarp, gwarp + nwarps, ...

  // the warp's k-th task (k warp-uniform); descriptors are loaded for 32 tasks at once
  int4 blk = make_int4(0, 0, 0, 0);
  int blk_id = -1;
  auto desc_of = [&](int k) -> int4 {
    if ((k >> 5) != blk_id) {
      blk_id = k >> 5;
      const int k2 = blk_id * 32 + lane;
      if (k2 < n_mine) {
        const long long t = gwarp + (long long)k2 * nwarps;
        if (a.dense) {
          const int c = (int)(t / a.n_arr);
          const int2 seg = a.colseg[c];
          blk = make_int4(c, a.arr_map[t - (long long)c * a.n_arr], seg.x, seg.y);
        } else {
          blk = a.tasks[t];
        }
      }
    }
    const int src = k & 31;
    return make_int4(__shfl_sync(FULL, blk.x, src), __shfl_sync(FULL, blk.y, src), __shfl_sync(FULL, blk.z, src), __shfl_sync(FULL, blk.w, src));
  };
  auto row_of = [&](int c, int arr) -> size_t { return arr == 0 ? (size_t)c * HP : ((size_t)(arr - 1) * a.n_cols + c) * HP; };

  int next = 0, head = 0, tail = 0;          // next task to look at; tasks consumed / issued (those not skipped)
  uint32_t phase = 0;
  int4 sd = make_int4(0, 0, 0, 0);           // lane s < RING: the task whose rows are in slot s
  auto fill = [&]() {
    while (tail - head < C::RING && next < n_mine) {
      const int4 d = desc_of(next);
      ++next;
      if (!a.dense && a.only != 0 && (a.only == 1) != (d.y == 0)) continue;
      const int slot = tail % C::RING;
      if (lane == slot) sd = d;
      if (lane == 0) {
        const size_t r = row_of(d.x, d.y);
        float4* dst = ring + (size_t)slot * C::NROWS * (HP / 4);
        mbar_expect_tx(&bars[warp][slot], C::NROWS * ROW_BYTES);
        bulk_row(dst, (d.y == 0 ? a.WdecT : a.Wenc) + r, ROW_BYTES, &bars[warp][slot]);
        if (KIND != OCF_OPT_SGD) bulk_row(dst + HP / 4, (d.y == 0 ? a.Wd_s1 : a.We_s1) + r, ROW_BYTES, &bars[warp][slot]);
        if (KIND == OCF_OPT_ADAM) bulk_row(dst + 2 * (HP / 4), (d.y == 0 ? a.Wd_s2 : a.We_s2) + r, ROW_BYTES, &bars[warp][slot]);
      }
      ++tail;
    }
  };
  fill();
  const float lr = __ldg(&a.opt.st->lr);
  while (head < tail) {
    const int slot = head % C::RING;
    const int c = __shfl_sync(FULL, sd.x, slot), arr = __shfl_sync(FULL, sd.y, slot);
    const int base = __shfl_sync(FULL, sd.z, slot), n = __shfl_sync(FULL, sd.w, slot);
    // g = sum over the task's matches of coef * X[b, :], in match order (L2 reads; the task's rows are on their way)
    float4 g[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) g[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    float cs = 0.f;
    {
      const float* X = (arr == 0 ? a.hdec : a.dz0) + lane * 4;
      const uint32_t bit = arr == 0 ? CODE_TGT : (arr == 1 ? a.bits.x : (arr == 2 ? a.bits.y : a.bits.z));
      for (int i0 = 0; i0 < n; i0 += 32) {
        uint32_t bc = 0; float coef = 0.f;
        if (i0 + lane < n) {
          const uint32_t* rec = a.matches + (size_t)(base + i0 + lane) * 3;
          bc = rec[0];
          coef = arr == 0 ? __ldg(a.dy + rec[2]) : (arr == 1 ? __uint_as_float(rec[1]) : __ldg(&a.bt_hdr->aux_value));
        }
        unsigned m = __ballot_sync(FULL, ((bc >> 16) & bit) != 0);
        while (m) {
          const int j = __ffs(m) - 1; m &= m - 1;
          const uint32_t b = __shfl_sync(FULL, bc, j) & 0xffffu;
          const float cf = __shfl_sync(FULL, coef, j);
          const float* x = X + (size_t)b * HP;
#pragma unroll
          for (int v = 0; v < NV; ++v) fma4(g[v], cf, ldg4(x + v * 128));
          cs += cf;
        }
      }
    }
    // the row and its state have landed (or land now)
    mbar_wait(&bars[warp][slot], (phase >> slot) & 1u); phase ^= 1u << slot;
    const float4* src = ring + (size_t)slot * C::NROWS * (HP / 4);
    float4 w[NV], t1[NV], t2[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      w[v] = src[v * 32 + lane];
      t1[v] = KIND != OCF_OPT_SGD ? src[HP / 4 + v * 32 + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
      t2[v] = KIND == OCF_OPT_ADAM ? src[2 * (HP / 4) + v * 32 + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncwarp();                            // every lane holds its part of the slot: refill it before the arithmetic
    ++head;
    fill();
    const size_t r = row_of(c, arr);
    float* Wrow = (arr == 0 ? a.WdecT : a.Wenc) + r + lane * 4;
    float* S1row = (arr == 0 ? a.Wd_s1 : a.We_s1) + r + lane * 4;
    float* S2row = (arr == 0 ? a.Wd_s2 : a.We_s2) + r + lane * 4;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      opt_apply_k<KIND>(a.opt, lr, g[v].x, w[v].x, t1[v].x, t2[v].x);
      opt_apply_k<KIND>(a.opt, lr, g[v].y, w[v].y, t1[v].y, t2[v].y);
      opt_apply_k<KIND>(a.opt, lr, g[v].z, w[v].z, t1[v].z, t2[v].z);
      opt_apply_k<KIND>(a.opt, lr, g[v].w, w[v].w, t1[v].w, t2[v].w);
      *reinterpret_cast<float4*>(Wrow + v * 128) = w[v];
      if (KIND != OCF_OPT_SGD) *reinterpret_cast<float4*>(S1row + v * 128) = t1[v];
      if (KIND == OCF_OPT_ADAM) *reinterpret_cast<float4*>(S2row + v * 128) = t2[v];
    }
    if (arr == 0 && lane == 0) {             // decoder bias: column-local gradient sum_b dy[b, c]
      OptDev ob = a.opt; ob.l2x2 = 0.f;      // Keras regularises kernels only
      float wb = a.bdec[c], b1 = KIND != OCF_OPT_SGD ? a.bd_s1[c] : 0.f, b2 = KIND == OCF_OPT_ADAM ? a.bd_s2[c] : 0.f;
      opt_apply_k<KIND>(ob, lr, cs, wb, b1, b2);
      a.bdec[c] = wb;
      if (KIND != OCF_OPT_SGD) a.bd_s1[c] = b1;
      if (KIND == OCF_OPT_ADAM) a.bd_s2[c] = b2;
    }
  }
}

}  // namespace st
}  // namespace ocf
