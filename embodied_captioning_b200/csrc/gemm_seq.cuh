// GEMM sequence: the run of small dependent GEMMs (+ LayerNorms) between two attention kernels of a decoder step as ONE
// persistent kernel (sm_100a, tcgen05 / TMEM / TMA; included by gemm_host.cu).
//
// Why.  A decoder step at 1024 crops is ~280 dependent kernels of 5-15 us; each graph node costs 1.15 us before its
// first instruction and ~2.2 us of prologue / teardown (barrier init, TMEM allocation, descriptor fetch, cold first
// loads) -- measured in DESIGN.md 3a: 70 % of the decoder is that chain, not bandwidth.  Every operation in the run
//     [V-projection] -> out_proj (+x) -> LayerNorm -> FC1 + GELU -> FC2 (+x) -> LayerNorm -> QKV / Q [-> absorbed Q']
// is local to a crop row, so the only true dependency between two stages is "the rows of MY 128-row block".  One CTA
// per SM stays resident for the whole run; stage s hands its rows to stage s + 1 through per-row-block counters in
// global memory (release: epilogue stores -> named barrier -> fence -> atomicAdd; acquire: ld.acquire spin ->
// fence.proxy.async -> TMA loads), never through a grid-wide barrier.  What it buys per removed launch: the node
// latency, the prologue / teardown, and -- because weights do not depend on anything -- the first ring of W tiles of
// the NEXT stage is already in flight while the CTA waits for its A rows.
//
// Stage = one GEMM of gemm_tcgen05_kernel's shape (128 x BN tiles, UMMA 128 x BN x 16, same epilogues, same fixed
// even / odd k-block split over two MMA-issuing warps into two TMEM accumulators -> results are bit-identical to the
// stand-alone launches), optionally followed by a LayerNorm tail: once all n-tiles of a row block are stored, each of
// the block's CTAs normalises its share of the 128 rows (same arithmetic as layernorm.cu).
// Tiles of a stage: tile t -> CTA t % gridDim; all CTAs are co-resident (grid <= SM count, 1 CTA / SM), so a spin wait
// on another CTA's counter always ends.  Counters: [stage][kind][row block], one 128-byte line each, zeroed by the
// caller before the launch.
//
// MEASURED (B200, 1024 crops, profiles/r02_gemm_seq.txt): bit-identical to the separate launches and SLOWER -- decoder
// alone 94.6 vs 86.6 ms.  The hand-over is cheap (0.4-1.1 us from "rows published" to the next stage's first operands)
// and every fused stage saves ~3 us, but publishing a tile costs the epilogue + a GPU-scope release (2.3-3 us), the
// LayerNorm tail 4-5 us, and a stage here is ~5 us slower than the stand-alone kernel + LayerNorm launch.  Opt-in
// (EC_SEQ=1 / option "decode_seq"); kept with its bit-identity test and the in-kernel stage clocks (-DEC_SEQ_TIMING).
#pragma once
#include "gemm_kernel.cuh"

namespace ec {

constexpr int kSeqMaxStages = 6;
constexpr int kSeqRing = 4;                                        // shared-memory ring depth (k-blocks)
constexpr int kSeqSlotBytes = kBM * kBK * 2 + 256 * kBK * 2;       // A tile + the widest W tile
constexpr int kSeqAccCols = 256;                                   // second accumulator's TMEM column offset
constexpr int kSeqSmemBytes = kSeqRing * kSeqSlotBytes + kEpiWarps * kEpiStageBytes + 1024 + 1024;
static_assert(kSeqSmemBytes <= 227 * 1024, "shared memory budget exceeded");
constexpr int kSeqMaxRowBlocks = 16;
constexpr int kSeqCounterStride = 32;   // ints: every counter has its own 128-byte line (pollers of one never delay another's atomics)
constexpr int kSeqCounterInts = kSeqMaxStages * 2 * kSeqMaxRowBlocks * kSeqCounterStride;   // per launch

struct SeqStageDev {
  EpiParams p;
  int K;          // reduction length (per group for the grouped form)
  int bn;         // tile width: 64 / 128 / 192 / 256
  int epi;        // GemmEpilogue
  int tiles_n;
  int dep;        // what the A operand waits for: 0 nothing, 1 every n-tile of the previous stage, 2 its LayerNorm rows
  int ln_tail;    // 1: this stage normalises its output rows into p.ln_out
};

struct SeqParams {
  int n_stages, tiles_m;
  int* counters;                       // [stage][2][kSeqMaxRowBlocks] x kSeqCounterStride: tiles stored / rows normalised
  const int* skip_flag;
  const unsigned char* pf_ptr; unsigned long long pf_bytes;     // the next kernel's weights -> L2
  int tag;                             // timing slot (EC_SEQ_TIMING builds)
  SeqStageDev st[kSeqMaxStages];
};

// In-kernel stage clocks of CTA 0 (build with -DEC_SEQ_TIMING; scripts/seq_timing_probe.py): [tag][stage][event], SM clock
// cycles since kernel entry.  Events: 0 prologue done, 1 producer: dependency met, 2 producer: last k-block issued,
// 3 issuer: first operands landed, 4 issuer: accumulator committed, 5 epilogue: accumulator ready, 6 epilogue: tile stored
// and published, 7 epilogue: row block complete (LayerNorm tail may start), 8 epilogue: rows normalised and published.
constexpr int kSeqTimingTags = 64, kSeqTimingEvents = 16;
__device__ long long g_seq_timing[kSeqTimingTags][kSeqMaxStages][kSeqTimingEvents];
#ifdef EC_SEQ_TIMING
#define SEQ_T(stage, ev) do { if (bid == 0 && lane == 0 && cp.tag >= 0 && cp.tag < kSeqTimingTags) \
    g_seq_timing[cp.tag][stage][ev] = clock64() - t_entry; } while (0)
#else
#define SEQ_T(stage, ev) do { } while (0)
#endif

struct SeqMaps { CUtensorMap a[kSeqMaxStages]; CUtensorMap w[kSeqMaxStages]; };

__device__ __forceinline__ uint32_t seq_idesc(int bn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
}

template <int EPI>
__device__ __forceinline__ void seq_epilogue_bn(const EpiParams& p, int bn, uint32_t tmem_acc, int row0, int n_blk,
                                                  int quad, int half, int lane, uint32_t stg, uint32_t acc2_off) {
  if (bn == 64) epilogue_tile<64, EPI>(p, tmem_acc, row0, n_blk, quad, half, lane, stg, 0, acc2_off, 0);
  else if (bn == 128) epilogue_tile<128, EPI>(p, tmem_acc, row0, n_blk, quad, half, lane, stg, 0, acc2_off, 0);
  else if constexpr (EPI != EPI_F32_RESID) {
    if (bn == 192) epilogue_tile<192, EPI>(p, tmem_acc, row0, n_blk, quad, half, lane, stg, 0, acc2_off, 0);
    else epilogue_tile<256, EPI>(p, tmem_acc, row0, n_blk, quad, half, lane, stg, 0, acc2_off, 0);
  }
}

__global__ void __launch_bounds__(kThreads1, 1)
gemm_seq_kernel(const __grid_constant__ SeqMaps maps, const __grid_constant__ SeqParams cp) {
#ifdef EC_SEQ_TIMING
  const long long t_entry = clock64();
#endif
  if (cp.skip_flag != nullptr && *cp.skip_flag != 0) return;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  auto slot_a = [&](int s) { return smem_base + (uint32_t)s * kSeqSlotBytes; };
  auto slot_w = [&](int s) { return smem_base + (uint32_t)s * kSeqSlotBytes + (uint32_t)(kBM * kBK * 2); };
  const uint32_t smem_epi = smem_base + kSeqRing * kSeqSlotBytes;
  const uint32_t bar0 = smem_epi + kEpiWarps * kEpiStageBytes;
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kSeqRing + s); };
  auto tfull_bar = [&](int w) { return bar0 + 8u * (2 * kSeqRing + w); };
  const uint32_t tempty_bar = bar0 + 8u * (2 * kSeqRing + 2);
  const uint32_t tmem_slot = bar0 + 8u * (2 * kSeqRing + 3);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int bid = (int)blockIdx.x, grid = (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < cp.n_stages; ++s) {
      ptx::prefetch_tmap(&maps.a[s]);
      ptx::prefetch_tmap(&maps.w[s]);
    }
    for (int s = 0; s < kSeqRing; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    ptx::mbar_init(tfull_bar(0), 1);
    ptx::mbar_init(tfull_bar(1), 1);
    ptx::mbar_init(tempty_bar, kEpiWarps);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  if (warp == 0) SEQ_T(0, 0);

  auto cnt_tiles = [&](int s, int m) { return cp.counters + ((s * 2 + 0) * kSeqMaxRowBlocks + m) * kSeqCounterStride; };
  auto cnt_rows = [&](int s, int m) { return cp.counters + ((s * 2 + 1) * kSeqMaxRowBlocks + m) * kSeqCounterStride; };

  if (warp == 2) {
    // ===================================================================== weights of the later stages (and of the
    // next kernel) -> L2 while stage 0 runs: they stream through L2 once per decoder step and would arrive at HBM latency
    for (int s = 1; s < cp.n_stages; ++s) {
      const SeqStageDev& st = cp.st[s];
      if (st.p.pf_ptr != nullptr) l2_prefetch_slice(st.p.pf_ptr, st.p.pf_bytes, lane);      // = this stage's own W
    }
    // biases and LayerNorm parameters are HBM-cold every step as well (a few KB each: CTA s takes stage s)
    if (bid < cp.n_stages && lane == 0) {
      const EpiParams& q = cp.st[bid].p;
      const unsigned int nb = ((unsigned int)q.N * 4u) & ~15u;
      if (q.bias != nullptr && nb) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(q.bias), "r"(nb) : "memory");
      if (q.ln_out != nullptr && nb) {
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(q.ln_g), "r"(nb) : "memory");
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(q.ln_b), "r"(nb) : "memory");
      }
    }
    if (cp.pf_ptr != nullptr) l2_prefetch_slice(cp.pf_ptr, cp.pf_bytes, lane);
  } else if (warp == 0) {
    // ===================================================================== TMA producer (warp-uniform, one elected lane)
    int g = 0;
    for (int s = 0; s < cp.n_stages; ++s) {
      struct { int K, bn, tiles_n, dep; struct { int grp_cols, M; } p; } st;     // scalars by value (see the epilogue warps)
      st.K = cp.st[s].K; st.bn = cp.st[s].bn; st.tiles_n = cp.st[s].tiles_n; st.dep = cp.st[s].dep;
      st.p.grp_cols = cp.st[s].p.grp_cols; st.p.M = cp.st[s].p.M;
      const int prev_tiles_n = s > 0 ? cp.st[s - 1].tiles_n : 0;
      const int num_kb = (st.K + kBK - 1) / kBK, tiles = cp.tiles_m * st.tiles_n;
      const uint32_t bytes = (uint32_t)(kBM * kBK * 2 + st.bn * kBK * 2);
      for (int tile = bid; tile < tiles; tile += grid) {
        const int m_blk = tile / st.tiles_n, n_blk = tile % st.tiles_n;
        const int a_k0 = (st.p.grp_cols > 0) ? (n_blk * st.bn / st.p.grp_cols) * st.K : 0;
        // the first ring of W tiles goes out BEFORE the dependency wait (weights depend on nothing)
        const int pre = (st.dep != 0) ? (num_kb < kSeqRing ? num_kb : kSeqRing) : 0;
        for (int kb = 0; kb < pre; ++kb) {
          const int slot = (g + kb) % kSeqRing;
          const uint32_t phase = (uint32_t)(((g + kb) / kSeqRing) & 1);
          if (!bounded_wait(empty_bar(slot), phase ^ 1u, 1)) goto done;
          if (ptx::elect_one()) {
            ptx::mbar_expect_tx(full_bar(slot), bytes);
            ptx::tma_load_2d(slot_w(slot), &maps.w[s], kb * kBK, n_blk * st.bn, full_bar(slot));
          }
          __syncwarp();
        }
        if (st.dep != 0) {
          const int rows = min(kBM, st.p.M - m_blk * kBM);
          const bool ok = (st.dep == 1) ? seq_wait_counter(cnt_tiles(s - 1, m_blk), prev_tiles_n)
                                        : seq_wait_counter(cnt_rows(s - 1, m_blk), rows);
          if (!ok) goto done;
          // the rows were written through the generic proxy by other CTAs; the loads below go through the async proxy
          asm volatile("fence.proxy.async.global;" ::: "memory");
          __syncwarp();
        }
        SEQ_T(s, 1);
        for (int kb = 0; kb < pre; ++kb) {
          const int slot = (g + kb) % kSeqRing;
          if (ptx::elect_one())
            ptx::tma_load_2d(slot_a(slot), &maps.a[s], a_k0 + kb * kBK, m_blk * kBM, full_bar(slot));
          __syncwarp();
        }
        for (int kb = pre; kb < num_kb; ++kb) {
          const int slot = (g + kb) % kSeqRing;
          const uint32_t phase = (uint32_t)(((g + kb) / kSeqRing) & 1);
          if (!bounded_wait(empty_bar(slot), phase ^ 1u, 1)) goto done;
          if (ptx::elect_one()) {
            ptx::mbar_expect_tx(full_bar(slot), bytes);
            ptx::tma_load_2d(slot_a(slot), &maps.a[s], a_k0 + kb * kBK, m_blk * kBM, full_bar(slot));
            ptx::tma_load_2d(slot_w(slot), &maps.w[s], kb * kBK, n_blk * st.bn, full_bar(slot));
          }
          __syncwarp();
        }
        g += num_kb;
        SEQ_T(s, 2);
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ===================================================================== MMA issuers: warp 1 the even k-blocks into
    // the accumulator at column 0, warp 3 the odd ones into the accumulator at column 256 (tiles with one k-block:
    // warp 1 alone).  Warp-uniform loops, one elected lane issues.
    const int w = (warp == 1) ? 0 : 1;
    int g = 0, ti = 0;
    const uint64_t adesc0 = make_smem_desc(slot_a(0)), bdesc0 = make_smem_desc(slot_w(0));
    const uint32_t tmem_d = tmem_base + (uint32_t)(w * kSeqAccCols);
    for (int s = 0; s < cp.n_stages; ++s) {
      const int st_K = cp.st[s].K, st_bn = cp.st[s].bn, st_tiles_n = cp.st[s].tiles_n;
      const int num_kb = (st_K + kBK - 1) / kBK, tiles = cp.tiles_m * st_tiles_n;
      const uint32_t idesc = seq_idesc(st_bn);
      const bool dual = num_kb >= 2;
      for (int tile = bid; tile < tiles; tile += grid, g += num_kb, ++ti) {
        if (w == 1 && !dual) continue;
        if (!bounded_wait(tempty_bar, (uint32_t)((ti & 1) ^ 1), 2)) goto done;      // the previous tile left the TMEM
        ptx::tc_fence_after();
        for (int kb = w; kb < num_kb; kb += (dual ? 2 : 1)) {
          const int slot = (g + kb) % kSeqRing;
          const uint32_t phase = (uint32_t)(((g + kb) / kSeqRing) & 1);
          if (!bounded_wait(full_bar(slot), phase, 3)) goto done;
          ptx::tc_fence_after();
          if (w == 0 && kb == 0) SEQ_T(s, 3);
          const uint64_t adesc = adesc0 + (uint64_t)((slot * kSeqSlotBytes) >> 4);
          const uint64_t bdesc = bdesc0 + (uint64_t)((slot * kSeqSlotBytes) >> 4);
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < kBK / kUmmaK; ++k)
              ptx::umma_bf16(tmem_d, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb > w || k > 0) ? 1u : 0u);
            ptx::umma_commit(empty_bar(slot));
          }
          __syncwarp();
        }
        if (ptx::elect_one()) ptx::umma_commit(tfull_bar(w));
        __syncwarp();
        if (w == 0) SEQ_T(s, 4);
      }
    }
  } else if (warp >= kFirstEpiWarp && warp < kFirstEpiWarp + kEpiWarps) {
    // ===================================================================== epilogue warps
    const int e = warp - kFirstEpiWarp;
    const int quad = warp & 3, half = e >> 2;
    const uint32_t stg = smem_epi + (uint32_t)e * kEpiStageBytes;
    int ti = 0, td = 0;
    for (int s = 0; s < cp.n_stages; ++s) {
      // BY VALUE: through a reference into the parameter block every p.field read is a load that ptxas may not hoist
      // across the asm statements of the epilogue (measured with EC_SEQ_TIMING: 2.7 us per 128 x 64 residual tile,
      // 6.5 us per LayerNorm tail)
      const EpiParams p = cp.st[s].p;
      const int st_K = cp.st[s].K, st_bn = cp.st[s].bn, st_epi = cp.st[s].epi, st_tiles_n = cp.st[s].tiles_n;
      const int st_ln = cp.st[s].ln_tail;
      const int num_kb = (st_K + kBK - 1) / kBK, tiles = cp.tiles_m * st_tiles_n;
      const bool dual = num_kb >= 2;
      for (int tile = bid; tile < tiles; tile += grid, ++ti, td += dual ? 1 : 0) {
        const int m_blk = tile / st_tiles_n, n_blk = tile % st_tiles_n;
        if (!bounded_wait(tfull_bar(0), (uint32_t)(ti & 1), 4)) goto done;
        if (dual && !bounded_wait(tfull_bar(1), (uint32_t)(td & 1), 4)) goto done;
        ptx::tc_fence_after();
        if (e == 0) SEQ_T(s, 5);
        const int row0 = m_blk * kBM + quad * 32;
        const uint32_t acc2 = dual ? (uint32_t)kSeqAccCols : 0u;
        switch (st_epi) {
          case EPI_BF16: seq_epilogue_bn<EPI_BF16>(p, st_bn, tmem_base, row0, n_blk, quad, half, lane, stg, acc2); break;
          case EPI_BF16_GELU: seq_epilogue_bn<EPI_BF16_GELU>(p, st_bn, tmem_base, row0, n_blk, quad, half, lane, stg, acc2); break;
          default: seq_epilogue_bn<EPI_F32_RESID>(p, st_bn, tmem_base, row0, n_blk, quad, half, lane, stg, acc2); break;
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(tempty_bar);
        // ---- publish the tile: every epilogue warp has stored its part -> release -> count
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
        if (e == 0 && lane == 0) seq_red_release(cnt_tiles(s, m_blk), 1);
        if (e == 0) SEQ_T(s, 6);
      }
      if (st_ln) {
        // ---- LayerNorm tail: the rows of the stage are dealt over EVERY epilogue warp of EVERY CTA (row r -> CTA
        // r % grid, warp (r / grid) % 8: one row per warp at 1024 crops), also the CTAs that had no tile in this stage.
        // A warp waits for its row's block to be complete, normalises the row and counts it: no CTA-wide barrier.
        for (int r = bid + e * grid; r < p.M; r += kEpiWarps * grid) {
          const int m_blk = r / kBM;
          if (!seq_wait_counter(cnt_tiles(s, m_blk), st_tiles_n)) goto done;
          if (e == 0) SEQ_T(s, 7);
          seq_ln_row(p, r, lane);
          __syncwarp();
          if (lane == 0) seq_red_release(cnt_rows(s, m_blk), 1);
          if (e == 0) SEQ_T(s, 8);
        }
      }
    }
  }
done:
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace ec
