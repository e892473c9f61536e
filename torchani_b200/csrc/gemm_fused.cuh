// The whole ensemble MLP of a step -- three forward layers (+ head) and three backward-to-input layers -- as
// ONE persistent launch of the tcgen05 GEMM of gemm_tc.cuh.
//
// The six launches of round 1 each paid a ramp-up, a tail and their own tile-count quantisation (4.2 tiles per
// CTA -> 84 % balance), and at small row counts (8 GPUs: 10 row tiles per rank) little else: 13.5 us per launch
// for 2-3 us of work.  The layers of one 128-row tile are a pure chain (layer k of row tile r needs layer k-1 of
// row tile r, nothing else), so the launch boundaries are replaced by DATA-FLOW: all (phase, row tile, member,
// column tile) units of the step form one list, phase-major and row-tile-major inside a phase; CTA c takes units
// c, c + G, c + 2G, ... in order, and before it touches the inputs of a unit it waits until the counter of
// (producer phase, row tile) has reached the number of arrivals that phase owes that row tile.  Every producer
// unit has a smaller list index than its consumers and every CTA walks its units in order, so the wait can
// always be satisfied (induction over the list index); row tiles finish in list order, so a phase's first units
// find their inputs long complete and the waits cost nothing except in the last few units.
//
//   * completion of a unit = its TMA bulk stores have fully completed (cp.async.bulk.wait_group, not .read),
//     then fence + one release-add per epilogue warp on sync[phase][row tile].  The wait for the stores is taken
//     one unit late (the stores of unit k have long landed when unit k+1's epilogue ends) or whenever the
//     epilogue would idle anyway (accumulator not ready) -- which also rules out the one cycle lagging could
//     create (a CTA whose next unit depends on its own previous one);
//   * consumers: the producer lane before its first bulk load of a unit, and the epilogue warps before they read
//     the stored activation (EPI_MUL_DCELU), acquire the counter and cross into the async proxy with a fence;
//   * the operand ring changes geometry between phases (stage = 16 KB of A + bn x 128 B of B): the producer
//     drains the ring at a phase change and both roles restart at slot 0; per-barrier parities are tracked in bit
//     masks; the store-staging buffers of the epilogue sit at the END of the dynamic shared memory, the ring at
//     its start, so they never overlap whatever the geometry;
//   * waits are bounded (status ANI_STATUS_INTERNAL instead of a hung GPU).
// The single-phase kernel of gemm_tc.cuh remains (ANI_B200_MLP_FUSED=0, and the CTA-pair experiment).
#pragma once
#include <type_traits>

#include "gemm_tc.cuh"
#include "gemm_epilogue.cuh"

namespace ani {
namespace tc {

constexpr int MAX_PHASES = 6;
constexpr int FUSED_SMEM_BYTES = 227 * 1024 - 12 * 1024;  // 12 KB left for the static part (six tile maps, staging)
constexpr long long FUSED_SPIN_LIMIT = 6000000000LL;      // ~3 s of clock64 ticks
constexpr int FTRACE_UNITS = 32;                          // units per CTA the role timeline records

struct FusedArgs {
  int n_phases;
  int epi[MAX_PHASES];    // epilogue of the phase (EPI_*)
  int dep[MAX_PHASES];    // phase whose row-tile counters the units of this phase wait on (-1: none)
  int32_t* sync;          // [MAX_PHASES][sync_stride] arrivals per (phase, row tile); zeroed by the launcher
  int sync_stride;
  int prefetch_b;         // issue the weight copies of a unit's first K-blocks before waiting for its inputs
  int epi_direct16;       // k_mlp_fused<16>: register-direct epilogue, a warp per (group parity, 16-column half)
  long long* trace;       // optional clock64 stamps [cta < 4][unit < FTRACE_UNITS][role 4][4] (ani_b200_debug_gemm_trace)
  Args ph[MAX_PHASES];
};

__device__ __forceinline__ int ld_acquire_gpu(const int32_t* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_relaxed_gpu(const int32_t* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void red_release_gpu(int32_t* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"   // non-blocking (try_wait may suspend the thread)
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
template <int N>
__device__ __forceinline__ void bulk_wait_done() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// bounded wait for `*counter >= need` (one lane); raises the status word instead of hanging
__device__ __forceinline__ void wait_counter(const int32_t* counter, int need, int32_t* status) {
  if (ld_acquire_gpu(counter) >= need) return;
  const long long t0 = clock64();
  while (ld_acquire_gpu(counter) < need) {
    if (clock64() - t0 > FUSED_SPIN_LIMIT) {
      if (status) atomicOr(status, ANI_STATUS_INTERNAL);
      break;
    }
  }
}

struct UnitRef {
  int phase, local;
};

// ---- the epilogue of one tile with EIGHT epilogue warps (thread = row, warps w and w + 4 alternate 32-column
// groups, TMEM loads one half group ahead): one instantiation per EPI, switch at tile granularity ----------------
template <int EPI>
__device__ __forceinline__ void tile_epilogue8(const Args& args, const TileMap& tm, const Tile& tl, const Species& sp,
                                              uint32_t taddr, unsigned char* sb0, int epi_bufs, uint32_t& buf,
                                              const float* __restrict__ bias, const float* __restrict__ w4,
                                              float* e_part, int warp, int lane, uint64_t* tfull_bar,
                                              uint32_t tfull_parity, float& omax, int& groups_committed,
                                              uint64_t* publish_bar = nullptr, uint64_t* ydep_bar = nullptr,
                                              uint32_t ydep_parity = 0) {
  // ydep_bar (data-flow launch): "the producer has acquired this unit's inputs" -- the stored activation may be read from
  // there on, i.e. during the main loop, instead of only after the accumulator barrier
  // publish_bar (gemm_chain.cuh): this warp's stores of the PREVIOUS unit are still in flight; once the first store group
  // of this tile has been committed, wait for everything older (cp.async.bulk.wait_group 1) and arrive there
  const CeluConst cc{args.alpha, 1.0f / args.alpha, 1.4426950408889634f / args.alpha};
  const int quad = warp & 3, half = warp >> 2;
  const int r_tile = quad * 32 + lane;
  // byte offset of 16-byte chunk ch of this thread's row: inside a 128-row piece (global) and inside this warp's
  // 32-row staging image (shared) -- swz_off() of common.cuh, its row part kept in two registers
  const uint32_t my_base = swz_off(r_tile, 0) & ~0x30u, st_base = swz_off(lane, 0) & ~0x30u;
  const uint32_t row_sw = (uint32_t)(lane >> 1) & 3u;   // (row >> 1) & 3 is the same for r_tile and lane
  auto my_off = [&](int ch) { return my_base + (((uint32_t)ch ^ row_sw) << 4); };
  auto st_off = [&](int ch) { return st_base + (((uint32_t)ch ^ row_sw) << 4); };
  const float acc_scale = sp.acc_scale;
  const bool tiled_out = EPI != EPI_PLAIN && (EPI != EPI_HEAD || args.want_backward);
  const int my_row = tl.rt * TM + r_tile;
  float e_acc = 0.f, seed = 0.f;
  bool row_valid = false;
  if (EPI == EPI_HEAD) {
    row_valid = args.row_atom[my_row] >= 0;
    seed = row_valid ? args.member_scale[tl.mem] : 0.f;
  }
  unsigned char* ct = reinterpret_cast<unsigned char*>(args.C) +
                      ((size_t)tl.rt * args.c_kblocks + (size_t)(tl.mem * sp.c_moff + tl.n0) / TK) * A_BLOCK_BYTES;
  float* cplain = reinterpret_cast<float*>(args.C) + (size_t)my_row * args.ldc + (size_t)tl.mem * sp.c_moff;
  const int ngroups = tl.bn / 32;
  // stored activation of this thread's row: one register set per 16-column half, each refilled for the SAME half of
  // the warp's next group as soon as it has been consumed -- the loads run two halves (one whole 32-column group of
  // epilogue math) ahead of their use; one half ahead left ~27 % of the backward epilogues' stall samples on them
  uint4 yq[2][2 * PARTS];
  auto load_y = [&](int g, int hh, uint4 (&q)[2 * PARTS]) {
    const unsigned char* blk = ct + (size_t)g * A_BLOCK_BYTES;
#pragma unroll
    for (int p = 0; p < PARTS; ++p) {
      // (plain cached loads: this SM has not read these lines before in this launch, so the L1 cannot hold a stale
      // copy of what another SM's TMA stores wrote in an earlier phase; the four chunks of a row share L1 lines)
      q[2 * p] = *reinterpret_cast<const uint4*>(blk + p * A_PART_BYTES + my_off(2 * hh));
      q[2 * p + 1] = *reinterpret_cast<const uint4*>(blk + p * A_PART_BYTES + my_off(2 * hh + 1));
    }
  };
  // The stored activation read here was written by an earlier phase, possibly moments ago: it may be read once this
  // warp is ordered after the producer's acquire of the row tile's inputs -- through the producer's `ydep` barrier
  // (then the loads fly during the main loop) or, without one, through the accumulator barrier.
  const bool y_early = EPI == EPI_MUL_DCELU && ydep_bar != nullptr;
  if (y_early) {
    mbar_wait(ydep_bar, ydep_parity);
    if (half < ngroups) {
      load_y(half, 0, yq[0]);
      load_y(half, 1, yq[1]);
    }
  }
  mbar_wait(tfull_bar, tfull_parity);
  tc_fence_after();
  if (EPI == EPI_MUL_DCELU && !y_early && half < ngroups) {
    load_y(half, 0, yq[0]);
    load_y(half, 1, yq[1]);
  }

  auto process = [&](int g, auto hh_c, const uint32_t (&r)[16]) {
    constexpr int hh = decltype(hh_c)::value;
    float y[16];
    if (EPI == EPI_MUL_DCELU) {
      join_chunk(yq[hh], y);
      join_chunk(yq[hh] + 1, y + 8);
      if (g + 2 < ngroups) load_y(g + 2, hh, yq[hh]);
    }
    unsigned char* sb = sb0 + buf * EPI_STAGE_BYTES;
    if (tiled_out && hh == 0) {
      if (lane == 0) {
        if (epi_bufs == 2)
          bulk_wait_read<1>();
        else
          bulk_wait_read<0>();
      }
      __syncwarp();
    }
    const int c0 = g * 32 + hh * 16;
    float o[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float v = __uint_as_float(r[j]) * acc_scale;
      if (EPI == EPI_BIAS_CELU) {
        v = celu(v + bias[c0 + j], cc);
      } else if (EPI == EPI_MUL_DCELU) {
        v *= dcelu_from_out(y[j] * args.y_inv_scale, cc);
      } else if (EPI == EPI_HEAD) {
        const float w = w4[c0 + j];
        const float a = celu(v + bias[c0 + j], cc);
        e_acc = fmaf(a, w, e_acc);
        v = seed * w * dcelu_from_out(a, cc);
      }
      if (EPI != EPI_PLAIN) {
        v *= args.out_scale;
        omax = fmaxf(omax, fabsf(v));
      }
      o[j] = v;
    }
    if (EPI == EPI_PLAIN) {
      const int col = (tm.nb_count >= 0 ? tm.nb[tl.n0 / 32 + g] * 32 : tl.n0 + g * 32) + hh * 16;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 v4 = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
        if (args.c_accumulate)
          red_add_v4(cplain + col + 4 * q, v4);
        else
          *reinterpret_cast<float4*>(cplain + col + 4 * q) = v4;
      }
    } else if (tiled_out) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t w[4][PARTS];
#pragma unroll
        for (int i = 0; i < 4; ++i) split_pair(o[8 * c + 2 * i], o[8 * c + 2 * i + 1], w[i]);
        const uint32_t off = st_off(2 * hh + c);
#pragma unroll
        for (int p = 0; p < PARTS; ++p)
          *reinterpret_cast<uint4*>(sb + p * EPI_PART_BYTES + off) = make_uint4(w[0][p], w[1][p], w[2][p], w[3][p]);
      }
      if (hh == 1) {
        unsigned char* blk = ct + (size_t)g * A_BLOCK_BYTES + quad * EPI_PART_BYTES;
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
#pragma unroll
          for (int p = 0; p < PARTS; ++p) bulk_s2g(blk + p * A_PART_BYTES, sb + p * EPI_PART_BYTES, EPI_PART_BYTES);
          bulk_commit();
          if (publish_bar && groups_committed == 0) {
            bulk_wait_done<1>();
            mbar_arrive(publish_bar);
          }
        }
        ++groups_committed;
        if (epi_bufs == 2) buf ^= 1;
      }
    }
  };

  {
    uint32_t r0[16], r1[16];
    if (half < ngroups) tmem_ld16_issue(taddr + half * 32, r0);
    for (int g = half; g < ngroups; g += 2) {
      tmem_ld_wait(r0);
      tmem_ld16_issue(taddr + g * 32 + 16, r1);
      process(g, std::integral_constant<int, 0>{}, r0);
      tmem_ld_wait(r1);
      if (g + 2 < ngroups) tmem_ld16_issue(taddr + (g + 2) * 32, r0);
      process(g, std::integral_constant<int, 1>{}, r1);
    }
  }
  if (EPI == EPI_HEAD) {
    e_part[warp * 32 + lane] = e_acc;
    asm volatile("bar.sync 1, %0;" ::"n"(NUM_EPI_WARPS * 32) : "memory");
    if (half == 0)
      args.e_member[(size_t)tl.mem * args.rows_cap + my_row] =
          row_valid ? e_part[warp * 32 + lane] + e_part[(warp + 4) * 32 + lane] + sp.b4[tl.mem] : 0.f;
    asm volatile("bar.sync 1, %0;" ::"n"(NUM_EPI_WARPS * 32) : "memory");
  }
}

// ---- the epilogue of one tile, one instantiation per EPI (switch at tile granularity) -----------------
// SIXTEEN epilogue warps (the single-phase kernel has eight): the step is bound by its epilogues -- 128 elements per
// thread and tile at ~12 dependent instructions each, two warps per scheduler -- so the fused kernel puts four warps
// on every scheduler.  Warp w: TMEM lane quadrant w & 3; the four warps of a quadrant form two PAIRS (w >> 3) that take
// alternate 32-column groups, and inside a pair the warps take one 16-column half each ((w >> 2) & 1).  A pair shares
// one store-staging buffer per group: both warps write their half of every row, meet at a 64-thread named barrier, and
// the first warp of the pair hands the finished 2 KB images to the TMA.
constexpr int fused_threads(int nw) { return (nw + 3) * 32; }  // epilogue warps + MMA + producer + signal warp

__device__ __forceinline__ void pair_barrier(int id) {
  asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
}

template <int EPI>
__device__ __forceinline__ void tile_epilogue16(const Args& args, const TileMap& tm, const Tile& tl, const Species& sp,
                                              uint32_t taddr, unsigned char* sb0, int epi_bufs, uint32_t& buf,
                                              const float* __restrict__ bias, const float* __restrict__ w4,
                                              float* e_part, int warp, int lane, uint64_t* tfull_bar,
                                              uint32_t tfull_parity, float& omax, int& groups_committed) {
  const CeluConst cc{args.alpha, 1.0f / args.alpha, 1.4426950408889634f / args.alpha};
  const int quad = warp & 3, hh = (warp >> 2) & 1, pairidx = warp >> 3;
  const int bar_id = 2 + quad * 2 + pairidx;  // named barriers 2..9 (0 = __syncthreads, 1 = all epilogue warps)
  const int r_tile = quad * 32 + lane;
  const float acc_scale = sp.acc_scale;
  const bool tiled_out = EPI != EPI_PLAIN && (EPI != EPI_HEAD || args.want_backward);
  const int my_row = tl.rt * TM + r_tile;
  float e_acc = 0.f, seed = 0.f;
  bool row_valid = false;
  if (EPI == EPI_HEAD) {
    row_valid = args.row_atom[my_row] >= 0;
    seed = row_valid ? args.member_scale[tl.mem] : 0.f;
  }
  unsigned char* ct = reinterpret_cast<unsigned char*>(args.C) +
                      ((size_t)tl.rt * args.c_kblocks + (size_t)(tl.mem * sp.c_moff + tl.n0) / TK) * A_BLOCK_BYTES;
  float* cplain = reinterpret_cast<float*>(args.C) + (size_t)my_row * args.ldc + (size_t)tl.mem * sp.c_moff;
  const int ngroups = tl.bn / 32;
  // byte offsets of this thread's two 16-byte chunks (columns 16 hh .. 16 hh + 15 of a 32-column group): inside a
  // 128-row piece in global memory, and inside the pair's 32-row staging image
  const uint32_t g_off0 = swz_off(r_tile, 2 * hh), g_off1 = swz_off(r_tile, 2 * hh + 1);
  const uint32_t s_off0 = swz_off(lane, 2 * hh), s_off1 = swz_off(lane, 2 * hh + 1);
  uint4 yq[2 * PARTS];
  auto load_y = [&](int g) {
    const unsigned char* blk = ct + (size_t)g * A_BLOCK_BYTES;
#pragma unroll
    for (int p = 0; p < PARTS; ++p) {
      // (plain cached loads: this SM has not read these lines before in this launch, so the L1 cannot hold a stale
      // copy of what another SM's TMA stores wrote in an earlier phase)
      yq[2 * p] = *reinterpret_cast<const uint4*>(blk + p * A_PART_BYTES + g_off0);
      yq[2 * p + 1] = *reinterpret_cast<const uint4*>(blk + p * A_PART_BYTES + g_off1);
    }
  };
  mbar_wait(tfull_bar, tfull_parity);
  tc_fence_after();
  // (only now: the accumulator barrier is what orders this warp after the producer's acquire of the row tile's inputs)
  if (EPI == EPI_MUL_DCELU && pairidx < ngroups) load_y(pairidx);

  for (int g = pairidx; g < ngroups; g += 2) {
    uint32_t r[16];
    tmem_ld16_issue(taddr + g * 32 + hh * 16, r);
    float y[16];
    if (EPI == EPI_MUL_DCELU) {
      join_chunk(yq, y);
      join_chunk(yq + 1, y + 8);
      if (g + 2 < ngroups) load_y(g + 2);  // prefetch the next group this warp handles
    }
    unsigned char* sb = sb0 + buf * EPI_STAGE_BYTES;
    if (tiled_out) {
      // the staging buffer was handed to the TMA epi_bufs groups ago by the first warp of the pair
      if (hh == 0 && lane == 0) {
        if (epi_bufs == 2)
          bulk_wait_read<1>();
        else
          bulk_wait_read<0>();
      }
      pair_barrier(bar_id);
    }
    tmem_ld_wait(r);
    const int c0 = g * 32 + hh * 16;
    float o[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float v = __uint_as_float(r[j]) * acc_scale;
      if (EPI == EPI_BIAS_CELU) {
        v = celu(v + bias[c0 + j], cc);
      } else if (EPI == EPI_MUL_DCELU) {
        v *= dcelu_from_out(y[j] * args.y_inv_scale, cc);
      } else if (EPI == EPI_HEAD) {
        const float w = w4[c0 + j];
        const float a = celu(v + bias[c0 + j], cc);
        e_acc = fmaf(a, w, e_acc);
        v = seed * w * dcelu_from_out(a, cc);
      }
      if (EPI != EPI_PLAIN) {
        v *= args.out_scale;
        omax = fmaxf(omax, fabsf(v));
      }
      o[j] = v;
    }
    if (EPI == EPI_PLAIN) {
      const int col = (tm.nb_count >= 0 ? tm.nb[tl.n0 / 32 + g] * 32 : tl.n0 + g * 32) + hh * 16;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 v4 = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
        if (args.c_accumulate)
          red_add_v4(cplain + col + 4 * q, v4);
        else
          *reinterpret_cast<float4*>(cplain + col + 4 * q) = v4;
      }
    } else if (tiled_out) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t w[4][PARTS];
#pragma unroll
        for (int i = 0; i < 4; ++i) split_pair(o[8 * c + 2 * i], o[8 * c + 2 * i + 1], w[i]);
        const uint32_t off = c == 0 ? s_off0 : s_off1;
#pragma unroll
        for (int p = 0; p < PARTS; ++p)
          *reinterpret_cast<uint4*>(sb + p * EPI_PART_BYTES + off) = make_uint4(w[0][p], w[1][p], w[2][p], w[3][p]);
      }
      fence_proxy_async();  // generic-proxy smem writes -> visible to the TMA
      pair_barrier(bar_id);
      if (hh == 0) {
        if (lane == 0) {
          unsigned char* blk = ct + (size_t)g * A_BLOCK_BYTES + quad * EPI_PART_BYTES;
#pragma unroll
          for (int p = 0; p < PARTS; ++p) bulk_s2g(blk + p * A_PART_BYTES, sb + p * EPI_PART_BYTES, EPI_PART_BYTES);
          bulk_commit();
        }
        ++groups_committed;
      }
      if (epi_bufs == 2) buf ^= 1;
    }
  }
  if (EPI == EPI_HEAD) {
    // the four warps of a row hold different column groups / halves: combine through shared memory
    e_part[warp * 32 + lane] = e_acc;
    asm volatile("bar.sync 1, %0;" ::"n"(16 * 32) : "memory");
    if (warp < 4)
      args.e_member[(size_t)tl.mem * args.rows_cap + my_row] =
          row_valid ? e_part[warp * 32 + lane] + e_part[(warp + 4) * 32 + lane] + e_part[(warp + 8) * 32 + lane] +
                          e_part[(warp + 12) * 32 + lane] + sp.b4[tl.mem]
                    : 0.f;
    asm volatile("bar.sync 1, %0;" ::"n"(16 * 32) : "memory");
  }
}

// (The register-direct epilogue of gemm_epilogue.cuh is used by the chained launches only.  Measured in this kernel
// on B200, 9 999 atoms: 277 us against 237 us with the staged stores -- here the stored activation may only be read
// after the accumulator barrier, which exposes its L2 latency on every tile (backward tiles 12 us instead of 5-6), and
// the extra live state of the data-flow bookkeeping pushed the inlined epilogue into spills.)
// The sixteen-warp register-direct epilogue as a real call: at 96 registers per thread the unit bookkeeping of the kernel
// around it must not stay live inside its loops
template <int EPI>
__device__ __noinline__ void tile_epilogue_direct16_call(const Args& args, const Tile& tl, const Species& sp, uint32_t taddr,
                                                         const float* bias, const float* w4, float* e_part, int warp,
                                                         int lane, uint64_t* tfull_bar, uint32_t tfull_parity, float* omax,
                                                         uint64_t* ydep_bar, uint32_t ydep_parity) {
  float om = *omax;
  tile_epilogue_direct16<EPI>(args, tl, sp, taddr, bias, w4, e_part, warp, lane, tfull_bar, tfull_parity, om, ydep_bar, ydep_parity);
  *omax = om;
}

// ---- the kernel -----------------------------------------------------------------------------
// NW = 8: thread = row, two warps per quadrant (tile_epilogue8).  NW = 16: warp pairs per 16-column half
// (tile_epilogue16).  Measured on B200 (profiles/): see DESIGN.md 4.1 for which one runs by default.
template <int NW>
__global__ void __launch_bounds__(fused_threads(NW), 1) k_mlp_fused(const __grid_constant__ FusedArgs F) {
  constexpr int FEPI_WARPS = NW, F_MMA_WARP = NW, F_PROD_WARP = NW + 1, F_SIGNAL_WARP = NW + 2;
  constexpr int EPI_SETS = NW == 16 ? 8 : NW;  // store-staging buffer sets: per warp pair (16) or per warp (8)
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ TileMap tms[MAX_PHASES];
  __shared__ int phase_base[MAX_PHASES + 1];
  __shared__ int s_epi_bufs;
  __shared__ int s_done;   // arrivals of epilogue warps: 8 per completed unit (stores landed), monotonic
  __shared__ float e_part[NW * 32];
  __shared__ __align__(16) float s_bias[2][TN_MAX];
  __shared__ __align__(16) float s_w4[2][TN_MAX];
  __shared__ __align__(8) uint64_t bars[2 * MAX_STAGES + 5];
  __shared__ __align__(8) uint64_t ydep[4];   // unit k of this CTA: the producer has acquired its inputs (barrier k & 3)
  uint64_t* full = bars;
  uint64_t* empty = bars + MAX_STAGES;
  uint64_t* tfull = bars + 2 * MAX_STAGES;
  uint64_t* tempty = bars + 2 * MAX_STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NP = F.n_phases;
  constexpr int AVAIL = FUSED_SMEM_BYTES - 1024;

  // the tile maps of the phases are built side by side (one lane of warp p builds phase p): a single thread building
  // all six was 9 % of this kernel's stall samples (every other warp at the barrier below; ncu, profiles/r02_*)
  if (lane == 0 && warp < NP) build_tile_map(F.ph[warp], tms[warp]);
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0, bufs = 2;
    for (int p = 0; p < NP; ++p) {
      phase_base[p] = run;
      run += tms[p].prefix[F.ph[p].num_species];
      bufs = min(bufs, tms[p].epi_bufs);
    }
    if (NW == 16 && F.epi_direct16) bufs = 0;   // no store staging at all (the plain layer-1 backward has none either)
    for (int p = NP; p <= MAX_PHASES; ++p) phase_base[p] = run;
    // one store-staging depth for all phases (the region is anchored at the end of the dynamic shared memory)
    s_epi_bufs = bufs;
    s_done = 0;
    for (int p = 0; p < NP; ++p)
      tms[p].stages = max(1, min(MAX_STAGES, (AVAIL - bufs * EPI_SETS * EPI_STAGE_BYTES) / tms[p].stage_bytes));
    for (int i = 0; i < MAX_STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], FEPI_WARPS);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&ydep[i], 1);
    fence_barrier_init();
  }
  if (warp == F_MMA_WARP) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int EPI_BUFS = s_epi_bufs;
  unsigned char* epi_stage = smem + AVAIL - EPI_BUFS * EPI_SETS * EPI_STAGE_BYTES;
  const int total_units = phase_base[NP];
  // timing experiments: role 0 producer, 1 MMA, 2 epilogue (clock64 stamps), 3 = what the unit is
  auto stamp = [&](int kloc, int role, int slot, long long v = -1) {
    if (F.trace && blockIdx.x < 4 && kloc < FTRACE_UNITS)
      F.trace[(((size_t)blockIdx.x * FTRACE_UNITS + kloc) * 4 + role) * 4 + slot] = v >= 0 ? v : clock64();
  };
  auto find_phase = [&](int g) {
    int p = 0;
    while (g >= phase_base[p + 1]) ++p;
    return p;
  };
  // arrivals the row tile `rt` is owed by phase d: one per unit of the row tile
  auto owed = [&](int d, int s) { return F.ph[d].members * tms[d].ntn[s]; };

  if (warp == F_PROD_WARP) {
    // ================================ producer (TMA) ================================
    uint32_t stage = 0, empty_par = 0xffffffffu;
    int cur_phase = -1, STAGES = 1, STAGE_BYTES = 0;
    int ahead = 0;   // counter of the NEXT unit's inputs, read (relaxed) while this unit's copies are issued
    bool ahead_valid = false;
    int kloc = -1;
    for (int g = blockIdx.x; g < total_units; g += gridDim.x) {
      ++kloc;
      if (lane == 0) stamp(kloc, 0, 0);
      const int p = find_phase(g);
      const Args& args = F.ph[p];
      const TileMap& tm = tms[p];
      if (p != cur_phase) {
        // new ring geometry: every slot of the old one must have been consumed
        for (int s = 0; s < STAGES && cur_phase >= 0; ++s) mbar_wait(&empty[s], (empty_par >> s) & 1u);
        cur_phase = p;
        STAGES = tm.stages;
        STAGE_BYTES = tm.stage_bytes;
        stage = 0;
      }
      const Tile tl = decode_tile(args, tm, g - phase_base[p]);
      const Species& sp = args.sp[tl.s];
      const int nkb = tm.kb_count >= 0 ? tm.kb_count : (sp.K + TK - 1) / TK;
      const int nkb_all = sp.b_kb_moff ? sp.b_kblocks : (sp.K + TK - 1) / TK;
      const int kb_boff = tl.mem * sp.b_kb_moff;
      const unsigned char* At =
          args.A + ((size_t)tl.rt * args.a_kblocks + (size_t)(tl.mem * sp.a_moff) / TK) * A_BLOCK_BYTES;
      const unsigned char* Bm =
          sp.Bt + (sp.b_kb_moff ? (size_t)0 : (size_t)tl.mem * sp.N * nkb_all * (PARTS * ROW_BYTES));
      const uint32_t b_bytes = (uint32_t)tl.bn * ROW_BYTES;
      const bool dense = tm.nb_count < 0 || args.b_compact;
      // the packed B tile this one lies in
      const int n0p = tl.n0 / TN_MAX * TN_MAX, bnp = min(TN_MAX, (args.b_compact ? tm.n_eff[tl.s] : sp.N) - n0p);
      const int gq = lane / PARTS, gpart = lane % PARTS;
      size_t g_src = 0;
      int g_bns = 0;
      const bool g_active = !dense && lane < PARTS * (tl.bn / 32);
      if (g_active) {
        const int row0 = tm.nb[tl.n0 / 32 + gq] * 32;
        const int n0s = row0 / TN_MAX * TN_MAX;
        g_bns = min(TN_MAX, sp.N - n0s);
        g_src = (size_t)n0s * nkb_all * (PARTS * ROW_BYTES) + (size_t)(row0 - n0s) * ROW_BYTES +
                (size_t)gpart * g_bns * ROW_BYTES;
      }
      // one K-block = A (16 / 24 KB, produced by an earlier phase) + B (weights, never produced in this launch).
      // `arm`: wait for the slot, announce the bytes; `load_b` / `load_a`: the two halves of the transaction
      auto arm = [&](uint32_t st_idx) {
        mbar_wait(&empty[st_idx], (empty_par >> st_idx) & 1u);
        empty_par ^= 1u << st_idx;
        if (lane == 0) mbar_arrive_expect_tx(&full[st_idx], A_BLOCK_BYTES + PARTS * b_bytes);
        __syncwarp();
      };
      auto load_b = [&](uint32_t st_idx, int kb) {
        unsigned char* st = smem + st_idx * STAGE_BYTES;
        const int kbb = (tm.kb_count >= 0 ? tm.kb[kb] : kb) + kb_boff;
        if (dense) {
          if (lane == 0) {
            if (tl.bn == bnp) {  // a whole packed B tile: the pieces are adjacent, one copy
              const unsigned char* bsrc = Bm + ((size_t)tl.n0 * nkb_all + (size_t)kbb * tl.bn) * (PARTS * ROW_BYTES);
              bulk_g2s(st + A_BLOCK_BYTES, bsrc, PARTS * b_bytes, &full[st_idx]);
            } else {             // a column tile inside a packed 256-row B tile: one copy per piece
#pragma unroll
              for (int pc = 0; pc < PARTS; ++pc)
                bulk_g2s(st + A_BLOCK_BYTES + pc * b_bytes,
                         Bm + ((size_t)n0p * nkb_all + (size_t)kbb * bnp) * (PARTS * ROW_BYTES) +
                             (size_t)pc * bnp * ROW_BYTES + (size_t)(tl.n0 - n0p) * ROW_BYTES,
                         b_bytes, &full[st_idx]);
            }
          }
        } else if (g_active) {
          bulk_g2s(st + A_BLOCK_BYTES + gpart * b_bytes + gq * 32 * ROW_BYTES,
                   Bm + g_src + (size_t)kbb * g_bns * (PARTS * ROW_BYTES), 32 * ROW_BYTES, &full[st_idx]);
        }
      };
      auto load_a = [&](uint32_t st_idx, int kb) {
        const int kbi = tm.kb_count >= 0 ? tm.kb[kb] : kb;
        if (lane == 0) bulk_g2s(smem + st_idx * STAGE_BYTES, At + (size_t)kbi * A_BLOCK_BYTES, A_BLOCK_BYTES, &full[st_idx]);
      };
      int kb_first = 0;
      if (F.dep[p] >= 0 && !F.prefetch_b) {
        // data-flow: the A operand of this unit is the output of phase dep[p] for this row tile.  The counter was
        // already looked at one unit ago (relaxed load, latency hidden behind that unit's copies): in the common
        // case it had reached its target and an acquire fence is all that is left to do
        if (lane == 0) {
          const int need = owed(F.dep[p], tl.s);
          if (ahead_valid && ahead >= need)
            __threadfence();
          else
            wait_counter(F.sync + (size_t)F.dep[p] * F.sync_stride + tl.rt, need, args.status);
          fence_proxy_async_all();  // the bulk copies below (async proxy) are ordered after the acquire
        }
        __syncwarp();
      } else if (F.dep[p] >= 0) {
        // data-flow: the A operand of this unit is the output of phase dep[p] for this row tile.  The weights do
        // not depend on anything: their copies of the first K-blocks are in flight while we wait
        const int pre = min(nkb, STAGES);
        for (int kb = 0; kb < pre; ++kb) {
          const uint32_t st_idx = (stage + kb) % (uint32_t)STAGES;
          arm(st_idx);
          load_b(st_idx, kb);
        }
        if (lane == 0) {
          wait_counter(F.sync + (size_t)F.dep[p] * F.sync_stride + tl.rt, owed(F.dep[p], tl.s), args.status);
          fence_proxy_async_all();  // the bulk copies below (async proxy) are ordered after the acquire
        }
        __syncwarp();
        for (int kb = 0; kb < pre; ++kb) load_a((stage + kb) % (uint32_t)STAGES, kb);
        stage = (stage + pre) % (uint32_t)STAGES;
        kb_first = pre;
      }
      if (lane == 0) {
        stamp(kloc, 0, 1);
        // (the producer is never more than two units ahead of the epilogue -- two accumulators -- so four barriers
        // cannot be re-armed before they have been waited on)
        mbar_arrive(&ydep[kloc & 3]);
      }
      // look ahead: the counter the next unit of this CTA will wait on
      ahead_valid = false;
      {
        const int g2 = g + (int)gridDim.x;
        if (lane == 0 && g2 < total_units) {
          const int p2 = find_phase(g2);
          if (F.dep[p2] >= 0) {
            const Tile t2 = decode_tile(F.ph[p2], tms[p2], g2 - phase_base[p2]);
            ahead = ld_relaxed_gpu(F.sync + (size_t)F.dep[p2] * F.sync_stride + t2.rt);
            ahead_valid = true;
          }
        }
      }
      for (int kb = kb_first; kb < nkb; ++kb) {
        arm(stage);
        load_a(stage, kb);
        load_b(stage, kb);
        __syncwarp();
        if (lane == 0 && kb == kb_first) stamp(kloc, 0, 2);
        if (++stage == (uint32_t)STAGES) stage = 0;
      }
      if (lane == 0) stamp(kloc, 0, 3);
    }
  } else if (warp == F_MMA_WARP) {
    // ================================ MMA issuer ================================
    uint32_t stage = 0, full_par = 0u, acc = 0, acc_phase = 0;
    int cur_phase = -1, STAGES = 1, STAGE_BYTES = 0;
    int kloc = -1;
    for (int g = blockIdx.x; g < total_units; g += gridDim.x) {
      ++kloc;
      if (lane == 0) stamp(kloc, 1, 0);
      const int p = find_phase(g);
      const Args& args = F.ph[p];
      const TileMap& tm = tms[p];
      if (p != cur_phase) {
        cur_phase = p;
        STAGES = tm.stages;
        STAGE_BYTES = tm.stage_bytes;
        stage = 0;
      }
      const Tile tl = decode_tile(args, tm, g - phase_base[p]);
      const int nkb = tm.kb_count >= 0 ? tm.kb_count : (args.sp[tl.s].K + TK - 1) / TK;
      const uint32_t idesc = make_idesc(tl.bn, TM);
      const uint32_t b_bytes = (uint32_t)tl.bn * ROW_BYTES;
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      if (lane == 0) stamp(kloc, 1, 1);
      const uint32_t d_tmem = tmem_base + acc * TN_MAX;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&full[stage], (full_par >> stage) & 1u);
        full_par ^= 1u << stage;
        tc_fence_after();
        if (lane == 0 && kb == 0) stamp(kloc, 1, 2);
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sb = sa + A_BLOCK_BYTES;
          const uint64_t a1 = make_desc(sa), a2 = make_desc(sa + A_PART_BYTES);
          const uint64_t b1 = make_desc(sb), b2 = make_desc(sb + b_bytes);
#if !ANI_OPND_FP16X2
          const uint64_t a3 = make_desc(sa + 2 * A_PART_BYTES), b3 = make_desc(sb + 2 * b_bytes);
#endif
#pragma unroll
          for (int k = 0; k < TK / 16; ++k) {
            const uint64_t adv = (uint64_t)(k * 2);
#if ANI_OPND_FP16X2
            umma_f16(d_tmem, a2 + adv, b1 + adv, idesc, (kb | k) != 0);
            umma_f16(d_tmem, a1 + adv, b2 + adv, idesc, 1);
            umma_f16(d_tmem, a1 + adv, b1 + adv, idesc, 1);
#else
            umma_f16(d_tmem, a3 + adv, b1 + adv, idesc, (kb | k) != 0);
            umma_f16(d_tmem, a1 + adv, b3 + adv, idesc, 1);
            umma_f16(d_tmem, a2 + adv, b2 + adv, idesc, 1);
            umma_f16(d_tmem, a2 + adv, b1 + adv, idesc, 1);
            umma_f16(d_tmem, a1 + adv, b2 + adv, idesc, 1);
            umma_f16(d_tmem, a1 + adv, b1 + adv, idesc, 1);
#endif
          }
          umma_commit(&empty[stage]);
          if (kb == nkb - 1) {
            umma_commit(&tfull[acc]);
            stamp(kloc, 1, 3);
          }
        }
        __syncwarp();
        if (++stage == (uint32_t)STAGES) stage = 0;
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else if (warp == F_SIGNAL_WARP) {
    // ================================ completion signals ================================
    // unit k of this CTA is complete when all eight epilogue warps have seen its stores land (s_done >= 8 (k+1));
    // one fence + release-add per unit, off the epilogue's critical path
    if (lane == 0) {
      volatile int* done = &s_done;
      int k = 0;
      for (int g = blockIdx.x; g < total_units; g += gridDim.x, ++k) {
        const int p = find_phase(g);
        const Tile tl = decode_tile(F.ph[p], tms[p], g - phase_base[p]);
        const long long t0 = clock64();
        while (*done < FEPI_WARPS * (k + 1)) {
          __nanosleep(64);
          if (clock64() - t0 > 4 * FUSED_SPIN_LIMIT) break;
        }
        bool consumed = false;   // does any later phase wait on this one?
        for (int q = p + 1; q < NP; ++q) consumed |= F.dep[q] == p;
        if (consumed) {
          __threadfence_block();   // the arrivals in s_done (and with them the epilogue warps' stores) happen-before ...
          fence_proxy_async_all();
          __threadfence();         // ... this gpu-scope fence and the release below
          red_release_gpu(F.sync + (size_t)p * F.sync_stride + tl.rt, 1);
        }
      }
    }
  } else {
    // ================================ epilogue ================================
    uint32_t acc = 0, acc_phase = 0, buf = 0;
    float omax = 0.f;
    const int quad = warp & 3;
    // store staging: one buffer set per warp (NW = 8) or per warp pair (NW = 16: pair id = quadrant * 2 + (w >> 3))
    unsigned char* sb0 = epi_stage + (NW == 16 ? quad * 2 + (warp >> 3) : warp) * (EPI_BUFS * EPI_STAGE_BYTES);
    // completion bookkeeping: the unit whose stores may still be in flight
    bool pending = false;
    int pend_groups = 0;  // bulk groups this warp committed for the pending unit (0: nothing to wait for)
    int pend_phase = -1, pend_rt = -1;
    // complete the pending unit: its bulk stores have landed (all of this lane's groups except the `newer`
    // most recent ones, which belong to the unit that was just finished), then one arrival in shared memory
    auto flush_pending = [&](int newer) {
      if (!pending) return;
      if (lane == 0) {
        if (pend_groups) {
          switch (newer) {
            case 0: bulk_wait_done<0>(); break;
            case 1: bulk_wait_done<1>(); break;
            case 2: bulk_wait_done<2>(); break;
            case 3: bulk_wait_done<3>(); break;
            default: bulk_wait_done<4>(); break;
          }
        }
        __threadfence_block();
        atomicAdd_block(&s_done, 1);   // the signal warp publishes the unit once all eight warps are here
      }
      __syncwarp();
      pending = false;
    };
    int kloc = -1;
    for (int g = blockIdx.x; g < total_units; g += gridDim.x) {
      ++kloc;
      const int p = find_phase(g);
      const Args& args = F.ph[p];
      const TileMap& tm = tms[p];
      const int epi = F.epi[p];
      const Tile tl = decode_tile(args, tm, g - phase_base[p]);
      const Species& sp = args.sp[tl.s];
      const bool direct = NW == 16 && F.epi_direct16 && epi != EPI_PLAIN;   // (tile_epilogue_direct16, gemm_epilogue.cuh)
      if (threadIdx.x == 0) {
        stamp(kloc, 2, 0);
        stamp(kloc, 3, 0, p);
        stamp(kloc, 3, 1, tl.rt);
        stamp(kloc, 3, 2, tl.mem);
        stamp(kloc, 3, 3, tl.bn);
      }
      if (epi == EPI_BIAS_CELU || epi == EPI_HEAD) {
        const int c = threadIdx.x;
        if (c < tl.bn) {
          s_bias[acc][c] = sp.bias[(size_t)tl.mem * sp.bias_mstride + tl.n0 + c] *
                           (direct && epi == EPI_BIAS_CELU ? args.out_scale : 1.0f);
          if (epi == EPI_HEAD) s_w4[acc][c] = sp.w4[(size_t)tl.mem * sp.N + c];
        }
        asm volatile("bar.sync 1, %0;" ::"n"(FEPI_WARPS * 32) : "memory");
      }
      // The pending unit is normally completed one unit late (at the end of this one).  It is completed NOW if this
      // unit waits on it (same row tile, producer phase: otherwise the CTA would wait for itself) or if the
      // accumulator has not shown up after a short grace period: whenever an epilogue warp blocks for real it has
      // published everything it owes, which is what makes the data-flow deadlock-free (a blocked CTA never holds
      // back a completion); the grace period keeps the 1-2 us store-completion wait off the critical path when the
      // main loop is only marginally behind
      if (pending) {
        bool now = pend_phase == F.dep[p] && pend_rt == tl.rt;
        if (!now) {
          now = true;
          for (int spin = 0; spin < 24; ++spin) {
            if (mbar_try(&tfull[acc], acc_phase)) {
              now = false;
              break;
            }
            __nanosleep(64);
          }
        }
        if (now) flush_pending(0);
      }
      // (EPI_MUL_DCELU reads the stored activation an earlier phase wrote for this row tile.  No counter poll here:
      // the producer lane acquired the counter before it issued this unit's copies, and this warp is ordered after
      // that through the full / accumulator mbarriers it waits on -- release.gpu -> acquire.gpu -> CTA-scope
      // synchronisation is a causality chain; the lines have not been in this SM's L1 before.)
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * TN_MAX;
      int groups = 0;
      if (threadIdx.x == 0) stamp(kloc, 2, 1);
      if (NW == 16 && direct) {
        uint64_t* yb = &ydep[kloc & 3];
        const uint32_t yp = (uint32_t)(kloc >> 2) & 1u;
        if (epi == EPI_BIAS_CELU)
          tile_epilogue_direct16_call<EPI_BIAS_CELU>(args, tl, sp, taddr, s_bias[acc], s_w4[acc], e_part, warp, lane, &tfull[acc],
                                                     acc_phase, &omax, yb, yp);
        else if (epi == EPI_MUL_DCELU)
          tile_epilogue_direct16_call<EPI_MUL_DCELU>(args, tl, sp, taddr, s_bias[acc], s_w4[acc], e_part, warp, lane, &tfull[acc],
                                                     acc_phase, &omax, yb, yp);
        else
          tile_epilogue_direct16_call<EPI_HEAD>(args, tl, sp, taddr, s_bias[acc], s_w4[acc], e_part, warp, lane, &tfull[acc],
                                                acc_phase, &omax, yb, yp);
      } else
      switch (epi) {
        case EPI_BIAS_CELU:
          if (NW == 16)
            tile_epilogue16<EPI_BIAS_CELU>(args, tm, tl, sp, taddr, sb0, EPI_BUFS, buf, s_bias[acc], s_w4[acc], e_part, warp, lane,
                                 &tfull[acc], acc_phase, omax, groups);
          else
            tile_epilogue8<EPI_BIAS_CELU>(args, tm, tl, sp, taddr, sb0, EPI_BUFS, buf, s_bias[acc], s_w4[acc], e_part, warp, lane,
                                &tfull[acc], acc_phase, omax, groups);
          break;
        case EPI_MUL_DCELU:
          if (NW == 16)
            tile_epilogue16<EPI_MUL_DCELU>(args, tm, tl, sp, taddr, sb0, EPI_BUFS, buf, s_bias[acc], s_w4[acc], e_part, warp, lane,
                                 &tfull[acc], acc_phase, omax, groups);
          else
            tile_epilogue8<EPI_MUL_DCELU>(args, tm, tl, sp, taddr, sb0, EPI_BUFS, buf, s_bias[acc], s_w4[acc], e_part, warp, lane,
                                &tfull[acc], acc_phase, omax, groups, nullptr, &ydep[kloc & 3], (uint32_t)(kloc >> 2) & 1u);
          break;
        case EPI_HEAD:
          if (NW == 16)
            tile_epilogue16<EPI_HEAD>(args, tm, tl, sp, taddr, sb0, EPI_BUFS, buf, s_bias[acc], s_w4[acc], e_part, warp, lane,
                                 &tfull[acc], acc_phase, omax, groups);
          else
            tile_epilogue8<EPI_HEAD>(args, tm, tl, sp, taddr, sb0, EPI_BUFS, buf, s_bias[acc], s_w4[acc], e_part, warp, lane,
                                &tfull[acc], acc_phase, omax, groups);
          break;
        default:
          if (NW == 16)
            tile_epilogue16<EPI_PLAIN>(args, tm, tl, sp, taddr, sb0, EPI_BUFS, buf, s_bias[acc], s_w4[acc], e_part, warp, lane,
                                 &tfull[acc], acc_phase, omax, groups);
          else
            tile_epilogue8<EPI_PLAIN>(args, tm, tl, sp, taddr, sb0, EPI_BUFS, buf, s_bias[acc], s_w4[acc], e_part, warp, lane,
                                &tfull[acc], acc_phase, omax, groups);
          break;
      }
      // hand the accumulator back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (threadIdx.x == 0) stamp(kloc, 2, 2);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
      // the previous unit's stores were committed a whole tile ago: complete it, then this unit becomes pending
      flush_pending(groups);   // (a tile has at most 8 column groups: 4 per warp)
      pending = true;
      pend_groups = groups;
      pend_phase = p;
      pend_rt = tl.rt;
      // register-direct stores: nothing to wait for here -- the warp arrives at once and the signal warp's gpu-scope
      // fence publishes the stores of all sixteen warps (cumulativity through the CTA-scope synchronisation)
      if (direct) flush_pending(0);
      if (threadIdx.x == 0) stamp(kloc, 2, 3);
    }
    flush_pending(0);
    if (ANI_OPND_FP16X2 && F.ph[0].status && !(omax <= OPND_HALF_MAX)) atomicOr(F.ph[0].status, ANI_STATUS_OPERAND_RANGE);
  }

  // ---- teardown
  if (warp < FEPI_WARPS && lane == 0) bulk_wait_all();
  tc_fence_before();
  __syncthreads();
  if (warp == F_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace tc
}  // namespace ani
