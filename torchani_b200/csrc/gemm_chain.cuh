// The ensemble MLP of a step as INDEPENDENT CHAINS: one persistent launch, no synchronisation between CTAs at all.
//
// Layer k of (row tile r, member m) needs layer k-1 of (r, m) and nothing else -- the six GEMMs of a (row tile, member)
// pair are a private chain (layer 1 reads the shared AEV tile, the layer-1 backward adds its member's share into
// dE/dAEV with vector REDs).  The phase-major data-flow launch of gemm_fused.cuh walks the step layer by layer over
// ALL row tiles: a tile written in layer k is read a whole layer later (by another SM, through device-wide counters),
// the 195 MB of activations of a 10 k-atom step stream through HBM (L2 hit rate 60 %), and a CTA's time is set by the
// slowest of the CTAs it waits for.  Here a CTA owns whole chains and runs D of them interleaved,
//     A.fwd1 B.fwd1 C.fwd1 | A.fwd2 B.fwd2 C.fwd2 | ... | A.bwd1 B.bwd1 C.bwd1 | next D chains,
// so that (i) a unit's input was written by this very CTA D-1 units earlier -- an mbarrier in shared memory is all
// the synchronisation there is, (ii) the epilogue of one chain's unit overlaps the main loop of the next chain's unit
// exactly as before (two accumulators in tensor memory), (iii) the live activations are D x 148 chains x 300 KB,
// which the L2 holds, (iv) short lists (multi-GPU shards, 1 k atoms: fewer chains than SMs) cost six units of
// latency instead of six launches.  Tile code (producer / MMA issue / epilogues) is that of gemm_tc.cuh/gemm_fused.cuh.
//
// Requirements (checked by the launcher): every phase runs per member (members == M; layer-1 weights packed per
// member), one column tile per member in phases 0-4 (h <= 256), any number in the layer-1 backward.
#pragma once
#include "gemm_fused.cuh"

namespace ani {
namespace tc {

constexpr int CHAIN_MAX_D = 8;
constexpr int CHAIN_STAGE_BYTES = A_BLOCK_BYTES + PARTS * TN_MAX * ROW_BYTES;   // one ring geometry for all phases
constexpr int CHAIN_WARPS = NUM_EPI_WARPS + 2;                                   // epilogue x 8, MMA, producer
constexpr int CHAIN_THREADS = CHAIN_WARPS * 32;

struct ChainArgs {
  int n_phases;            // 3 (forward only) or 6
  int depth;               // chains a CTA interleaves (1 .. CHAIN_MAX_D); 0: all of its chains, in equal rounds of <= CHAIN_MAX_D
  int epi[MAX_PHASES];
  long long* trace;        // optional clock64 stamps [cta < 4][unit < FTRACE_UNITS][role 4][4]
  Args ph[MAX_PHASES];
};

// what the n-th unit of a CTA is: rounds of `depth` chains, inside a round unit-major, chain-minor
struct ChainUnit {
  int slot;        // interleave slot (0 .. depth-1) of the chain inside its round
  int chain;       // global chain id = row tile index * M + member
  int u;           // unit of the chain: 0 .. U-1 (phase p = min(u, n_phases-1), column tile = u - p)
  bool first;      // first unit of its chain (no predecessor)
};

struct ChainWalk {
  int G, c, D, U, num_chains;
  int round, u, j;
  __device__ __forceinline__ void init(int G_, int c_, int D_, int U_, int nc) {
    G = G_; c = c_; D = D_; U = U_; num_chains = nc;
    round = 0; u = 0; j = -1;
  }
  // advance to the next unit of this CTA; false when there is none
  __device__ __forceinline__ bool next(ChainUnit& x) {
    while (true) {
      if (++j == D) {
        j = 0;
        if (++u == U) {
          u = 0;
          ++round;
        }
      }
      const long long first_chain = (long long)c + (long long)G * D * round;
      if (first_chain >= num_chains) return false;   // slot 0 of this round has no chain: nothing further
      const long long chain = first_chain + (long long)G * j;
      if (chain >= num_chains) continue;            // this slot is empty in the last round
      x.slot = j;
      x.chain = (int)chain;
      x.u = u;
      x.first = u == 0;
      return true;
    }
  }
};

template <int NW>
__global__ void __launch_bounds__(CHAIN_THREADS, 1) k_mlp_chain(const __grid_constant__ ChainArgs F) {
  static_assert(NW == NUM_EPI_WARPS, "eight epilogue warps");
  constexpr int C_MMA_WARP = NW, C_PROD_WARP = NW + 1;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ TileMap tms[MAX_PHASES];
  __shared__ int s_total_rt, s_units_per_chain, s_stages, s_epi_bufs;
  __shared__ float e_part[NW * 32];
  __shared__ __align__(16) float s_bias[2][TN_MAX];
  __shared__ __align__(16) float s_w4[2][TN_MAX];
  __shared__ __align__(8) uint64_t bars[2 * MAX_STAGES + 5 + CHAIN_MAX_D];
  uint64_t* full = bars;
  uint64_t* empty = bars + MAX_STAGES;
  uint64_t* tfull = bars + 2 * MAX_STAGES;
  uint64_t* tempty = bars + 2 * MAX_STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 4);
  uint64_t* cdone = bars + 2 * MAX_STAGES + 5;    // [depth]: "the previous unit of this slot's chain is in global memory"
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NP = F.n_phases;
  constexpr int AVAIL = FUSED_SMEM_BYTES - 1024;

  if (lane == 0 && warp < NP) build_tile_map(F.ph[warp], tms[warp]);   // (side by side, see k_mlp_fused)
  __syncthreads();
  if (threadIdx.x == 0) {
    int bufs = 2;
    for (int p = 0; p < NP; ++p)
      if (F.epi[p] != EPI_PLAIN) bufs = min(bufs, tms[p].epi_bufs);
    // one ring geometry for all phases (a CTA changes phase with every unit): stages of 16 KB + 256 rows of B
    bufs = min(bufs, (AVAIL - 2 * CHAIN_STAGE_BYTES) / (NW * EPI_STAGE_BYTES) >= 2 ? 2 : 1);
    s_epi_bufs = bufs;
    s_stages = max(1, min(MAX_STAGES, (AVAIL - bufs * NW * EPI_STAGE_BYTES) / CHAIN_STAGE_BYTES));
    s_total_rt = F.ph[0].layout_info[4 + F.ph[0].num_species];
    // units of a chain: one per phase, except the last phase of a backward step (live column tiles of dE/dAEV)
    int last_ntn = 1;
    for (int s = 0; s < F.ph[0].num_species; ++s)
      if (tms[NP - 1].cnt_rt[s] > 0) last_ntn = max(last_ntn, tms[NP - 1].ntn[s]);
    s_units_per_chain = NP - 1 + last_ntn;
    for (int i = 0; i < MAX_STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], NW);
    }
    for (int i = 0; i < CHAIN_MAX_D; ++i) mbar_init(&cdone[i], NW);
    fence_barrier_init();
  }
  if (warp == C_MMA_WARP) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int STAGES = s_stages, EPI_BUFS = s_epi_bufs;
  unsigned char* epi_stage = smem + AVAIL - EPI_BUFS * NW * EPI_STAGE_BYTES;
  const int M = F.ph[0].members;
  const int num_chains = s_total_rt * M, U = s_units_per_chain;
  // interleave depth: the chains of a CTA (at most ceil(chains / CTAs)) in as few, equally deep rounds as possible -- a
  // round with a single chain left runs its units back to back with every latency exposed
  int D = F.depth;
  if (D < 1) {
    const int per_cta = (num_chains + (int)gridDim.x - 1) / (int)gridDim.x;
    const int rounds = (per_cta + CHAIN_MAX_D - 1) / CHAIN_MAX_D;
    D = rounds > 0 ? (per_cta + rounds - 1) / rounds : 1;
  }
  D = max(1, min(D, CHAIN_MAX_D));
  auto stamp = [&](int kloc, int role, int slot, long long v = -1) {
    if (F.trace && blockIdx.x < 4 && kloc < FTRACE_UNITS)
      F.trace[(((size_t)blockIdx.x * FTRACE_UNITS + kloc) * 4 + role) * 4 + slot] = v >= 0 ? v : clock64();
  };
  // unit -> (phase, tile); valid == false: this chain has no such column tile (species with fewer live tiles)
  auto decode = [&](const ChainUnit& cu, int& p, Tile& tl) {
    p = min(cu.u, NP - 1);
    const int nt = cu.u - p;
    const int rt = cu.chain / M;
    const TileMap& tm = tms[p];
    int s = 0;
    while (s + 1 < F.ph[p].num_species && rt >= tm.first_rt[s] + tm.cnt_rt[s]) ++s;
    tl.s = s;
    tl.rt = rt;
    tl.mem = cu.chain - rt * M;
    tl.rt_last = tm.first_rt[s] + tm.cnt_rt[s] - 1;
    tl.n0 = nt * tm.tn[s];
    tl.bn = min(tm.tn[s], tm.n_eff[s] - tl.n0);
    return nt < tm.ntn[s] && rt >= tm.first_rt[s];
  };

  if (warp == C_PROD_WARP) {
    // ================================ producer (TMA) ================================
    uint32_t stage = 0, empty_par = 0xffffffffu, cdone_par = 0;
    ChainWalk w;
    w.init(gridDim.x, blockIdx.x, D, U, num_chains);
    ChainUnit cu;
    int kloc = -1;
    unsigned seen = 0;   // slots that have had a unit (a slot's first unit ever has nothing to wait for)
    while (w.next(cu)) {
      ++kloc;
      int p;
      Tile tl;
      const bool valid = decode(cu, p, tl);
      if (lane == 0) stamp(kloc, 0, 0);
      // every unit of a slot but the very first waits for the slot's previous unit: its outputs (this unit's A, or the
      // stored activation its epilogue reads) are complete in global memory -- or, across chains, simply consumed
      if ((seen >> cu.slot) & 1u) {
        mbar_wait(&cdone[cu.slot], (cdone_par >> cu.slot) & 1u);
        cdone_par ^= 1u << cu.slot;
      }
      seen |= 1u << cu.slot;
      if (lane == 0) stamp(kloc, 0, 1);
      if (!valid) continue;
      const Args& args = F.ph[p];
      const TileMap& tm = tms[p];
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
      if (lane == 0) fence_proxy_async_all();
      __syncwarp();
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&empty[stage], (empty_par >> stage) & 1u);
        empty_par ^= 1u << stage;
        unsigned char* st = smem + stage * CHAIN_STAGE_BYTES;
        const int kbi = tm.kb_count >= 0 ? tm.kb[kb] : kb;
        const int kbb = kbi + kb_boff;
        if (lane == 0) {
          mbar_arrive_expect_tx(&full[stage], A_BLOCK_BYTES + PARTS * b_bytes);
          bulk_g2s(st, At + (size_t)kbi * A_BLOCK_BYTES, A_BLOCK_BYTES, &full[stage]);
          if (dense) {
            if (tl.bn == bnp) {
              const unsigned char* bsrc = Bm + ((size_t)tl.n0 * nkb_all + (size_t)kbb * tl.bn) * (PARTS * ROW_BYTES);
              bulk_g2s(st + A_BLOCK_BYTES, bsrc, PARTS * b_bytes, &full[stage]);
            } else {
#pragma unroll
              for (int pc = 0; pc < PARTS; ++pc)
                bulk_g2s(st + A_BLOCK_BYTES + pc * b_bytes,
                         Bm + ((size_t)n0p * nkb_all + (size_t)kbb * bnp) * (PARTS * ROW_BYTES) +
                             (size_t)pc * bnp * ROW_BYTES + (size_t)(tl.n0 - n0p) * ROW_BYTES,
                         b_bytes, &full[stage]);
            }
          }
        }
        __syncwarp();
        if (g_active)
          bulk_g2s(st + A_BLOCK_BYTES + gpart * b_bytes + gq * 32 * ROW_BYTES,
                   Bm + g_src + (size_t)kbb * g_bns * (PARTS * ROW_BYTES), 32 * ROW_BYTES, &full[stage]);
        __syncwarp();
        if (lane == 0 && kb == 0) stamp(kloc, 0, 2);
        if (++stage == (uint32_t)STAGES) stage = 0;
      }
      if (lane == 0) stamp(kloc, 0, 3);
    }
  } else if (warp == C_MMA_WARP) {
    // ================================ MMA issuer ================================
    uint32_t stage = 0, full_par = 0u, acc = 0, acc_phase = 0;
    ChainWalk w;
    w.init(gridDim.x, blockIdx.x, D, U, num_chains);
    ChainUnit cu;
    int kloc = -1;
    while (w.next(cu)) {
      ++kloc;
      int p;
      Tile tl;
      if (!decode(cu, p, tl)) continue;
      const Args& args = F.ph[p];
      const TileMap& tm = tms[p];
      const int nkb = tm.kb_count >= 0 ? tm.kb_count : (args.sp[tl.s].K + TK - 1) / TK;
      const uint32_t idesc = make_idesc(tl.bn, TM);
      const uint32_t b_bytes = (uint32_t)tl.bn * ROW_BYTES;
      if (lane == 0) stamp(kloc, 1, 0);
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
          const uint32_t sa = smem_u32(smem + stage * CHAIN_STAGE_BYTES);
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
  } else {
    // ================================ epilogue ================================
    uint32_t acc = 0, acc_phase = 0, buf = 0;
    float omax = 0.f;
    const int quad = warp & 3;
    unsigned char* sb0 = epi_stage + warp * (EPI_BUFS * EPI_STAGE_BYTES);
    // the unit whose bulk stores may still be in flight: it is published (one arrival of this warp on its slot's
    // barrier) once they have fully completed
    bool pending = false;
    int pend_groups = 0, pend_slot = 0;
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
        mbar_arrive(&cdone[pend_slot]);   // (release at CTA scope: the producer lane acquires it before its bulk loads)
      }
      __syncwarp();
      pending = false;
    };
    ChainWalk w;
    w.init(gridDim.x, blockIdx.x, D, U, num_chains);
    ChainUnit cu;
    int kloc = -1;
    while (w.next(cu)) {
      ++kloc;
      int p;
      Tile tl;
      const bool valid = decode(cu, p, tl);
      if (!valid) {
        // nothing to compute, but the slot's barrier sees one arrival per unit and warp
        flush_pending(0);
        if (lane == 0) mbar_arrive(&cdone[cu.slot]);
        __syncwarp();
        continue;
      }
      const Args& args = F.ph[p];
      const TileMap& tm = tms[p];
      const int epi = F.epi[p];
      const Species& sp = args.sp[tl.s];
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
          s_bias[acc][c] = sp.bias[(size_t)tl.mem * sp.bias_mstride + tl.n0 + c];
          if (epi == EPI_HEAD) s_w4[acc][c] = sp.w4[(size_t)tl.mem * sp.N + c];
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory");
      }
      // The previous unit of this warp (stores possibly still in flight) is published as early as it can be without
      // idling: inside this tile's epilogue, right after its first store group has been committed (by then the older
      // stores have had the accumulator wait + one column group of math to land).  At once instead when this very unit
      // follows it in the same chain (the producer is waiting for it before it can feed the MMAs this warp waits for),
      // or when this warp commits no store group in this tile.
      uint64_t* publish = nullptr;
      if (pending) {
        const bool stores_here = epi != EPI_PLAIN && (epi != EPI_HEAD || args.want_backward) && (warp >> 2) < tl.bn / 32;
        if (pend_slot == cu.slot || !stores_here || pend_groups == 0)
          flush_pending(0);
        else
          publish = &cdone[pend_slot];
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * TN_MAX;
      int groups = 0;
      if (threadIdx.x == 0) stamp(kloc, 2, 1);
      switch (epi) {
        case EPI_BIAS_CELU:
          tile_epilogue8<EPI_BIAS_CELU>(args, tm, tl, sp, taddr, sb0, EPI_BUFS, buf, s_bias[acc], s_w4[acc], e_part, warp, lane,
                                        &tfull[acc], acc_phase, omax, groups, publish);
          break;
        case EPI_MUL_DCELU:
          tile_epilogue8<EPI_MUL_DCELU>(args, tm, tl, sp, taddr, sb0, EPI_BUFS, buf, s_bias[acc], s_w4[acc], e_part, warp, lane,
                                        &tfull[acc], acc_phase, omax, groups, publish);
          break;
        case EPI_HEAD:
          tile_epilogue8<EPI_HEAD>(args, tm, tl, sp, taddr, sb0, EPI_BUFS, buf, s_bias[acc], s_w4[acc], e_part, warp, lane,
                                   &tfull[acc], acc_phase, omax, groups, publish);
          break;
        default:
          tile_epilogue8<EPI_PLAIN>(args, tm, tl, sp, taddr, sb0, EPI_BUFS, buf, s_bias[acc], s_w4[acc], e_part, warp, lane,
                                    &tfull[acc], acc_phase, omax, groups, publish);
          break;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (threadIdx.x == 0) stamp(kloc, 2, 2);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
      if (publish) pending = false;   // published inside the tile's epilogue
      flush_pending(groups);
      pending = true;
      pend_groups = groups;
      pend_slot = cu.slot;
      if (threadIdx.x == 0) stamp(kloc, 2, 3);
    }
    flush_pending(0);
    if (ANI_OPND_FP16X2 && F.ph[0].status && !(omax <= OPND_HALF_MAX)) atomicOr(F.ph[0].status, ANI_STATUS_OPERAND_RANGE);
  }

  // ---- teardown
  if (warp < NW && lane == 0) bulk_wait_all();
  tc_fence_before();
  __syncthreads();
  if (warp == C_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace tc
}  // namespace ani
