// Persistent, warp-specialised tcgen05 grouped GEMM for the ensemble MLP (sm_100a).
//
//   C[128-row tile, bn] = epilogue( A[128, K] x B[bn, K]^T )        both operands K-major fp32
//
// fp32 accuracy on the tensor cores ("3xTF32"): every fp32 operand is split into
// hi = x & 0xffffe000 (exact TF32) and lo = x - hi (exact in fp32); the MMA thread issues three
// kind::tf32 products per K-step (lo*hi, hi*lo, hi*hi) into one fp32 accumulator in tensor
// memory.  The dropped lo*lo term is ~2^-22 relative.
//   * B (weights) is split, tiled and swizzled ONCE at model-pack time (ani_b200.h, "tiled B
//     operand"); a K-block of a tile is two contiguous byte ranges that one thread moves with
//     cp.async.bulk (TMA) straight into the SWIZZLE_128B shared-memory layout.
//   * A (activations) is split on the fly by the producer warps: coalesced 16-byte loads, the
//     loads of K-block k+1 are in flight while K-block k is split and stored.
//
// Roles (9 warps, one CTA per SM, persistent over the device-side tile list):
//   warps 0-3  epilogue : tcgen05.ld accumulator rows (warp w owns TMEM lanes 32w..32w+31);
//                         global traffic is staged through a 32x32 shared-memory transpose so
//                         that loads (old activation for CELU') and stores are 128-byte rows
//   warp  4    MMA      : TMEM alloc, one lane issues tcgen05.mma / tcgen05.commit
//   warps 5-8  producer : A split + swizzled st.shared + fence.proxy.async; thread 0 also
//                         issues the bulk copies of B (mbarrier expect_tx / complete_tx)
// Pipelines: smem full/empty (4 stages x 48 KB: three K-blocks in flight cover the L2/HBM latency
// of the one being multiplied) and TMEM full/empty (2 x 256 columns), so the
// epilogue of tile i overlaps the main loop of tile i+1.
#pragma once
#include "common.cuh"

namespace ani {
namespace tc {

constexpr int TM = ANI_TILE_ROWS;        // 128 rows per tile == UMMA M
constexpr int TN_MAX = 256;              // UMMA N (columns of one accumulator)
constexpr int TK = 16;                   // fp32 per K-block = one 64-byte swizzle row (SWIZZLE_64B)
constexpr int ROW_BYTES = TK * 4;        // 64
constexpr int GROUP_BYTES = 8 * ROW_BYTES;       // 8-row swizzle group = 512 B (descriptor SBO)
constexpr int STAGES = 4;
constexpr int A_TILE_BYTES = TM * ROW_BYTES;     // 8 KB
constexpr int B_TILE_BYTES = TN_MAX * ROW_BYTES; // 16 KB
constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;  // hi+lo of A and B = 48 KB
constexpr int EPI_LD = 36;                                  // padded row of the 32x32 transpose buffer (floats)
constexpr int EPI_BYTES = 4 * 32 * EPI_LD * 4;               // one buffer per epilogue warp
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int NUM_EPI_WARPS = 4, MMA_WARP = 4, FIRST_PROD_WARP = 5, NUM_PROD_WARPS = 4;
constexpr int NPT = NUM_PROD_WARPS * 32;  // producer threads
constexpr int THREADS = (NUM_EPI_WARPS + 1 + NUM_PROD_WARPS) * 32;  // 288
constexpr int TMEM_COLS = 512;

enum { EPI_BIAS_CELU = 0, EPI_MUL_DCELU = 1, EPI_PLAIN = 2, EPI_HEAD = 3 };

struct Species {
  const float* Bt;    // tiled B operand (hi/lo split, swizzled), see ani_b200.h
  const float* bias;  // [N] (+ member * bias_mstride) or nullptr
  int K, N;
  int a_moff, c_moff, bias_mstride;
  const float* w4;    // EPI_HEAD: final layer weights [M][N] and biases [M]
  const float* b4;
};

struct Args {
  const float* A;
  float* C;
  int lda, ldc;
  int members;                  // GEMMs per row tile (grid z of the SIMT version)
  const int32_t* layout_info;   // [4 + S + 1]: ..., first row tile of species s, total row tiles
  const int32_t* kblocks;       // optional list of live 32-wide K-blocks: [count, ids...]  (layer-1 forward)
  const int32_t* nblocks;       // optional list of live 32-wide column blocks of C          (layer-1 backward)
  int num_species;
  float alpha;
  // EPI_HEAD (layer 3 + final layer + gradient seed fused in the epilogue)
  float* e_member;              // [M][rows_cap]
  const int32_t* row_atom;      // [rows_cap], -1 for padding rows
  int rows_cap;
  int want_backward;
  int debug;                    // bring-up only (ANI_B200_GEMM_DEBUG): 1 no A stores, 2 no B copies, 4 no MMA, 8 no epilogue, 16 no A loads
  float member_scale[ANI_MAX_MEMBERS];
  Species sp[ANI_MAX_SPECIES];
};

// ---- PTX wrappers -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared, completion signalled on the mbarrier (complete_tx::bytes)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc], kind::tf32, single CTA
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued MMAs of this thread arrive on the mbarrier when they complete
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 16 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor: K-major, SWIZZLE_64B, 8-row groups 512 B apart
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);   // start address
  d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(GROUP_BYTES >> 4) << 32;       // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
  d |= (uint64_t)4 << 61;                        // SWIZZLE_64B
  return d;
}
// byte offset of 16-byte chunk `ch` (0..3) of row `row` inside a K-major SWIZZLE_64B tile
__device__ __forceinline__ uint32_t swz_off(int row, int ch) {
  return (uint32_t)(row >> 3) * GROUP_BYTES + (uint32_t)(row & 7) * ROW_BYTES + (uint32_t)((ch ^ ((row >> 1) & 3)) << 4);
}
// instruction descriptor: D=f32, A=B=tf32, both K-major, M=128, N=bn
__device__ __forceinline__ uint32_t make_idesc(int bn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}

// CELU(x) = max(0,x) + min(0, alpha*(exp(x/alpha)-1)) with exp via ex2.approx (rel. error ~2^-22:
// the negative branch is bounded by alpha, so the absolute error is < 1e-8 for alpha = 0.1)
struct CeluConst {
  float alpha, inv_alpha, inv_alpha_log2e;
};
__device__ __forceinline__ float celu(float x, const CeluConst& c) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * c.inv_alpha_log2e));
  return x > 0.f ? x : fmaf(c.alpha, e, -c.alpha);
}
__device__ __forceinline__ float dcelu_from_out(float y, const CeluConst& c) {
  return y > 0.f ? 1.0f : fmaf(y, c.inv_alpha, 1.0f);
}

// ---- tile enumeration -----------------------------------------------------------------------
// Row tiles of one species are contiguous; tile t -> (species, row tile, member, n0, bn).
constexpr int MAX_BLOCKS = 64;         // ldx / 32 <= 64
struct TileMap {
  int first_rt[ANI_MAX_SPECIES + 1];   // first row tile of each species (+ total)
  int ntn[ANI_MAX_SPECIES];            // N tiles per (row tile, member)
  int prefix[ANI_MAX_SPECIES + 1];     // exclusive prefix of tile counts
  int n_eff[ANI_MAX_SPECIES];          // columns actually computed (compacted when nblocks is given)
  int kb_count, nb_count;              // live K-blocks (-1: dense) / live column blocks (-1: dense)
  int kb[MAX_BLOCKS], nb[MAX_BLOCKS];
};

struct Tile {
  int s, rt, mem, n0, bn;
};

__device__ __forceinline__ void build_tile_map(const Args& a, TileMap& tm) {
  const int S = a.num_species;
  int run = 0;
  tm.kb_count = tm.nb_count = -1;
  if (a.kblocks) {
    tm.kb_count = min(a.kblocks[0], MAX_BLOCKS);
    for (int i = 0; i < tm.kb_count; ++i) tm.kb[i] = a.kblocks[1 + i];
  }
  if (a.nblocks) {
    tm.nb_count = min(a.nblocks[0], MAX_BLOCKS);
    for (int i = 0; i < tm.nb_count; ++i) tm.nb[i] = a.nblocks[1 + i];
  }
  for (int s = 0; s < S; ++s) {
    tm.first_rt[s] = a.layout_info[4 + s];
    tm.n_eff[s] = tm.nb_count >= 0 ? tm.nb_count * 32 : a.sp[s].N;
    tm.ntn[s] = (tm.n_eff[s] + TN_MAX - 1) / TN_MAX;
  }
  tm.first_rt[S] = a.layout_info[4 + S];
  for (int s = 0; s < S; ++s) {
    tm.prefix[s] = run;
    run += (tm.first_rt[s + 1] - tm.first_rt[s]) * a.members * tm.ntn[s];
  }
  tm.prefix[S] = run;
}

__device__ __forceinline__ Tile decode_tile(const Args& a, const TileMap& tm, int t) {
  Tile x;
  int s = 0;
  while (t >= tm.prefix[s + 1]) ++s;
  const int local = t - tm.prefix[s];
  const int ntn = tm.ntn[s];
  const int nt = local % ntn;
  const int rm = local / ntn;
  x.s = s;
  x.mem = rm % a.members;
  x.rt = tm.first_rt[s] + rm / a.members;
  x.n0 = nt * TN_MAX;
  x.bn = min(TN_MAX, tm.n_eff[s] - x.n0);
  return x;
}

// ---- the kernel -----------------------------------------------------------------------------
// position of a producer in the flattened (tile, K-block) sequence of its CTA
struct KItem {
  int t, kb, nkb, K;
  Tile tl;
  const float* A;
  bool valid;
};

template <int EPI>
__global__ void __launch_bounds__(THREADS, 1) k_gemm_tc(const __grid_constant__ Args args) {
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte aligned operand tiles (swizzle groups are 8 rows x 64 B)
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* epi_buf = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + EPI_BYTES);
  uint64_t* full = bars;                     // [STAGES]  producers (+ TMA bytes) -> MMA
  uint64_t* empty = bars + STAGES;           // [STAGES]  MMA (commit) -> producers
  uint64_t* tfull = bars + 2 * STAGES;       // [2]       MMA (commit) -> epilogue
  uint64_t* tempty = bars + 2 * STAGES + 2;  // [2]       epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  __shared__ TileMap tm;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    build_tile_map(args, tm);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], NUM_PROD_WARPS + 1);  // one arrival per producer warp + the expect_tx arrival
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], NUM_EPI_WARPS * 32);
    }
    fence_barrier_init();
  }
  if (warp == MMA_WARP) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int total_tiles = tm.prefix[args.num_species];
  // K-blocks are 16 floats; the optional live-block lists are in 32-column AEV blocks
  auto num_kb = [&](int K) { return tm.kb_count >= 0 ? 2 * tm.kb_count : (K + TK - 1) / TK; };
  auto kb_id = [&](int i) { return tm.kb_count >= 0 ? tm.kb[i >> 1] * 2 + (i & 1) : i; };

  if (warp >= FIRST_PROD_WARP) {
    // ================================ producers ================================
    // Thread pt owns 16-byte chunk (pt & 3) of rows (pt >> 2) + 32*i, i = 0..3, of every K-block:
    // its global pointer advances by k0 only and its four swizzled smem offsets are constants.
    const int pt = threadIdx.x - FIRST_PROD_WARP * 32;  // 0..NPT-1
    constexpr int A_IT = (TM * 4) / NPT;                // chunks per thread per K-block (4)
    const int my_row = pt >> 2, my_ch = pt & 3;
    const uint32_t my_off = swz_off(my_row, my_ch);     // + i * 32 rows = + i * 4 groups = + i * 2048 B
    const size_t row_stride = (size_t)32 * args.lda;

    auto set_tile = [&](KItem& it) {
      it.valid = it.t < total_tiles;
      if (it.valid) {
        it.tl = decode_tile(args, tm, it.t);
        it.K = args.sp[it.tl.s].K;
        it.nkb = num_kb(it.K);
        it.A = args.A + (size_t)it.tl.rt * TM * args.lda + (size_t)it.tl.mem * args.sp[it.tl.s].a_moff +
               (size_t)my_row * args.lda + my_ch * 4;
        it.kb = 0;
      }
    };
    auto advance = [&](KItem& it) {
      if (++it.kb >= it.nkb && it.valid) {
        it.t += gridDim.x;
        set_tile(it);
      }
    };
    auto load_a = [&](float4 (&dst)[A_IT], const KItem& it) {
      if (!it.valid) return;
      const int k0 = kb_id(it.kb) * TK;
      const bool in_k = (k0 + my_ch * 4 < it.K) && !(args.debug & 16);
      const float* p = it.A + k0;
#pragma unroll
      for (int i = 0; i < A_IT; ++i)
        dst[i] = in_k ? *reinterpret_cast<const float4*>(p + i * row_stride) : make_float4(0.f, 0.f, 0.f, 0.f);
    };

    KItem cur, pf;
    cur.t = blockIdx.x;
    set_tile(cur);
    pf = cur;
    // register ring with STATIC slots: slot k holds the A chunks of the K-block that is RING
    // iterations ahead; it is consumed (split + stored) and immediately refilled.  No register of an
    // in-flight load is ever moved, so RING K-blocks of HBM/L2 latency really overlap.
    constexpr int RING = 6;
    float4 ring[RING][A_IT];
#pragma unroll
    for (int d = 0; d < RING; ++d) {
      load_a(ring[d], pf);
      advance(pf);
    }
    uint32_t stage = 0, phase = 0;
    while (cur.valid) {
#pragma unroll
      for (int k = 0; k < RING; ++k) {
        if (cur.valid) {
          mbar_wait(&empty[stage], phase ^ 1);
          unsigned char* st = smem + stage * STAGE_BYTES;
          if (pt < 16) {
            // B: TMA bulk copies, one per thread (dense: hi, lo; gathered column blocks: 2 per block)
            const Species& sp = args.sp[cur.tl.s];
            const uint32_t b_bytes = (uint32_t)cur.tl.bn * ROW_BYTES;
            if (pt == 0) mbar_arrive_expect_tx(&full[stage], (args.debug & 2) ? 0u : 2 * b_bytes);
            const bool dense = tm.nb_count < 0;
            if (!(args.debug & 2) && pt < (dense ? 2 : 2 * (cur.tl.bn / 32))) {
              const int nkb_all = (cur.K + TK - 1) / TK;  // K-blocks of the stored operand
              const unsigned char* Bm = reinterpret_cast<const unsigned char*>(sp.Bt) +
                                        (size_t)cur.tl.mem * sp.N * nkb_all * (2 * ROW_BYTES);
              const int kbi = kb_id(cur.kb);
              if (dense) {
                // [member][n tile][k block][hi bn x 64 B | lo bn x 64 B]
                const unsigned char* src = Bm +
                                           ((size_t)cur.tl.n0 * nkb_all + (size_t)kbi * cur.tl.bn) * (2 * ROW_BYTES) +
                                           (size_t)pt * b_bytes;
                bulk_g2s(st + 2 * A_TILE_BYTES + pt * B_TILE_BYTES, src, b_bytes, &full[stage]);
              } else {
                // gathered column blocks: 32 rows (2 KB) of the stored operand per live block and part
                const int q = pt >> 1, part = pt & 1;
                const int row0 = tm.nb[cur.tl.n0 / 32 + q] * 32;
                const int n0s = row0 / TN_MAX * TN_MAX;
                const int bns = min(TN_MAX, sp.N - n0s);
                const unsigned char* src = Bm + ((size_t)n0s * nkb_all + (size_t)kbi * bns) * (2 * ROW_BYTES) +
                                           (size_t)(row0 - n0s) * ROW_BYTES + (size_t)part * bns * ROW_BYTES;
                bulk_g2s(st + 2 * A_TILE_BYTES + part * B_TILE_BYTES + q * 32 * ROW_BYTES, src, 32 * ROW_BYTES,
                         &full[stage]);
              }
            }
          }
          if (!(args.debug & 1)) {
#pragma unroll
            for (int i = 0; i < A_IT; ++i) {
              const float4 v = ring[k][i];
              float4 hi, lo;
              hi.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
              hi.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
              hi.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
              hi.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
              lo.x = v.x - hi.x;
              lo.y = v.y - hi.y;
              lo.z = v.z - hi.z;
              lo.w = v.w - hi.w;
              *reinterpret_cast<float4*>(st + my_off + i * (4 * GROUP_BYTES)) = hi;
              *reinterpret_cast<float4*>(st + A_TILE_BYTES + my_off + i * (4 * GROUP_BYTES)) = lo;
            }
          }
          fence_proxy_async();  // generic-proxy stores -> visible to the tensor core (async proxy)
          __syncwarp();
          if (lane == 0) mbar_arrive(&full[stage]);
          load_a(ring[k], pf);  // refill this slot for the K-block RING iterations ahead
          advance(pf);
          advance(cur);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ================================ MMA issuer ================================
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const Tile tl = decode_tile(args, tm, t);
      const int nkb = num_kb(args.sp[tl.s].K);
      const uint32_t idesc = make_idesc(tl.bn);
      mbar_wait(&tempty[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * TN_MAX;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t a_hi = make_desc(sa), a_lo = make_desc(sa + A_TILE_BYTES);
          const uint64_t b_hi = make_desc(sa + 2 * A_TILE_BYTES), b_lo = make_desc(sa + 2 * A_TILE_BYTES + B_TILE_BYTES);
#pragma unroll
          for (int k = 0; k < TK / 8; ++k) {
            if (args.debug & 4) break;
            const uint64_t adv = (uint64_t)(k * 2);  // 8 tf32 = 32 B = 2 x 16 B along the swizzle row
            umma_tf32(d_tmem, a_lo + adv, b_hi + adv, idesc, (kb | k) != 0);
            umma_tf32(d_tmem, a_hi + adv, b_lo + adv, idesc, 1);
            umma_tf32(d_tmem, a_hi + adv, b_hi + adv, idesc, 1);
          }
          umma_commit(&empty[stage]);                   // smem slot free once these MMAs retire
          if (kb == nkb - 1) umma_commit(&tfull[acc]);  // accumulator complete
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else {
    // ================================ epilogue ================================
    // Per 32-column chunk: (CELU' only) the stored activation arrives through a coalesced,
    // one-chunk-ahead prefetch into the warp's 32x32 buffer; tcgen05.ld (thread = row);
    // elementwise op; result back into the buffer; coalesced 128-byte-row stores.
    uint32_t acc = 0, acc_phase = 0;
    const CeluConst cc{args.alpha, 1.0f / args.alpha, 1.4426950408889634f / args.alpha};
    float* buf = epi_buf + warp * 32 * EPI_LD;
    const int cr = lane >> 3, cq = (lane & 7) * 4;  // coalesced phase: rows cr + 4*i, columns cq..cq+3
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const Tile tl = decode_tile(args, tm, t);
      const Species& sp = args.sp[tl.s];
      float* __restrict__ crow0 = args.C + (size_t)(tl.rt * TM + warp * 32) * args.ldc + (size_t)tl.mem * sp.c_moff;
      const float* __restrict__ bias = (EPI == EPI_BIAS_CELU || EPI == EPI_HEAD)
                                           ? sp.bias + (size_t)tl.mem * sp.bias_mstride + tl.n0
                                           : nullptr;
      // EPI_HEAD: this thread's row produces one atomic energy e = w4 . celu(z3) + b4
      const float* __restrict__ w4 = (EPI == EPI_HEAD) ? sp.w4 + (size_t)tl.mem * sp.N : nullptr;
      const int my_row = tl.rt * TM + warp * 32 + lane;
      float e_acc = 0.f, seed = 0.f;
      bool row_valid = false;
      if (EPI == EPI_HEAD) {
        row_valid = args.row_atom[my_row] >= 0;
        seed = row_valid ? args.member_scale[tl.mem] : 0.f;
      }
      // first column of chunk c0 in C (compacted column blocks map back to their place)
      auto chunk_ptr = [&](int c0) {
        return crow0 + (tm.nb_count >= 0 ? tm.nb[(tl.n0 + c0) / 32] * 32 : tl.n0 + c0);
      };
      float4 yreg[8];
      auto load_y = [&](int c0) {
        const int ncol = min(32, tl.bn - c0);
        if (cq < ncol) {
          const float* p = chunk_ptr(c0);
#pragma unroll
          for (int i = 0; i < 8; ++i) yreg[i] = *reinterpret_cast<const float4*>(p + (size_t)(cr + 4 * i) * args.ldc + cq);
        }
      };
      if (EPI == EPI_MUL_DCELU) load_y(0);  // independent of the accumulator: overlaps the MMA wait
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + acc * TN_MAX;
      for (int c0 = 0; c0 < tl.bn; c0 += 32) {
        if (args.debug & 8) break;
        const int ncol = min(32, tl.bn - c0);  // 32 or 16 (bn is a multiple of 16)
        float* __restrict__ cbase = chunk_ptr(c0);
        if (EPI == EPI_MUL_DCELU) {
          if (cq < ncol) {
#pragma unroll
            for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(&buf[(cr + 4 * i) * EPI_LD + cq]) = yreg[i];
          }
          __syncwarp();
          if (c0 + 32 < tl.bn) load_y(c0 + 32);
        }
        float v[32];
        {
          float lo16[16];
          tmem_ld16(taddr + c0, lo16);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = lo16[j];
          if (ncol > 16) {
            float hi16[16];
            tmem_ld16(taddr + c0 + 16, hi16);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[16 + j] = hi16[j];
          }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (4 * q < ncol) {
            float4 o = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            float4* slot = reinterpret_cast<float4*>(&buf[lane * EPI_LD + 4 * q]);
            if (EPI == EPI_BIAS_CELU) {
              const float4 b = *reinterpret_cast<const float4*>(bias + c0 + 4 * q);
              o.x = celu(o.x + b.x, cc);
              o.y = celu(o.y + b.y, cc);
              o.z = celu(o.z + b.z, cc);
              o.w = celu(o.w + b.w, cc);
            } else if (EPI == EPI_MUL_DCELU) {
              const float4 y = *slot;
              o.x *= dcelu_from_out(y.x, cc);
              o.y *= dcelu_from_out(y.y, cc);
              o.z *= dcelu_from_out(y.z, cc);
              o.w *= dcelu_from_out(y.w, cc);
            } else if (EPI == EPI_HEAD) {
              const float4 b = *reinterpret_cast<const float4*>(bias + c0 + 4 * q);
              const float4 w = *reinterpret_cast<const float4*>(w4 + c0 + 4 * q);
              float a;
              a = celu(o.x + b.x, cc); e_acc = fmaf(a, w.x, e_acc); o.x = seed * w.x * dcelu_from_out(a, cc);
              a = celu(o.y + b.y, cc); e_acc = fmaf(a, w.y, e_acc); o.y = seed * w.y * dcelu_from_out(a, cc);
              a = celu(o.z + b.z, cc); e_acc = fmaf(a, w.z, e_acc); o.z = seed * w.z * dcelu_from_out(a, cc);
              a = celu(o.w + b.w, cc); e_acc = fmaf(a, w.w, e_acc); o.w = seed * w.w * dcelu_from_out(a, cc);
            }
            *slot = o;
          }
        }
        __syncwarp();
        if (cq < ncol && (EPI != EPI_HEAD || args.want_backward)) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = cr + 4 * i;
            *reinterpret_cast<float4*>(cbase + (size_t)r * args.ldc + cq) =
                *reinterpret_cast<const float4*>(&buf[r * EPI_LD + cq]);
          }
        }
        __syncwarp();
      }
      if (EPI == EPI_HEAD)
        args.e_member[(size_t)tl.mem * args.rows_cap + my_row] = row_valid ? e_acc + sp.b4[tl.mem] : 0.f;
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace tc
}  // namespace ani
