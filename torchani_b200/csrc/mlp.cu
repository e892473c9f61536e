// Ensemble MLP forward + backward-to-input on the species-grouped AEV matrix.
//
// Replaces nn/_containers.py:377-421,608-651 (per-species gather / 8-member python loop),
// nn/_infer.py:61-216 (BmmEnsemble) and csrc/mnp.cpp:63-236 (mnp::run).  The rows of the AEV
// matrix are already grouped by species in 128-row tiles (ani_b200_species_layout), so every
// layer of every member of every species is one grouped GEMM launch:
//     C[tile rows, N] = epilogue(A[tile rows, K] x B[K, N])
// The 8 ensemble members share the layer-1 input, so layer 1 is one GEMM with the members
// concatenated along N (the AEV is read once, not 8x as in BmmLinear's expand()).
// Backward-to-input runs the same kernel on the natural-layout weights; CELU'(x) is recovered
// from the stored activation y as (y > 0 ? 1 : (y + alpha) / alpha)  (csrc/mnp.cpp:206-209).
//
// v1 arithmetic: fp32 FFMA register-tiled GEMM (128x128x16 CTA tile, 8x8 per thread).
#include "common.cuh"

namespace ani {

constexpr int BM = ANI_TILE_ROWS, BN = 128, BK = 16;
constexpr int GEMM_THREADS = 256;

enum { EPI_BIAS_CELU = 0, EPI_MUL_DCELU = 1, EPI_PLAIN = 2 };

struct GemmSpecies {
  const float* B;     // [K][ldb] (+ member * b_mstride)
  const float* bias;  // (+ member * bias_mstride) or nullptr
  int K, N, ldb;
  int a_moff, c_moff, bias_mstride;  // per-member column offsets into A / C, bias stride
  long long b_mstride;
};

struct GemmArgs {
  const float* A;
  float* C;
  int lda, ldc;
  const int32_t* tile_species;
  float alpha;
  GemmSpecies sp[ANI_MAX_SPECIES];
};

__device__ __forceinline__ float celu(float x, float alpha) {
  return x > 0.f ? x : alpha * (expf(x / alpha) - 1.0f);
}
__device__ __forceinline__ float dcelu_from_out(float y, float alpha) {
  return y > 0.f ? 1.0f : (y + alpha) / alpha;
}

template <int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 2) k_gemm(const __grid_constant__ GemmArgs args) {
  const int s = args.tile_species[blockIdx.y];
  if (s < 0) return;
  const GemmSpecies& sp = args.sp[s];
  const int n0 = blockIdx.x * BN;
  if (n0 >= sp.N) return;
  const int mem = blockIdx.z;
  const int row0 = blockIdx.y * BM;
  const float* __restrict__ A = args.A + (size_t)row0 * args.lda + (size_t)mem * sp.a_moff;
  const float* __restrict__ B = sp.B + (size_t)mem * sp.b_mstride;
  float* __restrict__ C = args.C + (size_t)row0 * args.ldc + (size_t)mem * sp.c_moff;
  const int K = sp.K, N = sp.N, lda = args.lda, ldb = sp.ldb;

  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN];

  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  // global -> register staging
  const int a_r0 = tid >> 2, a_kq = (tid & 3) * 4;            // rows a_r0 and a_r0 + 64
  const int b_kr0 = tid >> 5, b_c = (tid & 31) * 4;            // k rows b_kr0 and b_kr0 + 8
  const bool b_ok = (n0 + b_c) < N;
  float4 ra[2], rb[2];

  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      ra[it] = *reinterpret_cast<const float4*>(A + (size_t)(a_r0 + it * 64) * lda + k0 + a_kq);
      rb[it] = b_ok ? *reinterpret_cast<const float4*>(B + (size_t)(k0 + b_kr0 + it * 8) * ldb + n0 + b_c)
                    : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int r = a_r0 + it * 64;
      As[a_kq + 0][r] = ra[it].x;
      As[a_kq + 1][r] = ra[it].y;
      As[a_kq + 2][r] = ra[it].z;
      As[a_kq + 3][r] = ra[it].w;
      *reinterpret_cast<float4*>(&Bs[b_kr0 + it * 8][b_c]) = rb[it];
    }
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  load_tiles(0);
  store_tiles();
  __syncthreads();
  for (int k0 = 0; k0 < K; k0 += BK) {
    const bool more = (k0 + BK) < K;
    if (more) load_tiles(k0 + BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
    if (more) {
      store_tiles();
      __syncthreads();
    }
  }

  // ---- epilogue
  const float* bias = (EPI == EPI_BIAS_CELU) ? sp.bias + (size_t)mem * sp.bias_mstride : nullptr;
  const float alpha = args.alpha;
#pragma unroll
  for (int jh = 0; jh < 2; ++jh) {
    const int c = n0 + jh * 64 + tx * 4;
    if (c >= N) continue;
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (EPI == EPI_BIAS_CELU) bv = *reinterpret_cast<const float4*>(bias + c);
#pragma unroll
    for (int ih = 0; ih < 2; ++ih) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = ih * 64 + ty * 4 + i;
        float* cp = C + (size_t)r * args.ldc + c;
        float4 v = make_float4(acc[ih * 4 + i][jh * 4 + 0], acc[ih * 4 + i][jh * 4 + 1],
                               acc[ih * 4 + i][jh * 4 + 2], acc[ih * 4 + i][jh * 4 + 3]);
        if (EPI == EPI_BIAS_CELU) {
          v.x = celu(v.x + bv.x, alpha);
          v.y = celu(v.y + bv.y, alpha);
          v.z = celu(v.z + bv.z, alpha);
          v.w = celu(v.w + bv.w, alpha);
        } else if (EPI == EPI_MUL_DCELU) {
          const float4 y = *reinterpret_cast<const float4*>(cp);
          v.x *= dcelu_from_out(y.x, alpha);
          v.y *= dcelu_from_out(y.y, alpha);
          v.z *= dcelu_from_out(y.z, alpha);
          v.w *= dcelu_from_out(y.w, alpha);
        }
        *reinterpret_cast<float4*>(cp) = v;
      }
    }
  }
}

// final layer (h3 -> 1) + seed of the backward pass; one warp per (row, member)
struct HeadArgs {
  float* act3;
  int ld3;
  float* e_member;
  int rows_cap;
  const int32_t* tile_species;
  const int32_t* row_atom;
  float alpha;
  int num_members;
  int want_backward;
  float member_scale[ANI_MAX_MEMBERS];
  const float* w4[ANI_MAX_SPECIES];
  const float* b4[ANI_MAX_SPECIES];
  int h3[ANI_MAX_SPECIES];
};

__global__ void __launch_bounds__(256) k_mlp_head(const __grid_constant__ HeadArgs args) {
  const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int M = args.num_members;
  const int row = warp_global / M, m = warp_global % M;
  if (row >= args.rows_cap) return;
  const int s = args.tile_species[row / ANI_TILE_ROWS];
  if (s < 0) return;
  const bool valid = args.row_atom[row] >= 0;
  const int h3 = args.h3[s];
  float* a3 = args.act3 + (size_t)row * args.ld3 + (size_t)m * h3;
  const float* w4 = args.w4[s] + (size_t)m * h3;
  float e = 0.f;
  for (int c = lane; c < h3; c += 32) e += a3[c] * w4[c];
  e = warp_sum(e);
  if (lane == 0) args.e_member[(size_t)m * args.rows_cap + row] = valid ? e + args.b4[s][m] : 0.f;
  if (args.want_backward) {
    const float seed = valid ? args.member_scale[m] : 0.f;
    for (int c = lane; c < h3; c += 32) a3[c] = seed * w4[c] * dcelu_from_out(a3[c], args.alpha);
  }
}

// per-conformer reduction of atomic energies (fp64 accumulation) + self energies (sae.py:54-64)
struct ReduceArgs {
  const float* e_member;
  int rows_cap;
  const int32_t* row_of;
  const int32_t* orig_to_sorted;
  const int32_t* species;
  int n, lo, hi, n_per_conf;
  const double* sae;
  float* atomic_out;
  float* member_atomic_out;
  double* energies_out;
  int num_members;
  float member_scale[ANI_MAX_MEMBERS];
};

__global__ void __launch_bounds__(256) k_reduce_energies(const __grid_constant__ ReduceArgs args) {
  // one thread per atom; conformer sums by warp-aggregated float64 atomics (energies_out is
  // zeroed by the launcher)
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = a < args.n;
  double contrib = 0.0;
  int conf = -1;
  if (in_range) {
    conf = a / args.n_per_conf;
    const int sp = args.species[a];
    float e = 0.f;
    bool owned = false;
    int row = 0;
    if (sp >= 0) {
      const int i = args.orig_to_sorted[a];
      if (i >= args.lo && i < args.hi) {
        owned = true;
        row = args.row_of[i];
      }
    }
    for (int m = 0; m < args.num_members; ++m) {
      const float em = owned ? args.e_member[(size_t)m * args.rows_cap + row] : 0.f;
      if (args.member_atomic_out) args.member_atomic_out[(size_t)m * args.n + a] = em;
      e += args.member_scale[m] * em;
    }
    if (args.atomic_out) args.atomic_out[a] = e;
    if (owned) contrib = (double)e + (args.sae ? args.sae[sp] : 0.0);
  }
  const int conf0 = __shfl_sync(ANI_FULL_MASK, conf, 0);
  const bool uniform = __all_sync(ANI_FULL_MASK, conf == conf0 || !in_range);
  if (uniform) {
    for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(ANI_FULL_MASK, contrib, o);
    if ((threadIdx.x & 31) == 0 && conf0 >= 0 && contrib != 0.0) atomicAdd(&args.energies_out[conf0], contrib);
  } else if (in_range && contrib != 0.0) {
    atomicAdd(&args.energies_out[conf], contrib);
  }
}

template <int EPI>
static void launch_gemm(const GemmArgs& ga, int n_tiles_cap, int n_max, int members, cudaStream_t st) {
  dim3 grid((n_max + BN - 1) / BN, n_tiles_cap, members);
  k_gemm<EPI><<<grid, GEMM_THREADS, 0, st>>>(ga);
}

}  // namespace ani

using namespace ani;

extern "C" int ani_b200_mlp_forward_backward(const ani_mlp_model* model, float* x, int rows_cap,
                                             const int32_t* tile_species, const int32_t* row_atom, float* act1,
                                             float* act2, float* act3, float* e_member, int want_backward,
                                             void* stream) {
  if (!model || !x || !tile_species || !row_atom || !act1 || !act2 || !act3 || !e_member) return ANI_ERR_BAD_ARG;
  const int S = model->num_species, M = model->num_members;
  if (S < 1 || S > ANI_MAX_SPECIES || M < 1 || M > ANI_MAX_MEMBERS) return ANI_ERR_BAD_ARG;
  if (rows_cap < ANI_TILE_ROWS || rows_cap % ANI_TILE_ROWS) return ANI_ERR_BAD_ARG;
  const int ldx = model->ldx;
  if (ldx % BK || ldx < model->in_dim) return ANI_ERR_BAD_ARG;
  for (int s = 0; s < S; ++s) {
    const ani_mlp_species& p = model->sp[s];
    if (p.h1 % BK || p.h2 % BK || p.h3 % BK || p.h1 < BK || p.h2 < BK || p.h3 < BK) return ANI_ERR_UNSUPPORTED;
    if (p.h1 > model->h1_max || p.h2 > model->h2_max || p.h3 > model->h3_max) return ANI_ERR_BAD_ARG;
    if (!p.w1 || !p.b1 || !p.w2 || !p.b2 || !p.w3 || !p.b3 || !p.w4 || !p.b4 || !p.w3n || !p.w2n || !p.w1n)
      return ANI_ERR_BAD_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int n_tiles_cap = rows_cap / ANI_TILE_ROWS;
  const int ld1 = M * model->h1_max, ld2 = M * model->h2_max, ld3 = M * model->h3_max;
  GemmArgs ga;
  ga.tile_species = tile_species;
  ga.alpha = model->celu_alpha;
  for (int s = S; s < ANI_MAX_SPECIES; ++s) ga.sp[s] = GemmSpecies{nullptr, nullptr, 0, 0, 0, 0, 0, 0, 0};

  // ---- forward layer 1: all members at once (N = M*h1)
  ga.A = x; ga.lda = ldx; ga.C = act1; ga.ldc = ld1;
  for (int s = 0; s < S; ++s) {
    const ani_mlp_species& p = model->sp[s];
    ga.sp[s] = GemmSpecies{p.w1, p.b1, ldx, M * p.h1, M * p.h1, 0, 0, 0, 0};
  }
  launch_gemm<EPI_BIAS_CELU>(ga, n_tiles_cap, M * model->h1_max, 1, st);
  // ---- forward layer 2, 3: one GEMM per member (grid z)
  ga.A = act1; ga.lda = ld1; ga.C = act2; ga.ldc = ld2;
  for (int s = 0; s < S; ++s) {
    const ani_mlp_species& p = model->sp[s];
    ga.sp[s] = GemmSpecies{p.w2, p.b2, p.h1, p.h2, p.h2, p.h1, p.h2, p.h2, (long long)p.h1 * p.h2};
  }
  launch_gemm<EPI_BIAS_CELU>(ga, n_tiles_cap, model->h2_max, M, st);
  ga.A = act2; ga.lda = ld2; ga.C = act3; ga.ldc = ld3;
  for (int s = 0; s < S; ++s) {
    const ani_mlp_species& p = model->sp[s];
    ga.sp[s] = GemmSpecies{p.w3, p.b3, p.h2, p.h3, p.h3, p.h2, p.h3, p.h3, (long long)p.h2 * p.h3};
  }
  launch_gemm<EPI_BIAS_CELU>(ga, n_tiles_cap, model->h3_max, M, st);
  // ---- head: energies + gradient seed
  {
    HeadArgs ha;
    ha.act3 = act3; ha.ld3 = ld3; ha.e_member = e_member; ha.rows_cap = rows_cap;
    ha.tile_species = tile_species; ha.row_atom = row_atom; ha.alpha = model->celu_alpha;
    ha.num_members = M; ha.want_backward = want_backward;
    for (int m = 0; m < ANI_MAX_MEMBERS; ++m) ha.member_scale[m] = m < M ? model->member_scale[m] : 0.f;
    for (int s = 0; s < ANI_MAX_SPECIES; ++s) {
      ha.w4[s] = s < S ? model->sp[s].w4 : nullptr;
      ha.b4[s] = s < S ? model->sp[s].b4 : nullptr;
      ha.h3[s] = s < S ? model->sp[s].h3 : 0;
    }
    const long long warps = (long long)rows_cap * M;
    const int blocks = (int)((warps + 7) / 8);
    k_mlp_head<<<blocks, 256, 0, st>>>(ha);
  }
  if (want_backward) {
    // ---- backward: G2 = (G3 x W3) * celu'(A2), G1 = (G2 x W2) * celu'(A1), dX = G1 x W1
    ga.A = act3; ga.lda = ld3; ga.C = act2; ga.ldc = ld2;
    for (int s = 0; s < S; ++s) {
      const ani_mlp_species& p = model->sp[s];
      ga.sp[s] = GemmSpecies{p.w3n, nullptr, p.h3, p.h2, p.h2, p.h3, p.h2, 0, (long long)p.h3 * p.h2};
    }
    launch_gemm<EPI_MUL_DCELU>(ga, n_tiles_cap, model->h2_max, M, st);
    ga.A = act2; ga.lda = ld2; ga.C = act1; ga.ldc = ld1;
    for (int s = 0; s < S; ++s) {
      const ani_mlp_species& p = model->sp[s];
      ga.sp[s] = GemmSpecies{p.w2n, nullptr, p.h2, p.h1, p.h1, p.h2, p.h1, 0, (long long)p.h2 * p.h1};
    }
    launch_gemm<EPI_MUL_DCELU>(ga, n_tiles_cap, model->h1_max, M, st);
    ga.A = act1; ga.lda = ld1; ga.C = x; ga.ldc = ldx;
    for (int s = 0; s < S; ++s) {
      const ani_mlp_species& p = model->sp[s];
      ga.sp[s] = GemmSpecies{p.w1n, nullptr, M * p.h1, ldx, ldx, 0, 0, 0, 0};
    }
    launch_gemm<EPI_PLAIN>(ga, n_tiles_cap, ldx, 1, st);
  }
  ANI_CUDA_CHECK_LAUNCH();
  return ANI_OK;
}

extern "C" int ani_b200_reduce_energies(const ani_mlp_model* model, const float* e_member, int rows_cap,
                                        const int32_t* row_of, const int32_t* orig_to_sorted,
                                        const int32_t* species, int n, int lo, int hi, int n_conf, int n_per_conf,
                                        const double* sae, float* atomic_out, float* member_atomic_out,
                                        double* energies_out, void* stream) {
  if (!model || !e_member || !row_of || !orig_to_sorted || !species || !energies_out) return ANI_ERR_BAD_ARG;
  if (n != n_conf * n_per_conf || n_conf < 1) return ANI_ERR_BAD_ARG;
  ReduceArgs ra;
  ra.e_member = e_member; ra.rows_cap = rows_cap; ra.row_of = row_of; ra.orig_to_sorted = orig_to_sorted;
  ra.species = species; ra.n = n; ra.lo = lo; ra.hi = hi; ra.n_per_conf = n_per_conf; ra.sae = sae;
  ra.atomic_out = atomic_out; ra.member_atomic_out = member_atomic_out; ra.energies_out = energies_out;
  ra.num_members = model->num_members;
  for (int m = 0; m < ANI_MAX_MEMBERS; ++m)
    ra.member_scale[m] = m < model->num_members ? model->member_scale[m] : 0.f;
  cudaMemsetAsync(energies_out, 0, sizeof(double) * (size_t)n_conf, (cudaStream_t)stream);
  k_reduce_energies<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(ra);
  ANI_CUDA_CHECK_LAUNCH();
  return ANI_OK;
}
