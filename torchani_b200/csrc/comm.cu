// One-shot all-reduce of the per-rank partial forces / energies over NVLink peer memory.
//
// The path shards by central atom: every rank accumulates dE_owned/dx_j into a full-length
// [3N] float32 buffer and its share of the C conformer energies into float64 -- 120 KB at 10 k atoms.  The
// reference has no collective at all; a library all-reduce (NCCL) of that size is pure latency and sits
// outside the step's CUDA graph.  Here the partial buffers themselves live in CUDA-IPC memory that every peer
// maps (NVSwitch: full bandwidth to every peer), and ONE kernel at the tail of the step's stream does
//     barrier (flags in peer memory) -> every rank reads all W partials, sums them in rank order -> barrier
// so that all ranks end up with the bitwise identical total, nothing is copied or converted, and the whole step
// including the reduction is one CUDA graph.  Flags carry a monotonically increasing epoch kept in device memory
// and advanced by the kernel itself, so graph replays need no host-side argument changes.
//
// This is the one place where the C-ABI owns memory: IPC-exportable buffers must come from cudaMalloc (not from
// a caching allocator), so ani_b200_comm_create allocates them and ani_b200_comm_destroy frees them.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace ani {

constexpr int COMM_MAX_WORLD = 8;
constexpr int COMM_BLOCKS = 32;      // every block runs its own flag barrier with the same block of every peer
constexpr int COMM_THREADS = 512;

struct CommDev {
  float* f32[COMM_MAX_WORLD];        // partial buffers of every rank (own entry = local pointer)
  double* f64[COMM_MAX_WORLD];
  unsigned* flags[COMM_MAX_WORLD];   // [phase 2][COMM_BLOCKS][COMM_MAX_WORLD] per rank
  unsigned* epoch;                   // [COMM_BLOCKS] local
  int32_t* error;                    // local: 1 = a peer did not arrive in time
  long long n32, n64;
  int rank, world;
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer_f4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_peer_f1(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_peer_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// all ranks' block b meet: thread t < world signals rank t and waits for rank t's signal
__device__ __forceinline__ void peer_barrier(const CommDev& c, int phase, unsigned e) {
  __syncthreads();
  const int t = threadIdx.x;
  if (t < c.world) {
    const size_t slot = ((size_t)phase * COMM_BLOCKS + blockIdx.x) * COMM_MAX_WORLD;
    __threadfence_system();
    st_release_sys(c.flags[t] + slot + c.rank, e);
    const unsigned* mine = c.flags[c.rank] + slot + t;
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(mine) - e) < 0) {
      if (clock64() - t0 > 20000000000LL) {  // ~10 s: a peer is gone; give up instead of hanging the GPU
        *c.error = 1;
        break;
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(COMM_THREADS) k_allreduce_partials(const __grid_constant__ CommDev c,
                                                                     float* __restrict__ out32,
                                                                     double* __restrict__ out64) {
  const unsigned e = c.epoch[blockIdx.x] + 1u;
  // the partial sums of this rank were written by earlier kernels of this stream: complete and visible in this
  // GPU's L2 at the kernel boundary, which is where the peers' NVLink reads are served from; the release of the flag
  // store (peer_barrier) orders them for the peers
  peer_barrier(c, 0, e);
  const long long n4 = c.n32 >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < c.world; ++r) {  // fixed order: every rank computes the same bits
      const float4 v = ld_peer_f4(c.f32[r] + 4 * i);
      acc.x += v.x;
      acc.y += v.y;
      acc.z += v.z;
      acc.w += v.w;
    }
    *reinterpret_cast<float4*>(out32 + 4 * i) = acc;
  }
  if (blockIdx.x == 0) {
    for (long long i = (n4 << 2) + threadIdx.x; i < c.n32; i += blockDim.x) {
      float acc = 0.f;
      for (int r = 0; r < c.world; ++r) acc += ld_peer_f1(c.f32[r] + i);
      out32[i] = acc;
    }
    for (long long i = threadIdx.x; i < c.n64; i += blockDim.x) {
      double acc = 0.0;
      for (int r = 0; r < c.world; ++r) acc += ld_peer_f64(c.f64[r] + i);
      out64[i] = acc;
    }
  }
  // nobody may overwrite a partial buffer (the next step zero-fills it) before every peer has read it
  peer_barrier(c, 1, e);
  if (threadIdx.x == 0) c.epoch[blockIdx.x] = e;
}

struct Comm {
  CommDev dev;
  void* local = nullptr;     // cudaMalloc'ed: [f32 n32 (padded) | f64 n64 (padded) | flags | epoch | error]
  void* peers[COMM_MAX_WORLD] = {};
  size_t off64 = 0, off_flags = 0, off_epoch = 0, off_err = 0, bytes = 0;
  bool connected = false;
};

}  // namespace ani

using namespace ani;

extern "C" int ani_b200_comm_create(int rank, int world, long long n_f32, long long n_f64, void** comm_out) {
  if (!comm_out || world < 1 || world > COMM_MAX_WORLD || rank < 0 || rank >= world || n_f32 < 0 || n_f64 < 0)
    return ANI_ERR_BAD_ARG;
  Comm* c = new Comm();
  auto up = [](size_t v) { return (v + 255) / 256 * 256; };
  c->off64 = up((size_t)n_f32 * 4);
  c->off_flags = c->off64 + up((size_t)n_f64 * 8);
  c->off_epoch = c->off_flags + up((size_t)2 * COMM_BLOCKS * COMM_MAX_WORLD * 4);
  c->off_err = c->off_epoch + up((size_t)COMM_BLOCKS * 4);
  c->bytes = c->off_err + 256;
  cudaError_t e = cudaMalloc(&c->local, c->bytes);
  if (e == cudaSuccess) e = cudaMemset(c->local, 0, c->bytes);
  if (e != cudaSuccess) {
    set_cuda_error(e);
    delete c;
    return ANI_ERR_CUDA;
  }
  memset(&c->dev, 0, sizeof(c->dev));
  c->dev.rank = rank;
  c->dev.world = world;
  c->dev.n32 = n_f32;
  c->dev.n64 = n_f64;
  char* base = static_cast<char*>(c->local);
  c->dev.f32[rank] = reinterpret_cast<float*>(base);
  c->dev.f64[rank] = reinterpret_cast<double*>(base + c->off64);
  c->dev.flags[rank] = reinterpret_cast<unsigned*>(base + c->off_flags);
  c->dev.epoch = reinterpret_cast<unsigned*>(base + c->off_epoch);
  c->dev.error = reinterpret_cast<int32_t*>(base + c->off_err);
  c->connected = world == 1;
  *comm_out = c;
  return ANI_OK;
}

extern "C" int ani_b200_comm_handle(void* comm, void* handle64) {
  Comm* c = static_cast<Comm*>(comm);
  if (!c || !handle64) return ANI_ERR_BAD_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, c->local);
  if (e != cudaSuccess) {
    set_cuda_error(e);
    return ANI_ERR_CUDA;
  }
  memcpy(handle64, &h, 64);
  return ANI_OK;
}

extern "C" int ani_b200_comm_connect(void* comm, const void* handles) {
  Comm* c = static_cast<Comm*>(comm);
  if (!c || !handles) return ANI_ERR_BAD_ARG;
  for (int r = 0; r < c->dev.world; ++r) {
    if (r == c->dev.rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const char*>(handles) + 64 * (size_t)r, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      set_cuda_error(e);
      return ANI_ERR_CUDA;
    }
    c->peers[r] = p;
    char* base = static_cast<char*>(p);
    c->dev.f32[r] = reinterpret_cast<float*>(base);
    c->dev.f64[r] = reinterpret_cast<double*>(base + c->off64);
    c->dev.flags[r] = reinterpret_cast<unsigned*>(base + c->off_flags);
  }
  c->connected = true;
  return ANI_OK;
}

extern "C" int ani_b200_comm_buffers(void* comm, float** partial_f32, double** partial_f64, int32_t** error_word) {
  Comm* c = static_cast<Comm*>(comm);
  if (!c) return ANI_ERR_BAD_ARG;
  if (partial_f32) *partial_f32 = c->dev.f32[c->dev.rank];
  if (partial_f64) *partial_f64 = c->dev.f64[c->dev.rank];
  if (error_word) *error_word = c->dev.error;
  return ANI_OK;
}

extern "C" int ani_b200_comm_allreduce(void* comm, float* out_f32, double* out_f64, void* stream) {
  Comm* c = static_cast<Comm*>(comm);
  if (!c || !c->connected || (c->dev.n32 > 0 && !out_f32) || (c->dev.n64 > 0 && !out_f64)) return ANI_ERR_BAD_ARG;
  k_allreduce_partials<<<COMM_BLOCKS, COMM_THREADS, 0, (cudaStream_t)stream>>>(c->dev, out_f32, out_f64);
  ANI_CUDA_CHECK_LAUNCH();
  return ANI_OK;
}

extern "C" int ani_b200_comm_destroy(void* comm) {
  Comm* c = static_cast<Comm*>(comm);
  if (!c) return ANI_OK;
  for (int r = 0; r < c->dev.world; ++r)
    if (c->peers[r]) cudaIpcCloseMemHandle(c->peers[r]);
  if (c->local) cudaFree(c->local);
  delete c;
  return ANI_OK;
}
