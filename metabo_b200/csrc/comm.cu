// comm.cu -- the one exchange step of the path: the PPO gradient all-reduce (metabo/ppo/ppo.py:170-172 under data
// parallelism; the reference's learner is a single process, SURVEY 8e names the flat fp32 gradient all-reduce as the only
// collective) as ONE kernel over NVLink / NVSwitch peer memory, fused with the Adam update that consumes it.
//
// One process per GPU.  Every rank owns a region allocated HERE with cudaMalloc and exported as a CUDA IPC handle; the
// ranks exchange the 64-byte handles out of band (torch.distributed all_gather on the host) and map each other's regions.
//
//   region = | ctrl (256 B): epoch, arrive, done, err | flags[MAX_RANKS] (one 128 B line each) | slot 0 | slot 1 |
//
// One call (all ranks enqueue the same call, e.g. by replaying the same captured CUDA graph):
//   A  every CTA copies its share of the local payload into slot[epoch & 1] of the LOCAL region (so the producers of the
//      payload never write peer-visible memory and the slot parity follows the device-resident epoch: graph replays
//      alternate slots by themselves), fences at system scope and arrives on a local counter; the CTA that arrives last
//      publishes `epoch + 1` into flags[my_rank] of EVERY peer region (st.release.sys over NVLink);
//   B  every CTA spins (ld.acquire.sys on LOCAL memory, bounded by %globaltimer) until all peers published epoch + 1;
//   C  every CTA reads its share of every rank's slot in rank order 0..W-1 (ld.relaxed.sys over NVLink: one-shot "pull")
//      and sums in that fixed order -- every rank computes bit-identical sums -- then either stores the sum (plain
//      all-reduce, fp32 or fp64) or applies torch.optim.Adam's update to the parameters (fused form: the reduced
//      gradient is never written anywhere);
//   the CTA that finishes last advances the epoch.
// Slot reuse is safe without a second barrier: a rank can only overwrite slot[e & 1] in call e + 2, which it reaches
// after its own call e + 1 saw every peer's flag e + 2, i.e. after every peer FINISHED call e (stream order).
// The grid never exceeds one CTA per SM (all CTAs co-resident: the spin in B cannot starve phase A of the same grid).
// Payload <= slot_bytes (0.97 MB for the 4x200 policy + value nets): 8 ranks pull 7 x 0.97 MB each, ~10 us at NVLink speed.
#include "common.cuh"

namespace mbo {
namespace comm {

constexpr int MAX_RANKS = 16;
constexpr int CTRL_BYTES = 256;
constexpr int FLAG_STRIDE = 128;                    // one line per peer flag
constexpr int HDR_BYTES = CTRL_BYTES + MAX_RANKS * FLAG_STRIDE;
constexpr int THREADS = 512;
constexpr unsigned long long TIMEOUT_NS = 10000000000ull;   // one time-out latches `err`; later waits return at once

struct Ctrl {                 // first 256 bytes of a region
  unsigned long long epoch;   // number of completed calls
  unsigned int arrive;        // CTAs that finished phase A of the running call
  unsigned int done;          // CTAs that finished the running call
  unsigned int err;           // sticky: bit 0 = a wait on a peer flag timed out
  unsigned int calls_bad;
};

}  // namespace comm
}  // namespace mbo

struct mbo_comm {
  int rank = 0, world = 1, device = 0;
  size_t slot_bytes = 0, region_bytes = 0;
  char* local = nullptr;                       // this rank's region (cudaMalloc)
  char* peer[mbo::comm::MAX_RANKS] = {};       // every rank's region as mapped into this process (peer[rank] == local)
  bool opened[mbo::comm::MAX_RANKS] = {};
  bool connected = false;
  int num_sms = 148;
};

namespace mbo {
namespace comm {

struct Peers {
  char* region[MAX_RANKS];
  int rank, world;
  size_t slot_bytes;
};

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_sys(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_sys(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ char* slot_of(const Peers& P, int r, unsigned long long epoch) {
  return P.region[r] + HDR_BYTES + (size_t)(epoch & 1ull) * P.slot_bytes;
}

// phases A + B; returns the epoch of this call.  `src` = local payload (count elements of T).
template <typename T>
__device__ __forceinline__ unsigned long long publish_and_wait(const Peers& P, const T* __restrict__ src, long long count) {
  Ctrl* ctrl = (Ctrl*)P.region[P.rank];
  __shared__ unsigned long long s_epoch;
  if (threadIdx.x == 0) s_epoch = *(volatile unsigned long long*)&ctrl->epoch;
  __syncthreads();
  const unsigned long long epoch = s_epoch;
  T* mine = (T*)slot_of(P, P.rank, epoch);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) mine[i] = src[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(&ctrl->arrive, 1u);
    if (prev == gridDim.x - 1) {            // last CTA of phase A: the whole local slot is written and fenced
      ctrl->arrive = 0;
      __threadfence_system();
      for (int r = 0; r < P.world; ++r)
        if (r != P.rank)
          st_release_sys((unsigned long long*)(P.region[r] + CTRL_BYTES + (size_t)P.rank * FLAG_STRIDE), epoch + 1);
    }
  }
  // phase B: one thread per peer polls the LOCAL flag that peer writes
  if (threadIdx.x < P.world && (int)threadIdx.x != P.rank) {
    const unsigned long long* f = (const unsigned long long*)(P.region[P.rank] + CTRL_BYTES + (size_t)threadIdx.x * FLAG_STRIDE);
    const unsigned long long t0 = gtimer();
    while (ld_acquire_sys(f) < epoch + 1) {
      if (*(volatile unsigned int*)&ctrl->err) break;                        // fail fast after the first time-out
      if (gtimer() - t0 > TIMEOUT_NS) { atomicOr(&ctrl->err, 1u); break; }
      __nanosleep(64);
    }
  }
  __syncthreads();
  return epoch;
}

__device__ __forceinline__ void finish_call(const Peers& P, unsigned long long epoch) {
  Ctrl* ctrl = (Ctrl*)P.region[P.rank];
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(&ctrl->done, 1u);
    if (prev == gridDim.x - 1) {
      ctrl->done = 0;
      __threadfence();
      *(volatile unsigned long long*)&ctrl->epoch = epoch + 1;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(THREADS) allreduce_kernel(Peers P, const T* src, T* dst, long long count) {
  const unsigned long long epoch = publish_and_wait<T>(P, src, count);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
    T v[MAX_RANKS];
#pragma unroll
    for (int r = 0; r < MAX_RANKS; ++r)
      if (r < P.world) v[r] = ld_sys((const T*)slot_of(P, r, epoch) + i);
    T s = v[0];
#pragma unroll
    for (int r = 1; r < MAX_RANKS; ++r)
      if (r < P.world) s += v[r];
    dst[i] = s;
  }
  finish_call(P, epoch);
}

// ---- fused: gradient all-reduce + torch.optim.Adam step (ppo.py:93, :170-172) -----------------------------------------
constexpr int ADAM_MAX_TENSORS = 32;
struct AdamSeg {
  float* p[ADAM_MAX_TENSORS];
  float* m[ADAM_MAX_TENSORS];
  float* v[ADAM_MAX_TENSORS];
  long long end[ADAM_MAX_TENSORS];     // exclusive prefix end of tensor t in the flat gradient
  long long* st[ADAM_MAX_TENSORS];     // per-tensor {step, 1 - beta1^step, 1 - beta2^step}
  int n;
};
__global__ void adam_tick_kernel(AdamSeg a, double beta1, double beta2) {
  const int i = threadIdx.x;
  if (i >= a.n) return;
  long long* state = a.st[i];
  const long long t = state[0] + 1;
  state[0] = t;
  ((double*)state)[1] = 1.0 - pow(beta1, (double)t);
  ((double*)state)[2] = 1.0 - pow(beta2, (double)t);
}
__global__ void __launch_bounds__(THREADS) allreduce_adam_kernel(Peers P, const float* __restrict__ grad, float* __restrict__ grad_out,
                                                                 long long count, AdamSeg a, float lr, float beta1, float beta2,
                                                                 float eps) {
  const unsigned long long epoch = publish_and_wait<float>(P, grad, count);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
    float gv[MAX_RANKS];
#pragma unroll
    for (int r = 0; r < MAX_RANKS; ++r)
      if (r < P.world) gv[r] = ld_sys((const float*)slot_of(P, r, epoch) + i);
    float g = gv[0];
#pragma unroll
    for (int r = 1; r < MAX_RANKS; ++r)
      if (r < P.world) g += gv[r];
    if (grad_out) grad_out[i] = g;
    int t = 0;
    while (i >= a.end[t]) ++t;
    const long long j = i - (t ? a.end[t - 1] : 0);
    const double* state = (const double*)a.st[t];
    const float bc1 = (float)state[1], bc2s = (float)sqrt(state[2]);
    const float m = a.m[t][j] + (g - a.m[t][j]) * (1.f - beta1);        // exp_avg.lerp_(grad, 1 - beta1)
    const float v = a.v[t][j] * beta2 + (1.f - beta2) * g * g;
    a.m[t][j] = m;
    a.v[t][j] = v;
    a.p[t][j] -= (lr / bc1) * m / (sqrtf(v) / bc2s + eps);
  }
  finish_call(P, epoch);
}

static Peers peers_of(const mbo_comm* c) {
  Peers P;
  for (int r = 0; r < MAX_RANKS; ++r) P.region[r] = r < c->world ? c->peer[r] : nullptr;
  P.rank = c->rank; P.world = c->world; P.slot_bytes = c->slot_bytes;
  return P;
}
static int grid_for(const mbo_comm* c, long long count) {
  long long g = (count + THREADS - 1) / THREADS;
  if (g > c->num_sms) g = c->num_sms;          // <= 1 CTA per SM: all CTAs co-resident (see header comment)
  return g < 1 ? 1 : (int)g;
}

}  // namespace comm
}  // namespace mbo

using namespace mbo;
using namespace mbo::comm;

extern "C" int mbo_comm_create(int rank, int world, size_t slot_bytes, mbo_comm** out) {
  MBO_CHECK_ARG(out, "mbo_comm_create: null out");
  MBO_CHECK_ARG(world >= 1 && world <= MAX_RANKS && rank >= 0 && rank < world, "mbo_comm_create: rank %d / world %d (max %d)",
                rank, world, MAX_RANKS);
  MBO_CHECK_ARG(slot_bytes > 0, "mbo_comm_create: slot_bytes must be > 0");
  mbo_comm* c = new mbo_comm();
  c->rank = rank; c->world = world;
  c->slot_bytes = (slot_bytes + 255) & ~(size_t)255;
  c->region_bytes = HDR_BYTES + 2 * c->slot_bytes;
  cudaError_t e = cudaGetDevice(&c->device);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, c->device);
  if (e == cudaSuccess) e = cudaMalloc((void**)&c->local, c->region_bytes);
  if (e == cudaSuccess) e = cudaMemset(c->local, 0, c->region_bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    set_error("mbo_comm_create: %s", cudaGetErrorString(e));
    if (c->local) cudaFree(c->local);
    delete c;
    return MBO_ERR_CUDA;
  }
  c->peer[rank] = c->local;
  c->connected = (world == 1);
  *out = c;
  return MBO_OK;
}

extern "C" int mbo_comm_export(mbo_comm* c, void* handle_out) {
  MBO_CHECK_ARG(c && handle_out, "mbo_comm_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == MBO_COMM_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  MBO_CUDA(cudaIpcGetMemHandle(&h, c->local));
  memcpy(handle_out, &h, sizeof(h));
  return MBO_OK;
}

extern "C" int mbo_comm_connect(mbo_comm* c, const void* handles) {
  MBO_CHECK_ARG(c && handles, "mbo_comm_connect: null pointer");
  MBO_CHECK_ARG(!c->connected || c->world == 1, "mbo_comm_connect: already connected");
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + (size_t)r * MBO_COMM_HANDLE_BYTES, sizeof(h));
    void* p = nullptr;
    MBO_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->peer[r] = (char*)p;
    c->opened[r] = true;
  }
  c->connected = true;
  return MBO_OK;
}

extern "C" void mbo_comm_destroy(mbo_comm* c) {
  if (!c) return;
  for (int r = 0; r < c->world; ++r)
    if (c->opened[r] && c->peer[r]) cudaIpcCloseMemHandle(c->peer[r]);
  if (c->local) cudaFree(c->local);
  delete c;
}

extern "C" int mbo_comm_allreduce(mbo_comm* c, int dtype, const void* src, void* dst, size_t count, void* stream) {
  MBO_CHECK_ARG(c && src && dst && count > 0, "mbo_comm_allreduce: bad arguments");
  MBO_CHECK_ARG(c->connected, "mbo_comm_allreduce: call mbo_comm_connect first");
  MBO_CHECK_ARG(dtype == MBO_COMM_F32 || dtype == MBO_COMM_F64, "mbo_comm_allreduce: dtype must be MBO_COMM_F32 / MBO_COMM_F64");
  const size_t esz = dtype == MBO_COMM_F32 ? 4 : 8;
  MBO_CHECK_ARG(count * esz <= c->slot_bytes, "mbo_comm_allreduce: %zu bytes exceed the slot (%zu)", count * esz, c->slot_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  const Peers P = peers_of(c);
  const int grid = grid_for(c, (long long)count);
  if (dtype == MBO_COMM_F32)
    allreduce_kernel<float><<<grid, THREADS, 0, st>>>(P, (const float*)src, (float*)dst, (long long)count);
  else
    allreduce_kernel<double><<<grid, THREADS, 0, st>>>(P, (const double*)src, (double*)dst, (long long)count);
  MBO_LAUNCH_CHECK();
  return MBO_OK;
}

extern "C" int mbo_comm_allreduce_adam(mbo_comm* c, const float* grad, float* grad_out, int n, float* const* p, float* const* m,
                                       float* const* v, const int64_t* sizes, int64_t* const* state, float lr, float beta1,
                                       float beta2, float eps, void* stream) {
  MBO_CHECK_ARG(c && grad && p && m && v && sizes && state, "mbo_comm_allreduce_adam: null pointer");
  MBO_CHECK_ARG(c->connected, "mbo_comm_allreduce_adam: call mbo_comm_connect first");
  MBO_CHECK_ARG(n >= 1 && n <= ADAM_MAX_TENSORS, "mbo_comm_allreduce_adam: 1..%d tensors per call", ADAM_MAX_TENSORS);
  AdamSeg a;
  long long tot = 0;
  for (int t = 0; t < n; ++t) {
    MBO_CHECK_ARG(p[t] && m[t] && v[t] && state[t] && sizes[t] > 0, "mbo_comm_allreduce_adam: null / empty tensor %d", t);
    a.p[t] = p[t]; a.m[t] = m[t]; a.v[t] = v[t]; a.st[t] = (long long*)state[t];
    tot += sizes[t];
    a.end[t] = tot;
  }
  a.n = n;
  MBO_CHECK_ARG((size_t)tot * 4 <= c->slot_bytes, "mbo_comm_allreduce_adam: %lld floats exceed the slot (%zu bytes)", tot, c->slot_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  adam_tick_kernel<<<1, ADAM_MAX_TENSORS, 0, st>>>(a, (double)beta1, (double)beta2);
  allreduce_adam_kernel<<<grid_for(c, tot), THREADS, 0, st>>>(peers_of(c), grad, grad_out, tot, a, lr, beta1, beta2, eps);
  MBO_LAUNCH_CHECK();
  return MBO_OK;
}

extern "C" int mbo_comm_status(mbo_comm* c, unsigned long long* epoch_out, unsigned int* err_out, void* stream) {
  MBO_CHECK_ARG(c, "mbo_comm_status: null comm");
  Ctrl h;
  MBO_CUDA(cudaMemcpyAsync(&h, c->local, sizeof(Ctrl), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  MBO_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  if (epoch_out) *epoch_out = h.epoch;
  if (err_out) *err_out = h.err;
  if (h.err) {
    Ctrl z = h; z.err = 0;
    MBO_CUDA(cudaMemcpy(c->local + offsetof(Ctrl, err), &z.err, sizeof(unsigned int), cudaMemcpyHostToDevice));
  }
  return MBO_OK;
}
