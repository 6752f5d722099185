// xchg.cu -- the data-parallel step's gradient exchange over NVLink peer memory, without holding SMs.
//
// What it replaces: fairseq's LegacyDistributedDataParallel.all_reduce_grads (legacy_distributed_data_parallel.py:
// 76-165, called from trainer.py:928): ONE flat gradient buffer, divided by the world size, summed over the ranks.
// With NCCL that is a kernel of 16-32 thread blocks spinning on NVLink for the whole transfer; overlapped with the
// lattice kernels of the next step (which want all 148 SMs) it costs more SM time than link time.  Here the bytes
// are moved by the COPY ENGINES (cudaMemcpyAsync between peer-mapped allocations) and only the arithmetic -- one
// pass over 1/N of the buffer -- runs on SMs:
//
//   push                 rank r copies slice p of its buffer into row r of peer p's staging area   (N-1 CE copies)
//   barrier B            every rank's staging area is complete
//   reduce               buffer[slice r] = (own + sum of staging rows) / N, summed in rank order   (one short kernel)
//   push                 rank r copies its reduced slice r into every peer's buffer                (N-1 CE copies)
//   barrier C            every buffer is complete; nobody's staging area is still being read
//
// Copy-engine WRITES over NVLink run at ~1.5x the speed of reads (measured: 262 MB per phase and GPU at N=8 in 0.41 ms
// pulled), hence pushes.  A barrier is a one-warp kernel that stores a monotonically increasing epoch into each peer's
// flag word (system-scope release, stream-ordered after the rank's copies) and spins on its own flag words (bounded:
// ~4 s, then an error word is set instead of hanging the GPU).  Why two barriers suffice: a rank's pushes into my
// staging area for exchange i+1 come after it passed barrier C of exchange i, which I only reach after my reduce of
// exchange i; its push of a reduced slice into my buffer comes after barrier B, which I only reach after my own pushes
// out of that buffer; and the next step's gradients are written after barrier C.
// Each slice is reduced by exactly one rank in a fixed order, so all ranks end with bit-identical buffers.
//
// Optional (DAGB200_XCHG_SM_FRAC > 0, default 0): a share of every transfer is moved by a few thread blocks storing to
// peer memory (xchg_copy_kernel) while the copy engines move the rest; measured at N = 8: 0.53 -> 0.43 ms per phase with
// half of the bytes on 24 blocks -- NVLink saturates at ~580 GB/s per GPU and direction either way, which is why the
// default exchange of a box with NVLink SHARP is the in-switch form at the end of this file.
//
// Peer mapping: the buffers are plain cudaMalloc allocations exported with cudaIpcGetMemHandle and opened by the
// other ranks of the node (one process per GPU).  The handles travel through whatever the host side has
// (torch.distributed all_gather in daspeech_b200/dist.py).
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "../../include/dagb200.h"

namespace dagb200 {

constexpr int kMaxRanks = 16;
constexpr int kXchgStreams = 8;      // upper bound; the handle uses `ns` of them (DAGB200_XCHG_STREAMS, default 4)

struct Xchg {
  int rank = 0, world = 1, device = 0;
  size_t numel = 0, slice = 0;          // floats in the buffer / in one rank's slice (multiple of 4)
  float *buf[kMaxRanks] = {};           // buf[rank] local, others peer-mapped
  int *flags[kMaxRanks] = {};           // flags[p] = rank p's int32[kMaxRanks + 1] (last word: error)
  float *staging[kMaxRanks] = {};       // staging[p] = rank p's [world-1][slice] area (staging[rank] local)
  int epoch = 0;
  int ns = 4;
  float sm_frac = 0.f;                  // share of every transfer moved by thread blocks instead of the copy engines
  int sm_ctas = 16;                     //   (DAGB200_XCHG_SM_FRAC, DAGB200_XCHG_SM_CTAS); 0 = copy engines only
  cudaStream_t sm_stream = nullptr;
  cudaEvent_t sm_join = nullptr;
  cudaStream_t side[kXchgStreams] = {};
  cudaEvent_t fork = nullptr, join[kXchgStreams] = {};
  cudaEvent_t mark[6] = {};             // phase boundaries of the most recent exchange (dagb200_grad_exchange_phases)
  bool timing = false;
};

struct FlagTable { int *flags[kMaxRanks]; };

__global__ void xchg_barrier_kernel(FlagTable tab, int rank, int world, int epoch) {
  const int p = threadIdx.x;
  if (p >= world) return;
  __threadfence_system();
  volatile int *theirs = tab.flags[p] + rank;     // my arrival, in rank p's memory
  *theirs = epoch;
  __threadfence_system();
  volatile int *mine = tab.flags[rank] + p;       // rank p's arrival, in my memory
  const long long t0 = clock64();
  while (*mine - epoch < 0) {
    if (clock64() - t0 > (1ll << 33)) {            // ~4 s at 2 GHz: a peer died; do not hang the device
      tab.flags[rank][kMaxRanks] = epoch;
      break;
    }
    __nanosleep(200);
  }
  __threadfence_system();
}

// buffer[slice r] = (own + staging rows, in rank order) * inv.  Four independent float4 streams per thread and row keep
// enough loads in flight for the kernel to run at HBM speed with a grid that leaves most SM slots to the lattice kernels.
__global__ void __launch_bounds__(256) xchg_reduce_kernel(float4 *__restrict__ own, const float4 *__restrict__ staging,
                                                          size_t n4, size_t slice4, int rank, int world, float inv) {
  constexpr int U = 4;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t base = (size_t)blockIdx.x * blockDim.x + threadIdx.x; base < n4; base += stride * U) {
    float4 acc[U];
#pragma unroll
    for (int u = 0; u < U; u++) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    int row = 0;
    for (int p = 0; p < world; p++) {
      const float4 *src = p == rank ? own : staging + (size_t)(row++) * slice4;
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const size_t i = base + (size_t)u * stride;
        v[u] = i < n4 ? (p == rank ? src[i] : __ldcs(src + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < U; u++) { acc[u].x += v[u].x; acc[u].y += v[u].y; acc[u].z += v[u].z; acc[u].w += v[u].w; }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const size_t i = base + (size_t)u * stride;
      if (i < n4) own[i] = make_float4(acc[u].x * inv, acc[u].y * inv, acc[u].z * inv, acc[u].w * inv);
    }
  }
}

// The part of a phase's transfers that thread blocks move (peer stores over NVLink) while the copy engines move the rest:
// up to kMaxRanks - 1 segments, walked as one concatenated range with 8 float4 in flight per thread.
struct CopySegs {
  float4 *dst[kMaxRanks];
  const float4 *src[kMaxRanks];
  unsigned long long end4[kMaxRanks];   // cumulative length in float4
  int nseg;
};
__global__ void __launch_bounds__(512) xchg_copy_kernel(CopySegs cs) {
  constexpr int U = 8;
  const unsigned long long total = cs.nseg ? cs.end4[cs.nseg - 1] : 0ull;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long base = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; base < total; base += stride * U) {
    float4 v[U];
    int seg[U];
    unsigned long long off[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const unsigned long long i = base + (unsigned long long)u * stride;
      int sgm = 0;
      while (sgm < cs.nseg - 1 && i >= cs.end4[sgm]) sgm++;
      seg[u] = sgm;
      off[u] = i - (sgm ? cs.end4[sgm - 1] : 0ull);
      if (i < total) v[u] = __ldcs(cs.src[sgm] + off[u]);
    }
#pragma unroll
    for (int u = 0; u < U; u++)
      if (base + (unsigned long long)u * stride < total) cs.dst[seg[u]][off[u]] = v[u];
  }
}

// In-switch variant (NVLink SHARP): `mc` is a MULTICAST mapping of the same buffer on every rank.  multimem.ld_reduce
// returns the sum over all ranks computed inside the NVSwitch, multimem.st writes a value into every rank's copy; each
// rank handles its own 1/N slice, so a GPU receives its slice once and sends it once (plus what the switch reads from
// it) instead of N-1 staging rows each way: 2 * 300 MB per GPU and exchange at N = 8 against 2 * 525 MB for the
// point-to-point form -- and a handful of thread blocks are enough to keep the switch busy.
__global__ void __launch_bounds__(512) xchg_nvls_kernel(float *mc, size_t lo4, size_t n4, float inv) {
  constexpr int U = 4;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  float4 *p = reinterpret_cast<float4 *>(mc) + lo4;
  for (size_t base = (size_t)blockIdx.x * blockDim.x + threadIdx.x; base < n4; base += stride * U) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const size_t i = base + (size_t)u * stride;
      if (i < n4)
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(p + i) : "memory");
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const size_t i = base + (size_t)u * stride;
      if (i < n4)
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                     ::"l"(p + i), "f"(v[u].x * inv), "f"(v[u].y * inv), "f"(v[u].z * inv), "f"(v[u].w * inv) : "memory");
    }
  }
}

static void barrier(Xchg *x, cudaStream_t st) {
  FlagTable tab;
  for (int p = 0; p < kMaxRanks; p++) tab.flags[p] = x->flags[p];
  x->epoch++;
  xchg_barrier_kernel<<<1, 32, 0, st>>>(tab, x->rank, x->world, x->epoch);
}

// N-1 pushes spread over the side streams (each slice cut into pieces so that kXchgStreams copy engines work at any
// world size), joined back into st.  scatter: slice p of my buffer -> row(rank) of p's staging; else my reduced slice
// -> the same place in p's buffer.
static int pushes(Xchg *x, cudaStream_t st, bool scatter) {
  cudaError_t e = cudaEventRecord(x->fork, st);
  if (e != cudaSuccess) return cuda_fail(e, "xchg fork");
  const int ns = x->ns;
  for (int s = 0; s < ns; s++) cudaStreamWaitEvent(x->side[s], x->fork, 0);
  const int pieces = (ns + x->world - 2) / (x->world - 1);
  int q = 0;
  CopySegs cs;
  cs.nseg = 0;
  unsigned long long acc4 = 0;
  for (int k = 1; k < x->world; k++) {
    const int p = (x->rank + k) % x->world;       // every rank starts at a different peer: no hot destination
    const int owner = scatter ? p : x->rank;      // whose slice moves
    const size_t lo = (size_t)owner * x->slice;
    const size_t n = lo >= x->numel ? 0 : (x->numel - lo < x->slice ? x->numel - lo : x->slice);
    const int srow = x->rank < p ? x->rank : x->rank - 1;   // my row in p's staging: rows in rank order, p's own left out
    const float *src = x->buf[x->rank] + lo;
    float *dst = scatter ? x->staging[p] + (size_t)srow * x->slice : x->buf[p] + lo;
    // the tail of the transfer goes to the thread blocks (whole float4, 128-byte aligned start)
    size_t nsm = x->sm_frac > 0.f ? (size_t)((double)n * x->sm_frac) / 32 * 32 : 0;
    const size_t nce = n - nsm;
    if (nsm) {
      cs.dst[cs.nseg] = reinterpret_cast<float4 *>(dst + nce);
      cs.src[cs.nseg] = reinterpret_cast<const float4 *>(src + nce);
      acc4 += nsm / 4;
      cs.end4[cs.nseg++] = acc4;
    }
    const size_t step = ((nce + pieces - 1) / pieces + 31) / 32 * 32;
    for (size_t off = 0; off < nce; off += step, q++) {
      const size_t m = nce - off < step ? nce - off : step;
      e = cudaMemcpyAsync(dst + off, src + off, m * 4, cudaMemcpyDefault, x->side[q % ns]);
      if (e != cudaSuccess) return cuda_fail(e, "xchg peer copy");
    }
  }
  if (cs.nseg) {
    cudaStreamWaitEvent(x->sm_stream, x->fork, 0);
    xchg_copy_kernel<<<x->sm_ctas, 512, 0, x->sm_stream>>>(cs);
    cudaEventRecord(x->sm_join, x->sm_stream);
    cudaStreamWaitEvent(st, x->sm_join, 0);
  }
  for (int s = 0; s < ns; s++) {
    cudaEventRecord(x->join[s], x->side[s]);
    cudaStreamWaitEvent(st, x->join[s], 0);
  }
  return 0;
}

}  // namespace dagb200

using namespace dagb200;

extern "C" int dagb200_peer_alloc(size_t bytes, void **ptr) {
  if (!ptr || bytes == 0) { set_error("peer_alloc: bad arguments"); return DAGB200_EINVAL; }
  cudaError_t e = cudaMalloc(ptr, bytes);
  if (e != cudaSuccess) return cuda_fail(e, "peer_alloc");
  e = cudaMemset(*ptr, 0, bytes);
  if (e != cudaSuccess) return cuda_fail(e, "peer_alloc memset");
  return cudaDeviceSynchronize() == cudaSuccess ? 0 : cuda_fail(cudaGetLastError(), "peer_alloc sync");
}

extern "C" int dagb200_peer_free(void *ptr) {
  cudaError_t e = cudaFree(ptr);
  return e == cudaSuccess ? 0 : cuda_fail(e, "peer_free");
}

extern "C" int dagb200_peer_export(void *ptr, void *handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
  if (e != cudaSuccess) return cuda_fail(e, "peer_export");
  memcpy(handle64, &h, 64);
  return 0;
}

extern "C" int dagb200_peer_open(const void *handle64, void **ptr) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
  return e == cudaSuccess ? 0 : cuda_fail(e, "peer_open");
}

extern "C" int dagb200_peer_close(void *ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(ptr);
  return e == cudaSuccess ? 0 : cuda_fail(e, "peer_close");
}

extern "C" size_t dagb200_grad_exchange_slice(size_t numel, int world) {
  if (world <= 0) return 0;
  size_t s = (numel + world - 1) / world;
  return (s + 3) / 4 * 4;
}

extern "C" int dagb200_grad_exchange_create(void *const *bufs, void *const *flags, void *const *stagings, size_t numel,
                                            int rank, int world, void **handle) {
  if (!bufs || !flags || !stagings || !handle || world < 1 || world > kMaxRanks || rank < 0 || rank >= world || numel % 4) {
    set_error("grad_exchange_create: bad arguments (world <= %d, numel a multiple of 4)", kMaxRanks);
    return DAGB200_EINVAL;
  }
  Xchg *x = new Xchg();
  x->rank = rank; x->world = world; x->numel = numel;
  x->slice = dagb200_grad_exchange_slice(numel, world);
  cudaGetDevice(&x->device);
  if (const char *e = getenv("DAGB200_XCHG_SM_FRAC")) {
    const float v = (float)atof(e);
    if (v >= 0.f && v <= 1.f) x->sm_frac = v;
  }
  if (const char *e = getenv("DAGB200_XCHG_SM_CTAS")) {
    const int v = atoi(e);
    if (v >= 1 && v <= 148) x->sm_ctas = v;
  }
  if (const char *e = getenv("DAGB200_XCHG_STREAMS")) {
    const int v = atoi(e);
    if (v >= 1 && v <= kXchgStreams) x->ns = v;
  }
  for (int p = 0; p < world; p++) {
    x->buf[p] = (float *)bufs[p]; x->flags[p] = (int *)flags[p]; x->staging[p] = (float *)stagings[p];
  }
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  for (int s = 0; s < kXchgStreams; s++) {
    cudaError_t e = cudaStreamCreateWithPriority(&x->side[s], cudaStreamNonBlocking, hi);
    if (e != cudaSuccess) { delete x; return cuda_fail(e, "grad_exchange_create stream"); }
    cudaEventCreateWithFlags(&x->join[s], cudaEventDisableTiming);
  }
  cudaEventCreateWithFlags(&x->fork, cudaEventDisableTiming);
  cudaStreamCreateWithPriority(&x->sm_stream, cudaStreamNonBlocking, hi);
  cudaEventCreateWithFlags(&x->sm_join, cudaEventDisableTiming);
  *handle = x;
  return 0;
}

extern "C" int dagb200_grad_exchange_destroy(void *handle) {
  Xchg *x = (Xchg *)handle;
  if (!x) return 0;
  for (int s = 0; s < kXchgStreams; s++) {
    if (x->side[s]) cudaStreamDestroy(x->side[s]);
    if (x->join[s]) cudaEventDestroy(x->join[s]);
  }
  if (x->fork) cudaEventDestroy(x->fork);
  if (x->sm_stream) cudaStreamDestroy(x->sm_stream);
  if (x->sm_join) cudaEventDestroy(x->sm_join);
  delete x;
  return 0;
}

// Enqueue one exchange on `stream`: on return of the stream work, bufs[rank] holds sum over ranks / world on every rank.
extern "C" int dagb200_grad_exchange(void *handle, void *stream) {
  Xchg *x = (Xchg *)handle;
  if (!x) { set_error("grad_exchange: null handle"); return DAGB200_EINVAL; }
  cudaStream_t st = (cudaStream_t)stream;
  const float inv = 1.f / (float)x->world;
  const size_t lo = (size_t)x->rank * x->slice;
  const size_t n = lo >= x->numel ? 0 : (x->numel - lo < x->slice ? x->numel - lo : x->slice);
  if (x->world == 1) return 0;
  auto mark = [&](int i) { if (x->timing) cudaEventRecord(x->mark[i], st); };
  mark(0);
  if (int rc = pushes(x, st, true)) return rc;
  mark(1);
  barrier(x, st);
  mark(2);
  if (n) {
    const size_t n4 = n / 4;
    int grid = (int)((n4 + 1023) / 1024);
    const int cap = 4 * sm_count();
    if (grid > cap) grid = cap;
    // rows of the staging area are `slice` floats apart; the tail slice may be shorter but the pitch is the same
    xchg_reduce_kernel<<<grid, 256, 0, st>>>((float4 *)(x->buf[x->rank] + lo), (const float4 *)x->staging[x->rank], n4,
                                             x->slice / 4, x->rank, x->world, inv);
  }
  mark(3);
  if (int rc = pushes(x, st, false)) return rc;
  mark(4);
  barrier(x, st);
  mark(5);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "grad_exchange");
}

// Measurement aid: durations (ms) of the five phases of the most recent exchange -- push, barrier, reduce, push,
// barrier.  The first call switches the event recording on and returns -1 in every slot.
extern "C" int dagb200_grad_exchange_phases(void *handle, float *ms5) {
  Xchg *x = (Xchg *)handle;
  if (!x || !ms5) { set_error("grad_exchange_phases: bad arguments"); return DAGB200_EINVAL; }
  if (!x->timing) {
    for (int i = 0; i < 6; i++) cudaEventCreate(&x->mark[i]);
    x->timing = true;
    for (int i = 0; i < 5; i++) ms5[i] = -1.f;
    return 0;
  }
  cudaEventSynchronize(x->mark[5]);
  for (int i = 0; i < 5; i++)
    if (cudaEventElapsedTime(&ms5[i], x->mark[i], x->mark[i + 1]) != cudaSuccess) ms5[i] = -1.f;
  cudaGetLastError();
  return 0;
}

// The in-switch exchange of a buffer that is mapped through a multicast address on every rank (the mapping and the
// barriers around this call are the caller's: see dist.NvlsGradExchange).  Enqueues ONE kernel of `ctas` thread blocks
// on `stream`: this rank's slice of mc[0 .. numel) becomes (sum over ranks) / world on every rank.
extern "C" int dagb200_grad_exchange_nvls(void *mc, size_t numel, int rank, int world, int ctas, void *stream) {
  if (!mc || world < 1 || rank < 0 || rank >= world || numel % 4 || ctas < 1) {
    set_error("grad_exchange_nvls: bad arguments");
    return DAGB200_EINVAL;
  }
  const size_t slice = dagb200_grad_exchange_slice(numel, world);
  const size_t lo = (size_t)rank * slice;
  const size_t n = lo >= numel ? 0 : (numel - lo < slice ? numel - lo : slice);
  if (n) xchg_nvls_kernel<<<ctas, 512, 0, (cudaStream_t)stream>>>((float *)mc, lo / 4, n / 4, 1.f / (float)world);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : cuda_fail(e, "grad_exchange_nvls");
}

// 0 = all barriers so far completed; otherwise the epoch at which a peer failed to arrive (read after a synchronise)
extern "C" int dagb200_grad_exchange_status(void *handle, int *timed_out_epoch) {
  Xchg *x = (Xchg *)handle;
  if (!x || !timed_out_epoch) { set_error("grad_exchange_status: bad arguments"); return DAGB200_EINVAL; }
  cudaError_t e = cudaMemcpy(timed_out_epoch, x->flags[x->rank] + kMaxRanks, sizeof(int), cudaMemcpyDeviceToHost);
  return e == cudaSuccess ? 0 : cuda_fail(e, "grad_exchange_status");
}
