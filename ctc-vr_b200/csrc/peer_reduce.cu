// Data-parallel gradient exchange over NVLink peer memory (replaces the DDP all-reduce the
// reference's train loop gets from torch DistributedDataParallel, rnnt_train.py:60-75).
//
// ONE kernel per step, capturable in the step's CUDA graph (no NCCL call, no host round trip).  All NVLink traffic is
// POSTED STORES - a remote load is a ~2 us round trip per dependent step, a remote store is fire-and-forget:
//   scatter   the rank's gradient tensors (a table of segments) are read once; the part that falls into slice r is stored
//             into rank r's inbox, slot [this rank]
//   barrier A every rank's inbox is complete (flags written into the peers' buffers over NVLink)
//   reduce    rank r sums the N slots of its inbox (local loads, fixed order 0..N-1, so all ranks end with bit-identical
//             sums) and stores the sum into slice r of every rank's result area
//   barrier B every slice of the local result area has been written by its owner
//   unpack    the result area is copied back into the gradient tensors
// A payload of a few MB is latency-bound: NCCL's launch + protocol cost ~75-110 us per step on 2-8 B200s when called
// behind the step graph (profiles/README.md, round 1); this kernel is bounded by two flag round trips plus 2 x 7/8 of the
// payload in posted stores per rank.
//
// No grid-wide synchronisation: float4 unit i of a slice belongs to CTA (i / blockDim) % gridDim on EVERY rank, in all
// three phases, and CTA c only ever synchronises with CTA c of the peers.  Flags are monotonic step counters (no reset,
// replay-safe; a peer can run at most one barrier ahead, which the two areas make safe); waits are bounded and report
// through a mapped host word.
#include <algorithm>

#include "common.cuh"

namespace ctcvr {

namespace {

constexpr int PR_MAX_WORLD = 8;
constexpr int PR_MAX_SEG = 24;
constexpr int PR_MAX_CTAS = 128;
constexpr int PR_THREADS = 512;
constexpr size_t PR_HEADER = 8192;   // flags [PR_MAX_CTAS][PR_MAX_WORLD] u32 | epoch [PR_MAX_CTAS] u32

struct PeerCtx {
  int rank = 0, world = 1, device = 0;
  size_t cap_floats = 0;                 // payload capacity (floats), multiple of 4 * world; the buffer holds it twice
  uint8_t* local = nullptr;
  uint8_t* peer[PR_MAX_WORLD] = {};
  bool imported[PR_MAX_WORLD] = {};
  bool connected = false;
  unsigned int* err_h = nullptr;
  unsigned int* err_d = nullptr;
  long long timeout_ns = 10LL * 1000 * 1000 * 1000;
};

struct PeerArgs {
  uint8_t* peer[PR_MAX_WORLD];
  float* seg_ptr[PR_MAX_SEG];
  long seg_off[PR_MAX_SEG];              // offset in the flat payload (floats, multiple of 4)
  long seg_n[PR_MAX_SEG];                // floats
  int nseg, rank, world;
  long slice4;                           // float4 units per slice
  long result4;                          // float4 offset of the result area (fixed by the capacity: calls of different
                                         // payload sizes must not let one step's inbox overlap the previous step's result)
  unsigned int* err;
  long long timeout_ns;
  long long* prof;                       // dev tool: 8 globaltimer stamps per CTA (ctcvr_debug_set_prof), else NULL
  int mode;                              // dev tool: timing experiments (ctcvr_debug_set_mode bits 2..), 0 in production
};

static void* g_peer_prof = nullptr;
static int g_peer_mode = 0;

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// L1-bypassing loads: the lines are rewritten by other GPUs (and by earlier replays) behind this SM's back
__device__ __forceinline__ float4 ld_sys_f4(const float4* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ long long globaltimer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ unsigned int* flags_of(uint8_t* base, int cta) {
  return reinterpret_cast<unsigned int*>(base) + cta * PR_MAX_WORLD;
}
__device__ __forceinline__ unsigned int* epoch_of(uint8_t* base, int cta) {
  return reinterpret_cast<unsigned int*>(base + PR_MAX_CTAS * PR_MAX_WORLD * 4) + cta;
}
// inbox: [world][slice4] float4 (slot q = rank q's part of MY slice) | result: [world * slice4] float4 (the reduced payload)
__device__ __forceinline__ float4* inbox_of(uint8_t* base) { return reinterpret_cast<float4*>(base + PR_HEADER); }
__device__ __forceinline__ float4* result_of(uint8_t* base, long result4) { return inbox_of(base) + result4; }

// All of the CTA's earlier writes are ordered before the flag stores (bar.sync + release at system scope); the flag of
// this rank is raised in CTA c's row of every peer, then the CTA waits until all N flags of its own row reached `value`.
__device__ __forceinline__ void peer_barrier(const PeerArgs& a, unsigned int value, int site) {
  __syncthreads();
  const int t = threadIdx.x;
  if (t < a.world) {
    // the release store orders the CTA's earlier writes (bar.sync above + cumulativity); an extra __threadfence_system()
    // (fence.sc.sys) in front of it cost ~3 us per barrier (tools/peer_time.py, mode bit 0)
    if (a.mode & 1) __threadfence_system();
    st_release_sys(flags_of(a.peer[t], blockIdx.x) + a.rank, value);
    const unsigned int* mine = flags_of(a.peer[a.rank], blockIdx.x) + t;
    const long long t0 = globaltimer_ns();
    while ((int)(ld_acquire_sys(mine) - value) < 0) {
      if (globaltimer_ns() - t0 > a.timeout_ns) {
        if (a.err) *reinterpret_cast<volatile unsigned int*>(a.err) = 0x80000000u | ((unsigned)site << 16) | ((unsigned)t << 8) | blockIdx.x;
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
}

// Payload floats [4u, 4u+4) live in one segment (offsets and padded lengths are multiples of 4).
struct UnitRef { float* g; long left; bool whole; };   // left = floats of the segment from g on (<= 0: pure padding)
// `s` is the caller's segment cursor: a thread visits its units in increasing order, so the cursor only moves forward
// (a full scan of the table per 16-byte unit made the pack / unpack phases compute-bound at 13 MB payloads).
__device__ __forceinline__ UnitRef locate_unit(const PeerArgs& a, long u, int& s) {
  const long f = u * 4;
  while (s + 1 < a.nseg && f >= a.seg_off[s + 1]) ++s;
  const long j = f - a.seg_off[s];
  UnitRef r;
  r.g = a.seg_ptr[s] + j;
  r.left = a.seg_n[s] - j;
  r.whole = r.left >= 4 && ((reinterpret_cast<uintptr_t>(r.g) & 15u) == 0);
  return r;
}
__device__ __forceinline__ float4 load_unit(const UnitRef& r) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r.whole) return *reinterpret_cast<const float4*>(r.g);
  if (r.left > 0) v.x = r.g[0];
  if (r.left > 1) v.y = r.g[1];
  if (r.left > 2) v.z = r.g[2];
  if (r.left > 3) v.w = r.g[3];
  return v;
}
__device__ __forceinline__ void store_unit(const UnitRef& r, const float4 v) {
  if (r.whole) { *reinterpret_cast<float4*>(r.g) = v; return; }
  if (r.left > 0) r.g[0] = v.x;
  if (r.left > 1) r.g[1] = v.y;
  if (r.left > 2) r.g[2] = v.z;
  if (r.left > 3) r.g[3] = v.w;
}

// scatter (tensors -> the owners' inboxes) / unpack (local result area -> tensors) of this CTA's units of every slice,
// R loads in flight per thread.  Unit (slice r, index i) is flat payload unit r * slice4 + i.
template <bool kScatter>
__device__ __forceinline__ void move_units(const PeerArgs& a, int i0, int stride) {
  constexpr int R = 4;
  const float4* result = result_of(a.peer[a.rank], a.result4);
  int seg = 0;                                   // plain destination order 0..N-1: the units of a thread only increase
  for (int k0 = 0; k0 < a.world; ++k0) {
    // destinations in plain order; the rotated order (rank + 1, rank + 2, ..: every rank on a different peer at any
    // moment) measured no better on NVSwitch (tools/peer_time.py, mode bit 1)
    const int r = (a.mode & 2) ? (a.rank + 1 + k0) % a.world : k0;
    if (a.mode & 2) seg = 0;                     // rotated order (experiment switch): the cursor restarts per slice
    const long base = (long)r * a.slice4;
    float4* inbox = inbox_of(a.peer[r]) + (long)a.rank * a.slice4;
    for (long i = i0; i < a.slice4; i += (long)R * stride) {
      float4 v[R];
      UnitRef ref[R];
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const long ik = i + (long)k * stride;
        if (ik < a.slice4) {
          ref[k] = locate_unit(a, base + ik, seg);
          v[k] = kScatter ? load_unit(ref[k]) : ld_sys_f4(result + base + ik);
        }
      }
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const long ik = i + (long)k * stride;
        if (ik < a.slice4) {
          if (kScatter) inbox[ik] = v[k];
          else store_unit(ref[k], v[k]);
        }
      }
    }
  }
}

// Sum of the N inbox slots of this rank's slice (local, L1-bypassing: the peers wrote them), stored to every rank's result area.
template <int N>
__device__ __forceinline__ void reduce_slice(const PeerArgs& a, int i0, int stride) {
  constexpr int R = N <= 2 ? 4 : (N <= 4 ? 2 : 1);
  const float4* inbox = inbox_of(a.peer[a.rank]);
  const long base = (long)a.rank * a.slice4;
  for (long i = i0; i < a.slice4; i += (long)R * stride) {
    float4 v[R][N];
#pragma unroll
    for (int k = 0; k < R; ++k)
      if (i + (long)k * stride < a.slice4) {
#pragma unroll
        for (int q = 0; q < N; ++q) v[k][q] = ld_sys_f4(inbox + (long)q * a.slice4 + i + (long)k * stride);
      }
#pragma unroll
    for (int k = 0; k < R; ++k)
      if (i + (long)k * stride < a.slice4) {
        float4 s = v[k][0];
#pragma unroll
        for (int q = 1; q < N; ++q) { s.x += v[k][q].x; s.y += v[k][q].y; s.z += v[k][q].z; s.w += v[k][q].w; }
#pragma unroll
        for (int q0 = 0; q0 < N; ++q0) {
          const int q = (a.mode & 2) ? (a.rank + 1 + q0) % N : q0;
          result_of(a.peer[q], a.result4)[base + i + (long)k * stride] = s;
        }
      }
  }
}

__global__ void __launch_bounds__(PR_THREADS, 1) peer_allreduce_kernel(const PeerArgs a) {
  __shared__ unsigned int s_epoch;
  uint8_t* local = a.peer[a.rank];
  if (threadIdx.x == 0) s_epoch = *epoch_of(local, blockIdx.x);
  __syncthreads();
  const unsigned int e = s_epoch;
  const int i0 = blockIdx.x * PR_THREADS + threadIdx.x, stride = gridDim.x * PR_THREADS;
#define PEER_STAMP(k) do { if (a.prof && threadIdx.x == 0) a.prof[blockIdx.x * 8 + (k)] = globaltimer_ns(); } while (0)

  PEER_STAMP(0);
  move_units<true>(a, i0, stride);
  PEER_STAMP(1);
  peer_barrier(a, 2u * e + 1u, 1);
  PEER_STAMP(2);
  switch (a.world) {
    case 2: reduce_slice<2>(a, i0, stride); break;
    case 3: reduce_slice<3>(a, i0, stride); break;
    case 4: reduce_slice<4>(a, i0, stride); break;
    case 5: reduce_slice<5>(a, i0, stride); break;
    case 6: reduce_slice<6>(a, i0, stride); break;
    case 7: reduce_slice<7>(a, i0, stride); break;
    case 8: reduce_slice<8>(a, i0, stride); break;
    default: break;
  }
  PEER_STAMP(3);
  peer_barrier(a, 2u * e + 2u, 2);
  PEER_STAMP(4);
  move_units<false>(a, i0, stride);
  __syncthreads();
  PEER_STAMP(5);
  if (threadIdx.x == 0) *epoch_of(local, blockIdx.x) = e + 1u;
}

}  // namespace

int peer_create(int rank, int world, size_t max_floats, void** out_ctx, void* handle64) {
  CTCVR_REQUIRE(world >= 1 && world <= PR_MAX_WORLD && rank >= 0 && rank < world, "peer_create: world must be 1..%d and 0 <= rank < world (got rank %d of %d)", PR_MAX_WORLD, rank, world);
  CTCVR_REQUIRE(out_ctx && handle64 && max_floats > 0, "peer_create: NULL pointer or empty payload");
  PeerCtx* c = new PeerCtx();
  c->rank = rank; c->world = world;
  // every segment is padded to a multiple of 4 floats (<= 3 per segment) and the total to a multiple of 4 * world
  c->cap_floats = align_up(max_floats + 4 * PR_MAX_SEG, (size_t)4 * world);
  const size_t bytes = PR_HEADER + 2 * c->cap_floats * 4;      // inbox + result area
  if (cudaGetDevice(&c->device) != cudaSuccess || cudaMalloc(reinterpret_cast<void**>(&c->local), bytes) != cudaSuccess ||
      cudaMemset(c->local, 0, bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
    set_error("peer_create: cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
    delete c;
    return 1;
  }
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, c->local) != cudaSuccess) {
    set_error("peer_create: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(c->local);
    delete c;
    return 1;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  memcpy(handle64, &h, 64);
  if (cudaHostAlloc(reinterpret_cast<void**>(&c->err_h), sizeof(unsigned int), cudaHostAllocMapped) == cudaSuccess) {
    *c->err_h = 0u;
    if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&c->err_d), c->err_h, 0) != cudaSuccess) { c->err_d = nullptr; cudaGetLastError(); }
  } else { c->err_h = nullptr; cudaGetLastError(); }
  c->peer[rank] = c->local;
  *out_ctx = c;
  return 0;
}

// handles: world x 64 bytes (cudaIpcMemHandle_t of every rank, own slot ignored) when `local_ptrs` is NULL; otherwise
// local_ptrs[q] is rank q's buffer already addressable from this process (ranks that share a process, as in the tests).
int peer_connect(void* ctx, const void* handles, void* const* local_ptrs) {
  PeerCtx* c = static_cast<PeerCtx*>(ctx);
  CTCVR_REQUIRE(c && (handles || local_ptrs), "peer_connect: NULL pointer");
  for (int q = 0; q < c->world; ++q) {
    if (q == c->rank) continue;
    if (local_ptrs) {
      CTCVR_REQUIRE(local_ptrs[q], "peer_connect: NULL buffer for rank %d", q);
      c->peer[q] = static_cast<uint8_t*>(local_ptrs[q]);
    } else {
      cudaIpcMemHandle_t h;
      memcpy(&h, static_cast<const uint8_t*>(handles) + (size_t)q * 64, 64);
      void* p = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        set_error("peer_connect: cudaIpcOpenMemHandle for rank %d failed: %s (the ranks must be processes of one node with NVLink / PCIe peer access)", q, cudaGetErrorString(e));
        cudaGetLastError();
        return 1;
      }
      c->peer[q] = static_cast<uint8_t*>(p);
      c->imported[q] = true;
    }
  }
  c->connected = true;
  return 0;
}

void* peer_local_buffer(void* ctx) { return ctx ? static_cast<PeerCtx*>(ctx)->local : nullptr; }

int peer_set_timeout_ms(void* ctx, long ms) {
  PeerCtx* c = static_cast<PeerCtx*>(ctx);
  CTCVR_REQUIRE(c && ms > 0, "peer_set_timeout_ms: bad argument");
  c->timeout_ns = (long long)ms * 1000000LL;
  return 0;
}

int peer_allreduce(void* ctx, void* const* seg_ptrs, const long* seg_floats, int nseg, int ctas, cudaStream_t st) {
  PeerCtx* c = static_cast<PeerCtx*>(ctx);
  CTCVR_REQUIRE(c && seg_ptrs && seg_floats, "peer_allreduce: NULL pointer");
  CTCVR_REQUIRE(nseg >= 1 && nseg <= PR_MAX_SEG, "peer_allreduce: 1..%d gradient tensors per call (got %d)", PR_MAX_SEG, nseg);
  if (c->err_h && *reinterpret_cast<volatile unsigned int*>(c->err_h) != 0u) {
    const unsigned int code = *c->err_h;
    *c->err_h = 0u;
    set_error("peer_allreduce: an earlier exchange timed out waiting for a peer (flag 0x%08x: barrier %u, peer rank %u, CTA %u); its sums are invalid",
              code, (code >> 16) & 0x7fu, (code >> 8) & 0xffu, code & 0xffu);
    return 1;
  }
  if (c->world == 1) return 0;
  CTCVR_REQUIRE(c->connected, "peer_allreduce: peer_connect has not been called");
  PeerArgs a{};
  long off = 0;
  for (int i = 0; i < nseg; ++i) {
    CTCVR_REQUIRE(seg_ptrs[i] && seg_floats[i] > 0 && (reinterpret_cast<uintptr_t>(seg_ptrs[i]) & 3u) == 0, "peer_allreduce: segment %d is empty or misaligned", i);
    a.seg_ptr[i] = static_cast<float*>(seg_ptrs[i]);
    a.seg_off[i] = off;
    a.seg_n[i] = seg_floats[i];
    off += (long)align_up((size_t)seg_floats[i], 4);
  }
  // trailing pad units up to a multiple of 4 * world belong to the last segment (beyond its n: packed as zeros, not unpacked)
  const long total = (long)align_up((size_t)off, (size_t)4 * c->world);
  CTCVR_REQUIRE((size_t)total <= c->cap_floats, "peer_allreduce: payload of %ld floats exceeds the %zu the exchange was created for", total, c->cap_floats);
  a.nseg = nseg; a.rank = c->rank; a.world = c->world;
  a.slice4 = total / 4 / c->world;
  a.result4 = (long)(c->cap_floats / 4);
  for (int q = 0; q < c->world; ++q) a.peer[q] = c->peer[q];
  a.err = c->err_d;
  a.timeout_ns = c->timeout_ns;
  a.prof = static_cast<long long*>(g_peer_prof);
  a.mode = g_peer_mode;
  if (ctas <= 0) ctas = 128;
  ctas = std::min(ctas, PR_MAX_CTAS);
  peer_allreduce_kernel<<<ctas, PR_THREADS, 0, st>>>(a);
  CTCVR_LAUNCH_CHECK();
  return 0;
}

void peer_set_prof(void* buf) { g_peer_prof = buf; }
void peer_set_mode(int mode) { g_peer_mode = mode; }

int peer_destroy(void* ctx) {
  PeerCtx* c = static_cast<PeerCtx*>(ctx);
  if (!c) return 0;
  cudaDeviceSynchronize();
  for (int q = 0; q < c->world; ++q)
    if (c->imported[q] && c->peer[q]) cudaIpcCloseMemHandle(c->peer[q]);
  if (c->local) cudaFree(c->local);
  if (c->err_h) cudaFreeHost(c->err_h);
  cudaGetLastError();
  delete c;
  return 0;
}

}  // namespace ctcvr
