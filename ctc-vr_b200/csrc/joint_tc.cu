// bf16 tcgen05 path of the fused transducer joint (the performance path on B200).
//
//   z      = tanh(enc_proj[b,t,:] + pred_proj[b,u,:])        model/component/joint.py:57-67
//   logits = z . W_out^T + b_out                             model/component/joint.py:68
//   fwd   : per lattice cell lse / lp_blank / lp_label (log-softmax + gather of
//           torchaudio.functional.rnnt_loss, model/component/transducer.py:180-187)
//   bwd   : recompute logits, closed-form d cost/d logits, dZ = g.W, dH = dZ*(1-z^2),
//           d_enc = sum_u dH, d_pred = sum_t dH, dW = g^T.z, db = sum g
//
// Structure (one persistent CTA per SM, 512 threads, warp-specialised):
//   warp 0      TMA producer : streams bf16 W_out k-blocks (128B-swizzled, K-major) into a smem ring
//   warp 1      MMA issuer   : one elected thread issues tcgen05.mma (M=128, N=Vp/2 x2, K=16), fp32
//                              accumulators live in TMEM, commits release smem stages / signal the epilogue
//   warp 2      TMEM allocator
//   warps 4-7   epilogue     : tcgen05.ld (thread = lattice cell), online log-softmax, gathers
//   warps 8-15  A producers  : the A operand is COMPUTED, not loaded: tanh(e+p) -> bf16 -> smem in the
//                              UMMA canonical swizzled layout, fence.proxy.async, mbarrier arrive
// The [B,T,U+1,V] logits only ever exist as TMEM tiles.
#include <algorithm>
#include <stdlib.h>

#include "tc_common.cuh"

namespace ctcvr {
namespace tc {

constexpr int BM = 128;            // lattice cells (rows) per tile
constexpr int BK = 64;             // k-block: 64 bf16 = 128 B = one swizzle row
constexpr int A_STAGES = 3;
constexpr int W_STAGES = 5;
constexpr int NTHREADS = 512;
constexpr int A_STAGE_BYTES = BM * BK * 2;        // 16 KB
constexpr int SLAB_PITCH = 68;                    // floats; 272 B row pitch -> conflict-free LDS.128
constexpr int SLAB_ROWS_FLAT = 132;               // >= distinct u + distinct t of any 128-cell run (U1 <= 128)
constexpr int SLAB_ROWS_RECT = 24;                // 16 u + 8 t
constexpr int PROD_THREADS = 256;
constexpr int TMEM_COLS = 512;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

enum { TILE_FLAT = 0, TILE_RECT = 1 };
constexpr int PROF_CAP = 2048;
// timeline probe for CTA 0 (one lane per role); compiled in always, active only when p.prof != nullptr
#define TC_PROF(role, tag)                                                                         \
  do {                                                                                             \
    if (p.prof != nullptr && blockIdx.x == 0) {                                                    \
      int _n = prof_n++;                                                                           \
      if (_n < PROF_CAP) p.prof[(role)*PROF_CAP + _n] = ((long long)(tag) << 48) | (clock64() & 0xffffffffffffLL); \
    }                                                                                              \
  } while (0)


struct TcParams {
  const float* enc;        // [B,T,D] fp32
  const float* pred;       // [B,U1,D] fp32
  const float* bias_pad;   // [Vp] fp32, -inf beyond V
  const int32_t* targets;  // [B,U1-1]
  const int32_t* t_len;
  const int32_t* u_len;
  const int4* tiles;       // tile table: flat {b, c0, 0, 0} ; rect {b, s, tb, row0/128}
  const int* ntiles;
  int B, T, U1, D, V, Vp, NH, blank;
  // forward outputs
  float* lse;
  float* lp_blank;
  float* lp_label;
  // backward inputs
  const float* lse_in;
  const float* alpha;
  const float* beta;
  const float* costs;
  const float* grad_costs;
  float clamp;
  // backward outputs / scratch
  __nv_bfloat16* zt;       // [D][Rpad]   z^T spill (K-major operand of the dW GEMM)
  __nv_bfloat16* gt;       // [Vp][Rpad]  g^T spill
  long Rpad;
  float* d_enc_part;       // [S][B,T,D] partial d_enc per u-split
  float* d_pred;           // [B,U1,D] (atomic accumulate)
  float* d_bias;           // [Vp] (atomic accumulate)
  int S_max;
  long long* prof;   // optional timeline buffer (debug): [4 roles][PROF_CAP] of (tag<<48 | clock)
};

struct RowMap {
  int b, Tb, Ub, W;
  int t0;        // first frame covered by the tile
  int ubase;     // first label column covered by the tile's pred slab rows
  int np, ne;    // slab rows holding pred / enc vectors
};

// Opaque identity: stops the compiler from re-deriving a per-tile value from global memory inside the
// k-block loops (it otherwise rematerialises the tile-table / length loads there, a chain of two L2
// round trips in front of every operand fetch).
__device__ __forceinline__ void pin(int& x) { asm volatile("" : "+r"(x)); }

// ---- tile geometry ---------------------------------------------------------------------------
template <int TILE>
__device__ __forceinline__ RowMap tile_geometry(const int32_t* t_len, const int32_t* u_len, int T, int U1, int4 ti) {
  RowMap g;
  g.b = ti.x;
  g.Tb = min(t_len[g.b], T);
  g.Ub = min(u_len[g.b], U1 - 1);
  g.W = g.Ub + 1;
  if (TILE == TILE_FLAT) {
    int ncell = g.Tb * g.W;
    int c0 = ti.y;
    int clast = min(c0 + BM - 1, ncell - 1);
    g.t0 = c0 / g.W;
    g.ubase = 0;
    g.np = g.W;
    g.ne = clast / g.W - g.t0 + 1;
  } else {
    int S = (g.W + 15) >> 4;
    int us = (g.W + S - 1) / S;
    g.t0 = ti.z * 8;
    g.ubase = ti.y * us;
    g.np = 16;
    g.ne = 8;
  }
  pin(g.b); pin(g.Tb); pin(g.Ub); pin(g.W); pin(g.t0); pin(g.ubase); pin(g.np); pin(g.ne);
  return g;
}
template <int TILE>
__device__ __forceinline__ RowMap tile_geometry(const TcParams& p, int4 ti) {
  return tile_geometry<TILE>(p.t_len, p.u_len, p.T, p.U1, ti);
}

// row r of the tile -> lattice cell; returns false for padding rows
template <int TILE>
__device__ __forceinline__ bool row_cell(const RowMap& g, int4 ti, int r, int& t, int& u) {
  if (TILE == TILE_FLAT) {
    int c = ti.y + r;
    t = c / g.W;
    u = c - t * g.W;
    return c < g.Tb * g.W;
  } else {
    int S = (g.W + 15) >> 4;
    int us = (g.W + S - 1) / S;
    int tloc = r >> 4, ul = r & 15;
    t = g.t0 + tloc;
    u = g.ubase + ul;
    return t < g.Tb && ul < us && u <= g.Ub;
  }
}
template <int TILE>
__device__ __forceinline__ bool row_cell(const TcParams&, const RowMap& g, int4 ti, int r, int& t, int& u) {
  return row_cell<TILE>(g, ti, r, t, u);
}

struct SmemLayout {
  uint32_t a_base, w_base, w_bytes, bar_base;
  float* slab;
  float* bias;
  uint32_t* tmem_ptr;
  // addresses are computed, not tabulated: indexing a table with the runtime stage id would put it in local memory
  __device__ __forceinline__ uint32_t a_stage(int i) const { return a_base + i * A_STAGE_BYTES; }
  __device__ __forceinline__ uint32_t w_stage(int i) const { return w_base + i * w_bytes; }
  __device__ __forceinline__ uint32_t a_full(int i) const { return bar_base + i * 16; }
  __device__ __forceinline__ uint32_t a_empty(int i) const { return bar_base + i * 16 + 8; }
  __device__ __forceinline__ uint32_t w_full(int i) const { return bar_base + A_STAGES * 16 + i * 16; }
  __device__ __forceinline__ uint32_t w_empty(int i) const { return bar_base + A_STAGES * 16 + i * 16 + 8; }
  __device__ __forceinline__ uint32_t tmem_full() const { return bar_base + (A_STAGES + W_STAGES) * 16; }
  __device__ __forceinline__ uint32_t tmem_empty() const { return bar_base + (A_STAGES + W_STAGES) * 16 + 8; }
  __device__ __forceinline__ uint32_t aux_bar(int i) const { return bar_base + (A_STAGES + W_STAGES) * 16 + 16 + i * 8; }
};

__host__ __device__ inline size_t tc_smem_bytes(int NH, int Vp, int slab_rows) {
  size_t s = 1024;                                     // alignment slack
  s += (size_t)A_STAGES * A_STAGE_BYTES;
  s += (size_t)W_STAGES * NH * 128;
  s += (size_t)slab_rows * SLAB_PITCH * 4;
  s += (size_t)Vp * 4;
  s += 512;                                            // barriers + tmem ptr
  return s;
}

__device__ __forceinline__ void carve_smem(SmemLayout& L, uint8_t* raw, int NH, int Vp, int slab_rows) {
  uint32_t base = smem_u32(raw);
  uint32_t aligned = (base + 1023u) & ~1023u;
  uint32_t a = aligned;
  L.a_base = a; a += A_STAGES * A_STAGE_BYTES;
  L.w_base = a; L.w_bytes = NH * 128; a += W_STAGES * NH * 128;
  L.slab = reinterpret_cast<float*>(raw + (a - base)); a += slab_rows * SLAB_PITCH * 4;
  L.bias = reinterpret_cast<float*>(raw + (a - base)); a += Vp * 4;
  a = (a + 15u) & ~15u;
  L.bar_base = a; a += (A_STAGES + W_STAGES) * 16 + 16 + 64;
  L.tmem_ptr = reinterpret_cast<uint32_t*>(raw + (a - base));
}

// ---- A producer: slab staging + tanh tile -----------------------------------------------------
// Slab rows [0,np) hold pred_proj[b][ubase+i][k0..k0+64), rows [np,np+ne) hold enc_proj[b][t0+i][k0..k0+64).
__device__ __forceinline__ const float* slab_src(const TcParams& p, const RowMap& g, int srow) {
  if (srow < g.np) {
    int u = g.ubase + srow;
    return (u <= g.Ub) ? p.pred + ((size_t)g.b * p.U1 + u) * p.D : nullptr;
  }
  int t = g.t0 + srow - g.np;
  return (t < g.Tb) ? p.enc + ((size_t)g.b * p.T + t) * p.D : nullptr;
}

template <int NPRE>
__device__ __forceinline__ void slab_fetch(const TcParams& p, const RowMap& g, int k0, int pt, float4 (&pre)[NPRE]) {
  const int n4 = (g.np + g.ne) * 16;
#pragma unroll
  for (int i = 0; i < NPRE; ++i) {
    int idx = pt + i * PROD_THREADS;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (idx < n4) {
      const float* src = slab_src(p, g, idx >> 4);
      if (src) v = __ldg(reinterpret_cast<const float4*>(src + k0) + (idx & 15));
    }
    pre[i] = v;
  }
}
template <int NPRE>
__device__ __forceinline__ void slab_store(float* slab, const RowMap& g, int pt, const float4 (&pre)[NPRE]) {
  const int n4 = (g.np + g.ne) * 16;
#pragma unroll
  for (int i = 0; i < NPRE; ++i) {
    int idx = pt + i * PROD_THREADS;
    if (idx < n4) *reinterpret_cast<float4*>(slab + (idx >> 4) * SLAB_PITCH + (idx & 15) * 4) = pre[i];
  }
}

// One k-block of the A tile for row r, k sub-range khalf*32..+32: tanh(e+p) -> bf16 -> swizzled smem.
// Optionally spills z^T (bf16, [D][Rpad]) for the dW GEMM.
template <bool SPILL>
__device__ __forceinline__ void produce_a(const float* slab, uint32_t a_stage, int r, int khalf, bool valid,
                                          int prow, int erow, __nv_bfloat16* zt_col, long Rpad) {
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    const int chunk = khalf * 4 + ch;                 // 16-byte chunk within the 128-byte row
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    if (valid) {
      const float4 p0 = *reinterpret_cast<const float4*>(slab + prow * SLAB_PITCH + chunk * 8);
      const float4 p1 = *reinterpret_cast<const float4*>(slab + prow * SLAB_PITCH + chunk * 8 + 4);
      const float4 e0 = *reinterpret_cast<const float4*>(slab + erow * SLAB_PITCH + chunk * 8);
      const float4 e1 = *reinterpret_cast<const float4*>(slab + erow * SLAB_PITCH + chunk * 8 + 4);
      w[0] = pack_bf16(tanh_fast(e0.x + p0.x), tanh_fast(e0.y + p0.y));
      w[1] = pack_bf16(tanh_fast(e0.z + p0.z), tanh_fast(e0.w + p0.w));
      w[2] = pack_bf16(tanh_fast(e1.x + p1.x), tanh_fast(e1.y + p1.y));
      w[3] = pack_bf16(tanh_fast(e1.z + p1.z), tanh_fast(e1.w + p1.w));
    }
    const uint32_t dst = a_stage + r * 128 + ((chunk ^ (r & 7)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
                 : "memory");
    if (SPILL) {
      // zt_col points at zt[k0 + khalf*32][row]; lanes hold consecutive rows -> 64 B coalesced per k
      unsigned short* z = reinterpret_cast<unsigned short*>(zt_col) + (size_t)(ch * 8) * Rpad;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        z[(size_t)(2 * j) * Rpad] = (unsigned short)(w[j] & 0xffffu);
        z[(size_t)(2 * j + 1) * Rpad] = (unsigned short)(w[j] >> 16);
      }
    }
  }
}

struct Pipe {
  int stage = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void advance(int n) {
    if (++stage == n) { stage = 0; phase ^= 1u; }
  }
};

}  // namespace tc
}  // namespace ctcvr
#include "joint_tc_fwd.cuh"
#include "joint_tc_bwd.cuh"
namespace ctcvr {
namespace tc {

// =================================================================================================
// Forward kernel (v1, kept for A/B runs: CTCVR_FWD_V1=1)
// =================================================================================================
__global__ void __launch_bounds__(NTHREADS, 1)
joint_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_w, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  SmemLayout L;
  carve_smem(L, smem_raw, p.NH, p.Vp, SLAB_ROWS_FLAT);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KB = p.D / BK;
  const int ntiles = *p.ntiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    for (int i = 0; i < A_STAGES; ++i) { mbar_init(L.a_full(i), PROD_THREADS); mbar_init(L.a_empty(i), 1); }
    for (int i = 0; i < W_STAGES; ++i) { mbar_init(L.w_full(i), 1); mbar_init(L.w_empty(i), 1); }
    mbar_init(L.tmem_full(), 1);
    mbar_init(L.tmem_empty(), 128);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(L.tmem_ptr), TMEM_COLS);
  for (int i = tid; i < p.Vp; i += NTHREADS) L.bias[i] = p.bias_pad[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *L.tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (W_out)
    if (lane == 0) {
      Pipe wp;
      int prof_n = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
        for (int kb = 0; kb < KB; ++kb)
          for (int h = 0; h < 2; ++h) {
            mbar_wait(L.w_empty(wp.stage), wp.phase ^ 1u, 1);
            TC_PROF(0, kb * 2 + h);
            mbar_arrive_expect_tx(L.w_full(wp.stage), (uint32_t)p.NH * 128u);
            tma_load_2d(L.w_stage(wp.stage), &tmap_w, L.w_full(wp.stage), kb * BK, h * p.NH);
            wp.advance(W_STAGES);
          }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      Pipe ap, wp;
      uint32_t tphase = 0;
      int prof_n = 0;
      const uint32_t idesc = make_idesc_bf16(BM, p.NH);
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        TC_PROF(1, 100);
        mbar_wait(L.tmem_empty(), tphase ^ 1u, 2);
        TC_PROF(1, 101);
        tc_fence_after();
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(L.a_full(ap.stage), ap.phase, 3);
          TC_PROF(1, kb);
          for (int h = 0; h < 2; ++h) {
            mbar_wait(L.w_full(wp.stage), wp.phase, 4);
            TC_PROF(1, 50 + kb * 2 + h);
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < BK / 16; ++ks) {
              uint64_t ad = make_desc_sw128(L.a_stage(ap.stage) + ks * 32);
              uint64_t bd = make_desc_sw128(L.w_stage(wp.stage) + ks * 32);
              umma_bf16(tmem_base + h * p.NH, ad, bd, idesc, (kb | ks) ? 1u : 0u);
            }
            umma_commit(L.w_empty(wp.stage));
            wp.advance(W_STAGES);
          }
          umma_commit(L.a_empty(ap.stage));
          ap.advance(A_STAGES);
        }
        umma_commit(L.tmem_full());
        tphase ^= 1u;
      }
    }
    __syncwarp();
  } else if (warp >= 4 && warp < 8) {
    // ------------------------------------------------------------------ epilogue: online log-softmax
    const int q = warp & 3;
    const int r = q * 32 + lane;
    uint32_t tphase = 0;
    int prof_n = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      int4 ti = p.tiles[tile];
      pin(ti.x); pin(ti.y); pin(ti.z); pin(ti.w);
      const RowMap g = tile_geometry<TILE_FLAT>(p, ti);
      int t, u;
      const bool valid = row_cell<TILE_FLAT>(p, g, ti, r, t, u);
      int lab = -1;
      if (valid && u < g.Ub) lab = p.targets[(size_t)g.b * (p.U1 - 1) + u];
      mbar_wait(L.tmem_full(), tphase, 5);
      if (tid == 128) TC_PROF(2, 1);
      tc_fence_after();
      float m = kNegInf, s = 0.f, xb = 0.f, xl = 0.f;
      for (int c0 = 0; c0 < p.Vp; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c0, v);
        tmem_ld_wait();
        float cm[4] = {kNegInf, kNegInf, kNegInf, kNegInf};
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 bj = *reinterpret_cast<const float4*>(L.bias + c0 + j);
          v[j] += bj.x; v[j + 1] += bj.y; v[j + 2] += bj.z; v[j + 3] += bj.w;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            cm[e] = fmaxf(cm[e], v[j + e]);
            xb = (c0 + j + e == p.blank) ? v[j + e] : xb;
            xl = (c0 + j + e == lab) ? v[j + e] : xl;
          }
        }
        const float nm = fmaxf(m, fmaxf(fmaxf(cm[0], cm[1]), fmaxf(cm[2], cm[3])));
        const float nml = nm * LOG2E;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 32; j += 4)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[e] += ex2_fast(fmaf(v[j + e], LOG2E, -nml));
        s = s * ex2_fast((m - nm) * LOG2E) + ((acc[0] + acc[1]) + (acc[2] + acc[3]));
        m = nm;
      }
      tc_fence_before();
      mbar_arrive(L.tmem_empty());
      if (tid == 128) TC_PROF(2, 2);
      if (valid) {
        const size_t cell = ((size_t)g.b * p.T + t) * p.U1 + u;
        const float l = m + __logf(s);
        p.lse[cell] = l;
        p.lp_blank[cell] = xb - l;
        p.lp_label[cell] = (lab >= 0) ? xl - l : kNegInf;
      }
      tphase ^= 1u;
    }
  } else if (warp >= 8) {
    // ------------------------------------------------------------------ A producers
    const int pt = tid - 256;
    const int r = pt & 127, khalf = pt >> 7;
    Pipe ap;
    int prof_n = 0;
    constexpr int NPRE = (SLAB_ROWS_FLAT * 16 + PROD_THREADS - 1) / PROD_THREADS;   // 9
    float4 pre[NPRE];
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      int4 ti = p.tiles[tile];
      pin(ti.x); pin(ti.y); pin(ti.z); pin(ti.w);
      const RowMap g = tile_geometry<TILE_FLAT>(p, ti);
      int t, u;
      const bool valid = row_cell<TILE_FLAT>(p, g, ti, r, t, u);
      const int prow = u - g.ubase, erow = g.np + (t - g.t0);
      if (pt == 0) TC_PROF(3, 200);
      slab_fetch<NPRE>(p, g, 0, pt, pre);
      named_barrier_sync(1, PROD_THREADS);             // previous tile's last slab fully consumed
      slab_store<NPRE>(L.slab, g, pt, pre);
      named_barrier_sync(1, PROD_THREADS);
      if (pt == 0) TC_PROF(3, 201);
      for (int kb = 0; kb < KB; ++kb) {
        if (kb + 1 < KB) slab_fetch<NPRE>(p, g, (kb + 1) * BK, pt, pre);
        mbar_wait(L.a_empty(ap.stage), ap.phase ^ 1u, 6);
        if (pt == 0) TC_PROF(3, kb);
        produce_a<false>(L.slab, L.a_stage(ap.stage), r, khalf, valid, prow, erow, nullptr, 0);
        if (pt == 0) TC_PROF(3, 20 + kb);
        fence_proxy_async();
        mbar_arrive(L.a_full(ap.stage));
        if (pt == 0) TC_PROF(3, 40 + kb);
        ap.advance(A_STAGES);
        if (kb + 1 < KB) {
          named_barrier_sync(1, PROD_THREADS);
          if (pt == 0) TC_PROF(3, 60 + kb);
          slab_store<NPRE>(L.slab, g, pt, pre);
          named_barrier_sync(1, PROD_THREADS);
          if (pt == 0) TC_PROF(3, 80 + kb);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

// =================================================================================================
// Backward kernel 1: recompute logits -> g = d cost/d logits -> dZ^T = W^T g^T -> dH -> d_enc / d_pred
// Tiles are rectangles of 8 frames x 16 label columns (row = tloc*16 + ul) so that, with the transposed
// GEMM (TMEM lane = joint dim d, TMEM column = tile row), both reductions are thread-local and static:
//   d_enc[t]  = sum over the 16 columns of one tcgen05.ld.x16
//   d_pred[u] = sum over the 8 frame slots, kept in registers across the tiles of one (b, u-split) sweep
// Phases of one tile (TMEM is 512 columns, so logits [128 x Vp] and dZ^T [D x 128] cannot coexist):
//   P1 producers: tanh tile (+ z^T spill) | TMA: W_out k-blocks | MMA: logits -> TMEM
//   P2 epilogue : TMEM -> g (bf16) -> smem G tile (K-major over v) + g^T spill
//   P3 TMA: W_out^T tiles | MMA: dZ^T[mb] = W^T[mb] . G^T -> TMEM ; producers: column sums of G (d_bias)
//   P4 epilogue : TMEM -> dH = dZ*(1-z^2) -> d_enc partial (store), d_pred (registers)
// The smem of P2-P4 (G tile, W^T ring) overlays the smem of P1 (A ring, W ring, slab).
// =================================================================================================
constexpr int WT_STAGES = 6;
constexpr int WT_STAGE_BYTES = 128 * 128;          // [128 d][64 v] bf16

struct BwdSmem {
  uint32_t base;          // 1024-aligned
  uint32_t a_base, w_base, w_bytes, g_base, wt_base, bar_base;
  float* slab;
  float* bias;
  uint32_t* tmem_ptr;
  __device__ __forceinline__ uint32_t a_stage(int i) const { return a_base + i * A_STAGE_BYTES; }
  __device__ __forceinline__ uint32_t w_stage(int i) const { return w_base + i * w_bytes; }
  __device__ __forceinline__ uint32_t g_kblock(int i) const { return g_base + i * A_STAGE_BYTES; }
  __device__ __forceinline__ uint32_t wt_stage(int i) const { return wt_base + i * WT_STAGE_BYTES; }
  __device__ __forceinline__ uint32_t a_full(int i) const { return bar_base + i * 16; }
  __device__ __forceinline__ uint32_t a_empty(int i) const { return bar_base + i * 16 + 8; }
  __device__ __forceinline__ uint32_t w_full(int i) const { return bar_base + A_STAGES * 16 + i * 16; }
  __device__ __forceinline__ uint32_t w_empty(int i) const { return bar_base + A_STAGES * 16 + i * 16 + 8; }
  __device__ __forceinline__ uint32_t wt_full(int i) const { return bar_base + (A_STAGES + W_STAGES) * 16 + i * 16; }
  __device__ __forceinline__ uint32_t wt_empty(int i) const { return bar_base + (A_STAGES + W_STAGES) * 16 + i * 16 + 8; }
  __device__ __forceinline__ uint32_t misc(int i) const { return bar_base + (A_STAGES + W_STAGES + WT_STAGES) * 16 + i * 8; }
  __device__ __forceinline__ uint32_t tmem_full() const { return misc(0); }    // logits complete
  __device__ __forceinline__ uint32_t g_full() const { return misc(1); }       // G tile written (128 arrivals)
  __device__ __forceinline__ uint32_t dz_full() const { return misc(2); }      // dZ^T complete, smem free again
  __device__ __forceinline__ uint32_t tmem_empty() const { return misc(3); }   // epilogue done with TMEM (128)
};

__host__ __device__ inline size_t bwd_smem_bytes(int NH, int Vp) {
  size_t r1 = (size_t)A_STAGES * A_STAGE_BYTES + (size_t)W_STAGES * NH * 128 + (size_t)SLAB_ROWS_RECT * SLAB_PITCH * 4;
  size_t kbg = (Vp + 63) / 64;
  size_t r2 = kbg * A_STAGE_BYTES + (size_t)WT_STAGES * WT_STAGE_BYTES;
  size_t s = 1024 + (r1 > r2 ? r1 : r2);
  s = (s + 15) / 16 * 16;
  s += (size_t)Vp * 4 + 512;
  return s;
}

__device__ __forceinline__ void carve_bwd(BwdSmem& L, uint8_t* raw, int NH, int Vp) {
  uint32_t base = smem_u32(raw);
  uint32_t al = (base + 1023u) & ~1023u;
  L.base = al;
  L.a_base = al;
  L.w_base = al + A_STAGES * A_STAGE_BYTES;
  L.w_bytes = NH * 128;
  uint32_t slab_off = A_STAGES * A_STAGE_BYTES + W_STAGES * NH * 128;
  L.slab = reinterpret_cast<float*>(raw + (al - base) + slab_off);
  uint32_t r1 = slab_off + SLAB_ROWS_RECT * SLAB_PITCH * 4;
  uint32_t kbg = (Vp + 63) / 64;
  L.g_base = al;
  L.wt_base = al + kbg * A_STAGE_BYTES;
  uint32_t r2 = kbg * A_STAGE_BYTES + WT_STAGES * WT_STAGE_BYTES;
  uint32_t a = al + (r1 > r2 ? r1 : r2);
  a = (a + 15u) & ~15u;
  L.bias = reinterpret_cast<float*>(raw + (a - base));
  a += Vp * 4;
  a = (a + 15u) & ~15u;
  L.bar_base = a;
  a += (A_STAGES + W_STAGES + WT_STAGES) * 16 + 4 * 8;
  L.tmem_ptr = reinterpret_cast<uint32_t*>(raw + (a - base));
}

__global__ void __launch_bounds__(NTHREADS, 1)
joint_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_wt,
                    const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  BwdSmem L;
  carve_bwd(L, smem_raw, p.NH, p.Vp);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KB = p.D / BK;                 // k-blocks of the logits GEMM
  const int KBG = (p.Vp + 63) / 64;        // k-blocks (over v) of the dZ GEMM
  const int MB = p.D / 128;                // 128-lane blocks of dZ^T
  const int ntiles = *p.ntiles;
  const int tile_begin = (int)(((long)ntiles * blockIdx.x) / gridDim.x);
  const int tile_end = (int)(((long)ntiles * (blockIdx.x + 1)) / gridDim.x);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_wt);
    for (int i = 0; i < A_STAGES; ++i) { mbar_init(L.a_full(i), PROD_THREADS); mbar_init(L.a_empty(i), 1); }
    for (int i = 0; i < W_STAGES; ++i) { mbar_init(L.w_full(i), 1); mbar_init(L.w_empty(i), 1); }
    for (int i = 0; i < WT_STAGES; ++i) { mbar_init(L.wt_full(i), 1); mbar_init(L.wt_empty(i), 1); }
    mbar_init(L.tmem_full(), 1);
    mbar_init(L.g_full(), 128);
    mbar_init(L.dz_full(), 1);
    mbar_init(L.tmem_empty(), 128);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(L.tmem_ptr), TMEM_COLS);
  for (int i = tid; i < p.Vp; i += NTHREADS) L.bias[i] = p.bias_pad[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *L.tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      Pipe wp, tp;
      int prof_n = 0;
      uint32_t ph = 0;       // parity of tmem_full / dz_full for the current tile
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        TC_PROF(0, 1);
        if (tile > tile_begin) mbar_wait(L.dz_full(), ph ^ 1u, 10);       // previous tile released the overlay
        TC_PROF(0, 2);
        for (int kb = 0; kb < KB; ++kb)
          for (int h = 0; h < 2; ++h) {
            mbar_wait(L.w_empty(wp.stage), wp.phase ^ 1u, 11);
            mbar_arrive_expect_tx(L.w_full(wp.stage), (uint32_t)p.NH * 128u);
            tma_load_2d(L.w_stage(wp.stage), &tmap_w, L.w_full(wp.stage), kb * BK, h * p.NH);
            wp.advance(W_STAGES);
          }
        TC_PROF(0, 3);
        mbar_wait(L.tmem_full(), ph, 12);                                  // logits done: W ring is dead
        TC_PROF(0, 4);
        for (int mb = 0; mb < MB; ++mb)
          for (int kb = 0; kb < KBG; ++kb) {
            mbar_wait(L.wt_empty(tp.stage), tp.phase ^ 1u, 13);
            mbar_arrive_expect_tx(L.wt_full(tp.stage), (uint32_t)WT_STAGE_BYTES);
            tma_load_2d(L.wt_stage(tp.stage), &tmap_wt, L.wt_full(tp.stage), kb * 64, mb * 128);
            tp.advance(WT_STAGES);
          }
        TC_PROF(0, 5);
        ph ^= 1u;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      Pipe ap, wp, tp;
      int prof_n = 0;
      uint32_t ph = 0;
      const uint32_t idesc1 = make_idesc_bf16(BM, p.NH);
      const uint32_t idesc2 = make_idesc_bf16(128, BM);
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        TC_PROF(1, 1);
        mbar_wait(L.tmem_empty(), ph ^ 1u, 20);
        TC_PROF(1, 2);
        tc_fence_after();
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(L.a_full(ap.stage), ap.phase, 21);
          for (int h = 0; h < 2; ++h) {
            mbar_wait(L.w_full(wp.stage), wp.phase, 22);
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < BK / 16; ++ks)
              umma_bf16(tmem_base + h * p.NH, make_desc_sw128(L.a_stage(ap.stage) + ks * 32),
                        make_desc_sw128(L.w_stage(wp.stage) + ks * 32), idesc1, (kb | ks) ? 1u : 0u);
            umma_commit(L.w_empty(wp.stage));
            wp.advance(W_STAGES);
          }
          umma_commit(L.a_empty(ap.stage));
          ap.advance(A_STAGES);
        }
        umma_commit(L.tmem_full());
        TC_PROF(1, 3);
        // ---- P3: dZ^T[mb] (128 d x 128 rows) = W^T[mb] (128 x Vp) . G^T (Vp x 128)
        mbar_wait(L.g_full(), ph, 23);
        TC_PROF(1, 4);
        tc_fence_after();
        for (int mb = 0; mb < MB; ++mb)
          for (int kb = 0; kb < KBG; ++kb) {
            mbar_wait(L.wt_full(tp.stage), tp.phase, 24);
            tc_fence_after();
            const int nks = min(4, (p.Vp - kb * 64) / 16);
            for (int ks = 0; ks < nks; ++ks)
              umma_bf16(tmem_base + mb * 128, make_desc_sw128(L.wt_stage(tp.stage) + ks * 32),
                        make_desc_sw128(L.g_kblock(kb) + ks * 32), idesc2, (kb | ks) ? 1u : 0u);
            umma_commit(L.wt_empty(tp.stage));
            tp.advance(WT_STAGES);
          }
        umma_commit(L.dz_full());
        TC_PROF(1, 5);
        ph ^= 1u;
      }
    }
    __syncwarp();
  } else if (warp >= 4 && warp < 8) {
    // ------------------------------------------------------------------ epilogue warps
    const int q = warp & 3;
    const int r = q * 32 + lane;               // P2: tile row ; P4: lane of the d block
    uint32_t ph = 0;
    int prof_n = 0;
    float pacc[4][16];                          // d_pred partial sums: [d block][label slot]
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 16; ++j) pacc[i][j] = 0.f;
    int cur_b = -1, cur_ubase = 0;
    auto flush_pred = [&]() {
      if (cur_b < 0) return;
      const int Ub = min(p.u_len[cur_b], p.U1 - 1);
#pragma unroll
      for (int mb = 0; mb < 4; ++mb) {
        if (mb < MB) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int u = cur_ubase + j;
            if (u <= Ub) atomicAdd(p.d_pred + ((size_t)cur_b * p.U1 + u) * p.D + mb * 128 + r, pacc[mb][j]);
            pacc[mb][j] = 0.f;
          }
        }
      }
    };
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      int4 ti = p.tiles[tile];
      pin(ti.x); pin(ti.y); pin(ti.z); pin(ti.w);
      const RowMap g = tile_geometry<TILE_RECT>(p, ti);
      if (g.b != cur_b || g.ubase != cur_ubase) { flush_pred(); cur_b = g.b; cur_ubase = g.ubase; }
      const size_t row0 = (size_t)ti.w * BM;
      // ---------------- P2: g = d cost / d logits for row r
      int t, u;
      const bool valid = row_cell<TILE_RECT>(p, g, ti, r, t, u);
      float k_all = kNegInf, k_blank = kNegInf, k_label = kNegInf, scale = 0.f;
      int lab = -1;
      if (valid) {
        const size_t cell = ((size_t)g.b * p.T + t) * p.U1 + u;
        const float al = p.alpha[cell], be = p.beta[cell], cost = p.costs[g.b], l = p.lse_in[cell];
        k_all = al + be + cost - l;
        float bnext = kNegInf;
        if (t + 1 < g.Tb) bnext = p.beta[cell + p.U1];
        else if (u == g.Ub) bnext = 0.f;
        k_blank = al + bnext + cost - l;
        if (u < g.Ub) { k_label = al + p.beta[cell + 1] + cost - l; lab = p.targets[(size_t)g.b * (p.U1 - 1) + u]; }
        scale = p.grad_costs[g.b];
      }
      if (tid == 128) TC_PROF(2, 1);
      mbar_wait(L.tmem_full(), ph, 30);
      if (tid == 128) TC_PROF(2, 2);
      tc_fence_after();
      for (int c0 = 0; c0 < p.Vp; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c0, v);
        tmem_ld_wait();
        if (tid == 128) TC_PROF(2, 10);
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          float gg[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int col = c0 + j + e;
            const float x = v[j + e] + L.bias[col];
            float gv = ex2_fast((x + k_all) * LOG2E);
            if (col == p.blank) gv -= ex2_fast((x + k_blank) * LOG2E);
            if (col == lab) gv -= ex2_fast((x + k_label) * LOG2E);
            if (p.clamp > 0.f) gv = fminf(fmaxf(gv, -p.clamp), p.clamp);
            gg[e] = valid ? gv * scale : 0.f;
          }
          pk[j >> 1] = pack_bf16(gg[0], gg[1]);
        }
        if (tid == 128) TC_PROF(2, 11);
        // G tile: k-block c0/64, row r, 16-byte chunks (c0%64)/8 .. +3, 128B swizzle
        const uint32_t gb = L.g_kblock(c0 >> 6) + r * 128;
        const int ch0 = (c0 & 63) >> 3;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t dst = gb + (((ch0 + i) ^ (r & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk[4 * i]), "r"(pk[4 * i + 1]),
                       "r"(pk[4 * i + 2]), "r"(pk[4 * i + 3]) : "memory");
        }
        // g^T spill: gt[col][row0 + r] (lanes = consecutive rows -> 64 B per column)
        unsigned short* gt = reinterpret_cast<unsigned short*>(p.gt) + (size_t)c0 * p.Rpad + row0 + r;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          gt[(size_t)(2 * j) * p.Rpad] = (unsigned short)(pk[j] & 0xffffu);
          gt[(size_t)(2 * j + 1) * p.Rpad] = (unsigned short)(pk[j] >> 16);
        }
        if (tid == 128) TC_PROF(2, 12);
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(L.g_full());
      if (tid == 128) TC_PROF(2, 3);
      // ---------------- P4: dH = dZ * (1 - z^2); reductions
      mbar_wait(L.dz_full(), ph, 31);
      if (tid == 128) TC_PROF(2, 4);
      tc_fence_after();
#pragma unroll
      for (int mb = 0; mb < 4; ++mb) {
        if (mb < MB) {
          const int d = mb * 128 + r;
          const uint4* zrow = reinterpret_cast<const uint4*>(p.zt + (size_t)d * p.Rpad + row0);
          uint4 zn0 = __ldcg(zrow), zn1 = __ldcg(zrow + 1);
          float v[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + mb * 128, v);
#pragma unroll
          for (int tloc = 0; tloc < 8; ++tloc) {
            const uint4 z0 = zn0, z1 = zn1;
            tmem_ld_wait();
            float w[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) w[j] = v[j];
            if (tloc < 7) {      // software pipeline: next frame slot's z row and TMEM columns are in flight
              zn0 = __ldcg(zrow + 2 * (tloc + 1));
              zn1 = __ldcg(zrow + 2 * (tloc + 1) + 1);
              tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + mb * 128 + (tloc + 1) * 16, v);
            }
            const uint32_t zw[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
            float es0 = 0.f, es1 = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float za = __uint_as_float(zw[j] << 16), zb = __uint_as_float(zw[j] & 0xffff0000u);
              const float ha = w[2 * j] * fmaf(-za, za, 1.f), hb = w[2 * j + 1] * fmaf(-zb, zb, 1.f);
              es0 += ha;
              es1 += hb;
              pacc[mb][2 * j] += ha;
              pacc[mb][2 * j + 1] += hb;
            }
            const int tt = g.t0 + tloc;
            if (tt < g.Tb) p.d_enc_part[(((size_t)ti.y * p.B + g.b) * p.T + tt) * p.D + d] = es0 + es1;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(L.tmem_empty());
      if (tid == 128) TC_PROF(2, 5);
      ph ^= 1u;
    }
    flush_pred();
  } else if (warp >= 8) {
    // ------------------------------------------------------------------ A producers (+ d_bias column sums)
    const int pt = tid - 256;
    const int r = pt & 127, khalf = pt >> 7;
    Pipe ap;
    uint32_t ph = 0;
    int prof_n = 0;
    constexpr int NPRE = (SLAB_ROWS_RECT * 16 + PROD_THREADS - 1) / PROD_THREADS;   // 2
    float4 pre[NPRE];
    float db0 = 0.f, db1 = 0.f;                  // columns pt and pt + 256
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      int4 ti = p.tiles[tile];
      pin(ti.x); pin(ti.y); pin(ti.z); pin(ti.w);
      const RowMap g = tile_geometry<TILE_RECT>(p, ti);
      const size_t row0 = (size_t)ti.w * BM;
      int t, u;
      const bool valid = row_cell<TILE_RECT>(p, g, ti, r, t, u);
      const int prow = r & 15, erow = 16 + (r >> 4);
      if (pt == 0) TC_PROF(3, 1);
      slab_fetch<NPRE>(p, g, 0, pt, pre);
      if (tile > tile_begin) mbar_wait(L.dz_full(), ph ^ 1u, 40);
      if (pt == 0) TC_PROF(3, 2);          // overlay (G tile / W^T ring) released
      named_barrier_sync(1, PROD_THREADS);
      slab_store<NPRE>(L.slab, g, pt, pre);
      named_barrier_sync(1, PROD_THREADS);
      for (int kb = 0; kb < KB; ++kb) {
        if (kb + 1 < KB) slab_fetch<NPRE>(p, g, (kb + 1) * BK, pt, pre);
        mbar_wait(L.a_empty(ap.stage), ap.phase ^ 1u, 41);
        produce_a<true>(L.slab, L.a_stage(ap.stage), r, khalf, valid, prow, erow,
                        p.zt + (size_t)(kb * BK + khalf * 32) * p.Rpad + row0 + r, p.Rpad);
        fence_proxy_async();
        mbar_arrive(L.a_full(ap.stage));
        ap.advance(A_STAGES);
        if (kb + 1 < KB) {
          named_barrier_sync(1, PROD_THREADS);
          slab_store<NPRE>(L.slab, g, pt, pre);
          named_barrier_sync(1, PROD_THREADS);
        }
      }
      // d_bias: column sums of the bf16 G tile (written by the epilogue warps)
      if (pt == 0) TC_PROF(3, 3);
      mbar_wait(L.g_full(), ph, 42);
      if (pt == 0) TC_PROF(3, 4);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int col = pt + c * 256;
        if (col < p.Vp) {
          const uint32_t gb = L.g_kblock(col >> 6) + (col & 7) * 2;
          const int ch = (col & 63) >> 3;
          float acc = 0.f;
          for (int rr = 0; rr < BM; ++rr) {
            unsigned short hv;
            asm volatile("ld.shared.u16 %0, [%1];" : "=h"(hv) : "r"(gb + rr * 128 + ((ch ^ (rr & 7)) << 4)));
            acc += __uint_as_float((uint32_t)hv << 16);
          }
          if (c == 0) db0 += acc; else db1 += acc;
        }
      }
      if (pt == 0) TC_PROF(3, 5);
      ph ^= 1u;
    }
    if (pt < p.V && tile_end > tile_begin) atomicAdd(p.d_bias + pt, db0);
    if (pt + 256 < p.V && tile_end > tile_begin) atomicAdd(p.d_bias + pt + 256, db1);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

// d_enc[b,t,:] = sum over the u-splits of the partial sums (zero for padded frames)
__global__ void reduce_denc_kernel(const float* __restrict__ part, const int32_t* __restrict__ t_len,
                                   const int32_t* __restrict__ u_len, float* __restrict__ d_enc, int B, int T, int U1,
                                   int D) {
  const int bt = blockIdx.x;
  const int b = bt / T, t = bt - b * T;
  const int Tb = min(t_len[b], T), W = min(u_len[b], U1 - 1) + 1;
  const int S = (W + 15) >> 4;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float s = 0.f;
    if (t < Tb)
      for (int i = 0; i < S; ++i) s += part[(((size_t)i * B + b) * T + t) * D + d];
    d_enc[((size_t)b * T + t) * D + d] = s;
  }
}

// =================================================================================================
// Backward kernel 2: dW^T[d][v] = sum_rows z^T[d][row] * g^T[v][row]  (plain TMA-fed tcgen05 GEMM, split-K)
// =================================================================================================
constexpr int DW_STAGES = 3;
constexpr int DW_THREADS = 256;

__global__ void __launch_bounds__(DW_THREADS, 1)
dw_gemm_kernel(const __grid_constant__ CUtensorMap tmap_zt, const __grid_constant__ CUtensorMap tmap_gt,
               const int* __restrict__ ntiles_ptr, float* __restrict__ partials, int D, int Vp, int NH, int KS) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  const uint32_t al = (base + 1023u) & ~1023u;
  const uint32_t stage_bytes = A_STAGE_BYTES + 2 * NH * 128;
  const uint32_t bar = al + DW_STAGES * stage_bytes;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem_raw + (bar + 128 - base));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mb = blockIdx.x, ks = blockIdx.y;
  const int kblocks = (*ntiles_ptr) * 2;                      // 64-row k-blocks actually written by kernel 1
  const int kb_begin = (int)(((long)kblocks * ks) / KS), kb_end = (int)(((long)kblocks * (ks + 1)) / KS);
  auto full = [&](int i) { return bar + i * 16; };
  auto empty = [&](int i) { return bar + i * 16 + 8; };
  const uint32_t done = bar + DW_STAGES * 16;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_zt);
    tma_prefetch_desc(&tmap_gt);
    for (int i = 0; i < DW_STAGES; ++i) { mbar_init(full(i), 1); mbar_init(empty(i), 1); }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(tmem_ptr), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      Pipe sp;
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(empty(sp.stage), sp.phase ^ 1u, 50);
        const uint32_t st = al + sp.stage * stage_bytes;
        mbar_arrive_expect_tx(full(sp.stage), stage_bytes);
        tma_load_2d(st, &tmap_zt, full(sp.stage), kb * 64, mb * 128);
        tma_load_2d(st + A_STAGE_BYTES, &tmap_gt, full(sp.stage), kb * 64, 0);
        tma_load_2d(st + A_STAGE_BYTES + NH * 128, &tmap_gt, full(sp.stage), kb * 64, NH);
        sp.advance(DW_STAGES);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      Pipe sp;
      const uint32_t idesc = make_idesc_bf16(128, NH);
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(full(sp.stage), sp.phase, 51);
        tc_fence_after();
        const uint32_t st = al + sp.stage * stage_bytes;
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            umma_bf16(tmem_base + h * NH, make_desc_sw128(st + k4 * 32),
                      make_desc_sw128(st + A_STAGE_BYTES + h * NH * 128 + k4 * 32), idesc,
                      (kb > kb_begin || k4 > 0) ? 1u : 0u);
        umma_commit(empty(sp.stage));
        sp.advance(DW_STAGES);
      }
      umma_commit(done);
    }
    __syncwarp();
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int d = mb * 128 + q * 32 + lane;
    float* out = partials + ((size_t)ks * D + d) * Vp;
    if (kb_end > kb_begin) {
      mbar_wait(done, 0, 52);
      tc_fence_after();
      for (int c0 = 0; c0 < Vp; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c0, v);
        tmem_ld_wait();
        if (d < D) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(out + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
      }
    } else if (d < D) {
      for (int c0 = 0; c0 < Vp; c0 += 4) *reinterpret_cast<float4*>(out + c0) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

// d_w[v][d] = sum_ks partials[ks][d][v]
__global__ void reduce_dw_kernel(const float* __restrict__ partials, float* __restrict__ d_w, int D, int V, int Vp,
                                 int KS) {
  __shared__ float tile[32][33];
  const int d0 = blockIdx.x * 32, v0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int d = d0 + i, v = v0 + tx;
    float s = 0.f;
    if (d < D && v < Vp)
      for (int k = 0; k < KS; ++k) s += partials[((size_t)k * D + d) * Vp + v];
    tile[i][tx] = s;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int v = v0 + i, d = d0 + tx;
    if (v < V && d < D) d_w[(size_t)v * D + d] = tile[tx][i];
  }
}

// =================================================================================================
// Helper kernels: weight conversion, tile tables
// =================================================================================================
// wb [Vp][D] bf16 (zero rows beyond V), wtb [D][Vp] bf16 (optional), bias_pad [Vp] (-inf beyond V)
__global__ void prep_weights_kernel(const float* __restrict__ w, const float* __restrict__ bias,
                                    __nv_bfloat16* __restrict__ wb, __nv_bfloat16* __restrict__ wtb,
                                    float* __restrict__ bias_pad, int V, int Vp, int D) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Vp * D) {
    int v = i / D, d = i - v * D;
    float x = (v < V) ? w[(size_t)v * D + d] : 0.f;
    wb[i] = __float2bfloat16(x);
    if (wtb) wtb[(size_t)d * Vp + v] = __float2bfloat16(x);
  }
  if (i < Vp) bias_pad[i] = (i < V) ? bias[i] : kNegInf;
}

// Single-CTA tile table builder (B is small).  flat: 128 consecutive valid cells per tile.
// rect: tiles of 8 frames x 16 label columns, ordered (b, u-split, frame block) so that one CTA sweeps
// consecutive frame blocks of the same (b, u-split).
__global__ void build_tiles_kernel(const int32_t* __restrict__ t_len, const int32_t* __restrict__ u_len, int B, int T,
                                   int U1, int rect, int4* __restrict__ tiles, int* __restrict__ ntiles,
                                   int max_tiles) {
  __shared__ int s_off[1025];
  if (threadIdx.x == 0) {
    int off = 0;
    for (int b = 0; b < B; ++b) {
      s_off[b & 1023] = off;   // (only used when B <= 1024; larger batches recompute below)
      int Tb = min(t_len[b], T), W = min(u_len[b], U1 - 1) + 1;
      if (Tb > 0) {
        if (rect) { int S = (W + 15) >> 4; off += S * ((Tb + 7) >> 3); }
        else off += (Tb * W + BM - 1) / BM;
      }
    }
    *ntiles = min(off, max_tiles);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    int off;
    if (B <= 1024) off = s_off[b];
    else {
      off = 0;
      for (int bb = 0; bb < b; ++bb) {
        int Tb = min(t_len[bb], T), W = min(u_len[bb], U1 - 1) + 1;
        if (Tb > 0) off += rect ? ((W + 15) >> 4) * ((Tb + 7) >> 3) : (Tb * W + BM - 1) / BM;
      }
    }
    int Tb = min(t_len[b], T), W = min(u_len[b], U1 - 1) + 1;
    if (Tb <= 0) continue;
    if (rect) {
      int S = (W + 15) >> 4, NTB = (Tb + 7) >> 3;
      for (int s = 0; s < S; ++s)
        for (int tb = 0; tb < NTB; ++tb) {
          int idx = off + s * NTB + tb;
          if (idx < max_tiles) tiles[idx] = make_int4(b, s, tb, idx);
        }
    } else {
      int n = (Tb * W + BM - 1) / BM;
      for (int i = 0; i < n; ++i)
        if (off + i < max_tiles) tiles[off + i] = make_int4(b, i * BM, 0, off + i);
    }
  }
}

// fp32 -> bf16 copies of the activations (n4 float4 groups each); the bf16 path rounds enc_proj / pred_proj
// to bf16 (under autocast they already are bf16 values, so this is lossless there).
__global__ void to_bf16_kernel(const float4* __restrict__ a, uint2* __restrict__ ab, long na4,
                               const float4* __restrict__ b, uint2* __restrict__ bb, long nb4) {
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < na4 + nb4; i += stride) {
    const bool first = i < na4;
    const float4 v = first ? __ldg(a + i) : __ldg(b + (i - na4));
    const uint2 o = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    if (first) ab[i] = o; else bb[i - na4] = o;
  }
}

// wb [Vp][D] bf16 (zero rows beyond V), wtb [D][Vp] bf16 (optional), bias_pad [Vp] (-inf beyond V),
// bias_l2 [Vp] = bias * log2(e) (-inf beyond V)
__global__ void prep_weights2_kernel(const float* __restrict__ w, const float* __restrict__ bias,
                                     __nv_bfloat16* __restrict__ wb, __nv_bfloat16* __restrict__ wtb,
                                     float* __restrict__ bias_pad, float* __restrict__ bias_l2, int V, int Vp, int D) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Vp * D) {
    int v = i / D, d = i - v * D;
    float x = (v < V) ? w[(size_t)v * D + d] : 0.f;
    wb[i] = __float2bfloat16(x);
    if (wtb) wtb[(size_t)d * Vp + v] = __float2bfloat16(x);
  }
  if (i < Vp) {
    if (bias_pad) bias_pad[i] = (i < V) ? bias[i] : kNegInf;
    if (bias_l2) bias_l2[i] = (i < V) ? bias[i] * LOG2E : kNegInf;
  }
}

// Forward tile table (see joint_tc_fwd.cuh): one thread per utterance, block-wide exclusive scan of the
// per-utterance tile counts (B is processed in chunks of blockDim.x with a running base).
__global__ void build_tiles_fwd_kernel(const int32_t* __restrict__ t_len, const int32_t* __restrict__ u_len, int B,
                                       int T, int U1, int4* __restrict__ tiles, int* __restrict__ ntiles,
                                       int max_tiles) {
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int b0 = 0; b0 < B; b0 += blockDim.x) {
    const int b = b0 + threadIdx.x;
    int Tb = 0, W = 1, n = 0;
    if (b < B) {
      Tb = min(t_len[b], T);
      W = min(u_len[b], U1 - 1) + 1;
      n = fwd_tiles_of(Tb, W);
    }
    int incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int x = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += x;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int woff = 0;
    for (int i = 0; i < warp; ++i) woff += s_warp[i];
    int total = 0;
    for (int i = 0; i < nwarp; ++i) total += s_warp[i];
    const int off = s_base + woff + incl - n;
    if (n > 0) {
      const int nb32 = (Tb + 31) >> 5, n4 = (W >> 2) * nb32;
      const int n2 = (W & 2) ? ((Tb + 63) >> 6) : 0;
      for (int i = 0; i < n; ++i) {
        int4 e;
        if (i < n4) { const int g = i / nb32; e = make_int4(b, 4 * g, 32 * (i - g * nb32), 4); }
        else if (i < n4 + n2) e = make_int4(b, (W >> 2) * 4, 64 * (i - n4), 2);
        else e = make_int4(b, W - 1, 128 * (i - n4 - n2), 1);
        if (off + i < max_tiles) tiles[off + i] = e;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) s_base += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *ntiles = min(s_base, max_tiles);
}

int make_tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems,
                      uint32_t box_rows) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qr;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qr);
    if (e != cudaSuccess || qr != cudaDriverEntryPointSuccess || !sym) {
      set_error("cuTensorMapEncodeTiled is not available from the driver (%s)", cudaGetErrorString(e));
      return 1;
    }
    fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu box_rows=%u)", (int)r,
              (unsigned long long)rows, (unsigned long long)cols, box_rows);
    return 1;
  }
  return 0;
}

static int check_tc_error(const char* where) {
  // asynchronous: reads the flag left by a PREVIOUS launch (no host sync on the hot path)
  (void)where;
  return 0;
}

}  // namespace tc

// =================================================================================================
// Host entry points
// =================================================================================================
using namespace tc;

static long long* g_prof_buf = nullptr;
static int pad_v(int V) { return (V + 31) / 32 * 32; }
static int max_tiles_flat(int B, int T, int U1) { return B * (int)(((long)T * U1 + BM - 1) / BM); }
static int max_tiles_rect(int B, int T, int U1) { return B * ((U1 + 15) / 16) * ((T + 7) / 8); }

bool joint_tc_supported(int U1, int D, int V) { return D % 64 == 0 && D >= 64 && D <= 1024 && pad_v(V) <= 512 && U1 <= 128; }

static int max_tiles_fwd2(int B, int T, int U1) {
  return B * ((U1 >> 2) * ((T + 31) / 32) + (T + 63) / 64 + (T + 127) / 128);
}

struct FwdWs {
  __nv_bfloat16 *wb, *eb, *pb;
  float *bias_pad, *bias_l2;
  int4* tiles;
  int* ntiles;
  size_t bytes;
};
static FwdWs carve_fwd_ws(void* ws, int B, int T, int U1, int D, int V) {
  int Vp = pad_v(V);
  uint8_t* p = reinterpret_cast<uint8_t*>(ws);
  size_t off = 0;
  FwdWs w;
  auto take = [&](size_t n) { void* r = p + off; off = align_up(off + n, 1024); return r; };
  w.wb = reinterpret_cast<__nv_bfloat16*>(take((size_t)Vp * D * 2));
  w.eb = reinterpret_cast<__nv_bfloat16*>(take((size_t)B * T * D * 2));
  w.pb = reinterpret_cast<__nv_bfloat16*>(take((size_t)B * U1 * D * 2));
  w.bias_pad = reinterpret_cast<float*>(take((size_t)Vp * 4));
  w.bias_l2 = reinterpret_cast<float*>(take((size_t)Vp * 4));
  const int mt = max_tiles_flat(B, T, U1) > max_tiles_fwd2(B, T, U1) ? max_tiles_flat(B, T, U1) : max_tiles_fwd2(B, T, U1);
  w.tiles = reinterpret_cast<int4*>(take((size_t)mt * 16));
  w.ntiles = reinterpret_cast<int*>(take(4));
  w.bytes = off;
  return w;
}

size_t joint_fwd_tc_ws_bytes(int B, int T, int U1, int D, int V) { return carve_fwd_ws(nullptr, B, T, U1, D, V).bytes; }

static int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

static bool env_flag(const char* name) {
  const char* v = getenv(name);
  return v && v[0] && v[0] != '0';
}

int joint_fwd_f32(const float*, const float*, const float*, const float*, const int32_t*, const int32_t*,
                  const int32_t*, float*, float*, float*, int, int, int, int, int, int, cudaStream_t);

static int joint_fwd_tc_v1(const float* enc, const float* pred, const float* w, const float* bias,
                           const int32_t* targets, const int32_t* t_len, const int32_t* u_len, float* lse,
                           float* lp_blank, float* lp_label, int B, int T, int U1, int D, int V, int blank, void* ws,
                           cudaStream_t st) {
  const int Vp = pad_v(V), NH = Vp / 2;
  FwdWs W = carve_fwd_ws(ws, B, T, U1, D, V);
  prep_weights_kernel<<<cdiv((long)Vp * D, 256), 256, 0, st>>>(w, bias, W.wb, nullptr, W.bias_pad, V, Vp, D);
  CTCVR_LAUNCH_CHECK();
  const int mt = max_tiles_flat(B, T, U1);
  build_tiles_kernel<<<1, 256, 0, st>>>(t_len, u_len, B, T, U1, 0, W.tiles, W.ntiles, mt);
  CTCVR_LAUNCH_CHECK();
  CUtensorMap tmap;
  if (make_tmap_bf16_2d(&tmap, W.wb, Vp, D, D, NH)) return 1;
  TcParams p{};
  p.enc = enc; p.pred = pred; p.bias_pad = W.bias_pad; p.targets = targets; p.t_len = t_len; p.u_len = u_len;
  p.tiles = W.tiles; p.ntiles = W.ntiles;
  p.B = B; p.T = T; p.U1 = U1; p.D = D; p.V = V; p.Vp = Vp; p.NH = NH; p.blank = blank;
  p.lse = lse; p.lp_blank = lp_blank; p.lp_label = lp_label;
  p.prof = g_prof_buf;
  size_t smem = tc_smem_bytes(NH, Vp, SLAB_ROWS_FLAT);
  CTCVR_REQUIRE(smem <= 232448, "joint_rnnt_fwd bf16: shared memory budget exceeded (%zu B)", smem);
  CTCVR_CHECK_CUDA(cudaFuncSetAttribute(joint_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = min(sm_count(), mt);
  joint_fwd_tc_kernel<<<grid, NTHREADS, smem, st>>>(tmap, p);
  CTCVR_LAUNCH_CHECK();
  return 0;
}

int joint_fwd_tc(const float* enc, const float* pred, const float* w, const float* bias, const int32_t* targets,
                 const int32_t* t_len, const int32_t* u_len, float* lse, float* lp_blank, float* lp_label, int B, int T,
                 int U1, int D, int V, int blank, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!joint_tc_supported(U1, D, V))   // shapes outside the tensor-core tiling use the fp32 kernels (same GPU)
    return joint_fwd_f32(enc, pred, w, bias, targets, t_len, u_len, lse, lp_blank, lp_label, B, T, U1, D, V, blank, st);
  CTCVR_REQUIRE(ws && ws_bytes >= joint_fwd_tc_ws_bytes(B, T, U1, D, V), "joint_rnnt_fwd bf16: workspace too small");
  CTCVR_REQUIRE(((uintptr_t)enc & 15) == 0 && ((uintptr_t)pred & 15) == 0, "joint_rnnt_fwd bf16: enc_proj / pred_proj must be 16-byte aligned");
  if (env_flag("CTCVR_FWD_V1"))
    return joint_fwd_tc_v1(enc, pred, w, bias, targets, t_len, u_len, lse, lp_blank, lp_label, B, T, U1, D, V, blank, ws, st);
  const int Vp = pad_v(V), NH = Vp / 2;
  FwdWs W = carve_fwd_ws(ws, B, T, U1, D, V);
  prep_weights2_kernel<<<cdiv((long)Vp * D, 256), 256, 0, st>>>(w, bias, W.wb, nullptr, nullptr, W.bias_l2, V, Vp, D);
  CTCVR_LAUNCH_CHECK();
  {
    const long na4 = (long)B * T * D / 4, nb4 = (long)B * U1 * D / 4;
    const int blocks = (int)std::min<long>((na4 + nb4 + 255) / 256, 148L * 8);
    to_bf16_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(enc), reinterpret_cast<uint2*>(W.eb), na4,
                                           reinterpret_cast<const float4*>(pred), reinterpret_cast<uint2*>(W.pb), nb4);
    CTCVR_LAUNCH_CHECK();
  }
  const int mt = max_tiles_fwd2(B, T, U1);
  build_tiles_fwd_kernel<<<1, 256, 0, st>>>(t_len, u_len, B, T, U1, W.tiles, W.ntiles, mt);
  CTCVR_LAUNCH_CHECK();
  CUtensorMap tmap_w, tmap_e, tmap_p;
  if (make_tmap_bf16_2d(&tmap_w, W.wb, Vp, D, D, NH)) return 1;
  if (make_tmap_bf16_2d(&tmap_e, W.eb, (uint64_t)B * T, D, D, 32)) return 1;
  if (make_tmap_bf16_2d(&tmap_p, W.pb, (uint64_t)B * U1, D, D, 4)) return 1;
  FwdParams p{};
  p.bias = bias; p.bias_l2 = W.bias_l2; p.targets = targets; p.t_len = t_len; p.u_len = u_len;
  p.tiles = W.tiles; p.ntiles = W.ntiles;
  p.B = B; p.T = T; p.U1 = U1; p.D = D; p.V = V; p.Vp = Vp; p.NH = NH; p.blank = blank;
  p.lse = lse; p.lp_blank = lp_blank; p.lp_label = lp_label;
  p.prof = g_prof_buf;
  int ws_n = F_MAX_W_STAGES;
  while (ws_n > 2 && fwd2_smem_bytes(NH, Vp, ws_n) > 232448) --ws_n;
  p.w_stages = ws_n;
  const size_t smem = fwd2_smem_bytes(NH, Vp, ws_n);
  CTCVR_REQUIRE(smem <= 232448, "joint_rnnt_fwd bf16: shared memory budget exceeded (%zu B)", smem);
  CTCVR_CHECK_CUDA(cudaFuncSetAttribute(joint_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = min(sm_count(), mt);
  joint_fwd2_kernel<<<grid, NTHREADS, smem, st>>>(tmap_w, tmap_e, tmap_p, p);
  CTCVR_LAUNCH_CHECK();
  return 0;
}

size_t joint_bwd_f32_ws_bytes(int, int, int, int, int);
int joint_bwd_f32(const float*, const float*, const float*, const float*, const int32_t*, const int32_t*,
                  const int32_t*, const float*, const float*, const float*, const float*, const float*, float, float*,
                  float*, float*, float*, int, int, int, int, int, int, void*, size_t, cudaStream_t);

static bool joint_tc_bwd_supported(int U1, int D, int V) {
  return joint_tc_supported(U1, D, V) && D % 128 == 0 && D <= 512;
}

struct BwdWs {
  __nv_bfloat16 *wb, *wtb, *zt, *gt, *eb, *pb;
  float *bias_pad, *bias_l2, *d_enc_part, *partials;
  int4* tiles;
  int* ntiles;
  long Rpad;
  int KS, S_max, mt;
  size_t bytes;
};
static BwdWs carve_bwd_ws(void* ws, int B, int T, int U1, int D, int V) {
  const int Vp = pad_v(V);
  uint8_t* p = reinterpret_cast<uint8_t*>(ws);
  size_t off = 0;
  BwdWs w;
  w.mt = max_tiles_rect(B, T, U1);
  w.Rpad = (long)(w.mt + 1) * BM;           // + one scratch row tile for the dummy iterations of the lock-step loop
  w.S_max = (U1 + 15) / 16;
  const int MB = D / 128;
  w.KS = MB > 0 ? sm_count() / MB : 1;
  if (w.KS < 1) w.KS = 1;
  auto take = [&](size_t n) { void* r = p + off; off = align_up(off + n, 1024); return r; };
  w.wb = reinterpret_cast<__nv_bfloat16*>(take((size_t)Vp * D * 2));
  w.wtb = reinterpret_cast<__nv_bfloat16*>(take((size_t)Vp * D * 2));
  w.eb = reinterpret_cast<__nv_bfloat16*>(take((size_t)B * T * D * 2));
  w.pb = reinterpret_cast<__nv_bfloat16*>(take((size_t)B * U1 * D * 2));
  w.bias_pad = reinterpret_cast<float*>(take((size_t)Vp * 4));
  w.bias_l2 = reinterpret_cast<float*>(take((size_t)Vp * 4));
  w.tiles = reinterpret_cast<int4*>(take((size_t)w.mt * 16));
  w.ntiles = reinterpret_cast<int*>(take(4));
  w.zt = reinterpret_cast<__nv_bfloat16*>(take((size_t)D * w.Rpad * 2));
  w.gt = reinterpret_cast<__nv_bfloat16*>(take((size_t)Vp * w.Rpad * 2));
  w.d_enc_part = reinterpret_cast<float*>(take((size_t)w.S_max * B * T * D * 4));
  w.partials = reinterpret_cast<float*>(take((size_t)w.KS * D * Vp * 4));
  w.bytes = off;
  return w;
}

size_t joint_bwd_tc_ws_bytes(int B, int T, int U1, int D, int V) {
  if (!joint_tc_bwd_supported(U1, D, V)) return joint_bwd_f32_ws_bytes(B, T, U1, D, V);
  return carve_bwd_ws(nullptr, B, T, U1, D, V).bytes;
}

int joint_bwd_tc(const float* enc, const float* pred, const float* w, const float* bias, const int32_t* targets,
                 const int32_t* t_len, const int32_t* u_len, const float* lse, const float* alpha, const float* beta,
                 const float* costs, const float* grad_costs, float clamp, float* d_enc, float* d_pred, float* d_w,
                 float* d_b, int B, int T, int U1, int D, int V, int blank, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!joint_tc_bwd_supported(U1, D, V))   // shapes outside the tensor-core tiling use the fp32 kernels (same GPU)
    return joint_bwd_f32(enc, pred, w, bias, targets, t_len, u_len, lse, alpha, beta, costs, grad_costs, clamp, d_enc,
                         d_pred, d_w, d_b, B, T, U1, D, V, blank, ws, ws_bytes, st);
  CTCVR_REQUIRE(ws && ws_bytes >= joint_bwd_tc_ws_bytes(B, T, U1, D, V), "joint_rnnt_bwd bf16: workspace too small");
  CTCVR_REQUIRE(((uintptr_t)enc & 15) == 0 && ((uintptr_t)pred & 15) == 0, "joint_rnnt_bwd bf16: enc_proj / pred_proj must be 16-byte aligned");
  const int Vp = pad_v(V), NH = Vp / 2, MB = D / 128;
  BwdWs W = carve_bwd_ws(ws, B, T, U1, D, V);
  const bool v1 = env_flag("CTCVR_BWD_V1");
  prep_weights2_kernel<<<cdiv((long)Vp * D, 256), 256, 0, st>>>(w, bias, W.wb, W.wtb, W.bias_pad, W.bias_l2, V, Vp, D);
  CTCVR_LAUNCH_CHECK();
  const int mt = W.mt;
  build_tiles_kernel<<<1, 256, 0, st>>>(t_len, u_len, B, T, U1, 1, W.tiles, W.ntiles, mt);
  CTCVR_LAUNCH_CHECK();
  CTCVR_CHECK_CUDA(cudaMemsetAsync(d_pred, 0, (size_t)B * U1 * D * sizeof(float), st));
  CTCVR_CHECK_CUDA(cudaMemsetAsync(d_b, 0, (size_t)V * sizeof(float), st));
  CUtensorMap tmap_zt, tmap_gt;
  if (make_tmap_bf16_2d(&tmap_zt, W.zt, D, W.Rpad, W.Rpad, 128)) return 1;
  if (make_tmap_bf16_2d(&tmap_gt, W.gt, Vp, W.Rpad, W.Rpad, NH)) return 1;
  if (v1) {
    CUtensorMap tmap_w, tmap_wt;
    if (make_tmap_bf16_2d(&tmap_w, W.wb, Vp, D, D, NH)) return 1;
    if (make_tmap_bf16_2d(&tmap_wt, W.wtb, D, Vp, Vp, 128)) return 1;
    TcParams p{};
    p.enc = enc; p.pred = pred; p.bias_pad = W.bias_pad; p.targets = targets; p.t_len = t_len; p.u_len = u_len;
    p.tiles = W.tiles; p.ntiles = W.ntiles;
    p.B = B; p.T = T; p.U1 = U1; p.D = D; p.V = V; p.Vp = Vp; p.NH = NH; p.blank = blank;
    p.lse_in = lse; p.alpha = alpha; p.beta = beta; p.costs = costs; p.grad_costs = grad_costs; p.clamp = clamp;
    p.zt = W.zt; p.gt = W.gt; p.Rpad = W.Rpad; p.d_enc_part = W.d_enc_part; p.d_pred = d_pred; p.d_bias = d_b;
    p.S_max = W.S_max;
    p.prof = g_prof_buf;
    size_t smem = bwd_smem_bytes(NH, Vp);
    CTCVR_REQUIRE(smem <= 232448, "joint_rnnt_bwd bf16: shared memory budget exceeded (%zu B)", smem);
    CTCVR_CHECK_CUDA(cudaFuncSetAttribute(joint_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = min(sm_count(), mt);
    joint_bwd_tc_kernel<<<grid, NTHREADS, smem, st>>>(tmap_w, tmap_wt, p);
    CTCVR_LAUNCH_CHECK();
  } else {
    {
      const long na4 = (long)B * T * D / 4, nb4 = (long)B * U1 * D / 4;
      const int blocks = (int)std::min<long>((na4 + nb4 + 255) / 256, 148L * 8);
      to_bf16_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(enc), reinterpret_cast<uint2*>(W.eb), na4,
                                             reinterpret_cast<const float4*>(pred), reinterpret_cast<uint2*>(W.pb), nb4);
      CTCVR_LAUNCH_CHECK();
    }
    const int grid = min(sm_count(), mt);
    CUtensorMap tmap_w, tmap_wt, tmap_e, tmap_p;
    if (make_tmap_bf16_2d(&tmap_w, W.wb, Vp, D, D, NH)) return 1;
    if (make_tmap_bf16_2d(&tmap_wt, W.wtb, D, Vp, Vp, 128)) return 1;
    if (make_tmap_bf16_2d(&tmap_e, W.eb, (uint64_t)B * T, D, D, 8)) return 1;
    if (make_tmap_bf16_2d(&tmap_p, W.pb, (uint64_t)B * U1, D, D, 16)) return 1;
    BwdParams p{};
    p.bias = bias; p.bias_l2 = W.bias_l2; p.targets = targets; p.t_len = t_len; p.u_len = u_len;
    p.tiles = W.tiles; p.ntiles = W.ntiles;
    p.B = B; p.T = T; p.U1 = U1; p.D = D; p.V = V; p.Vp = Vp; p.NH = NH; p.blank = blank;
    p.lse = lse; p.alpha = alpha; p.beta = beta; p.costs = costs; p.grad_costs = grad_costs; p.clamp = clamp;
    p.zt = W.zt; p.gt = W.gt; p.Rpad = W.Rpad; p.scratch_tile = mt;
    p.d_enc_part = W.d_enc_part; p.d_pred = d_pred; p.d_bias = d_b;
    p.prof = g_prof_buf;
    const size_t smem = bwd2_smem_bytes(NH, Vp, D);
    CTCVR_REQUIRE(smem <= 232448, "joint_rnnt_bwd bf16: shared memory budget exceeded (%zu B)", smem);
    CTCVR_CHECK_CUDA(cudaFuncSetAttribute(joint_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    joint_bwd2_kernel<<<grid, NTHREADS, smem, st>>>(tmap_w, tmap_wt, tmap_e, tmap_p, tmap_zt, p);
    CTCVR_LAUNCH_CHECK();
  }
  reduce_denc_kernel<<<B * T, 128, 0, st>>>(W.d_enc_part, t_len, u_len, d_enc, B, T, U1, D);
  CTCVR_LAUNCH_CHECK();
  {
    size_t smem = 1024 + (size_t)DW_STAGES * (A_STAGE_BYTES + 2 * NH * 128) + 256;
    CTCVR_REQUIRE(smem <= 232448, "dW GEMM: shared memory budget exceeded (%zu B)", smem);
    CTCVR_CHECK_CUDA(cudaFuncSetAttribute(dw_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dw_gemm_kernel<<<dim3(MB, W.KS), DW_THREADS, smem, st>>>(tmap_zt, tmap_gt, W.ntiles, W.partials, D, Vp, NH, W.KS);
    CTCVR_LAUNCH_CHECK();
  }
  reduce_dw_kernel<<<dim3(cdiv(D, 32), cdiv(Vp, 32)), 256, 0, st>>>(W.partials, d_w, D, V, Vp, W.KS);
  CTCVR_LAUNCH_CHECK();
  return 0;
}

void tc_set_prof(void* buf) { g_prof_buf = reinterpret_cast<long long*>(buf); }

unsigned int tc_error_flag() {
  unsigned int v = 0;
  cudaMemcpyFromSymbol(&v, tc::g_tc_error, sizeof(v));
  unsigned int zero = 0;
  if (v) cudaMemcpyToSymbol(tc::g_tc_error, &zero, sizeof(zero));
  return v;
}

}  // namespace ctcvr
