// Backward kernel 1 of the bf16 tcgen05 path (included by joint_tc.cu):
//   recompute logits -> g = d cost / d logits -> dZ^T = W^T g^T -> dH = dZ (1 - z^2) -> d_enc / d_pred partials,
//   d_bias, and the bf16 spill of the G tiles consumed by the dW GEMM (kernel 2, which recomputes z).
//   (model/component/joint.py:57-68 + torchaudio rnnt_loss backward, model/component/transducer.py:180-187)
//
// Tiles are rectangles of TT frames x P label columns of one utterance (row = tloc*P + ul; <21,6> or <16,8>).
// With the transposed second GEMM (TMEM lane = joint dim d, TMEM column = tile row) both reductions are
// thread-local:
//   d_enc[t]  = sum over the P columns of a frame
//   d_pred[u] = sum over the frame slots, kept in registers across the tiles of one (b, u-split) sweep
// TMEM holds 512 columns, so the logits [128 x Vp] and dZ^T [D x 128] cannot coexist: a tile runs four phases
//   P1  12 warps : tanh k-half slots -> A operand in TMEM (tcgen05.st) + z^T staged in smem and bulk-stored to the
//                  CTA's zt scratch | bulk copies: W_out k-blocks | MMA (TS): logits -> TMEM
//   P2  12 warps : TMEM -> g (bf16) -> smem G tile (K-major over v); exact fp32 blank/label entries
//   P3  bulk copies: W_out^T blocks | MMA (SS): dZ^T[mb] = W^T[mb] . G^T -> TMEM ; one thread bulk-stores the G tile,
//                  the 12 warps sum its columns (d_bias)
//   P4  8 warps  : z^T boxes bulk-loaded back | TMEM -> dH -> d_enc partial (store), d_pred (registers)
// W_out / W_out^T stream through ONE smem ring (3 x NH*128 B in P1, 5 x 16 KB in P3); the 128 KB "GZ" region is the
// z^T staging (P1), the G tile (P2/P3) and the z^T boxes (P4) in turn.
//
// Roles (512 threads): warp 0 bulk-copy ring | warp 1 MMA issuer | warp 2 TMEM alloc | warp 3 TMA slabs |
// warps 4-15: P1 A producers, P2 workers, (4-11) P4 workers.  Role loops run warp-wide with elect.sync around
// the single-thread instructions (tc_common.cuh: elect_one).
#pragma once
#include "tc_common.cuh"

namespace ctcvr {
namespace tc {

constexpr int B_A_STAGES = 3;
constexpr int B_Z_STAGES = 3;                  // z^T staging (P1): 3 warp groups x 3 buffers x 8 KB in the GZ region, one
                                               // (k-block, k-half) slot ([2 row halves][32 d][64 rows]) each
constexpr int B_ACC_COLS = 416;                // TMEM: logits [0, 416) | A stages 416 + 32*stage (P1) ; dZ^T [0, 512) (P3/P4)
constexpr int B_S_STAGES = 2;
constexpr int B_R1_STAGES = 3;                 // W_out ring view   (P1): NH x 128 B per stage
constexpr int B_R3_STAGES = 5;                 // W_out^T ring view (P3): 16 KB per stage, same memory
constexpr int B_SLAB_MAX = 4096;               // slab stage of the largest tile geometry: 24 pred rows (3 KB) + 8 enc rows

// Tile geometry of the single-CTA kernel: TT frames x P label columns, row = tloc*P + ul (TT*P <= 128).
//   <16, 8> : 128 rows, u-splits of <= 16 columns          <21, 6> : 126 rows, u-splits of <= 21 columns
// The host picks the variant that wastes fewer rows for the batch's U (U+1 = 41 -> 2 x 21: 2.4% padding vs 14.6%).
template <int P, int TT>
struct BwdGeom {
  int b, Tb, Ub, W, us, t0, ubase;
  __device__ __forceinline__ void init(const int32_t* t_len, const int32_t* u_len, int T, int U1, int4 ti) {
    b = ti.x;
    Tb = min(t_len[b], T);
    Ub = min(u_len[b], U1 - 1);
    W = Ub + 1;
    const int S = (W + P - 1) / P;
    us = (W + S - 1) / S;
    t0 = ti.z * TT;
    ubase = ti.y * us;
    pin(b); pin(Tb); pin(Ub); pin(W); pin(us); pin(t0); pin(ubase);
  }
  // tile row -> lattice cell; false for padding rows
  __device__ __forceinline__ bool cell(int r, int& t, int& u, int& ul) const {
    const int tloc = r / P;
    ul = r - tloc * P;
    t = t0 + tloc;
    u = ubase + ul;
    return tloc < TT && t < Tb && ul < us && u <= Ub;
  }
};
template <int P>
__host__ __device__ constexpr uint32_t bwd_pred_region() { return (uint32_t)((P * 128 + 1023) / 1024 * 1024); }

constexpr int WORKERS = 384;

struct BwdParams {
  const __nv_bfloat16* w_t;  // tiled W_out   [KB][2][NH][64]     (prep_weights3_kernel)
  const __nv_bfloat16* wt_t; // tiled W_out^T [MB][KBG][128][64]
  const float* bias;        // [V]
  const float* bias_l2;     // [Vp] bias*log2e, -inf beyond V
  const int32_t* targets;
  const int32_t* t_len;
  const int32_t* u_len;
  const int4* tiles;        // {b, u-split, frame block, tile index}
  const int* ntiles;
  int B, T, U1, D, V, Vp, NH, blank;
  const float* lse;
  const float* alpha;
  const float* beta;
  const float* costs;
  const float* grad_costs;
  float clamp;
  // zt [CTA][MB][2][128 d][64 rows] : z^T of the CTA's current tile, pre-swizzled boxes for P4's 1-D bulk loads (element
  //                                   (d, row rr) at box (d>>7, rr>>6), row d&127, chunk ((rr&63)>>3) ^ (d&7), element
  //                                   rr&7).  Reused every tile, so it lives in L2; kernel 2 recomputes z instead.
  // gt [tile][2][KBG][64 rows][64 v] : the G tile as it lies in shared memory (bulk stores), read MN-major by kernel 2
  __nv_bfloat16* zt;
  __nv_bfloat16* gt;
  long Rpad;
  float* d_enc_part;        // [S][B,T,D]
  float* d_pred;            // [B,U1,D] atomic accumulate
  float* d_bias;            // [V] atomic accumulate
  long long* prof;
  unsigned int* err_host;   // mapped host word for bounded-wait time-outs (tc_common.cuh)
};

// Shared memory: [GZ region: G tile (P2/P3) = A ring (P1) = z^T tile (P4)] [weight ring: 3 W stages = 5 W^T stages]
// [slab ring] [bias] [column-sum partials] [barriers]
struct Bwd2Smem {
  uint32_t g_base, r_base, r1_bytes, s_base, bar_base;
  float* bias_l2;
  float* dbp;               // [4][Vp] column-sum partials
  uint32_t* tmem_ptr;
  __device__ __forceinline__ uint32_t g_kblock(int i) const { return g_base + i * A_STAGE_BYTES; }
  __device__ __forceinline__ uint32_t z_stage(int grp, int i) const { return g_base + (uint32_t)(grp * B_Z_STAGES + i) * 8192u; }
  __device__ __forceinline__ uint32_t z_box(int i) const { return g_base + i * A_STAGE_BYTES; }    // (mb*2 + half)
  __device__ __forceinline__ uint32_t r1_stage(int i) const { return r_base + i * r1_bytes; }
  __device__ __forceinline__ uint32_t r3_stage(int i) const { return r_base + i * 16384; }
  __device__ __forceinline__ uint32_t s_stage(int i) const { return s_base + i * B_SLAB_MAX; }
  __device__ __forceinline__ uint32_t a_full(int i) const { return bar_base + i * 16; }
  __device__ __forceinline__ uint32_t a_empty(int i) const { return bar_base + i * 16 + 8; }
  __device__ __forceinline__ uint32_t s_full(int i) const { return bar_base + 48 + i * 16; }
  __device__ __forceinline__ uint32_t s_empty(int i) const { return bar_base + 48 + i * 16 + 8; }
  __device__ __forceinline__ uint32_t r1_full(int i) const { return bar_base + 96 + i * 16; }
  __device__ __forceinline__ uint32_t r1_empty(int i) const { return bar_base + 96 + i * 16 + 8; }
  __device__ __forceinline__ uint32_t r3_full(int i) const { return bar_base + 144 + i * 16; }
  __device__ __forceinline__ uint32_t r3_empty(int i) const { return bar_base + 144 + i * 16 + 8; }
  __device__ __forceinline__ uint32_t z_full(int i) const { return bar_base + 224 + i * 8; }
  __device__ __forceinline__ uint32_t tmem_full() const { return bar_base + 256; }
  __device__ __forceinline__ uint32_t g_full() const { return bar_base + 264; }
  __device__ __forceinline__ uint32_t dz_full() const { return bar_base + 272; }
  __device__ __forceinline__ uint32_t tmem_empty() const { return bar_base + 280; }
  __device__ __forceinline__ uint32_t gs_done() const { return bar_base + 288; }   // column sums have read G
};

__host__ __device__ inline uint32_t bwd2_ring_bytes(int NH) {
  uint32_t a = (uint32_t)B_R1_STAGES * (uint32_t)NH * 128u, b = (uint32_t)B_R3_STAGES * 16384u;
  return a > b ? a : b;
}
__host__ __device__ inline uint32_t bwd2_gz_blocks(int Vp, int D) {
  const uint32_t kbg = (Vp + 63) / 64, zb = 2 * (D / 128), ring = (3 * B_Z_STAGES * 8192 + A_STAGE_BYTES - 1) / A_STAGE_BYTES;
  const uint32_t m = kbg > zb ? kbg : zb;
  return m > ring ? m : ring;               // P1 view: z^T staging
}

__host__ __device__ inline size_t bwd2_smem_bytes(int NH, int Vp, int D) {
  size_t s = 1024;
  s += (size_t)bwd2_gz_blocks(Vp, D) * A_STAGE_BYTES;
  s += bwd2_ring_bytes(NH);
  s = (s + 1023) / 1024 * 1024;
  s += (size_t)B_S_STAGES * B_SLAB_MAX;
  s += (size_t)Vp * 4 + (size_t)4 * Vp * 4;
  s += 304 + 16 + 16;                  // barriers + tmem pointer
  return s;
}

__device__ __forceinline__ void carve_bwd2(Bwd2Smem& L, uint8_t* raw, int NH, int Vp, int D) {
  const uint32_t base = smem_u32(raw);
  uint32_t a = (base + 1023u) & ~1023u;
  L.g_base = a; a += bwd2_gz_blocks(Vp, D) * A_STAGE_BYTES;
  L.r_base = a; L.r1_bytes = (uint32_t)NH * 128u; a += bwd2_ring_bytes(NH);
  a = (a + 1023u) & ~1023u;
  L.s_base = a; a += B_S_STAGES * B_SLAB_MAX;
  L.bias_l2 = reinterpret_cast<float*>(raw + (a - base)); a += Vp * 4;
  L.dbp = reinterpret_cast<float*>(raw + (a - base)); a += 4 * Vp * 4;
  a = (a + 15u) & ~15u;
  L.bar_base = a; a += 304;
  L.tmem_ptr = reinterpret_cast<uint32_t*>(raw + (a - base));
}

template <int P, int TT>
__global__ void __launch_bounds__(NTHREADS, 1)
joint_bwd2_kernel(const __grid_constant__ CUtensorMap tmap_e, const __grid_constant__ CUtensorMap tmap_p,
                  const BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  Bwd2Smem L;
  carve_bwd2(L, smem_raw, p.NH, p.Vp, p.D);
  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
  const int KB = p.D / BK;                 // k-blocks of the logits GEMM
  const int KBG = (p.Vp + 63) / 64;        // k-blocks (over v) of the dZ GEMM
  const int MB = p.D / 128;                // 128-lane blocks of dZ^T
  const int ntiles = *p.ntiles;
  const int tile_begin = (int)(((long)ntiles * blockIdx.x) / gridDim.x);
  const int tile_end = (int)(((long)ntiles * (blockIdx.x + 1)) / gridDim.x);

  if (warp == 0 && lane == 0) {
    g_tc_error_host = p.err_host;
    tma_prefetch_desc(&tmap_e);
    tma_prefetch_desc(&tmap_p);
    for (int i = 0; i < B_A_STAGES; ++i) { mbar_init(L.a_full(i), 8); mbar_init(L.a_empty(i), 1); }
    for (int i = 0; i < B_S_STAGES; ++i) { mbar_init(L.s_full(i), 1); mbar_init(L.s_empty(i), 8); }
    for (int i = 0; i < B_R1_STAGES; ++i) { mbar_init(L.r1_full(i), 1); mbar_init(L.r1_empty(i), 1); }
    for (int i = 0; i < B_R3_STAGES; ++i) { mbar_init(L.r3_full(i), 1); mbar_init(L.r3_empty(i), 1); }
    for (int i = 0; i < 4; ++i) mbar_init(L.z_full(i), 1);
    mbar_init(L.tmem_full(), 1);
    mbar_init(L.g_full(), WORKERS / 32);
    mbar_init(L.dz_full(), 1);
    mbar_init(L.tmem_empty(), 8);
    mbar_init(L.gs_done(), 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32(L.tmem_ptr), TMEM_COLS);
  for (int i = tid; i < p.Vp; i += NTHREADS) L.bias_l2[i] = p.bias_l2[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *L.tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA: W_out (P1), W_out^T (P3), z^T tile (P4)
    Pipe r1, r3;
    int prof_n = 0;
    uint32_t ph = 0;
    const uint32_t r1_bytes = (uint32_t)p.NH * 128u;
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      if (lane == 0) TC_PROF(0, 1);
      // the ring is drained here: the previous tile's dz_full was observed below
      for (int i = 0; i < 2 * KB; ++i) {
        mbar_wait(L.r1_empty(r1.stage), r1.phase ^ 1u, 11);
        if (lane == 0) TC_PROF(0, 10 + i);
        if (elect_one()) {
          mbar_arrive_expect_tx(L.r1_full(r1.stage), r1_bytes);
          bulk_load(L.r1_stage(r1.stage), p.w_t + (size_t)i * p.NH * 64, r1_bytes, L.r1_full(r1.stage));
        }
        __syncwarp();
        r1.advance(B_R1_STAGES);
      }
      if (lane == 0) TC_PROF(0, 2);
      mbar_wait(L.tmem_full(), ph, 12);           // every P1 MMA has completed: the W view of the ring is dead
      if (lane == 0) TC_PROF(0, 3);
      for (int i = 0; i < MB * KBG; ++i) {
        mbar_wait(L.r3_empty(r3.stage), r3.phase ^ 1u, 13);
        if (elect_one()) {
          mbar_arrive_expect_tx(L.r3_full(r3.stage), 16384u);
          bulk_load(L.r3_stage(r3.stage), p.wt_t + (size_t)i * 8192, 16384u, L.r3_full(r3.stage));
        }
        __syncwarp();
        r3.advance(B_R3_STAGES);
      }
      if (lane == 0) TC_PROF(0, 4);
      mbar_wait(L.dz_full(), ph, 14);             // every P3 MMA has completed: G tile and the W^T view are dead
      mbar_wait(L.gs_done(), ph, 16);             // ... and the d_bias column sums have read G
      if (lane == 0) TC_PROF(0, 5);
      if (elect_one()) {
        for (int mb = 0; mb < MB; ++mb) {
          mbar_arrive_expect_tx(L.z_full(mb), 32768u);
          bulk_load(L.z_box(2 * mb), p.zt + (((size_t)blockIdx.x * MB + mb) * 2) * 8192, 32768u, L.z_full(mb));
        }
      }
      __syncwarp();
      ph ^= 1u;
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ TMA: enc / pred slabs
    Pipe sp;
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      const int4 ti = p.tiles[tile];
      const int b = ti.x;
      const int W = min(p.u_len[b], p.U1 - 1) + 1;
      const int S = (W + P - 1) / P, us = (W + S - 1) / S;
      const int prow = b * p.U1 + ti.y * us;
      const int erow = b * p.T + ti.z * TT;
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(L.s_empty(sp.stage), sp.phase ^ 1u, 15);
        if (elect_one()) {
          const uint32_t st = L.s_stage(sp.stage);
          mbar_arrive_expect_tx(L.s_full(sp.stage), (uint32_t)(P + TT) * 128u);
          tma_load_2d(st, &tmap_p, L.s_full(sp.stage), kb * BK, prow);
          tma_load_2d(st + bwd_pred_region<P>(), &tmap_e, L.s_full(sp.stage), kb * BK, erow);
        }
        __syncwarp();
        sp.advance(B_S_STAGES);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (warp-wide loop, one elected lane issues)
    Pipe ap, r1, r3;
    int prof_n = 0;
    uint32_t ph = 0;
    const uint32_t idesc1 = make_idesc_bf16(BM, p.NH);
    const uint32_t idesc2 = make_idesc_bf16(128, BM);
    const uint64_t r1_desc0 = make_desc_sw128(L.r1_stage(0));     // + stage * NH * 8
    const uint64_t r3_desc0 = make_desc_sw128(L.r3_stage(0));     // + stage * 1024
    const uint64_t g_desc0 = make_desc_sw128(L.g_kblock(0));      // + kb * 1024
    const uint32_t r1_step = (uint32_t)p.NH * 8u;
    const int last_nks = (p.Vp - (KBG - 1) * 64) / 16;
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      if (lane == 0) TC_PROF(1, 1);
      mbar_wait(L.tmem_empty(), ph ^ 1u, 20);
      if (lane == 0) TC_PROF(1, 2);
      tc_fence_after();
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(L.a_full(ap.stage), ap.phase, 21);
        if (lane == 0) TC_PROF(1, 50 + kb);
        for (int h = 0; h < 2; ++h) {
          mbar_wait(L.r1_full(r1.stage), r1.phase, 22);
          if (lane == 0) TC_PROF(1, 100 + kb * 2 + h);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a = tmem_base + B_ACC_COLS + ap.stage * 32;
            const uint64_t bd = r1_desc0 + (uint64_t)(r1.stage * r1_step);
            const uint32_t d = tmem_base + h * p.NH;
            umma_bf16_ts(d, a, bd, idesc1, kb ? 1u : 0u);
            umma_bf16_ts(d, a + 8, bd + 2, idesc1, 1u);
            umma_bf16_ts(d, a + 16, bd + 4, idesc1, 1u);
            umma_bf16_ts(d, a + 24, bd + 6, idesc1, 1u);
            umma_commit(L.r1_empty(r1.stage));
            if (h == 1) umma_commit(L.a_empty(ap.stage));
          }
          __syncwarp();
          r1.advance(B_R1_STAGES);
        }
        ap.advance(B_A_STAGES);
      }
      if (elect_one()) umma_commit(L.tmem_full());
      __syncwarp();
      if (lane == 0) TC_PROF(1, 3);
      // ---- P3: dZ^T[mb] (128 d x 128 rows) = W^T[mb] (128 x Vp) . G^T (Vp x 128)
      mbar_wait(L.g_full(), ph, 23);
      if (lane == 0) TC_PROF(1, 4);
      tc_fence_after();
      for (int mb = 0; mb < MB; ++mb)
        for (int kb = 0; kb < KBG; ++kb) {
          mbar_wait(L.r3_full(r3.stage), r3.phase, 24);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t ad = r3_desc0 + (uint64_t)(r3.stage * 1024);
            const uint64_t bd = g_desc0 + (uint64_t)(kb * 1024);
            const uint32_t d = tmem_base + mb * 128;
            const int nks = kb == KBG - 1 ? last_nks : 4;
            if (nks > 0) umma_bf16(d, ad, bd, idesc2, kb ? 1u : 0u);
            if (nks > 1) umma_bf16(d, ad + 2, bd + 2, idesc2, 1u);
            if (nks > 2) umma_bf16(d, ad + 4, bd + 4, idesc2, 1u);
            if (nks > 3) umma_bf16(d, ad + 6, bd + 6, idesc2, 1u);
            umma_commit(L.r3_empty(r3.stage));
          }
          __syncwarp();
          r3.advance(B_R3_STAGES);
        }
      if (elect_one()) umma_commit(L.dz_full());
      __syncwarp();
      if (lane == 0) TC_PROF(1, 5);
      ph ^= 1u;
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ workers (warps 4-15): P1 producers, P2, P4
    const int q = warp & 3;
    const int wg = (warp - 4) >> 2;            // 0..2
    const int wt = tid - 128;                  // 0..383
    const int r = q * 32 + lane;               // P2: tile row ; P4: lane of the d block
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t ph = 0;
    int prof_n = 0;
    int zp = 0;
    float db0 = 0.f, db1 = 0.f;                // d_bias of columns wt and wt + 384
    float pacc[2][P];                          // d_pred sums of d blocks 2wg, 2wg+1 over the tiles of one (b, u-split) sweep
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < P; ++j) pacc[i][j] = 0.f;
    int cur_b = -1, cur_ubase = 0;
    auto flush_pred = [&]() {
      if (cur_b < 0 || wg >= 2) return;
      const int Ub = min(p.u_len[cur_b], p.U1 - 1);
#pragma unroll
      for (int mbl = 0; mbl < 2; ++mbl) {
        const int mb = 2 * wg + mbl;
        if (mb < MB) {
#pragma unroll
          for (int j = 0; j < P; ++j) {
            const int u = cur_ubase + j;
            if (u <= Ub) atomicAdd(p.d_pred + ((size_t)cur_b * p.U1 + u) * p.D + mb * 128 + r, pacc[mbl][j]);
            pacc[mbl][j] = 0.f;
          }
        }
      }
    };
    // producer addressing (A operand lives in TMEM): thread = tile row r = 32q + lane (its own TMEM lane).  All 12
    // worker warps produce: a k-block is two slots (k-half kh: 32 of its 64 k -> 16 packed bf16x2 -> one tcgen05.st);
    // warp group wg takes the slots s = 2 kb + kh with s % 3 == wg.  rows >= TT*P are padding
    const int p_tloc = min(r / P, TT - 1), p_ul = r % P;
    const uint32_t e_row = bwd_pred_region<P>() + (uint32_t)p_tloc * 128u, e_sw = (uint32_t)(p_tloc & 7);
    const uint32_t p_row = (uint32_t)p_ul * 128u, p_sw = (uint32_t)(p_ul & 7);
    // z^T staging (one 8 KB buffer per slot: [2 row halves][32 d][64 rows]): row d of the slot, half r>>6, 16-byte
    // chunk ((r&63)>>3) ^ (d&7), element r&7 - written by stmatrix.trans from the A stage in TMEM (see P1)
    uint32_t kb_base = 0;                       // k-blocks produced before this tile (ring stages follow it)

    for (int tile = tile_begin; tile < tile_end; ++tile) {
      int4 ti = p.tiles[tile];
      pin(ti.x); pin(ti.y); pin(ti.z); pin(ti.w);
      BwdGeom<P, TT> g;
      g.init(p.t_len, p.u_len, p.T, p.U1, ti);
      if (g.b != cur_b || g.ubase != cur_ubase) { flush_pred(); cur_b = g.b; cur_ubase = g.ubase; }
      const size_t rowtile = (size_t)ti.w;

      // ---------------- P1: A k-blocks into TMEM + z^T spill (staged in shared memory, bulk-stored)
      {
        // the staging blocks overlay the z^T tile of the previous iteration, the A columns its dZ^T accumulator:
        // wait until its readers (P4) are done
        mbar_wait(L.tmem_empty(), ph ^ 1u, 40);
        if (tid == 128) TC_PROF(3, 1);
        for (int s = wg; s < 2 * KB; s += 3) {
          const int kb = s >> 1, kh = s & 1;
          const uint32_t kbc = kb_base + (uint32_t)kb;
          const uint32_t a_stg = kbc % B_A_STAGES, a_ph = (kbc / B_A_STAGES) & 1u;
          const uint32_t s_stg = kbc % B_S_STAGES, s_ph = (kbc / B_S_STAGES) & 1u;
          mbar_wait(L.s_full(s_stg), s_ph, 41);
          mbar_wait(L.a_empty(a_stg), a_ph ^ 1u, 42);
          tc_fence_after();
          const uint32_t sb = L.s_stage(s_stg);
          uint4 ev[4], pv[4];
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            const uint32_t c = (uint32_t)(kh * 4 + c4);
            ev[c4] = lds128(sb + e_row + ((c ^ e_sw) << 4));
            pv[c4] = lds128(sb + p_row + ((c ^ p_sw) << 4));
          }
          uint32_t w[16];
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            w[4 * c4 + 0] = tanh_add_bf16x2_packed(ev[c4].x, pv[c4].x);
            w[4 * c4 + 1] = tanh_add_bf16x2_packed(ev[c4].y, pv[c4].y);
            w[4 * c4 + 2] = tanh_add_bf16x2_packed(ev[c4].z, pv[c4].z);
            w[4 * c4 + 3] = tanh_add_bf16x2_packed(ev[c4].w, pv[c4].w);
          }
          const uint32_t acol = (uint32_t)(B_ACC_COLS + a_stg * 32 + kh * 16);
          tmem_st16(tq + acol, w);
          tmem_st_wait();
          // z^T staging straight from the A stage: read the quarter's 32 lanes x 16 columns back in fragment layout
          // (two 16-lane halves) and let stmatrix.trans write the 8x8 tiles transposed - row d of the box gets the
          // 16-byte chunk of 8 consecutive tile rows, the same bytes the per-element stores produced
          const uint32_t zbuf = L.z_stage(wg, zp);
          {
            uint32_t f0[8], f1[8];
            tmem_ld_16x128b_x4(tq + acol, f0);
            tmem_ld_16x128b_x4(tq + (16u << 16) + acol, f1);
            tmem_ld_wait();
            const uint32_t m = (uint32_t)lane >> 3, j = (uint32_t)lane & 7u;
            const uint32_t zrow = zbuf + (uint32_t)(q >> 1) * 4096u + ((m >> 1) * 8u + j) * 128u;
            const uint32_t ch = (uint32_t)(q & 1) * 4u + (m & 1u);                 // + 2 for the second half
            // call c: column groups 2c, 2c+1 -> staging rows 16c + (m>>1)*8 + j
            stmatrix_x4_trans(zrow + ((ch ^ j) << 4), f0[0], f0[1], f0[2], f0[3]);
            stmatrix_x4_trans(zrow + 2048u + ((ch ^ j) << 4), f0[4], f0[5], f0[6], f0[7]);
            stmatrix_x4_trans(zrow + (((ch + 2u) ^ j) << 4), f1[0], f1[1], f1[2], f1[3]);
            stmatrix_x4_trans(zrow + 2048u + (((ch + 2u) ^ j) << 4), f1[4], f1[5], f1[6], f1[7]);
          }
          tc_fence_before();
          fence_proxy_async();
          warp_arrive(L.a_full(a_stg));
          warp_arrive(L.s_empty(s_stg));
          named_barrier_sync(4 + wg, 128);            // the group's staged slot is complete (and fenced) in shared memory
          if (q == 0 && lane == 0) {
            // box (kb>>1, half hh): rows (kb&1)*64 + kh*32 .. +31 of 128 B
            __nv_bfloat16* zdst = p.zt + (((size_t)blockIdx.x * MB + (kb >> 1)) * 2) * 8192 + ((kb & 1) * 64 + kh * 32) * 64;
            bulk_store(zdst, zbuf, 4096u);
            bulk_store(zdst + 8192, zbuf + 4096u, 4096u);
            bulk_commit();
            bulk_wait_read<1>();                      // the slot staged before this one has left shared memory
          }
          zp = (zp + 1 == B_Z_STAGES) ? 0 : zp + 1;
        }
        kb_base += (uint32_t)KB;
        if (q == 0 && lane == 0) bulk_wait_all<0>();  // z^T is in global memory before P4's bulk loads (ordered via g_full)
        if (tid == 128) TC_PROF(3, 2);
      }

      // ---------------- P2: g = d cost / d logits for row r, column chunks wg, wg+3, ...
      int t, u, ul;
      const bool valid = g.cell(r, t, u, ul);
      float k_all = kNegInf, k_blank = kNegInf, k_label = kNegInf, scale = 0.f;
      int lab = -1;
      if (valid) {
        const size_t cell = ((size_t)g.b * p.T + t) * p.U1 + u;
        const float al = p.alpha[cell], be = p.beta[cell], cost = p.costs[g.b], l = p.lse[cell];
        k_all = al + be + cost - l;
        float bnext = kNegInf;
        if (t + 1 < g.Tb) bnext = p.beta[cell + p.U1];
        else if (u == g.Ub) bnext = 0.f;
        k_blank = al + bnext + cost - l;
        if (u < g.Ub) { k_label = al + p.beta[cell + 1] + cost - l; lab = p.targets[(size_t)g.b * (p.U1 - 1) + u]; }
        scale = p.grad_costs[g.b];
      }
      // fast path: no clamp and a positive cost gradient (uniform per tile): fold log2(scale) into the exponent
      const float sc_tile = p.grad_costs[g.b];
      const bool fast = !(p.clamp > 0.f) && sc_tile > 0.f;
      const float kr = (valid && fast) ? fmaf(k_all, LOG2E, lg2_fast(scale)) : kNegInf;
      if (tid == 128) TC_PROF(2, 1);
      mbar_wait(L.tmem_full(), ph, 30);
      named_barrier_sync(7, WORKERS);            // the z^T staging blocks (overlaid by G) have been stored
      if (tid == 128) TC_PROF(2, 2);
      tc_fence_after();
      // 16-column pieces (two per 32-column chunk wg, wg+3, ..); the next piece's TMEM load is in flight during the math
      const int npieces = 2 * ((p.Vp / 32 - wg + 2) / 3);
      auto piece_col = [&](int i) { return wg * 32 + (i >> 1) * 96 + (i & 1) * 16; };
      float v[16];
      if (npieces > 0) tmem_ld16(tq + piece_col(0), v);
      for (int pi = 0; pi < npieces; ++pi) {
        const int c0 = piece_col(pi);
        tmem_ld_wait();
        float y[16];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 bj = *reinterpret_cast<const float4*>(L.bias_l2 + c0 + j);
          y[j] = fmaf(v[j], LOG2E, bj.x);
          y[j + 1] = fmaf(v[j + 1], LOG2E, bj.y);
          y[j + 2] = fmaf(v[j + 2], LOG2E, bj.z);
          y[j + 3] = fmaf(v[j + 3], LOG2E, bj.w);
        }
        if (pi + 1 < npieces) tmem_ld16(tq + piece_col(pi + 1), v);
        uint32_t pk[8];
        if (fast) {
#pragma unroll
          for (int j = 0; j < 16; j += 2) pk[j >> 1] = pack_bf16(ex2_fast(y[j] + kr), ex2_fast(y[j + 1] + kr));
        } else {
          const float ka2 = k_all * LOG2E;
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            float gg[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              float gv = ex2_fast(y[j + e] + ka2);
              if (p.clamp > 0.f) gv = fminf(gv, p.clamp);
              gg[e] = valid ? gv * scale : 0.f;
            }
            pk[j >> 1] = pack_bf16(gg[0], gg[1]);
          }
        }
        // G tile: k-block c0/64, row r, 16-byte chunks (c0%64)/8, +1, 128B swizzle
        const uint32_t gb = L.g_kblock(c0 >> 6) + r * 128;
        const int ch0 = (c0 & 63) >> 3;
#pragma unroll
        for (int i = 0; i < 2; ++i)
          sts128(gb + (((ch0 + i) ^ (r & 7)) << 4), pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
      }
      named_barrier_sync(2, WORKERS);            // every generic entry of G is written
      if (wg == 0) {
        // exact (fp32, single rounding) blank and label entries of row r
        const float xb = tmem_ld1(tq + p.blank);
        float xl = 0.f;
        for (int i = 0; i < P; ++i) {
          const int ui = g.ubase + i;
          int col = 0;
          if (ui < g.Ub) col = p.targets[(size_t)g.b * (p.U1 - 1) + ui];
          const float xi = tmem_ld1(tq + col);
          if (ul == i) xl = xi;
        }
        tmem_ld_wait();
        if (valid) {
          auto entry = [&](float x, float kc1, float kc2) {
            float gv = __expf(x + k_all) - __expf(x + kc1);
            if (kc2 != kNegInf) gv -= __expf(x + kc2);
            if (p.clamp > 0.f) gv = fminf(fmaxf(gv, -p.clamp), p.clamp);
            return gv * scale;
          };
          auto put = [&](int col, float val) {
            const unsigned short h = __bfloat16_as_ushort(__float2bfloat16(val));
            const uint32_t a = L.g_kblock(col >> 6) + r * 128 + ((((col & 63) >> 3) ^ (r & 7)) << 4) + (col & 7) * 2;
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"(h) : "memory");
          };
          const float xbb = xb + __ldg(p.bias + p.blank);
          put(p.blank, entry(xbb, k_blank, (lab == p.blank) ? k_label : kNegInf));
          if (lab >= 0 && lab != p.blank) put(lab, entry(xl + __ldg(p.bias + lab), k_label, kNegInf));
        }
      }
      fence_proxy_async();
      tc_fence_before();
      warp_arrive(L.g_full());
      if (tid == 128) TC_PROF(2, 3);

      // ---------------- P3 (MMA busy): d_bias = column sums of the final G tile
      mbar_wait(L.g_full(), ph, 31);
      if (wt == 0) {
        // spill the finished G tile as it lies in shared memory: k-block i (64 label columns) -> two 8 KB boxes
        // [64 rows][64 v] (rows 0-63 / 64-127), read back MN-major by the dW GEMM.  No per-thread stores.
        __nv_bfloat16* gdst = p.gt + (rowtile * 2) * (size_t)KBG * 4096;
        for (int i = 0; i < KBG; ++i) {
          bulk_store(gdst + (size_t)i * 4096, L.g_kblock(i), 8192u);
          bulk_store(gdst + (size_t)(KBG + i) * 4096, L.g_kblock(i) + 8192u, 8192u);
        }
        bulk_commit();
      }
      {
        const int nchunk = p.Vp >> 3;
        if (wt < 4 * nchunk) {
          const int rg = wt / nchunk, c = wt - rg * nchunk;
          const uint32_t gb = L.g_kblock(c >> 3) + (uint32_t)(rg * 32) * 128u;
          float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
          for (int rr = 0; rr < 32; ++rr) {
            const uint4 x = lds128(gb + rr * 128 + ((((c & 7) ^ (rr & 7))) << 4));
            acc[0] += __uint_as_float(x.x << 16); acc[1] += __uint_as_float(x.x & 0xffff0000u);
            acc[2] += __uint_as_float(x.y << 16); acc[3] += __uint_as_float(x.y & 0xffff0000u);
            acc[4] += __uint_as_float(x.z << 16); acc[5] += __uint_as_float(x.z & 0xffff0000u);
            acc[6] += __uint_as_float(x.w << 16); acc[7] += __uint_as_float(x.w & 0xffff0000u);
          }
          float* o = L.dbp + rg * p.Vp + c * 8;
          *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
          *reinterpret_cast<float4*>(o + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
        }
        named_barrier_sync(2, WORKERS);
        if (wt == 0) {
          bulk_wait_read<0>();                        // the bulk stores have read G
          mbar_arrive(L.gs_done());
        }
        if (wt < p.Vp) db0 += (L.dbp[wt] + L.dbp[p.Vp + wt]) + (L.dbp[2 * p.Vp + wt] + L.dbp[3 * p.Vp + wt]);
        if (wt + WORKERS < p.Vp)
          db1 += (L.dbp[wt + WORKERS] + L.dbp[p.Vp + wt + WORKERS]) + (L.dbp[2 * p.Vp + wt + WORKERS] + L.dbp[3 * p.Vp + wt + WORKERS]);
      }

      // ---------------- P4 (warps 4-11): dH = dZ * (1 - z^2); reductions.  Warp group wg owns d blocks 2wg, 2wg+1.
      // z^T arrives in shared memory (bulk copy issued after dz_full): box (mb, half) = [128 d][64 rows], 128B swizzle.
      // TMEM column c of a d block = tile row c = (frame slot c / P, label slot c % P): everything is static after
      // unrolling, so any (P, TT) works with aligned 32-column loads.
      if (wg < 2) {
        mbar_wait(L.dz_full(), ph, 32);
        if (tid == 128) TC_PROF(2, 4);
        tc_fence_after();
#pragma unroll
        for (int mbl = 0; mbl < 2; ++mbl) {
          const int mb = 2 * wg + mbl;
          if (mb < MB) {
            mbar_wait(L.z_full(mb), ph, 33);
            if (tid == 128) TC_PROF(2, 40 + mb);
            const int d = mb * 128 + r;
            float es[TT];
#pragma unroll
            for (int i = 0; i < TT; ++i) es[i] = 0.f;
            float v[32];
            tmem_ld32(tq + mb * 128, v);
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
              const uint32_t zb = L.z_box(2 * mb + (ch >> 1)) + (uint32_t)r * 128u;
              uint4 zc[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) zc[i] = lds128(zb + ((((ch & 1) * 4 + i) ^ (r & 7)) << 4));
              tmem_ld_wait();
              float w[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) w[j] = v[j];
              if (ch < 3) tmem_ld32(tq + mb * 128 + (ch + 1) * 32, v);
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int c = ch * 32 + j;                     // tile row
                if (c < TT * P) {
                  const uint4 zq = zc[j >> 3];
                  const uint32_t zw = ((j & 7) >> 1) == 0 ? zq.x : ((j & 7) >> 1) == 1 ? zq.y : ((j & 7) >> 1) == 2 ? zq.z : zq.w;
                  const float z = (j & 1) ? __uint_as_float(zw & 0xffff0000u) : __uint_as_float(zw << 16);
                  const float h = w[j] * fmaf(-z, z, 1.f);
                  es[c / P] += h;
                  pacc[mbl][c % P] += h;
                }
              }
            }
#pragma unroll
            for (int i = 0; i < TT; ++i) {
              const int tt = g.t0 + i;
              if (tt < g.Tb) p.d_enc_part[(((size_t)ti.y * p.B + g.b) * p.T + tt) * p.D + d] = es[i];
            }
          }
        }
        tc_fence_before();
        warp_arrive(L.tmem_empty());
        if (tid == 128) TC_PROF(2, 5);
      }
      ph ^= 1u;
    }
    flush_pred();
    if (tile_end > tile_begin) {
      if (wt < p.V) atomicAdd(p.d_bias + wt, db0);
      if (wt + WORKERS < p.V) atomicAdd(p.d_bias + wt + WORKERS, db1);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace tc
}  // namespace ctcvr
