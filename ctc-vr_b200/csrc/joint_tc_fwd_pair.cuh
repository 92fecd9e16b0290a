// Forward kernel of the bf16 tcgen05 path on CTA PAIRS (included by joint_tc.cu).  PARKED: correct (tools/fwd_pair_check.py,
// 1e-6 against the single-CTA kernel) but slower at cfg2 (199-232 us against 142 us), so joint_fwd_tc only takes it after
// ctcvr_debug_set_mode(0); DESIGN.md section 5 has the measurements.  It needs the CTA's half of W_out resident in shared
// memory (V = 412, D = 512: 208 KB).
//
//   z      = tanh(enc_proj[b,t,:] + pred_proj[b,u,:])        model/component/joint.py:57-67
//   logits = z . W_out^T + b_out                             model/component/joint.py:68
//   out    : lse, lp_blank, lp_label per lattice cell         log-softmax + gather of torchaudio rnnt_loss
//                                                             (model/component/transducer.py:180-187)
//
// Why pairs.  A 128 x Vp fp32 accumulator is 416 of the 512 TMEM columns, so in the single-CTA kernel the softmax sweep
// of a tile and the MMAs of the next one can never overlap (143 us: 6.7 k MMA + 5.3 k epilogue cycles per tile).
// tcgen05.mma.cta_group::2 with M = 128 gives each CTA 64 rows of the tile at the full tensor rate and lays its
// 64 x N accumulator out over 128 lanes x N/2 columns (tools/pair_probe.cu): one tile is 208 columns per CTA, two
// accumulators fit, and the epilogue of tile i runs under the MMAs of tile i+1.
// W_out stays RESIDENT in shared memory (each CTA holds the 104 rows per N-half that the pair MMA reads from it, all of
// K): no weight streaming at all - the copy engine and L2 carry nothing in steady state, and shared memory only serves
// the tensor core's B reads.  The producers read their enc / pred rows straight from global memory (L2) into registers.
//
// Tiles as in joint_tc_fwd.cuh: 128 cells = nu label columns x 128/nu frames, row R = 32 q' + lane, cell
// (t0 + (q' / nu) * 32 + lane, u0 + q' % nu); CTA r of the pair owns q' = 2r, 2r+1 (rows 64r .. 64r+63).  A one-column
// tile (nu = 1) spans 64 frames only (CTA r: frames 32r .., its second quarter idle).
// TMEM lane l of a CTA: row l & 63, column half l >> 6 (logits v = h*NH + (l >> 6)*NH/2 + j at column h*NH/2 + j).
//
// Roles (640 threads): warp 0 W_out load (once) | warp 1 MMA issuer (leader CTA) | warp 2 TMEM alloc |
// warps 4-11 epilogue (TMEM quarter q = warp & 3, column group eg) | warps 12-19 A producers (quarter q, j):
// tanh(e+p) -> packed bf16 -> tcgen05.st into the A stage; lanes l and l + 64 hold the same row (the pair MMA needs the
// A rows in both lane halves): the two quarters split the k range and swap their halves through shared memory.
#pragma once
#include "tc_common.cuh"

#ifndef CTCVR_EXP
#define CTCVR_EXP 0        // tools/exp_build.sh: timing experiments that drop a piece of the kernel (results invalid)
#endif

namespace ctcvr {
namespace tc {

constexpr int FP_A_STAGES = 3;
constexpr int FP_EPI_GROUPS = 2;
constexpr int FP_EPI_WARPS = 4 * FP_EPI_GROUPS;
constexpr int FP_PROD_WARP0 = 4 + FP_EPI_WARPS;
constexpr int FP_PROD_WARPS = 8;
constexpr int FP_THREADS = (FP_PROD_WARP0 + FP_PROD_WARPS) * 32;
constexpr int FP_A_COL = 416;                    // A stages behind the two accumulators (2 * NH <= 416)
#ifndef FP_POLY_MASK
#define FP_POLY_MASK 0x0                         // which of every 4 softmax exponentials run on the FMA pipe (ex2_poly)
#endif
constexpr int FP_PF = 3;    static_assert(FP_PF == 3, "the producer's slot dispatch is written for 3");                         // k-blocks of enc / pred operands in flight per producer thread (registers)

struct FwdPairParams {
  const __nv_bfloat16* w_t; // tiled W_out: [KB][2][NH][64] bf16, pre-swizzled (prep_weights3_part)
  const __nv_bfloat16* eb;  // bf16 enc_proj [B*T][D]
  const __nv_bfloat16* pb;  // bf16 pred_proj [B*U1][D]
  const float* bias;        // [V]
  const float* bias_l2;     // [Vp] bias * log2(e), -inf beyond V
  const int32_t* targets;   // [B,U1-1]
  const int32_t* t_len;
  const int32_t* u_len;
  const int4* tiles;        // {b, u0, t0, nu}
  const int* ntiles;
  int B, T, U1, D, V, Vp, NH, blank;
  float* lse;
  float* lp_blank;
  float* lp_label;
  long long* prof;
  unsigned int* err_host;
};

struct FwdPairSmem {
  uint32_t w_base, x_base, bar_base;
  float* bias_r;            // [2 lane halves][NH] bias_l2 in TMEM column order
  float2* epi_x;            // [2*groups - 1][64] (max, sum) partials of the other (lane half, group) threads of a row
  float* gx;                // [2][64] gathered blank / label logits
  uint32_t* tmem_ptr;
  __device__ __forceinline__ uint32_t a_full(int i) const { return bar_base + i * 16; }
  __device__ __forceinline__ uint32_t a_empty(int i) const { return bar_base + i * 16 + 8; }
  __device__ __forceinline__ uint32_t acc_full(int i) const { return bar_base + 48 + i * 16; }
  __device__ __forceinline__ uint32_t acc_empty(int i) const { return bar_base + 48 + i * 16 + 8; }
  __device__ __forceinline__ uint32_t w_full() const { return bar_base + 80; }
  __device__ __forceinline__ uint32_t w_ready() const { return bar_base + 88; }
};

__host__ __device__ inline size_t fwdp_smem_bytes(int NH, int D) {
  size_t s = 0;                                                // the dynamic shared memory is declared 1024-byte aligned
  s += (size_t)(D / BK) * 2 * (NH / 2) * 128;                  // resident W half
  s += (size_t)FP_PROD_WARPS * 1024;                           // producer exchange slots
  s += (size_t)2 * NH * 4;
  s += (size_t)(2 * FP_EPI_GROUPS - 1) * 64 * 8 + 2 * 64 * 4;
  s += 16 + 144 + 16;
  return s;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FP_THREADS, 1)
joint_fwd3_kernel(const FwdPairParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  FwdPairSmem L;
  const int NH = p.NH, NQ = NH >> 1, KB = p.D / BK;
  {
    const uint32_t base = smem_u32(smem_raw);
    uint32_t a = base;                    // 1024-byte aligned (W blocks are multiples of 1 KB: NQ % 8 == 0)
    L.w_base = a; a += (uint32_t)KB * 2u * (uint32_t)NQ * 128u;
    L.x_base = a; a += FP_PROD_WARPS * 1024;
    L.bias_r = reinterpret_cast<float*>(smem_raw + (a - base)); a += 2 * NH * 4;
    L.epi_x = reinterpret_cast<float2*>(smem_raw + (a - base)); a += (2 * FP_EPI_GROUPS - 1) * 64 * 8;
    L.gx = reinterpret_cast<float*>(smem_raw + (a - base)); a += 2 * 64 * 4;
    a = (a + 15u) & ~15u;
    L.bar_base = a; a += 144;
    L.tmem_ptr = reinterpret_cast<uint32_t*>(smem_raw + (a - base));
  }
  const int tid = threadIdx.x, warp = uniform_warp_idx(), lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int ntiles = *p.ntiles;
  const int ncl = (int)gridDim.x >> 1, cl = (int)blockIdx.x >> 1;

  if (warp == 0 && lane == 0) {
    g_tc_error_host = p.err_host;
    // a_full / acc_empty collect the CTA's own warps; the leader's also take ONE arrival per phase forwarded by the
    // peer's warps 1 / 0 (a remote arrive per producer warp and k-block cost ~200 cycles each and serialised)
    const uint32_t fwd = rank == 0 ? 1u : 0u;
    for (int i = 0; i < FP_A_STAGES; ++i) { mbar_init(L.a_full(i), FP_PROD_WARPS + fwd); mbar_init(L.a_empty(i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(L.acc_full(i), 1); mbar_init(L.acc_empty(i), FP_EPI_WARPS + fwd); }
    mbar_init(L.w_full(), 1);
    mbar_init(L.w_ready(), 2);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc2(smem_u32(L.tmem_ptr), TMEM_COLS);
  // bias in the column order of a lane half: column cc of half hv is logit v = (cc / NQ) * NH + hv * NQ + cc % NQ
  for (int i = tid; i < 2 * NH; i += FP_THREADS) {
    const int hv = i / NH, cc = i - hv * NH;
    L.bias_r[i] = p.bias_l2[(cc / NQ) * NH + hv * NQ + cc % NQ];
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                       // the peer's barriers exist before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *L.tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ W_out: this CTA's NQ rows of every (k-block, N-half)
    const uint32_t bytes = (uint32_t)NQ * 128u;
    if (elect_one()) {
      mbar_arrive_expect_tx(L.w_full(), bytes * (uint32_t)(2 * KB));
      for (int i = 0; i < 2 * KB; ++i)
        bulk_load(L.w_base + (uint32_t)i * bytes, p.w_t + ((size_t)i * NH + (size_t)rank * NQ) * 64, bytes, L.w_full());
    }
    __syncwarp();
    mbar_wait(L.w_full(), 0, 1);
    warp_arrive_leader(L.w_ready(), rank);
    if (rank == 1) {
      // peer: forward "my epilogue has drained accumulator acc" to the leader, one remote arrive per tile
      int it = 0;
      for (int tile = cl; tile < ntiles; tile += ncl, ++it) {
        mbar_wait(L.acc_empty(it & 1), (uint32_t)(it >> 1) & 1u, 9);
        if (lane == 0) mbar_arrive_remote_relaxed(L.acc_empty(it & 1), 0);
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA; warp-wide loop, one lane issues)
    if (rank == 0) {
      Pipe ap;
      int prof_n = 0;
      const uint32_t idesc = make_idesc_bf16(128, NH);
      const uint64_t w_desc0 = make_desc_sw128(L.w_base);
      const uint32_t w_step = (uint32_t)NQ * 8u;          // one (k-block, N-half) block, in 16-byte descriptor units
      mbar_wait(L.w_ready(), 0, 2);
      int it = 0;
      for (int tile = cl; tile < ntiles; tile += ncl, ++it) {
        const int acc = it & 1;
        const uint32_t aph = (uint32_t)(it >> 1) & 1u;
        if (lane == 0) TC_PROF(1, 100);
        mbar_wait(L.acc_empty(acc), aph ^ 1u, 3);
        if (lane == 0) TC_PROF(1, 101);
        tc_fence_after();
        for (int kb = 0; kb < KB; ++kb) {
          if (!(CTCVR_EXP & 4)) mbar_wait(L.a_full(ap.stage), ap.phase, 4);
          if (lane == 0) TC_PROF(1, 50 + kb);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a = tmem_base + FP_A_COL + ap.stage * 32;
#pragma unroll
            for (int h = 0; h < ((CTCVR_EXP & 16) ? 4 : 2); ++h) {
              const uint64_t bd = w_desc0 + (uint64_t)((kb * 2 + (h & 1)) * w_step);
              const uint32_t d = tmem_base + (uint32_t)(acc * NH + (h & 1) * NQ);
              umma2_bf16_ts(d, a, bd, idesc, kb ? 1u : 0u);
              umma2_bf16_ts(d, a + 8, bd + 2, idesc, 1u);
              umma2_bf16_ts(d, a + 16, bd + 4, idesc, 1u);
              umma2_bf16_ts(d, a + 24, bd + 6, idesc, 1u);
            }
            umma2_commit_mc(L.a_empty(ap.stage), 3);
            if (kb == KB - 1) umma2_commit_mc(L.acc_full(acc), 3);
          }
          __syncwarp();
          ap.advance(FP_A_STAGES);
        }
        if (lane == 0) TC_PROF(1, 102);
      }
    } else {
      // peer: forward "my half of A stage s is in tensor memory" to the leader, one remote arrive per k-block
      Pipe ap;
      for (int tile = cl; tile < ntiles; tile += ncl)
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(L.a_full(ap.stage), ap.phase, 10);
          if (lane == 0) mbar_arrive_remote_relaxed(L.a_full(ap.stage), 0);
          __syncwarp();
          ap.advance(FP_A_STAGES);
        }
    }
  } else if (warp >= 4 && warp < FP_PROD_WARP0) {
    // ------------------------------------------------------------------ epilogue: online log-softmax (base 2)
    // TMEM quarter q: lanes 32q..32q+31 = rows 32(q&1) + lane of this CTA, column half hv = q >> 1; group eg takes a
    // contiguous run of the half's 16-column pieces.  The 2*groups threads of a row keep their own (max, sum) and hand
    // them to thread (hv = 0, eg = 0) through shared memory, together with the gathered blank / label logits.
    const int q = warp & 3, eg = (warp - 4) >> 2;
    const int hv = q >> 1, rloc = ((q & 1) << 5) + lane;
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
    const float* bias_h = L.bias_r + hv * NH;
    int prof_n = 0;
    const float bias_blank = __ldg(p.bias + p.blank);
    const int nch = NH / 16;
    const int c_begin = (nch * eg / FP_EPI_GROUPS) * 16, c_end = (nch * (eg + 1) / FP_EPI_GROUPS) * 16;
    // where logit v lives: lane half (v % NH) / NQ, column (v / NH) * NQ + v % NQ'
    const int bl_w = p.blank % NH, bl_half = bl_w / NQ, bl_col = (p.blank / NH) * NQ + bl_w % NQ;
    int it = 0;
    for (int tile = cl; tile < ntiles; tile += ncl, ++it) {
      const int acc = it & 1;
      const uint32_t aph = (uint32_t)(it >> 1) & 1u;
      int4 ti = p.tiles[tile];
      pin(ti.x); pin(ti.y); pin(ti.z); pin(ti.w);
      const int b = ti.x, nu = ti.w;
      const int lognu = nu >> 1;                          // 4 -> 2, 2 -> 1, 1 -> 0
      const int qp = 2 * (int)rank + (q & 1);             // quarter of the 128-row pair tile
      const int u = ti.y + (qp & (nu - 1));
      const int t = ti.z + (((nu == 1) ? (int)rank : (qp >> lognu)) << 5) + lane;
      int Tb = max(min(p.t_len[b], p.T), 0), Ub = max(min(p.u_len[b], p.U1 - 1), 0);
      pin(Tb); pin(Ub);
      const bool valid = t < Tb && !(nu == 1 && (q & 1));
      int lab = -1;
      if (eg == 0 && u < Ub) {
        lab = p.targets[(size_t)b * (p.U1 - 1) + u];
        if ((unsigned)lab >= (unsigned)p.V) lab = p.blank;  // out-of-range ids cannot index outside the tile
      }
      pin(lab);
      const int lb_w = (lab >= 0 ? lab : 0) % NH, lb_half = lb_w / NQ, lb_col = ((lab >= 0 ? lab : 0) / NH) * NQ + lb_w % NQ;
      mbar_wait(L.acc_full(acc), aph, 6);
      if (tid == 128) TC_PROF(2, 1);
      tc_fence_after();
      const uint32_t ta = tq + (uint32_t)(acc * NH);
      float xb = 0.f, xl = 0.f;
      if (eg == 0) {
        if (hv == bl_half) xb = tmem_ld1(ta + bl_col);
        if (hv == lb_half) xl = tmem_ld1(ta + lb_col);
      }
      float m = kNegInf, s = 0.f;
      float v[16];
      if (c_begin < c_end) tmem_ld16(ta + c_begin, v);
      for (int c0 = c_begin; c0 < c_end; c0 += 16) {
        tmem_ld_wait();
        float y[16];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 bj = *reinterpret_cast<const float4*>(bias_h + c0 + j);
          y[j] = fmaf(v[j], LOG2E, bj.x);
          y[j + 1] = fmaf(v[j + 1], LOG2E, bj.y);
          y[j + 2] = fmaf(v[j + 2], LOG2E, bj.z);
          y[j + 3] = fmaf(v[j + 3], LOG2E, bj.w);
        }
        if (c0 + 16 < c_end) tmem_ld16(ta + c0 + 16, v);    // next piece in flight during the math below
        float cm[4] = {kNegInf, kNegInf, kNegInf, kNegInf};
#pragma unroll
        for (int j = 0; j < 16; j += 4)
#pragma unroll
          for (int e = 0; e < 4; ++e) cm[e] = fmaxf(cm[e], y[j + e]);
        const float nm = fmaxf(m, fmaxf(fmaxf(cm[0], cm[1]), fmaxf(cm[2], cm[3])));
        // a run whose columns so far are all padding (bias -inf) has nm = -inf: subtract 0 instead (every term is 0)
        const float nz = (nm == kNegInf) ? 0.f : nm;
        float ac[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 16; j += 4)
#pragma unroll
          for (int e = 0; e < 4; ++e) ac[e] += ((FP_POLY_MASK >> e) & 1) ? ex2_poly(y[j + e] - nz) : ex2_fast(y[j + e] - nz);
        s = s * ex2_fast(m - nz) + ((ac[0] + ac[1]) + (ac[2] + ac[3]));
        m = nm;
      }
      tmem_ld_wait();
      tc_fence_before();
      warp_arrive(L.acc_empty(acc));
      if (tid == 128) TC_PROF(2, 2);
      const int slot = hv * FP_EPI_GROUPS + eg;            // 0 = the combining thread
      if (slot > 0) L.epi_x[(slot - 1) * 64 + rloc] = make_float2(m, s);
      if (eg == 0) {
        if (hv == bl_half) L.gx[rloc] = xb;
        if (hv == lb_half) L.gx[64 + rloc] = xl;
      }
      named_barrier_sync(3, FP_EPI_WARPS * 32);            // partials and gathers are visible
      if (slot == 0) {
#pragma unroll
        for (int gi = 0; gi < 2 * FP_EPI_GROUPS - 1; ++gi) {
          const float2 o = L.epi_x[gi * 64 + rloc];
          const float nm = fmaxf(m, o.x);
          const float nz = (nm == kNegInf) ? 0.f : nm;
          s = s * ex2_fast(m - nz) + o.y * ex2_fast(o.x - nz);
          m = nm;
        }
        if (valid) {
          const size_t cell = ((size_t)b * p.T + t) * p.U1 + u;
          const float l = (m + lg2_fast(s)) * LN2;
          p.lse[cell] = l;
          p.lp_blank[cell] = L.gx[rloc] + bias_blank - l;
          p.lp_label[cell] = (lab >= 0) ? L.gx[64 + rloc] + __ldg(p.bias + lab) - l : kNegInf;
        }
      }
      named_barrier_sync(3, FP_EPI_WARPS * 32);            // consumed before the next tile overwrites them
    }
  } else if (warp >= FP_PROD_WARP0) {
    // ------------------------------------------------------------------ A producers (A operand lives in TMEM)
    // The pair MMA wants row m of the CTA in TMEM lanes m AND m + 64, so quarters q and q^2 hold the same 32 rows.  The
    // four warps (q, j), (q^2, j), j = 0, 1 of a row group split the 64 k of a k-block into quarters: warp (q, j)
    // computes kq = 2 (q >> 1) + j - tanh(e + p) for 16 k -> 8 packed bf16x2 - writes them into its own lanes
    // (tcgen05.st) and hands them to warp (q^2, j) through shared memory, which writes them into the other lane half:
    // no tanh is computed twice (MUFU is the busiest pipe of this kernel).  The operands come straight from global
    // memory (L2) into registers, FP_PF k-blocks ahead: 32 contiguous bytes of the thread's enc row and of the warp's
    // pred row per k-block - the resident W_out leaves no room for a slab ring, and shared memory has no bandwidth to
    // spare for it either.
    const int pw = warp - FP_PROD_WARP0;
    const int q = warp & 3, j = pw >> 2;
    const int kq = 2 * (q >> 1) + j;                       // my k-quarter; the partner's is kq ^ 2
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)FP_A_COL;
    const uint32_t x_mine = L.x_base + (uint32_t)(pw * 1024 + lane * 16);
    const uint32_t x_peer = L.x_base + (uint32_t)((pw ^ 2) * 1024 + lane * 16);
    const uint32_t bar_id = 4u + (uint32_t)((q & 1) * 2 + j);   // the pair (q, j) / (q^2, j)
    Pipe ap;
    int prof_n = 0;
    // load stream: runs FP_PF k-blocks ahead of the compute stream, across tile boundaries
    int l_tile = cl, l_kb = 0;
    const __nv_bfloat16 *l_e = nullptr, *l_p = nullptr;
    auto l_rows = [&]() {
      if (l_tile >= ntiles) return;                        // past the end: keep re-reading the last rows (discarded)
      const int4 ti = p.tiles[l_tile];
      const int nu = ti.w, lognu = nu >> 1;
      const int qp = 2 * (int)rank + (q & 1);
      const int u = min(ti.y + (qp & (nu - 1)), p.U1 - 1);
      const int t = min(ti.z + (((nu == 1) ? (int)rank : (qp >> lognu)) << 5) + lane, p.T - 1);
      l_e = p.eb + ((size_t)ti.x * p.T + t) * p.D + kq * 16;
      l_p = p.pb + ((size_t)ti.x * p.U1 + u) * p.D + kq * 16;
    };
    U32x8 eq[FP_PF], pq[FP_PF];
    auto l_next = [&](U32x8& e2, U32x8& p2) {
      e2 = ldg256(l_e + l_kb * 64);
      p2 = ldg256(l_p + l_kb * 64);
      if (++l_kb == KB) { l_kb = 0; l_tile += ncl; l_rows(); }
    };
    if (cl < ntiles) {
      l_rows();
#pragma unroll
      for (int i = 0; i < FP_PF; ++i) l_next(eq[i], pq[i]);
    }
    // one k-block; S = its slot of the register ring (compile-time, so the ring is never copied: a register move would
    // wait for the load it moves)
    auto body = [&](auto SLOT, int kb) {
      constexpr int S = decltype(SLOT)::value;
      uint32_t w[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) w[i] = tanh_add_bf16x2_packed(eq[S].v[i], pq[S].v[i]);
      l_next(eq[S], pq[S]);                                // the operands FP_PF k-blocks ahead
      if (tid == FP_PROD_WARP0 * 32) TC_PROF(3, 60 + kb);
      named_barrier_sync(bar_id, 64);                      // the partner has read my previous slot
      sts128(x_mine, w[0], w[1], w[2], w[3]);
      sts128(x_mine + 512u, w[4], w[5], w[6], w[7]);
      mbar_wait(L.a_empty(ap.stage), ap.phase ^ 1u, 8);
      if (tid == FP_PROD_WARP0 * 32) TC_PROF(3, kb);
      tc_fence_after();
      tmem_st8(tq + (uint32_t)(ap.stage * 32 + kq * 8), w);
      named_barrier_sync(bar_id, 64);                      // the partner's quarter is in shared memory
      {
        const uint4 a0 = lds128(x_peer), a1 = lds128(x_peer + 512u);
        const uint32_t o[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        tmem_st8(tq + (uint32_t)(ap.stage * 32 + (kq ^ 2) * 8), o);
      }
      tmem_st_wait();
      tc_fence_before();
      warp_arrive(L.a_full(ap.stage));
      if (tid == FP_PROD_WARP0 * 32) TC_PROF(3, 40 + kb);
      ap.advance(FP_A_STAGES);
    };
    int slot = 0;
    for (int tile = cl; tile < ntiles; tile += ncl) {
      for (int kb = 0; kb < KB; ++kb) {
        if (slot == 0) body(std::integral_constant<int, 0>{}, kb);
        else if (slot == 1) body(std::integral_constant<int, 1>{}, kb);
        else body(std::integral_constant<int, 2>{}, kb);
        slot = (slot + 1 == FP_PF) ? 0 : slot + 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                       // no MMA / remote arrive may target a CTA that has left
  if (warp == 2) tmem_dealloc2(tmem_base, TMEM_COLS);
}

}  // namespace tc
}  // namespace ctcvr
