// Backward kernel 1 on CTA PAIRS (cta_group::2) - the default for D % 256 == 0.  Same phases and data flow as
// joint_tc_bwd.cuh; the two CTAs of a cluster process tiles 2i and 2i+1 (same utterance, same u-split) in lock step:
//   P1  one tcgen05.mma of M = 256 covers both CTAs' A tiles; each CTA loads only HALF of every W_out stage
//   P3  dZ^T for BOTH tiles at once: M = 256 joint dims (CTA r owns d blocks r, r+2), N = 256 rows (each CTA's own G
//       tile is its half of the B operand); each CTA loads only ITS W_out^T blocks (a quarter of the traffic per tile)
//   P4  CTA r reduces its joint dims for both tiles (z^T of the peer's tile comes from the global spill)
// Shared-memory bandwidth (tensor-core operand reads + copy-engine writes + producers) is what bounds the single-CTA
// kernel; pairing halves the operand traffic per flop.  Only the leader CTA issues MMAs; the peer's warp 1 forwards
// its local "full" barriers to the leader with remote mbarrier arrives; commits are multicast to both CTAs.
#pragma once
#include "joint_tc_bwd.cuh"

namespace ctcvr {
namespace tc {

constexpr int BP_R1_STAGES = 6;                // half W_out stages: (NH/2) x 128 B
constexpr int BP_A_STAGES = 8;                 // the whole A tile: the GZ region is free during P1, and the producers never
                                               // wait for the (cross-CTA, high-latency) release of a stage inside a tile
struct BwdPairBars {
  uint32_t base;
  __device__ __forceinline__ uint32_t peer_a_full(int i) const { return base + i * 8; }            // 8
  __device__ __forceinline__ uint32_t peer_r1_full(int i) const { return base + 64 + i * 8; }      // 6
  __device__ __forceinline__ uint32_t peer_r3_full(int i) const { return base + 112 + i * 8; }     // 5
  __device__ __forceinline__ uint32_t peer_g_full() const { return base + 152; }
  __device__ __forceinline__ uint32_t peer_tmem_empty() const { return base + 160; }
  __device__ __forceinline__ uint32_t r1p_full(int i) const { return base + 168 + i * 16; }        // 6 own half-stage rings
  __device__ __forceinline__ uint32_t r1p_empty(int i) const { return base + 168 + i * 16 + 8; }
  __device__ __forceinline__ uint32_t a_full(int i) const { return base + 264 + i * 16; }          // 8
  __device__ __forceinline__ uint32_t a_empty(int i) const { return base + 264 + i * 16 + 8; }
};

__global__ void __launch_bounds__(NTHREADS, 1)
joint_bwd2p_kernel(const __grid_constant__ CUtensorMap tmap_e, const __grid_constant__ CUtensorMap tmap_p,
                  const BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  Bwd2Smem L;
  carve_bwd2(L, smem_raw, p.NH, p.Vp, p.D);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KB = p.D / BK;                 // k-blocks of the logits GEMM
  const int KBG = (p.Vp + 63) / 64;        // k-blocks (over v) of the dZ GEMM
  const int MB = p.D / 128;                // 128-lane blocks of dZ^T
  const int ntiles = *p.ntiles;                          // even: every (b, u-split) sweep is padded to an even count
  const uint32_t rank = cluster_ctarank();               // 0 = leader (issues the MMAs), 1 = peer
  const int npairs = ntiles >> 1, ncl = (int)gridDim.x >> 1, cl = (int)blockIdx.x >> 1;
  const int pair_begin = (int)(((long)npairs * cl) / ncl), pair_end = (int)(((long)npairs * (cl + 1)) / ncl);
  const int tile_begin = 2 * pair_begin + (int)rank, tile_end = 2 * pair_end;      // my tiles: tile_begin, +2, +4, ...
  BwdPairBars X;
  X.base = L.bar_base + 304;
  const uint32_t r1p_bytes = (uint32_t)(p.NH / 2) * 128u;
  auto r1p_stage = [&](int i) { return L.r_base + (uint32_t)i * r1p_bytes; };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_e);
    tma_prefetch_desc(&tmap_p);
    for (int i = 0; i < BP_A_STAGES; ++i) { mbar_init(X.a_full(i), PROD_THREADS); mbar_init(X.a_empty(i), 1); mbar_init(X.peer_a_full(i), 1); }
    for (int i = 0; i < B_S_STAGES; ++i) { mbar_init(L.s_full(i), 1); mbar_init(L.s_empty(i), PROD_THREADS); }
    for (int i = 0; i < BP_R1_STAGES; ++i) { mbar_init(X.r1p_full(i), 1); mbar_init(X.r1p_empty(i), 1); mbar_init(X.peer_r1_full(i), 1); }
    for (int i = 0; i < B_R3_STAGES; ++i) mbar_init(X.peer_r3_full(i), 1);
    mbar_init(X.peer_g_full(), 1);
    mbar_init(X.peer_tmem_empty(), 1);
    for (int i = 0; i < B_R3_STAGES; ++i) { mbar_init(L.r3_full(i), 1); mbar_init(L.r3_empty(i), 1); }
    for (int i = 0; i < 4; ++i) mbar_init(L.z_full(i), 1);
    mbar_init(L.tmem_full(), 1);
    mbar_init(L.g_full(), WORKERS);
    mbar_init(L.dz_full(), 1);
    mbar_init(L.tmem_empty(), 256);
    mbar_init(L.gs_done(), 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc2(smem_u32(L.tmem_ptr), TMEM_COLS);
  for (int i = tid; i < p.Vp; i += NTHREADS) L.bias_l2[i] = p.bias_l2[i];
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                       // the peer's barriers exist before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *L.tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA: W_out (P1), W_out^T (P3), z^T tile (P4)
    if (lane == 0) {
      Pipe r1, r3;
      int prof_n = 0;
      uint32_t ph = 0;
      for (int tile = tile_begin; tile < tile_end; tile += 2) {
        const size_t rt_own = (size_t)p.tiles[tile].w, rt_peer = (size_t)p.tiles[tile ^ 1].w;
        const size_t rt_x[2] = {rank == 0 ? rt_own : rt_peer, rank == 0 ? rt_peer : rt_own};   // x = 0: leader's tile
        TC_PROF(0, 1);
        for (int kb = 0; kb < KB; ++kb)
          for (int h = 0; h < 2; ++h) {
            mbar_wait(X.r1p_empty(r1.stage), r1.phase ^ 1u, 11);
            mbar_arrive_expect_tx(X.r1p_full(r1.stage), r1p_bytes);
            bulk_load(r1p_stage(r1.stage), p.w_t + ((size_t)(kb * 2 + h) * p.NH + rank * (p.NH / 2)) * 64, r1p_bytes,
                      X.r1p_full(r1.stage));
            r1.advance(BP_R1_STAGES);
          }
        TC_PROF(0, 2);
        mbar_wait(L.tmem_full(), ph, 12);           // every P1 MMA has completed: the W view of the ring is dead
        TC_PROF(0, 3);
        for (int pb = 0; pb < MB / 2; ++pb)
          for (int kb = 0; kb < KBG; ++kb) {
            mbar_wait(L.r3_empty(r3.stage), r3.phase ^ 1u, 13);
            mbar_arrive_expect_tx(L.r3_full(r3.stage), 16384u);
            bulk_load(L.r3_stage(r3.stage), p.wt_t + (size_t)((2 * pb + (int)rank) * KBG + kb) * 8192, 16384u,
                      L.r3_full(r3.stage));
            r3.advance(B_R3_STAGES);
          }
        TC_PROF(0, 4);
        mbar_wait(L.dz_full(), ph, 14);             // every P3 MMA has completed: G tile and the W^T view are dead
        mbar_wait(L.gs_done(), ph, 16);             // ... and the d_bias column sums have read G
        TC_PROF(0, 5);
        fence_proxy_async_global();                 // z^T was written with st.global (by this CTA and by its peer)
        for (int pb = 0; pb < MB / 2; ++pb) {
          mbar_arrive_expect_tx(L.z_full(pb), 65536u);
          for (int x = 0; x < 2; ++x)
            bulk_load(L.z_box((pb * 2 + x) * 2), p.zt + ((rt_x[x] * MB + 2 * pb + rank) * 2) * 8192, 32768u, L.z_full(pb));
        }
        ph ^= 1u;
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    // ------------------------------------------------------------------ TMA: enc / pred slabs
    if (lane == 0) {
      Pipe sp;
      for (int tile = tile_begin; tile < tile_end; tile += 2) {
        const int4 ti = p.tiles[tile];
        const int b = ti.x;
        const int W = min(p.u_len[b], p.U1 - 1) + 1;
        const int S = (W + 15) >> 4, us = (W + S - 1) / S;
        const int prow = b * p.U1 + ti.y * us;
        const int erow = b * p.T + ti.z * 8;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(L.s_empty(sp.stage), sp.phase ^ 1u, 15);
          const uint32_t st = L.s_stage(sp.stage);
          mbar_arrive_expect_tx(L.s_full(sp.stage), (uint32_t)B_SLAB_BYTES);
          tma_load_2d(st, &tmap_p, L.s_full(sp.stage), kb * BK, prow);
          tma_load_2d(st + 2048, &tmap_e, L.s_full(sp.stage), kb * BK, erow);
          sp.advance(B_S_STAGES);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      Pipe ap, r1, r3;
      int prof_n = 0;
      uint32_t ph = 0;
      if (rank == 0) {
        const uint32_t idesc1 = make_idesc_bf16(256, p.NH);
        const uint32_t idesc2 = make_idesc_bf16(256, 256);
        for (int tile = tile_begin; tile < tile_end; tile += 2) {
          TC_PROF(1, 1);
          mbar_wait(L.tmem_empty(), ph ^ 1u, 20);
          mbar_wait(X.peer_tmem_empty(), ph ^ 1u, 25);
          TC_PROF(1, 2);
          tc_fence_after();
          for (int kb = 0; kb < KB; ++kb) {
            mbar_wait(X.a_full(ap.stage), ap.phase, 21);
            mbar_wait(X.peer_a_full(ap.stage), ap.phase, 26);
            TC_PROF(1, 50 + kb);
            for (int h = 0; h < 2; ++h) {
              mbar_wait(X.r1p_full(r1.stage), r1.phase, 22);
              mbar_wait(X.peer_r1_full(r1.stage), r1.phase, 27);
              TC_PROF(1, 100 + kb * 2 + h);
              tc_fence_after();
#pragma unroll
              for (int ks = 0; ks < BK / 16; ++ks)
                umma2_bf16(tmem_base + h * p.NH, make_desc_sw128(L.a_stage(ap.stage) + ks * 32),
                           make_desc_sw128(r1p_stage(r1.stage) + ks * 32), idesc1, (kb | ks) ? 1u : 0u);
              umma2_commit_mc(X.r1p_empty(r1.stage), 3);
              r1.advance(BP_R1_STAGES);
            }
            umma2_commit_mc(X.a_empty(ap.stage), 3);
            ap.advance(BP_A_STAGES);
          }
          umma2_commit_mc(L.tmem_full(), 3);
          TC_PROF(1, 3);
          // ---- P3: dZ^T (256 joint dims x 256 rows) per pair block: A = both CTAs' W^T blocks, B = both CTAs' G tiles
          mbar_wait(L.g_full(), ph, 23);
          mbar_wait(X.peer_g_full(), ph, 28);
          TC_PROF(1, 4);
          tc_fence_after();
          for (int pb = 0; pb < MB / 2; ++pb)
            for (int kb = 0; kb < KBG; ++kb) {
              mbar_wait(L.r3_full(r3.stage), r3.phase, 24);
              mbar_wait(X.peer_r3_full(r3.stage), r3.phase, 29);
              TC_PROF(1, 200 + pb * KBG + kb);
              tc_fence_after();
              const int nks = min(4, (p.Vp - kb * 64) / 16);
              for (int ks = 0; ks < nks; ++ks)
                umma2_bf16(tmem_base + pb * 256, make_desc_sw128(L.r3_stage(r3.stage) + ks * 32),
                           make_desc_sw128(L.g_kblock(kb) + ks * 32), idesc2, (kb | ks) ? 1u : 0u);
              umma2_commit_mc(L.r3_empty(r3.stage), 3);
              r3.advance(B_R3_STAGES);
            }
          umma2_commit_mc(L.dz_full(), 3);
          TC_PROF(1, 5);
          ph ^= 1u;
        }
      } else {
        // peer: forward "my operands are ready" / "my TMEM is drained" to the leader's barriers, in the leader's order
        for (int tile = tile_begin; tile < tile_end; tile += 2) {
          if (tile != tile_begin) {                 // the leader's first wait on this barrier passes by parity
            mbar_wait(L.tmem_empty(), ph ^ 1u, 20);
            mbar_arrive_remote(X.peer_tmem_empty(), 0);
          }
          for (int kb = 0; kb < KB; ++kb) {
            mbar_wait(X.a_full(ap.stage), ap.phase, 21);
            mbar_arrive_remote(X.peer_a_full(ap.stage), 0);
            for (int h = 0; h < 2; ++h) {
              mbar_wait(X.r1p_full(r1.stage), r1.phase, 22);
              mbar_arrive_remote(X.peer_r1_full(r1.stage), 0);
              r1.advance(BP_R1_STAGES);
            }
            ap.advance(BP_A_STAGES);
          }
          mbar_wait(L.g_full(), ph, 23);
          mbar_arrive_remote(X.peer_g_full(), 0);
          for (int pb = 0; pb < MB / 2; ++pb)
            for (int kb = 0; kb < KBG; ++kb) {
              mbar_wait(L.r3_full(r3.stage), r3.phase, 24);
              mbar_arrive_remote(X.peer_r3_full(r3.stage), 0);
              r3.advance(B_R3_STAGES);
            }
          ph ^= 1u;
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ workers (warps 4-15), producers (8-15)
    const int q = warp & 3;
    const int wg = (warp - 4) >> 2;            // 0..2
    const int wt = tid - 128;                  // 0..383
    const int r = q * 32 + lane;               // P2: tile row ; P4: lane of the d block
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool producer = warp >= 8;
    const int pc = (warp - 8) & 7;             // producer: 16-byte chunk of the k-block handled by this warp
    uint32_t ph = 0;
    int prof_n = 0;
    Pipe ap, sp;
    float db0 = 0.f, db1 = 0.f;                // d_bias of columns wt and wt + 384
    float pacc[16];                            // d_pred sums of d block 2*wg + rank over the tiles of one (b, u-split) sweep
#pragma unroll
    for (int j = 0; j < 16; ++j) pacc[j] = 0.f;
    int cur_b = -1, cur_ubase = 0;
    auto flush_pred = [&]() {
      if (cur_b < 0 || wg >= 2 || wg >= MB / 2) return;
      const int Ub = min(p.u_len[cur_b], p.U1 - 1);
      const int mb = 2 * wg + (int)rank;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int u = cur_ubase + j;
        if (u <= Ub) atomicAdd(p.d_pred + ((size_t)cur_b * p.U1 + u) * p.D + mb * 128 + r, pacc[j]);
        pacc[j] = 0.f;
      }
    };
    // producer addressing (rows lane, lane+32, lane+64, lane+96 of the tile; row = tloc*16 + ul)
    const int ul_p = lane & 15;
    const uint32_t p_off = (uint32_t)ul_p * 128u + (uint32_t)((pc ^ (ul_p & 7)) << 4);
    uint32_t e_off[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int tloc = (lane >> 4) + 2 * j;
      e_off[j] = 2048u + (uint32_t)tloc * 128u + (uint32_t)((pc ^ tloc) << 4);
    }
    const uint32_t a_off = (uint32_t)lane * 128u + (uint32_t)((pc ^ (lane & 7)) << 4);

    for (int tile = tile_begin; tile < tile_end; tile += 2) {
      int4 ti = p.tiles[tile];
      pin(ti.x); pin(ti.y); pin(ti.z); pin(ti.w);
      int tz_peer = p.tiles[tile ^ 1].z;                     // frame block of the peer's tile (same b, same u-split)
      pin(tz_peer);
      const RowMap g = tile_geometry<TILE_RECT>(p.t_len, p.u_len, p.T, p.U1, ti);
      if (g.b != cur_b || g.ubase != cur_ubase) { flush_pred(); cur_b = g.b; cur_ubase = g.ubase; }
      const size_t rowtile = (size_t)ti.w;

      // ---------------- P1 (warps 8-15): A tile k-blocks + z^T spill
      if (producer) {
        // the A ring overlays the z^T tile of the previous iteration: wait until its readers (P4) are done
        mbar_wait(L.tmem_empty(), ph ^ 1u, 40);
        if (tid == 256) TC_PROF(3, 1);
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(L.s_full(sp.stage), sp.phase, 41);
          mbar_wait(X.a_empty(ap.stage), ap.phase ^ 1u, 42);
          const uint32_t sb = L.s_stage(sp.stage);
          const uint32_t ab = L.a_stage(ap.stage) + a_off;
          const uint4 pv = lds128(sb + p_off);
          // d = kb*64 + pc*8 + e -> box row (kb&1)*64 + pc*8 + e of d block kb>>1; tile row lane + 32j -> half j>>1
          unsigned short* z = reinterpret_cast<unsigned short*>(p.zt) + ((rowtile * MB + (kb >> 1)) * 2) * 8192 +
                              ((kb & 1) * 64 + pc * 8) * 64 + (lane & 7);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 ev = lds128(sb + e_off[j]);
            uint32_t w[4];
            w[0] = tanh_add_bf16x2_packed(ev.x, pv.x);
            w[1] = tanh_add_bf16x2_packed(ev.y, pv.y);
            w[2] = tanh_add_bf16x2_packed(ev.z, pv.z);
            w[3] = tanh_add_bf16x2_packed(ev.w, pv.w);
            sts128(ab + j * 4096, w[0], w[1], w[2], w[3]);
            const int chunk = (lane >> 3) + 4 * (j & 1);           // 8-row chunk of the 64-row half
            unsigned short* zj = z + (j >> 1) * 8192;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              zj[(2 * e) * 64 + ((chunk ^ (2 * e)) << 3)] = (unsigned short)(w[e] & 0xffffu);
              zj[(2 * e + 1) * 64 + ((chunk ^ (2 * e + 1)) << 3)] = (unsigned short)(w[e] >> 16);
            }
          }
          fence_proxy_async();
          mbar_arrive(X.a_full(ap.stage));
          mbar_arrive(L.s_empty(sp.stage));
          ap.advance(BP_A_STAGES);
          sp.advance(B_S_STAGES);
        }
        if (tid == 256) TC_PROF(3, 2);
        __threadfence();                             // the peer CTA reads this z^T spill too
        fence_proxy_async_global();                  // z^T (st.global above) is read back by the P4 bulk copies
        if (tid == 256) TC_PROF(3, 3);
      }

      // ---------------- P2: g = d cost / d logits for row r, column chunks wg, wg+3, ...
      int t, u;
      const bool valid = row_cell<TILE_RECT>(g, ti, r, t, u);
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
      if (tid == 128) TC_PROF(2, 2);
      tc_fence_after();
      for (int c0 = wg * 32; c0 < p.Vp; c0 += 96) {
        float v[32];
        tmem_ld32(tq + c0, v);
        tmem_ld_wait();
        uint32_t pk[16];
        if (fast) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 bj = *reinterpret_cast<const float4*>(L.bias_l2 + c0 + j);
            const float g0 = ex2_fast(fmaf(v[j], LOG2E, bj.x) + kr);
            const float g1 = ex2_fast(fmaf(v[j + 1], LOG2E, bj.y) + kr);
            const float g2 = ex2_fast(fmaf(v[j + 2], LOG2E, bj.z) + kr);
            const float g3 = ex2_fast(fmaf(v[j + 3], LOG2E, bj.w) + kr);
            pk[j >> 1] = pack_bf16(g0, g1);
            pk[(j >> 1) + 1] = pack_bf16(g2, g3);
          }
        } else {
          const float ka2 = k_all * LOG2E;
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float gg[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              float gv = ex2_fast(fmaf(v[j + e], LOG2E, L.bias_l2[c0 + j + e]) + ka2);
              if (p.clamp > 0.f) gv = fminf(gv, p.clamp);
              gg[e] = valid ? gv * scale : 0.f;
            }
            pk[j >> 1] = pack_bf16(gg[0], gg[1]);
          }
        }
        // G tile: k-block c0/64, row r, 16-byte chunks (c0%64)/8 .. +3, 128B swizzle
        const uint32_t gb = L.g_kblock(c0 >> 6) + r * 128;
        const int ch0 = (c0 & 63) >> 3;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          sts128(gb + (((ch0 + i) ^ (r & 7)) << 4), pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
        // g^T spill: gt[col][row0 + r] (lanes = consecutive rows -> 64 B per column)
        // g^T spill (tiled): column v of this row -> gt[tile][r>>6][v][chunk ((r&63)>>3) ^ (v&7)][r&7]
        unsigned short* gt = reinterpret_cast<unsigned short*>(p.gt) + ((rowtile * 2 + (r >> 6)) * p.Vp + c0) * 64 + (r & 7);
        const int gch = (r & 63) >> 3;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          gt[(2 * j) * 64 + ((gch ^ ((2 * j) & 7)) << 3)] = (unsigned short)(pk[j] & 0xffffu);
          gt[(2 * j + 1) * 64 + ((gch ^ ((2 * j + 1) & 7)) << 3)] = (unsigned short)(pk[j] >> 16);
        }
      }
      named_barrier_sync(2, WORKERS);            // every generic entry of G / g^T is written
      if (wg == 0) {
        // exact (fp32, single rounding) blank and label entries of row r
        const float xb = tmem_ld1(tq + p.blank);
        float xl = 0.f;
        for (int i = 0; i < 16; ++i) {
          const int ui = g.ubase + i;
          int col = 0;
          if (ui < g.Ub) col = p.targets[(size_t)g.b * (p.U1 - 1) + ui];
          const float xi = tmem_ld1(tq + col);
          if ((r & 15) == i) xl = xi;
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
            reinterpret_cast<unsigned short*>(p.gt)[((rowtile * 2 + (r >> 6)) * p.Vp + col) * 64 +
                                                    ((((r & 63) >> 3) ^ (col & 7)) << 3) + (r & 7)] = h;
          };
          const float xbb = xb + __ldg(p.bias + p.blank);
          put(p.blank, entry(xbb, k_blank, (lab == p.blank) ? k_label : kNegInf));
          if (lab >= 0 && lab != p.blank) put(lab, entry(xl + __ldg(p.bias + lab), k_label, kNegInf));
        }
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(L.g_full());
      if (tid == 128) TC_PROF(2, 3);

      // ---------------- P3 (MMA busy): d_bias = column sums of the final G tile
      mbar_wait(L.g_full(), ph, 31);
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
        if (wt == 0) mbar_arrive(L.gs_done());
        if (wt < p.Vp) db0 += (L.dbp[wt] + L.dbp[p.Vp + wt]) + (L.dbp[2 * p.Vp + wt] + L.dbp[3 * p.Vp + wt]);
        if (wt + WORKERS < p.Vp)
          db1 += (L.dbp[wt + WORKERS] + L.dbp[p.Vp + wt + WORKERS]) + (L.dbp[2 * p.Vp + wt + WORKERS] + L.dbp[3 * p.Vp + wt + WORKERS]);
      }

      // ---------------- P4 (warps 4-11): dH = dZ * (1 - z^2); reductions.  This CTA owns d blocks mb = 2 pb + rank;
      // warp group wg = pair block pb.  TMEM columns pb*256 + x*128 + tloc*16 + ul: x = 0 the leader's tile, x = 1 the
      // peer's.  z^T boxes (pb, x, half) = [128 d][64 rows] arrive in shared memory (bulk copies after dz_full).
      if (wg < 2) {
        mbar_wait(L.dz_full(), ph, 32);
        if (tid == 128) TC_PROF(2, 4);
        tc_fence_after();
        const int pb = wg;
        if (pb < MB / 2) {
          mbar_wait(L.z_full(pb), ph, 33);
          if (tid == 128) TC_PROF(2, 40 + pb);
          const int d = (2 * pb + (int)rank) * 128 + r;
          const int t0x[2] = {(rank == 0 ? ti.z : tz_peer) * 8, (rank == 0 ? tz_peer : ti.z) * 8};
          float v[16];
          tmem_ld16(tq + pb * 256, v);
#pragma unroll
          for (int it = 0; it < 16; ++it) {
            const int x = it >> 3, tloc = it & 7;
            const uint32_t zb = L.z_box((pb * 2 + x) * 2 + (tloc >> 2)) + (uint32_t)r * 128u;
            const uint4 z0 = lds128(zb + ((((tloc & 3) * 2) ^ (r & 7)) << 4));
            const uint4 z1 = lds128(zb + ((((tloc & 3) * 2 + 1) ^ (r & 7)) << 4));
            tmem_ld_wait();
            float w[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) w[j] = v[j];
            if (it < 15) tmem_ld16(tq + pb * 256 + (it + 1) * 16, v);
            const uint32_t zw[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
            float es0 = 0.f, es1 = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float za = __uint_as_float(zw[j] << 16), zb2 = __uint_as_float(zw[j] & 0xffff0000u);
              const float ha = w[2 * j] * fmaf(-za, za, 1.f), hb = w[2 * j + 1] * fmaf(-zb2, zb2, 1.f);
              es0 += ha;
              es1 += hb;
              pacc[2 * j] += ha;
              pacc[2 * j + 1] += hb;
            }
            const int tt = t0x[x] + tloc;
            if (tt < g.Tb) p.d_enc_part[(((size_t)ti.y * p.B + g.b) * p.T + tt) * p.D + d] = es0 + es1;
          }
        }
        tc_fence_before();
        mbar_arrive(L.tmem_empty());
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
  cluster_sync_all();                       // no CTA exits while the pair may still touch its barriers / TMEM
  if (warp == 2) tmem_dealloc2(tmem_base, TMEM_COLS);
}

}  // namespace tc
}  // namespace ctcvr
