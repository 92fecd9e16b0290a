// Device building blocks shared by the on-device RNN-T decoders (greedy / online beam / prefix beam):
// one predictor step (embedding + LSTM cell(s) + projection, model/component/predictor.py:79-98 with
// padding == 0) and one joint step (model/component/joint.py:48-69 with T=U=1) for NB hypotheses that
// advance in lock-step inside one CTA.  Weights are streamed from L2 in the transposed layouts of
// ctcvr_decoder_weights so that consecutive threads read consecutive addresses; per-hypothesis
// vectors live in shared memory as [k][NB] so that one 16/32-byte broadcast load feeds NB FMAs.
#pragma once
#include "common.cuh"

namespace ctcvr {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// two IEEE fp32 FMAs in one instruction (FFMA2, sm_100+): {d.lo, d.hi} = {a.lo * b.lo + c.lo, a.hi * b.hi + c.hi} -
// bit-identical to two fmaf() calls, half the issue slots
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
  return d;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

// out[j][n] = init(j,n) + sum_k Wt[k*J + j] * xs[k*NB + n]   for j in [0,J)
// Per output row the sum runs over k in ascending order with one fused multiply-add per term - the same arithmetic for
// every NB and for the packed form, so hypotheses do not depend on how many advance in lock-step.
// 16 weight loads in flight per thread: with 4, a 512-thread CTA kept ~8 KB in flight against an L2 latency of ~600
// cycles (~25 GB/s per SM) and the weight stream, not the FMAs, set the step time.  The loop is then issue-bound
// (NB FMAs + NB/4 shared loads per weight and thread): the packed FFMA2 halves the FMA issue slots.
template <int NB, class Init, class Store>
__device__ __forceinline__ void gemv_t(const float* __restrict__ Wt, int J, int K, const float* xs, Init init,
                                       Store store) {
  if constexpr (NB % 2 == 0) {
    constexpr int NP = NB / 2;
    for (int j = threadIdx.x; j < J; j += blockDim.x) {
      unsigned long long acc[NP];
#pragma unroll
      for (int m = 0; m < NP; ++m) acc[m] = pack2(init(j, 2 * m), init(j, 2 * m + 1));
      int k = 0;
      for (; k + 16 <= K; k += 16) {
        float w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = __ldg(Wt + (size_t)(k + i) * J + j);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const unsigned long long wp = pack2(w[i], w[i]);
          const unsigned long long* xr = reinterpret_cast<const unsigned long long*>(xs + (size_t)(k + i) * NB);
#pragma unroll
          for (int m = 0; m < NP; ++m) acc[m] = fma2(wp, xr[m], acc[m]);
        }
      }
      for (; k < K; ++k) {
        const float w = __ldg(Wt + (size_t)k * J + j);
        const unsigned long long wp = pack2(w, w);
        const unsigned long long* xr = reinterpret_cast<const unsigned long long*>(xs + (size_t)k * NB);
#pragma unroll
        for (int m = 0; m < NP; ++m) acc[m] = fma2(wp, xr[m], acc[m]);
      }
#pragma unroll
      for (int m = 0; m < NP; ++m) {
        float lo, hi;
        unpack2(acc[m], lo, hi);
        store(j, 2 * m, lo);
        store(j, 2 * m + 1, hi);
      }
    }
  } else {
    for (int j = threadIdx.x; j < J; j += blockDim.x) {
      float acc[NB];
#pragma unroll
      for (int n = 0; n < NB; ++n) acc[n] = init(j, n);
      int k = 0;
      for (; k + 16 <= K; k += 16) {
        float w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = __ldg(Wt + (size_t)(k + i) * J + j);
#pragma unroll
        for (int i = 0; i < 16; ++i)
#pragma unroll
          for (int n = 0; n < NB; ++n) acc[n] = fmaf(w[i], xs[(k + i) * NB + n], acc[n]);
      }
      for (; k < K; ++k) {
        const float w = __ldg(Wt + (size_t)k * J + j);
#pragma unroll
        for (int n = 0; n < NB; ++n) acc[n] = fmaf(w, xs[k * NB + n], acc[n]);
      }
#pragma unroll
      for (int n = 0; n < NB; ++n) store(j, n, acc[n]);
    }
  }
}

// Shared-memory working set of one lock-step group.
template <int NB>
struct DecodeSmem {
  float* hs;     // [L][H][NB] committed h
  float* cs;     // [L][H][NB] committed c
  float* hn;     // [L][H][NB] state after feeding `tok` (pending)
  float* cn;     // [L][H][NB]
  float* gates;  // [4H][NB]
  float* pout;   // [P][NB]
  float* pproj;  // [D][NB]  pred_ffn(projection(h'))
  float* z;      // [D][NB]
  float* logit;  // [V][NB]  (only used by the beam decoders)
  __host__ __device__ static size_t floats(const ctcvr_decoder_weights& w, bool with_logits) {
    return (size_t)NB * (4 * w.L * w.H + 4 * w.H + w.P + 2 * w.D + (with_logits ? w.V : 0));
  }
  __device__ void carve(float* base, const ctcvr_decoder_weights& w, bool with_logits) {
    size_t lh = (size_t)w.L * w.H * NB;
    hs = base; cs = hs + lh; hn = cs + lh; cn = hn + lh;
    gates = cn + lh; pout = gates + (size_t)4 * w.H * NB; pproj = pout + (size_t)w.P * NB;
    z = pproj + (size_t)w.D * NB; logit = z + (size_t)w.D * NB;
    (void)with_logits;
  }
};

// Predictor step for all NB slots: reads committed (hs,cs) and tok[], writes pending (hn,cn) and pproj.
// Must be called by the whole CTA; ends with a __syncthreads().
template <int NB>
__device__ void predictor_step(const ctcvr_decoder_weights& w, DecodeSmem<NB>& s, const int* tok) {
  const int H = w.H, G = 4 * w.H;
  for (int l = 0; l < w.L; ++l) {
    const float* hprev = s.hs + (size_t)l * H * NB;
    if (l == 0) {
      gemv_t<NB>(w.w_hh_t, G, H, hprev,
                 [&](int j, int n) { return __ldg(w.gate_tok + (size_t)tok[n] * G + j); },
                 [&](int j, int n, float v) { s.gates[j * NB + n] = v; });
    } else {
      gemv_t<NB>(w.w_ih_t + (size_t)(l - 1) * H * G, G, H, s.hn + (size_t)(l - 1) * H * NB,
                 [&](int j, int n) { return __ldg(w.b_gate + (size_t)(l - 1) * G + j); },
                 [&](int j, int n, float v) { s.gates[j * NB + n] = v; });
      __syncthreads();
      gemv_t<NB>(w.w_hh_t + (size_t)l * H * G, G, H, hprev,
                 [&](int j, int n) { return s.gates[j * NB + n]; },
                 [&](int j, int n, float v) { s.gates[j * NB + n] = v; });
    }
    __syncthreads();
    for (int i = threadIdx.x; i < H * NB; i += blockDim.x) {
      int k = i / NB, n = i - k * NB;
      float gi = sigmoidf_(s.gates[(k)*NB + n]);
      float gf = sigmoidf_(s.gates[(H + k) * NB + n]);
      float gg = tanhf(s.gates[(2 * H + k) * NB + n]);
      float go = sigmoidf_(s.gates[(3 * H + k) * NB + n]);
      float c = gf * s.cs[((size_t)l * H + k) * NB + n] + gi * gg;
      s.cn[((size_t)l * H + k) * NB + n] = c;
      s.hn[((size_t)l * H + k) * NB + n] = go * tanhf(c);
    }
    __syncthreads();
  }
  gemv_t<NB>(w.proj_t, w.P, H, s.hn + (size_t)(w.L - 1) * H * NB,
             [&](int j, int n) { return __ldg(w.proj_b + j); },
             [&](int j, int n, float v) { s.pout[j * NB + n] = v; });
  __syncthreads();
  gemv_t<NB>(w.pred_ffn_t, w.D, w.P, s.pout,
             [&](int j, int n) { return __ldg(w.pred_ffn_b + j); },
             [&](int j, int n, float v) { s.pproj[j * NB + n] = v; });
  __syncthreads();
}

// (value, index) max with lowest-index tie-break
__device__ __forceinline__ void argmax_combine(float& v, int& i, float ov, int oi) {
  if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}

}  // namespace ctcvr
