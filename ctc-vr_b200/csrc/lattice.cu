// RNN-T lattice: alpha / beta anti-diagonal wavefront and costs (the recursion
// torchaudio.functional.rnnt_loss runs inside rnnt_loss_forward, called from
// model/component/transducer.py:180-187).
//
// One CTA per utterance.  Warp 0 walks alpha forward, warp 1 walks beta backward, concurrently.
// Lane l owns lattice columns u = l, l+32, ... (NJ per lane); one anti-diagonal t+u=d per step.
// The neighbour needed from column u-1 (alpha) / u+1 (beta) lives in the adjacent lane one
// diagonal earlier, so the exchange is a single warp shuffle and no barrier is needed.
// lp_blank / lp_label for the utterance are staged once into shared memory with coalesced loads
// (pitch chosen so that a diagonal read is bank-conflict free); utterances whose log-probs do not
// fit in shared memory are read through L2.
//
// Algorithmic HBM bytes: 24 B per cell (2 log-probs read for alpha, again for beta, alpha and beta
// written).  The kernel is bound by the T+U dependent diagonals, not by bandwidth.
#include "common.cuh"

namespace ctcvr {

constexpr int LAT_THREADS = 128;

// log(exp(a)+exp(b)) on the MUFU fast path (ex2/lg2.approx): the wavefront is a chain of T+U dependent
// steps, so the latency of this function IS the kernel time.  Absolute error ~1e-7 per step.
__device__ __forceinline__ float lae_fast(float a, float b) {
  const float m = fmaxf(a, b);
  const float r = m + __logf(1.f + __expf(-fabsf(a - b)));
  return (m == kNegInf) ? kNegInf : r;
}

template <int NJ>
__global__ void __launch_bounds__(LAT_THREADS) rnnt_lattice_kernel(
    const float* __restrict__ lp_blank, const float* __restrict__ lp_label, const int32_t* __restrict__ t_len,
    const int32_t* __restrict__ u_len, float* __restrict__ alpha, float* __restrict__ beta,
    float* __restrict__ costs, int T, int U1, int pitch, int use_smem) {
  extern __shared__ float sm[];
  const int b = blockIdx.x;
  const int Tb = t_len[b], Ub = u_len[b];
  const size_t base = (size_t)b * T * U1;
  const float* gb = lp_blank + base;
  const float* gl = lp_label + base;
  const float* sb = gb;
  const float* sl = gl;
  int sp = U1;
  if (Tb <= 0) { if (threadIdx.x == 0) costs[b] = 0.f; return; }
  if (use_smem) {
    float* s0 = sm;
    float* s1 = sm + (size_t)T * pitch;
    const int W = Ub + 1;
    for (int i = threadIdx.x; i < Tb * W; i += LAT_THREADS) {
      int t = i / W, u = i - t * W;
      s0[t * pitch + u] = gb[t * U1 + u];
      s1[t * pitch + u] = gl[t * U1 + u];
    }
    sb = s0; sl = s1; sp = pitch;
    __syncthreads();
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp > 1) return;
  const int ndiag = Tb + Ub;           // diagonals 0 .. Tb+Ub-1
  float prev[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) prev[j] = kNegInf;

  if (warp == 0) {
    // ---------------- alpha: alpha(t,u) = LSE(alpha(t-1,u)+lpb(t-1,u), alpha(t,u-1)+lpl(t,u-1))
    float* ab = alpha + base;
    // log-probs needed by the NEXT diagonal are fetched from smem before the dependent chain of this one
    float nb[NJ], nl[NJ];
    auto fetch_a = [&](int d) {
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int u = lane + 32 * j, t = d - u;
        const bool ok = (u <= Ub && t >= 0 && t < Tb);
        nb[j] = (ok && t > 0) ? sb[(t - 1) * sp + u] : kNegInf;
        nl[j] = (ok && u > 0) ? sl[t * sp + u - 1] : kNegInf;
      }
    };
    fetch_a(0);
    for (int d = 0; d < ndiag; ++d) {
      float cb[NJ], cl[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) { cb[j] = nb[j]; cl[j] = nl[j]; }
      if (d + 1 < ndiag) fetch_a(d + 1);
      float cur[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int u = lane + 32 * j;
        const int t = d - u;
        float left = __shfl_up_sync(0xffffffffu, prev[j], 1);
        float wrap = __shfl_sync(0xffffffffu, prev[j > 0 ? j - 1 : 0], 31);
        if (lane == 0) left = (j > 0) ? wrap : kNegInf;
        float v = kNegInf;
        if (u <= Ub && t >= 0 && t < Tb) {
          v = (d == 0) ? 0.f : lae_fast(prev[j] + cb[j], left + cl[j]);
          ab[t * U1 + u] = v;
        }
        cur[j] = v;
      }
#pragma unroll
      for (int j = 0; j < NJ; ++j) prev[j] = cur[j];
    }
  } else {
    // ---------------- beta: beta(t,u) = LSE(beta(t+1,u)+lpb(t,u), beta(t,u+1)+lpl(t,u))
    float* bb = beta + base;
    float nb[NJ], nl[NJ];
    auto fetch_b = [&](int d) {
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int u = lane + 32 * j, t = d - u;
        const bool ok = (u <= Ub && t >= 0 && t < Tb);
        nb[j] = ok ? sb[t * sp + u] : kNegInf;
        nl[j] = (ok && u < Ub) ? sl[t * sp + u] : kNegInf;
      }
    };
    fetch_b(ndiag - 1);
    for (int d = ndiag - 1; d >= 0; --d) {
      float cb[NJ], cl[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) { cb[j] = nb[j]; cl[j] = nl[j]; }
      if (d > 0) fetch_b(d - 1);
      float cur[NJ];
#pragma unroll
      for (int j = NJ - 1; j >= 0; --j) {
        const int u = lane + 32 * j;
        const int t = d - u;
        float right = __shfl_down_sync(0xffffffffu, prev[j], 1);
        float wrap = __shfl_sync(0xffffffffu, prev[(j + 1 < NJ) ? j + 1 : j], 0);
        if (lane == 31) right = (j + 1 < NJ) ? wrap : kNegInf;
        float v = kNegInf;
        if (u <= Ub && t >= 0 && t < Tb) {
          if (t == Tb - 1 && u == Ub) v = cb[j];
          else {
            const float a = (t + 1 < Tb) ? prev[j] + cb[j] : kNegInf;
            v = lae_fast(a, right + cl[j]);
          }
          bb[t * U1 + u] = v;
        }
        cur[j] = v;
      }
#pragma unroll
      for (int j = 0; j < NJ; ++j) prev[j] = cur[j];
    }
    if (lane == 0) costs[b] = -prev[0];     // beta(0,0)
  }
}

// Fallback for very long targets (U1 > 256): one CTA per utterance, block-wide diagonal sweep.
__global__ void rnnt_lattice_generic_kernel(const float* __restrict__ lp_blank, const float* __restrict__ lp_label,
                                            const int32_t* __restrict__ t_len, const int32_t* __restrict__ u_len,
                                            float* __restrict__ alpha, float* __restrict__ beta,
                                            float* __restrict__ costs, int T, int U1) {
  const int b = blockIdx.x;
  const int Tb = t_len[b], Ub = u_len[b];
  const size_t base = (size_t)b * T * U1;
  const float* lb = lp_blank + base;
  const float* ll = lp_label + base;
  float* ab = alpha + base;
  float* bb = beta + base;
  if (Tb <= 0) { if (threadIdx.x == 0) costs[b] = 0.f; return; }
  const int ndiag = Tb + Ub;
  for (int d = 0; d < ndiag; ++d) {
    for (int u = threadIdx.x; u <= Ub; u += blockDim.x) {
      int t = d - u;
      if (t < 0 || t >= Tb) continue;
      float v;
      if (d == 0) v = 0.f;
      else {
        float a = (t > 0) ? ab[(t - 1) * U1 + u] + lb[(t - 1) * U1 + u] : kNegInf;
        float c = (u > 0) ? ab[t * U1 + u - 1] + ll[t * U1 + u - 1] : kNegInf;
        v = log_add_exp(a, c);
      }
      ab[t * U1 + u] = v;
    }
    __syncthreads();
  }
  for (int d = ndiag - 1; d >= 0; --d) {
    for (int u = threadIdx.x; u <= Ub; u += blockDim.x) {
      int t = d - u;
      if (t < 0 || t >= Tb) continue;
      float v;
      if (t == Tb - 1 && u == Ub) v = lb[t * U1 + u];
      else {
        float a = (t + 1 < Tb) ? bb[(t + 1) * U1 + u] + lb[t * U1 + u] : kNegInf;
        float c = (u < Ub) ? bb[t * U1 + u + 1] + ll[t * U1 + u] : kNegInf;
        v = log_add_exp(a, c);
      }
      bb[t * U1 + u] = v;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) costs[b] = -bb[0];
}

template <int NJ>
static int launch_lattice(const float* lpb, const float* lpl, const int32_t* t_len, const int32_t* u_len,
                          float* alpha, float* beta, float* costs, int B, int T, int U1, cudaStream_t st) {
  int pitch = (U1 % 2 == 0) ? U1 : U1 + 1;       // pitch-1 odd => diagonal reads hit distinct banks
  size_t smem = (size_t)2 * T * pitch * sizeof(float);
  int use_smem = smem <= 200 * 1024;
  if (!use_smem) smem = 0;
  CTCVR_CHECK_CUDA(cudaFuncSetAttribute(rnnt_lattice_kernel<NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(200 * 1024)));
  rnnt_lattice_kernel<NJ><<<B, LAT_THREADS, smem, st>>>(lpb, lpl, t_len, u_len, alpha, beta, costs, T, U1, pitch,
                                                         use_smem);
  CTCVR_LAUNCH_CHECK();
  return 0;
}

int rnnt_lattice(const float* lpb, const float* lpl, const int32_t* t_len, const int32_t* u_len, float* alpha,
                 float* beta, float* costs, int B, int T, int U1, cudaStream_t st) {
  if (U1 <= 32) return launch_lattice<1>(lpb, lpl, t_len, u_len, alpha, beta, costs, B, T, U1, st);
  if (U1 <= 64) return launch_lattice<2>(lpb, lpl, t_len, u_len, alpha, beta, costs, B, T, U1, st);
  if (U1 <= 128) return launch_lattice<4>(lpb, lpl, t_len, u_len, alpha, beta, costs, B, T, U1, st);
  if (U1 <= 256) return launch_lattice<8>(lpb, lpl, t_len, u_len, alpha, beta, costs, B, T, U1, st);
  rnnt_lattice_generic_kernel<<<B, 256, 0, st>>>(lpb, lpl, t_len, u_len, alpha, beta, costs, T, U1);
  CTCVR_LAUNCH_CHECK();
  return 0;
}

}  // namespace ctcvr
