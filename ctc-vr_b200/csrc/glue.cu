// Glue around the fused loss and the evaluation loop.
//
//  rnnt_prologue : add_blank (model/component/transducer.py:8-19), the ignore_id -> 0 remap and the int64 -> int32 casts
//                  of transducer.py:168,174-178 in ONE launch (the reference spends ~8 tiny kernels on them per step):
//                    ys_in [B,U+1] int64 = [blank, text]          (predictor input; padding is NOT remapped there)
//                    targets [B,U] int32 = text with ignore_id -> 0
//                    t_len / u_len [B] int32 from the int64 encoder lengths / the int32 or int64 text lengths
//  loss_combine  : loss = transducer_weight * mean_b(costs) + ctc_weight * loss_ctc (transducer.py:122-128) and the two
//                  gradient seeds, one launch
//  cer_batch     : calculate_cer (rnnt_eval.py:11-56) for N (hypothesis, reference) pairs at once: Levenshtein table
//                  filled along anti-diagonals by one warp per pair, then the reference's backtrace with its
//                  tie-breaking order (match, substitution, deletion, insertion) -> S, D, I, N.  Integer work: bit-exact.
#include "common.cuh"

namespace ctcvr {

__global__ void rnnt_prologue_kernel(const int64_t* __restrict__ text, const void* __restrict__ text_lens, int lens64,
                                     const int64_t* __restrict__ enc_lens, int B, int U, int blank, int ignore_id,
                                     int64_t* __restrict__ ys_in, int32_t* __restrict__ targets, int32_t* __restrict__ t_len,
                                     int32_t* __restrict__ u_len) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (long)B * (U + 1)) {
    const int b = (int)(i / (U + 1)), u = (int)(i - (long)b * (U + 1));
    if (u == 0) {
      ys_in[i] = blank;
    } else {
      const int64_t v = text[(long)b * U + u - 1];
      ys_in[i] = v;
      targets[(long)b * U + u - 1] = (v == (int64_t)ignore_id) ? 0 : (int32_t)v;
    }
  }
  if (i < B) {
    t_len[i] = (int32_t)enc_lens[i];
    u_len[i] = lens64 ? (int32_t)reinterpret_cast<const int64_t*>(text_lens)[i]
                      : reinterpret_cast<const int32_t*>(text_lens)[i];
  }
}

int rnnt_prologue(const int64_t* text, const void* text_lens, int lens64, const int64_t* enc_lens, int B, int U, int blank,
                  int ignore_id, int64_t* ys_in, int32_t* targets, int32_t* t_len, int32_t* u_len, cudaStream_t st) {
  const long n = (long)B * (U + 1);
  rnnt_prologue_kernel<<<cdiv(n > B ? n : B, 256), 256, 0, st>>>(text, text_lens, lens64, enc_lens, B, U, blank, ignore_id,
                                                                  ys_in, targets, t_len, u_len);
  CTCVR_LAUNCH_CHECK();
  return 0;
}

// loss = tw * sum_b costs[b] / B + cw * loss_ctc ; seeds: d loss / d costs[b] = tw / B, d loss / d loss_ctc = cw
__global__ void loss_combine_kernel(const float* __restrict__ costs, int B, const float* __restrict__ loss_ctc, float tw,
                                    float cw, float* __restrict__ out) {
  float s = 0.f;
  for (int i = threadIdx.x; i < B; i += 32) s += costs[i];
  s = warp_sum(s);
  if (threadIdx.x == 0) {
    const float lr = s / (float)B;
    out[0] = tw * lr + (loss_ctc ? cw * loss_ctc[0] : 0.f);
    out[1] = lr;
  }
}

int loss_combine(const float* costs, int B, const float* loss_ctc, float tw, float cw, float* out, cudaStream_t st) {
  loss_combine_kernel<<<1, 32, 0, st>>>(costs, B, loss_ctc, tw, cw, out);
  CTCVR_LAUNCH_CHECK();
  return 0;
}

// One warp per pair.  dp is [(m+1) x (n+1)] uint16 in the workspace (pitch = Ln + 1): dp[i][j] = edit distance between
// the first i hypothesis tokens and the first j reference tokens.  Anti-diagonal d = i + j: cells of a diagonal depend
// on diagonals d-1 and d-2 only.
__global__ void __launch_bounds__(128) cer_kernel(const int32_t* __restrict__ hyp, const int32_t* __restrict__ hyp_len, int Lh,
                                                  const int32_t* __restrict__ ref, const int32_t* __restrict__ ref_len, int Lr,
                                                  int N, unsigned short* __restrict__ ws, int32_t* __restrict__ out) {
  const int pair = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (pair >= N) return;
  const int m = max(min(hyp_len[pair], Lh), 0), n = max(min(ref_len[pair], Lr), 0);
  const int32_t* h = hyp + (size_t)pair * Lh;
  const int32_t* r = ref + (size_t)pair * Lr;
  const int pitch = Lr + 1;
  unsigned short* dp = ws + (size_t)pair * (size_t)(Lh + 1) * pitch;
  for (int i = lane; i <= m; i += 32) dp[(size_t)i * pitch] = (unsigned short)i;
  for (int j = lane; j <= n; j += 32) dp[j] = (unsigned short)j;
  __syncwarp();
  for (int d = 2; d <= m + n; ++d) {
    const int i_lo = max(1, d - n), i_hi = min(m, d - 1);
    for (int i = i_lo + lane; i <= i_hi; i += 32) {
      const int j = d - i;
      const int cost = (h[i - 1] == r[j - 1]) ? 0 : 1;
      const int a = dp[(size_t)(i - 1) * pitch + j] + 1, b = dp[(size_t)i * pitch + j - 1] + 1,
                c = dp[(size_t)(i - 1) * pitch + j - 1] + cost;
      dp[(size_t)i * pitch + j] = (unsigned short)min(a, min(b, c));
    }
    __syncwarp();
  }
  if (lane == 0) {
    int i = m, j = n, S = 0, D = 0, I = 0;
    while (i > 0 && j > 0) {
      if (h[i - 1] == r[j - 1]) { --i; --j; }
      else {
        const int cur = dp[(size_t)i * pitch + j];
        if (cur == dp[(size_t)(i - 1) * pitch + j - 1] + 1) { ++S; --i; --j; }
        else if (cur == dp[(size_t)(i - 1) * pitch + j] + 1) { ++D; --i; }
        else { ++I; --j; }
      }
    }
    D += i;
    I += j;
    out[pair * 4 + 0] = S; out[pair * 4 + 1] = D; out[pair * 4 + 2] = I; out[pair * 4 + 3] = n;
  }
}

size_t cer_ws_bytes(int N, int Lh, int Lr) { return (size_t)N * (size_t)(Lh + 1) * (size_t)(Lr + 1) * sizeof(unsigned short); }

int cer_batch(const int32_t* hyp, const int32_t* hyp_len, int Lh, const int32_t* ref, const int32_t* ref_len, int Lr, int N,
              void* ws, size_t ws_bytes, int32_t* out, cudaStream_t st) {
  CTCVR_REQUIRE(Lh + Lr < 65535, "cer_batch: sequences too long for the 16-bit table (%d + %d)", Lh, Lr);
  CTCVR_REQUIRE(ws_bytes >= cer_ws_bytes(N, Lh, Lr), "cer_batch: workspace too small");
  cer_kernel<<<cdiv(N, 4), 128, 0, st>>>(hyp, hyp_len, Lh, ref, ref_len, Lr, N, reinterpret_cast<unsigned short*>(ws), out);
  CTCVR_LAUNCH_CHECK();
  return 0;
}

}  // namespace ctcvr
