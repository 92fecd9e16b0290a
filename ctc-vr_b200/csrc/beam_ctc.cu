// CTC prefix beam search on the device: wenet/transformer/search.py:125-247 with
// context_graph=None, PrefixScore (:62-104), log_add (wenet/utils/common.py:302-310).
//
// One CTA per utterance; the beam (<= BEAM_MAX prefixes) and the per-frame candidate table live in shared
// memory, token / time lists in a global workspace (ping-pong per frame).  Per frame:
//   1. top-k (k = beam) of the frame's log-probs: k rounds of a block-wide argmax (lowest index wins ties)
//   2. candidate slots: slot p < nb is "prefix p unchanged", slot nb + q*k + i is "prefix q + token i";
//      an extension that spells an existing prefix (q+u == p) aliases slot p, exactly like the reference's
//      dict keyed by the token tuple (detected with a rolling hash, verified token by token)
//   3. one thread per slot replays the reference's double loop (for u in topk: for prefix in beam) in order
//      and applies the updates that land on its slot - fp64 log_add in the reference's order, Viterbi score /
//      token-time bookkeeping included; the first-touch index reproduces the dict insertion order
//   4. rank by (score desc, insertion order asc) = Python's stable sort, keep `beam`, materialise the token and
//      time lists of the survivors (copy parent list, append / replace last).
#include "common.cuh"

namespace ctcvr {

constexpr int BEAM_MAX = 16;
constexpr int CAND_MAX = BEAM_MAX * (BEAM_MAX + 1);
constexpr int PB_THREADS = 288;                     // >= CAND_MAX

struct PbEntry {
  double s, ns, v_s, v_ns;
  unsigned long long hash;
  int len, last;
};
struct PbCand {
  double s, ns, v_s, v_ns, score;
  float cur_token_prob;
  int first_touch;            // INT_MAX = not a key of the reference's dict
  int parent, ext;            // token list = tokens(parent) [+ ext]
  int ts_src, ts_which;       // times_s  = copy of entry ts_src's list (0 = s, 1 = ns); -1 = []
  int tn_src, tn_which, tn_op, tn_t;   // times_ns = copy, then op: 0 none, 1 append t, 2 replace last with t
};

__device__ __forceinline__ double pb_log_add(double a, double b) {
  if (a == -INFINITY && b == -INFINITY) return -INFINITY;
  const double m = fmax(a, b);
  return m + log(exp(a - m) + exp(b - m));
}
__device__ __forceinline__ unsigned long long pb_hash_ext(unsigned long long h, int u) {
  return h * 1099511628211ULL + (unsigned long long)(u + 1);
}

__global__ void __launch_bounds__(PB_THREADS) ctc_prefix_beam_kernel(
    const float* __restrict__ probs, const int32_t* __restrict__ lens, int T, int V, int beam, int blank,
    int32_t* __restrict__ out_n, int32_t* __restrict__ out_tokens, int32_t* __restrict__ out_lens,
    double* __restrict__ out_scores, int32_t* __restrict__ out_times, int32_t* __restrict__ ws) {
  extern __shared__ __align__(16) unsigned char pb_smem[];
  float* row = reinterpret_cast<float*>(pb_smem);                               // [V]
  PbEntry* cur = reinterpret_cast<PbEntry*>(pb_smem + (((size_t)V * 4 + 15) / 16) * 16);   // [BEAM_MAX]
  PbCand* cand = reinterpret_cast<PbCand*>(cur + BEAM_MAX);                     // [CAND_MAX]
  __shared__ float red_v[PB_THREADS / 32];
  __shared__ int red_i[PB_THREADS / 32];
  __shared__ int top_tok[BEAM_MAX];
  __shared__ float top_p[BEAM_MAX];
  __shared__ int alias[BEAM_MAX * BEAM_MAX];      // extension (q, i) -> existing prefix p or -1
  __shared__ int sel[BEAM_MAX];                   // slots that survive the frame, best first
  __shared__ int s_nb, s_nsel;
  __shared__ double cur_score[BEAM_MAX];         // log_add(s, ns) of the beam entries, once per frame

  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Tb = min(lens[b], T);
  // workspace of this utterance: [2 buffers][3 lists: tokens, times_s, times_ns][BEAM_MAX][T], + lengths
  int32_t* wsb = ws + (size_t)b * (2 * 3 * BEAM_MAX * (size_t)T + 2 * 3 * BEAM_MAX);
  auto list = [&](int buf, int which, int e) { return wsb + (((size_t)buf * 3 + which) * BEAM_MAX + e) * T; };
  int32_t* llen = wsb + 2 * 3 * BEAM_MAX * (size_t)T;           // [2][3][BEAM_MAX]
  auto len_of = [&](int buf, int which, int e) -> int32_t& { return llen[(buf * 3 + which) * BEAM_MAX + e]; };

  if (tid == 0) {
    cur[0].s = 0.0; cur[0].ns = -INFINITY; cur[0].v_s = 0.0; cur[0].v_ns = 0.0;
    cur[0].hash = 1469598103934665603ULL; cur[0].len = 0; cur[0].last = -1;
    s_nb = 1;
    for (int w3 = 0; w3 < 3; ++w3) len_of(0, w3, 0) = 0;
  }
  __syncthreads();
  int buf = 0;
  const int k = min(beam, V);

  for (int t = 0; t < Tb; ++t) {
    const int nb = s_nb;
    const float* src = probs + ((size_t)b * T + t) * V;
    for (int i = tid; i < V; i += PB_THREADS) row[i] = src[i];
    __syncthreads();
    // ---- 1. top-k
    for (int r = 0; r < k; ++r) {
      float bv = -INFINITY;
      int bi = 0x7fffffff;
      for (int i = tid; i < V; i += PB_THREADS) {
        const float x = row[i];
        if (x > bv || (x == bv && i < bi)) { bv = x; bi = i; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if (lane == 0) { red_v[warp] = bv; red_i[warp] = bi; }
      __syncthreads();
      if (tid == 0) {
        float v = red_v[0];
        int ix = red_i[0];
        for (int w2 = 1; w2 < PB_THREADS / 32; ++w2)
          if (red_v[w2] > v || (red_v[w2] == v && red_i[w2] < ix)) { v = red_v[w2]; ix = red_i[w2]; }
        top_tok[r] = ix;
        top_p[r] = v;
        row[ix] = -INFINITY;            // the row is re-staged next frame
      }
      __syncthreads();
    }
    // ---- 2. aliases of extensions, candidate init
    for (int pr = tid; pr < nb * k; pr += PB_THREADS) {
      const int q = pr / k, i = pr - q * k, u = top_tok[i];
      int a = -1;
      if (u != blank) {
        const unsigned long long h = pb_hash_ext(cur[q].hash, u);
        for (int p2 = 0; p2 < nb && a < 0; ++p2) {
          if (cur[p2].len == cur[q].len + 1 && cur[p2].last == u && cur[p2].hash == h) {
            const int32_t* tp = list(buf, 0, p2);
            const int32_t* tq = list(buf, 0, q);
            bool same = true;
            for (int j = 0; j < cur[q].len && same; ++j) same = (tp[j] == tq[j]);
            if (same) a = p2;
          }
        }
      }
      alias[q * BEAM_MAX + i] = a;
    }
    if (tid < nb) cur_score[tid] = pb_log_add(cur[tid].s, cur[tid].ns);
    const int nslot = nb + nb * k;
    for (int c = tid; c < nslot; c += PB_THREADS) {
      PbCand& n = cand[c];
      n.s = n.ns = n.v_s = n.v_ns = -INFINITY;
      n.cur_token_prob = -INFINITY;
      n.first_touch = 0x7fffffff;
      n.ts_src = -1; n.ts_which = 0;
      n.tn_src = -1; n.tn_which = 0; n.tn_op = 0; n.tn_t = 0;
      if (c < nb) { n.parent = c; n.ext = -1; }
      else { const int q = (c - nb) / k, i = (c - nb) - q * k; n.parent = q; n.ext = top_tok[i]; }
    }
    __syncthreads();
    // ---- 3. one thread per slot replays the reference's loops in order
    if (tid < nslot) {
      PbCand n = cand[tid];
      const int me = tid;
      for (int i = 0; i < k; ++i) {
        const int u = top_tok[i];
        const double prob = (double)top_p[i];
        for (int q = 0; q < nb; ++q) {
          const PbEntry ps = cur[q];
          const int order = 2 * (i * nb + q);
          const double score = cur_score[q];
          const bool s_wins = ps.v_s > ps.v_ns;
          const double vit = s_wins ? ps.v_s : ps.v_ns;
          int ext_slot = -1;
          if (u != blank) { const int a = alias[q * BEAM_MAX + i]; ext_slot = (a >= 0) ? a : nb + q * k + i; }
          if (u == blank) {
            if (me == q) {
              n.s = pb_log_add(n.s, score + prob);
              n.v_s = vit + prob;
              n.ts_src = q; n.ts_which = s_wins ? 0 : 1;
              n.first_touch = min(n.first_touch, order);
            }
          } else if (u == ps.last) {
            if (me == q) {
              n.ns = pb_log_add(n.ns, ps.ns + prob);
              if (n.v_ns < ps.v_ns + prob) {
                n.v_ns = ps.v_ns + prob;
                if (n.cur_token_prob < (float)prob) {
                  n.cur_token_prob = (float)prob;
                  n.tn_src = q; n.tn_which = 1; n.tn_op = 2; n.tn_t = t;
                }
              }
              n.first_touch = min(n.first_touch, order);
            }
            if (me == ext_slot) {
              n.ns = pb_log_add(n.ns, ps.s + prob);
              if (n.v_ns < ps.v_s + prob) {
                n.v_ns = ps.v_s + prob;
                n.cur_token_prob = (float)prob;
                n.tn_src = q; n.tn_which = 0; n.tn_op = 1; n.tn_t = t;
              }
              n.first_touch = min(n.first_touch, order + 1);
            }
          } else {
            if (me == ext_slot) {
              n.ns = pb_log_add(n.ns, score + prob);
              if (n.v_ns < vit + prob) {
                n.v_ns = vit + prob;
                n.cur_token_prob = (float)prob;
                n.tn_src = q; n.tn_which = s_wins ? 0 : 1; n.tn_op = 1; n.tn_t = t;
              }
              n.first_touch = min(n.first_touch, order);
            }
          }
        }
      }
      n.score = pb_log_add(n.s, n.ns);
      cand[tid] = n;
    }
    __syncthreads();
    // ---- 4. rank (score desc, insertion order asc), keep `beam`
    if (tid < nslot) {
      const PbCand& n = cand[tid];
      if (n.first_touch != 0x7fffffff) {
        int rank = 0;
        for (int j = 0; j < nslot; ++j) {
          const PbCand& m = cand[j];
          if (m.first_touch == 0x7fffffff || j == tid) continue;
          if (m.score > n.score || (m.score == n.score && m.first_touch < n.first_touch)) ++rank;
        }
        if (rank < beam) sel[rank] = tid;
      }
    }
    if (tid == 0) {
      int cnt = 0;
      for (int j = 0; j < nslot; ++j) cnt += (cand[j].first_touch != 0x7fffffff);
      s_nsel = min(cnt, beam);
    }
    __syncthreads();
    const int nsel = s_nsel;
    // materialise the survivors into the other buffer
    const int nbuf = buf ^ 1;
    for (int e = 0; e < nsel; ++e) {
      const PbCand& n = cand[sel[e]];
      const int plen = cur[n.parent].len;
      const int32_t* ptok = list(buf, 0, n.parent);
      int32_t* dtok = list(nbuf, 0, e);
      for (int j = tid; j < plen; j += PB_THREADS) dtok[j] = ptok[j];
      if (tid == 0 && n.ext >= 0) dtok[plen] = n.ext;
      // times_s
      int ls = 0;
      if (n.ts_src >= 0) {
        ls = len_of(buf, 1 + n.ts_which, n.ts_src);
        const int32_t* sp = list(buf, 1 + n.ts_which, n.ts_src);
        int32_t* dp = list(nbuf, 1, e);
        for (int j = tid; j < ls; j += PB_THREADS) dp[j] = sp[j];
      }
      // times_ns
      int ln = 0;
      if (n.tn_src >= 0) {
        const int l0 = len_of(buf, 1 + n.tn_which, n.tn_src);
        const int32_t* sp = list(buf, 1 + n.tn_which, n.tn_src);
        int32_t* dp = list(nbuf, 2, e);
        for (int j = tid; j < l0; j += PB_THREADS) dp[j] = sp[j];
        ln = l0;
        if (n.tn_op == 1) { if (tid == 0) dp[l0] = n.tn_t; ln = l0 + 1; }
        // op 2 (replace last) is applied after the copy, below
      }
      __syncthreads();
      if (tid == 0) {
        if (n.tn_src >= 0 && n.tn_op == 2 && ln > 0) list(nbuf, 2, e)[ln - 1] = n.tn_t;
        len_of(nbuf, 0, e) = plen + (n.ext >= 0 ? 1 : 0);
        len_of(nbuf, 1, e) = ls;
        len_of(nbuf, 2, e) = ln;
      }
    }
    __syncthreads();
    buf = nbuf;
    {
      // two-phase publish: compute into registers (parents read from `cur`), barrier, then write
      PbEntry e;
      bool have = false;
      if (tid < nsel) {
        const PbCand& n = cand[sel[tid]];
        const PbEntry par = cur[n.parent];
        e.s = n.s; e.ns = n.ns; e.v_s = n.v_s; e.v_ns = n.v_ns;
        e.len = par.len + (n.ext >= 0 ? 1 : 0);
        e.last = (n.ext >= 0) ? n.ext : par.last;
        e.hash = (n.ext >= 0) ? pb_hash_ext(par.hash, n.ext) : par.hash;
        have = true;
      }
      __syncthreads();
      if (have) cur[tid] = e;
      if (tid == 0) s_nb = nsel;
      __syncthreads();
    }
  }

  // ---- results: beam entries are already sorted by score
  const int nb = s_nb;
  if (tid == 0) out_n[b] = nb;
  for (int e = 0; e < nb; ++e) {
    const PbEntry& ce = cur[e];
    const bool s_wins = ce.v_s > ce.v_ns;
    const int tl = len_of(buf, s_wins ? 1 : 2, e);
    const int32_t* tp = list(buf, 0, e);
    const int32_t* tm = list(buf, s_wins ? 1 : 2, e);
    for (int j = tid; j < ce.len; j += PB_THREADS) out_tokens[((size_t)b * beam + e) * T + j] = tp[j];
    for (int j = tid; j < tl; j += PB_THREADS) out_times[((size_t)b * beam + e) * T + j] = tm[j];
    if (tid == 0) {
      out_lens[(size_t)b * beam + e] = ce.len;
      out_scores[(size_t)b * beam + e] = pb_log_add(ce.s, ce.ns);
    }
  }
}

size_t ctc_prefix_beam_ws_bytes(int B, int T, int V, int beam) {
  (void)V; (void)beam;
  return (size_t)B * (2 * 3 * BEAM_MAX * (size_t)T + 2 * 3 * BEAM_MAX) * sizeof(int32_t);
}

int ctc_prefix_beam(const float* probs, const int32_t* lens, int B, int T, int V, int beam, int blank, int32_t* out_n,
                    int32_t* out_tokens, int32_t* out_lens, double* out_scores, int32_t* out_times, void* ws,
                    size_t ws_bytes, cudaStream_t st) {
  CTCVR_REQUIRE(beam >= 1 && beam <= BEAM_MAX, "ctc_prefix_beam: beam_size %d must be within [1, %d]", beam, BEAM_MAX);
  CTCVR_REQUIRE(ws && ws_bytes >= ctc_prefix_beam_ws_bytes(B, T, V, beam), "ctc_prefix_beam: workspace too small");
  const size_t smem = (((size_t)V * 4 + 15) / 16) * 16 + BEAM_MAX * sizeof(PbEntry) + CAND_MAX * sizeof(PbCand);
  CTCVR_REQUIRE(smem <= 200 * 1024, "ctc_prefix_beam: vocabulary too large for shared memory (V=%d)", V);
  CTCVR_CHECK_CUDA(cudaFuncSetAttribute(ctc_prefix_beam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ctc_prefix_beam_kernel<<<B, PB_THREADS, smem, st>>>(probs, lens, T, V, beam, blank, out_n, out_tokens, out_lens,
                                                      out_scores, out_times, reinterpret_cast<int32_t*>(ws));
  CTCVR_LAUNCH_CHECK();
  return 0;
}

}  // namespace ctcvr
