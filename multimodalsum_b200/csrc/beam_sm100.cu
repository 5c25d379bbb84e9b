// Device-side candidate selection of one beam-search token (BASELINE config 5).
//
// Replaces, per hypothesis row, the chain the reference runs on [N, V] tensors plus host loops every token:
//   adjust_logits_during_generation (force BOS after the start token / EOS at the end of the frame,
//   src/transformer/modeling_multimodalsum.py:3084-3101), log_softmax (:2893), the EOS ban below min_length and the n-gram
//   ban (postprocess_next_token_scores / calc_banned_ngram_tokens, src/transformer/generation_utils.py:57-99, 848-868 —
//   Python loops over batch x beam with .tolist()), `scores + beam_scores` and the top-2k selection (:2919-2925).
// One CTA per row: pass 1 = log-sum-exp of the UNBANNED row (as the reference: bans are applied to the log-probabilities),
// the banned tokens of the row are found from its own history, pass 2 = top-K of the remaining tokens.  The global top-2k
// of a business is a subset of the union of its rows' top-2k, so the host side only merges k x 2k candidates.
// HBM/L2-bound: the fp32 logits row (201 KB) is read twice.
#include "common.cuh"
#include "../../include/mmsum_b200.h"

namespace mmsum {

static constexpr int kBeamThreads = 256;
static constexpr int kMaxBan = 160;

struct Cand { float v; int tok; int src; };
__device__ __forceinline__ bool cand_better(float v, int tok, float v2, int tok2) { return v > v2 || (v == v2 && tok < tok2); }

template <int K>
__global__ void __launch_bounds__(kBeamThreads)
beam_topk_kernel(const float* __restrict__ logits, long long ld, int V, const float* __restrict__ beam_scores,
                 const long long* __restrict__ ids, int L, const long long* __restrict__ cur_dev, int min_length, int ngram,
                 int bos, int eos, float* __restrict__ out_val, int* __restrict__ out_tok) {
  __shared__ float red_m[8], red_s[8];
  __shared__ int s_ban[kMaxBan];
  __shared__ int s_nban;
  __shared__ float s_cv[8];
  __shared__ int s_ct[8], s_cs[8];
  const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* z = logits + (long long)row * ld;
  const long long* hist = ids + (long long)row * L;
  const int cur = (int)cur_dev[0];
  const float bscore = beam_scores[row];
  float* ov = out_val + (long long)row * K;
  int* ot = out_tok + (long long)row * K;
  const int forced = (cur == 1) ? bos : ((cur == L - 1) ? eos : -1);
  if (forced >= 0) {
    // a forced token carries log-probability 0, every other token -inf (their identity is irrelevant; keep them distinct,
    // unfinished and different from the forced token as torch.topk would)
    if (tid < K) {
      ov[tid] = (tid == 0) ? bscore : -INFINITY;
      ot[tid] = (tid == 0) ? forced : (tid + 2 + (forced > 2 && tid + 2 >= forced ? 1 : 0));
    }
    return;
  }
  if (tid == 0) s_nban = 0;
  // ---- pass 1: log-sum-exp of the row
  float m = -INFINITY, s = 0.f;
  const int nvec = V >> 2;
  for (int i = tid; i < nvec; i += kBeamThreads) {
    const float4 f = *reinterpret_cast<const float4*>(z + 4 * i);
    const float bm = fmaxf(fmaxf(f.x, f.y), fmaxf(f.z, f.w));
    const float mn = fmaxf(m, bm);
    s = s * __expf(m - mn) + __expf(f.x - mn) + __expf(f.y - mn) + __expf(f.z - mn) + __expf(f.w - mn);
    m = mn;
  }
  for (int i = nvec * 4 + tid; i < V; i += kBeamThreads) {
    const float v = z[i];
    const float mn = fmaxf(m, v);
    s = s * __expf(m - mn) + __expf(v - mn);
    m = mn;
  }
  const float wm = warp_max(m);
  s = (m == -INFINITY) ? 0.f : s * __expf(m - wm);
  s = warp_sum(s);
  if (lane == 0) { red_m[warp] = wm; red_s[warp] = s; }
  __syncthreads();
  float M = red_m[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) M = fmaxf(M, red_m[w]);
  float S = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) S += (red_m[w] == -INFINITY) ? 0.f : red_s[w] * __expf(red_m[w] - M);
  const float lse = M + logf(S);
  // ---- banned tokens of this row: the EOS below min_length, and every token that would complete an n-gram already present
  if (tid == 0 && cur < min_length) s_ban[atomicAdd(&s_nban, 1)] = eos;
  if (ngram > 0 && cur >= ngram) {
    for (int i = tid; i <= cur - ngram; i += kBeamThreads) {
      bool hit = true;
      for (int j = 0; j < ngram - 1; ++j) hit = hit && (hist[i + j] == hist[cur - ngram + 1 + j]);
      if (hit) {
        const int slot = atomicAdd(&s_nban, 1);
        if (slot < kMaxBan) s_ban[slot] = (int)hist[i + ngram - 1];
      }
    }
  }
  __syncthreads();
  const int nban = min(s_nban, kMaxBan);
  // ---- pass 2: per-thread top-K (sorted, descending), then K rounds of block arg-max over the list heads
  float tv[K];
  int tt[K];
#pragma unroll
  for (int j = 0; j < K; ++j) { tv[j] = -INFINITY; tt[j] = 0x7fffffff; }
  auto offer = [&](float v, int tok) {
    if (cand_better(v, tok, tv[K - 1], tt[K - 1])) {
      for (int b = 0; b < nban; ++b) if (s_ban[b] == tok) return;
      tv[K - 1] = v; tt[K - 1] = tok;
#pragma unroll
      for (int j = K - 1; j > 0; --j) {
        if (cand_better(tv[j], tt[j], tv[j - 1], tt[j - 1])) {
          const float fv = tv[j]; tv[j] = tv[j - 1]; tv[j - 1] = fv;
          const int ft = tt[j]; tt[j] = tt[j - 1]; tt[j - 1] = ft;
        }
      }
    }
  };
  for (int i = tid; i < nvec; i += kBeamThreads) {
    const float4 f = *reinterpret_cast<const float4*>(z + 4 * i);
    offer(f.x, 4 * i); offer(f.y, 4 * i + 1); offer(f.z, 4 * i + 2); offer(f.w, 4 * i + 3);
  }
  for (int i = nvec * 4 + tid; i < V; i += kBeamThreads) offer(z[i], i);
  int head = 0;
  for (int r = 0; r < K; ++r) {
    float bv = -INFINITY; int bt = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < K; ++j) if (j == head) { bv = tv[j]; bt = tt[j]; }
    int bsrc = tid;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float v2 = __shfl_xor_sync(0xffffffffu, bv, o);
      const int t2 = __shfl_xor_sync(0xffffffffu, bt, o);
      const int s2 = __shfl_xor_sync(0xffffffffu, bsrc, o);
      if (cand_better(v2, t2, bv, bt)) { bv = v2; bt = t2; bsrc = s2; }
    }
    if (lane == 0) { s_cv[warp] = bv; s_ct[warp] = bt; s_cs[warp] = bsrc; }
    __syncthreads();
    float gv = s_cv[0]; int gt = s_ct[0], gs = s_cs[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) if (cand_better(s_cv[w], s_ct[w], gv, gt)) { gv = s_cv[w]; gt = s_ct[w]; gs = s_cs[w]; }
    if (tid == gs && head < K) ++head;
    if (tid == 0) {
      ov[r] = (gv == -INFINITY) ? -INFINITY : (gv - lse + bscore);
      ot[r] = (gt == 0x7fffffff) ? (r + 3) : gt;
    }
    __syncthreads();
  }
}

// ----------------------------------------------------------------------------------------------
// Per-business beam update: merge the k x K row candidates, move finished hypotheses into the business' pool, pick the next
// k beams, permute the token histories (and the decoder's self-attention slot table) accordingly.
// One warp per business; the decisions are a few dozen scalar steps (lane 0), the row copies are warp-wide.
// (_generate_beam_search, modeling_multimodalsum.py:2933-3010; BeamHypotheses.add / is_done, generation_utils.py:962-993.)
// ----------------------------------------------------------------------------------------------
static constexpr int kBuMaxK = 8, kBuMaxL = 160, kBuMaxCand = 16;
struct BeamUpdateArgs {
  const float* cand_val; const int* cand_tok;          // [B*k, K]
  long long* ids;                                      // [B*k, L] token histories (in place)
  float* beam_scores;                                  // [B*k]
  unsigned char* done;                                 // [B]
  float* pool_score; long long* pool_tok; long long* pool_len; long long* pool_n;   // [B,k], [B,k,L], [B,k], [B]
  const long long* cur_dev;
  long long* beam_idx; long long* next_tok;            // [B*k] outputs
  int* next_tok32;                                     // optional [B*k]: the decoder's token input of the next step
  int* hist;                                           // optional [B*k, 128]: self-attention slot table, permuted like ids
  int k, K, L, eos, pad, early_stopping;
  float length_penalty;
};

__global__ void __launch_bounds__(32) beam_update_kernel(const BeamUpdateArgs a) {
  __shared__ long long s_ids[kBuMaxK][kBuMaxL];
  __shared__ int s_hist[kBuMaxK][128];
  __shared__ float s_cv[kBuMaxK * kBuMaxCand];
  __shared__ int s_ct[kBuMaxK * kBuMaxCand];
  __shared__ float s_sel_v[kBuMaxCand];
  __shared__ int s_sel_t[kBuMaxCand], s_sel_b[kBuMaxCand];
  __shared__ int s_admit_slot[kBuMaxK], s_admit_src[kBuMaxK], s_n_admit;
  __shared__ int s_nsrc[kBuMaxK], s_ntok[kBuMaxK];
  __shared__ float s_nscore[kBuMaxK];
  const int b = blockIdx.x, lane = threadIdx.x;
  const int k = a.k, K = a.K, L = a.L;
  const int row0 = b * k;
  const int cur = (int)a.cur_dev[0];
  for (int i = lane; i < k * L; i += 32) s_ids[i / L][i % L] = a.ids[(long long)row0 * L + i];
  if (a.hist != nullptr)
    for (int i = lane; i < k * 128; i += 32) s_hist[i >> 7][i & 127] = a.hist[(long long)row0 * 128 + i];
  for (int i = lane; i < k * K; i += 32) { s_cv[i] = a.cand_val[(long long)row0 * K + i]; s_ct[i] = a.cand_tok[(long long)row0 * K + i]; }
  __syncwarp();
  if (lane == 0) {
    // ---- the K best of the k x K row candidates, by descending score (rows are sorted: ties go to the lower flat index)
    for (int r = 0; r < K; ++r) {
      int best = -1;
      for (int i = 0; i < k * K; ++i)
        if (s_ct[i] >= 0 && (best < 0 || s_cv[i] > s_cv[best])) best = i;
      s_sel_v[r] = s_cv[best]; s_sel_t[r] = s_ct[best]; s_sel_b[r] = best / K;
      s_ct[best] = -1;
    }
    const bool live = a.done[b] == 0;
    const float norm = powf((float)cur, a.length_penalty);
    float* ps = a.pool_score + (long long)b * k;
    long long* pl = a.pool_len + (long long)b * k;
    int pn = (int)a.pool_n[b];
    // ---- finished candidates among the first k ranks enter the pool (a full pool admits only a better-than-worst one)
    int n_admit = 0;
    for (int r = 0; r < k; ++r) {
      if (!(live && s_sel_t[r] == a.eos)) continue;
      const float score = s_sel_v[r] / norm;
      int slot = -1;
      if (pn < k) { slot = pn; ++pn; }
      else {
        int w = 0;
        for (int j = 1; j < k; ++j) if (ps[j] < ps[w]) w = j;
        if (score > ps[w]) slot = w;
      }
      if (slot >= 0) {
        ps[slot] = score; pl[slot] = cur;
        // an earlier admission of this step into the same slot is superseded
        for (int q = 0; q < n_admit; ++q) if (s_admit_slot[q] == slot) s_admit_slot[q] = -1;
        s_admit_slot[n_admit] = slot; s_admit_src[n_admit] = s_sel_b[r]; ++n_admit;
      }
    }
    s_n_admit = n_admit;
    a.pool_n[b] = pn;
    // ---- the first k unfinished candidates, in rank order, are the next beams
    int nb = 0;
    for (int r = 0; r < K && nb < k; ++r) {
      if (s_sel_t[r] == a.eos) continue;
      s_nscore[nb] = s_sel_v[r]; s_ntok[nb] = s_sel_t[r]; s_nsrc[nb] = s_sel_b[r]; ++nb;
    }
    for (; nb < k; ++nb) { s_nscore[nb] = -INFINITY; s_ntok[nb] = a.pad; s_nsrc[nb] = nb; }   // cannot happen with K = 2k
    bool finished = pn >= k;
    if (finished && !a.early_stopping) {
      float w = ps[0];
      for (int j = 1; j < k; ++j) w = fminf(w, ps[j]);
      finished = w >= s_sel_v[0] / norm;
    }
    if (!live) for (int j = 0; j < k; ++j) { s_nscore[j] = 0.f; s_ntok[j] = a.pad; s_nsrc[j] = j; }
    if (finished) a.done[b] = 1;
  }
  __syncwarp();
  // ---- pool token rows (from the histories as they were before the permutation)
  for (int q = 0; q < s_n_admit; ++q) {
    const int slot = s_admit_slot[q];
    if (slot < 0) continue;
    long long* dst = a.pool_tok + ((long long)b * k + slot) * L;
    for (int t = lane; t < L; t += 32) dst[t] = s_ids[s_admit_src[q]][t];
  }
  // ---- next beams: scores, source rows, tokens; histories and slot table permuted in place (business-local)
  for (int j = lane; j < k; j += 32) {
    a.beam_scores[row0 + j] = s_nscore[j];
    a.beam_idx[row0 + j] = row0 + s_nsrc[j];
    a.next_tok[row0 + j] = s_ntok[j];
    if (a.next_tok32 != nullptr) a.next_tok32[row0 + j] = s_ntok[j];
  }
  for (int i = lane; i < k * L; i += 32) {
    const int j = i / L, t = i % L;
    a.ids[(long long)(row0 + j) * L + t] = (t == cur) ? (long long)s_ntok[j] : s_ids[s_nsrc[j]][t];
  }
  if (a.hist != nullptr)
    for (int i = lane; i < k * 128; i += 32) a.hist[(long long)(row0 + (i >> 7)) * 128 + (i & 127)] = s_hist[s_nsrc[i >> 7]][i & 127];
}

}  // namespace mmsum

using namespace mmsum;

extern "C" int mmsum_beam_update(const float* cand_val, const int32_t* cand_tok, int64_t* ids, float* beam_scores, uint8_t* done,
                                 float* pool_score, int64_t* pool_tok, int64_t* pool_len, int64_t* pool_n, const int64_t* cur_dev,
                                 int64_t* beam_idx, int64_t* next_tok, int32_t* next_tok32, int32_t* hist, int32_t B, int32_t k,
                                 int32_t K, int32_t L, int32_t eos, int32_t pad, int32_t early_stopping, float length_penalty,
                                 void* stream_v) {
  if (!cand_val || !cand_tok || !ids || !beam_scores || !done || !pool_score || !pool_tok || !pool_len || !pool_n || !cur_dev ||
      !beam_idx || !next_tok) return MMSUM_ERR_INVALID;
  if (B <= 0 || k <= 0 || k > kBuMaxK || K < k || K > kBuMaxCand || L < 2 || L > kBuMaxL) return MMSUM_ERR_INVALID;
  BeamUpdateArgs a;
  a.cand_val = cand_val; a.cand_tok = cand_tok; a.ids = reinterpret_cast<long long*>(ids); a.beam_scores = beam_scores; a.done = done;
  a.pool_score = pool_score; a.pool_tok = reinterpret_cast<long long*>(pool_tok); a.pool_len = reinterpret_cast<long long*>(pool_len);
  a.pool_n = reinterpret_cast<long long*>(pool_n); a.cur_dev = reinterpret_cast<const long long*>(cur_dev);
  a.beam_idx = reinterpret_cast<long long*>(beam_idx); a.next_tok = reinterpret_cast<long long*>(next_tok); a.next_tok32 = next_tok32;
  a.hist = hist; a.k = k; a.K = K; a.L = L; a.eos = eos; a.pad = pad; a.early_stopping = early_stopping; a.length_penalty = length_penalty;
  beam_update_kernel<<<B, 32, 0, reinterpret_cast<cudaStream_t>(stream_v)>>>(a);
  MMSUM_CHECK_LAUNCH();
  return 0;
}

extern "C" int mmsum_beam_topk(const float* logits, int64_t ld, int32_t rows, int32_t V, const float* beam_scores,
                               const int64_t* ids, int32_t L, const int64_t* cur_dev, int32_t min_length, int32_t ngram,
                               int32_t bos, int32_t eos, int32_t K, float* out_val, int32_t* out_tok, void* stream_v) {
  if (!logits || !beam_scores || !ids || !cur_dev || !out_val || !out_tok || rows <= 0 || V < 32 || L < 2) return MMSUM_ERR_INVALID;
  if ((ld % 4) || ld < V || (reinterpret_cast<uintptr_t>(logits) & 15)) return MMSUM_ERR_INVALID;
  if (ngram < 0 || ngram > 8 || L - 1 + 1 > kMaxBan) return MMSUM_ERR_INVALID;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  const long long* ids_ll = reinterpret_cast<const long long*>(ids);
  const long long* cur_ll = reinterpret_cast<const long long*>(cur_dev);
  if (K == 8) beam_topk_kernel<8><<<rows, kBeamThreads, 0, stream>>>(logits, ld, V, beam_scores, ids_ll, L, cur_ll, min_length, ngram, bos, eos, out_val, out_tok);
  else if (K == 16) beam_topk_kernel<16><<<rows, kBeamThreads, 0, stream>>>(logits, ld, V, beam_scores, ids_ll, L, cur_ll, min_length, ngram, bos, eos, out_val, out_tok);
  else if (K == 2) beam_topk_kernel<2><<<rows, kBeamThreads, 0, stream>>>(logits, ld, V, beam_scores, ids_ll, L, cur_ll, min_length, ngram, bos, eos, out_val, out_tok);
  else if (K == 4) beam_topk_kernel<4><<<rows, kBeamThreads, 0, stream>>>(logits, ld, V, beam_scores, ids_ll, L, cur_ll, min_length, ngram, bos, eos, out_val, out_tok);
  else return MMSUM_ERR_INVALID;
  MMSUM_CHECK_LAUNCH();
  return 0;
}
