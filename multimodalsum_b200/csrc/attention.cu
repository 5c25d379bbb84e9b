// Multi-entity attention (forward, dQ backward, dK/dV backward) for the MultimodalSum step.
//
// One family of kernels covers
//   * encoder self-attention       (modeling_multimodalsum.py:746-749, 783-853; key-pad mask)
//   * decoder causal self-attention (same call path, + causal triu mask)
//   * the multi-entity, multi-modal cross-attention (:722-745, :768-869): for every modality an independent
//     softmax PER ENTITY (review / table / image), then the mean over the entities that have at least one valid
//     key; the leave-one-out target of multimodal_train.py:150-163 is just "entity i is excluded for target i".
// Self-attention is the special case "one modality, one entity, memory = own sequence".
//
// Masking semantics: pad keys of a valid entity get probability exactly 0 (the reference fills -2^16 / -inf, both
// underflow to 0 in fp32 next to any valid key); entities without a valid key are skipped (the reference zeroes
// them and removes them from the divisor); a modality without any valid entity yields 0.
//
// Shapes are fixed to the model's: 128 query positions per sequence, head_dim 64.  Scores never leave the SM:
// per (sequence, head) the kernel walks 64-key blocks with an online softmax.  Math is bf16 mma.sync m16n8k16 with
// fp32 accumulation (the tensor-core GEMMs of the step run on tcgen05 — gemm_sm100.cu; these small 128x64xK
// problems are the next candidate for a tcgen05 port).
#include "common.cuh"
#include "../../include/mmsum_b200.h"

namespace mmsum {

static constexpr int SQ = 128;   // query rows per sequence
static constexpr int HD = 64;    // head dim
static constexpr int KB = 64;    // keys per block
static constexpr int kMaxItems = 64;
static constexpr float kLog2e = 1.4426950408889634f;

struct Item {
  long long kv_row0;   // global KV row of the first key of this block
  int nkeys;           // valid key slots in this block (<= 64)
  int key0;            // index of the first key within the entity
  short mod, ent;      // modality index, global entity index (for LSE / DELTA)
  short first, last;   // first / last block of its entity
};

// [rows][64] bf16 tile, 16-byte chunks XOR-swizzled by row so ldmatrix is conflict-free
__device__ __forceinline__ uint32_t tile_addr(uint32_t base, int row, int chunk) {
  return base + row * 128 + (((chunk ^ row) & 7) << 4);
}

template <int NT>
__device__ __forceinline__ void load_tile_async(uint32_t sbase, const bf16* g, long long ld, int rows, int rows_valid) {
  for (int c = threadIdx.x; c < rows * 8; c += NT) {
    const int r = c >> 3, ch = c & 7;
    const bool ok = r < rows_valid;
    const bf16* src = g + (long long)(ok ? r : 0) * ld + ch * 8;
    cp_async_16(tile_addr(sbase, r, ch), src, ok);
  }
}

// A fragments (16 rows x 64 cols) of a row-major tile: rows r0..r0+15
__device__ __forceinline__ void load_a_frags(uint32_t (&f)[4][4], uint32_t sbase, int r0, int lane) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
    ldmatrix_x4(f[ks], tile_addr(sbase, r0 + (lane & 7) + (((lane >> 3) & 1) << 3), ks * 2 + (lane >> 4)));
}

// acc[16 x 64] += A(16 x 64 over k) * T^T where T is a [64 n][64 k] row-major tile  (n = tile row, k = tile col)
__device__ __forceinline__ void mma_a_tileT(float (&acc)[8][4], const uint32_t (&a)[4][4], uint32_t sbase, int lane) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b[4];
      ldmatrix_x4(b, tile_addr(sbase, np * 16 + (lane & 7) + ((lane >> 4) << 3), ks * 2 + ((lane >> 3) & 1)));
      const uint32_t b0[2] = {b[0], b[1]}, b1[2] = {b[2], b[3]};
      mma_bf16_16816(acc[2 * np], a[ks], b0);
      mma_bf16_16816(acc[2 * np + 1], a[ks], b1);
    }
  }
}

// acc[16 x 64] += P(16 x 64 over k, from C-layout registers) * T where T is a [64 k][64 n] row-major tile
__device__ __forceinline__ void pack_p(uint32_t (&a)[4][4], const float (&p)[8][4]) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    a[kk][0] = pack_bf16(p[2 * kk][0], p[2 * kk][1]);
    a[kk][1] = pack_bf16(p[2 * kk][2], p[2 * kk][3]);
    a[kk][2] = pack_bf16(p[2 * kk + 1][0], p[2 * kk + 1][1]);
    a[kk][3] = pack_bf16(p[2 * kk + 1][2], p[2 * kk + 1][3]);
  }
}
__device__ __forceinline__ void mma_p_tile(float (&acc)[8][4], const uint32_t (&a)[4][4], uint32_t sbase, int lane) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
    for (int dp = 0; dp < 4; ++dp) {
      uint32_t b[4];
      ldmatrix_x4_trans(b, tile_addr(sbase, kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3), dp * 2 + (lane >> 4)));
      const uint32_t b0[2] = {b[0], b[1]}, b1[2] = {b[2], b[3]};
      mma_bf16_16816(acc[2 * dp], a[kk], b0);
      mma_bf16_16816(acc[2 * dp + 1], a[kk], b1);
    }
  }
}
// same, but two accumulators sharing the B fragments
__device__ __forceinline__ void mma_p_tile2(float (&acc0)[8][4], const uint32_t (&a0)[4][4], float (&acc1)[8][4],
                                            const uint32_t (&a1)[4][4], uint32_t sbase, int lane) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
    for (int dp = 0; dp < 4; ++dp) {
      uint32_t b[4];
      ldmatrix_x4_trans(b, tile_addr(sbase, kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3), dp * 2 + (lane >> 4)));
      const uint32_t b0[2] = {b[0], b[1]}, b1[2] = {b[2], b[3]};
      mma_bf16_16816(acc0[2 * dp], a0[kk], b0);
      mma_bf16_16816(acc0[2 * dp + 1], a0[kk], b1);
      mma_bf16_16816(acc1[2 * dp], a1[kk], b0);
      mma_bf16_16816(acc1[2 * dp + 1], a1[kk], b1);
    }
  }
}

__device__ __forceinline__ void zero_acc(float (&a)[8][4]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i][0] = 0.f; a[i][1] = 0.f; a[i][2] = 0.f; a[i][3] = 0.f; }
}

// Enumerate the (modality, entity, key block) work items of one query sequence into shared memory.
__device__ int build_items(const MmsumAttnArgs& p, int qseq, Item* items, int* mod_begin) {
  const int biz = qseq / p.R;
  const int tgt = qseq - biz * p.R;
  int n = 0;
  for (int m = 0; m < p.n_mod; ++m) {
    mod_begin[m] = n;
    const MmsumAttnMod& md = p.mods[m];
    for (int e = 0; e < md.E; ++e) {
      if (md.loo && e == tgt) continue;
      const int ge = md.ent_base + e;
      if (p.ent_valid != nullptr && p.ent_valid[(long long)biz * p.E_total + ge] == 0) continue;
      const long long row0 = md.kv_row_base + ((long long)biz * md.E + e) * md.Sk;
      const int nb = (md.Sk + KB - 1) / KB;
      for (int kb = 0; kb < nb; ++kb) {
        if (p.causal && kb * KB > SQ - 1) break;
        Item it;
        it.kv_row0 = row0 + kb * KB;
        it.key0 = kb * KB;
        it.nkeys = min(KB, md.Sk - kb * KB);
        it.mod = (short)m; it.ent = (short)ge;
        it.first = (kb == 0); it.last = (kb == nb - 1) || (p.causal && (kb + 1) * KB > SQ - 1);
        if (n < kMaxItems) items[n++] = it;
      }
    }
  }
  mod_begin[p.n_mod] = n;
  return n;
}

struct FwdSmem {
  uint8_t q[SQ * 128];
  uint8_t k[2][KB * 128];
  uint8_t v[2][KB * 128];
  uint8_t kvalid[2][KB];
  Item items[kMaxItems];
  int mod_begin[4];
  int n_items;
};

template <int NT>
__device__ __forceinline__ void prefetch_kv(const MmsumAttnArgs& p, const Item& it, int h, uint32_t sk, uint32_t sv,
                                            uint8_t* kvalid) {
  const bf16* kv = reinterpret_cast<const bf16*>(p.KV);
  load_tile_async<NT>(sk, kv + it.kv_row0 * p.ldkv + p.k_col + h * HD, p.ldkv, KB, it.nkeys);
  load_tile_async<NT>(sv, kv + it.kv_row0 * p.ldkv + p.v_col + h * HD, p.ldkv, KB, it.nkeys);
  if (threadIdx.x < KB) {
    const int j = threadIdx.x;
    uint8_t ok = (j < it.nkeys) ? 1 : 0;
    if (ok && p.key_valid != nullptr) ok = p.key_valid[it.kv_row0 + j];
    kvalid[j] = ok;
  }
}

// ----------------------------------------------------------------------------------------------
// forward: O_mod[qrow, h*64:...] = (1/n) sum_e softmax_e(scale * Q K_e^T) V_e ;  LSE[qseq,h,e,row]
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2) attn_fwd_kernel(const MmsumAttnArgs p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  FwdSmem& sm = *reinterpret_cast<FwdSmem*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  // block order keeps the R targets of one (business, head) adjacent so their shared K/V stays in L2
  const int tgt = blockIdx.x % p.R;
  const int h = (blockIdx.x / p.R) % p.H;
  const int biz = blockIdx.x / (p.R * p.H);
  const int qseq = biz * p.R + tgt;
  const long long qrow0 = (long long)qseq * SQ;

  if (threadIdx.x == 0) sm.n_items = build_items(p, qseq, sm.items, sm.mod_begin);
  const uint32_t sq = smem_u32(sm.q);
  load_tile_async<256>(sq, reinterpret_cast<const bf16*>(p.Q) + qrow0 * p.ldq + p.q_col + h * HD, p.ldq, SQ, SQ);
  cp_async_commit();
  __syncthreads();
  const int n_items = sm.n_items;
  if (n_items > 0) prefetch_kv<256>(p, sm.items[0], h, smem_u32(sm.k[0]), smem_u32(sm.v[0]), sm.kvalid[0]);
  cp_async_commit();
  cp_async_wait<1>();
  __syncthreads();
  uint32_t qf[4][4];
  const int r0 = warp * 16;
  load_a_frags(qf, sq, r0, lane);

  const float sc = p.scale * kLog2e;
  float o[8][4], acc[8][4];
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  zero_acc(o); zero_acc(acc);
  int cur_mod = 0;
  bf16* O = reinterpret_cast<bf16*>(p.O);

  auto flush_mod = [&](int m) {
    // write the entity-mean of modality m (zeros when it had no valid entity) and reset the accumulator
    bf16* dst = O + p.mods[m].o_off + (qrow0 + r0) * p.ldo + h * HD;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      *reinterpret_cast<uint32_t*>(dst + (long long)g * p.ldo + nt * 8 + 2 * t) = pack_bf16(acc[nt][0], acc[nt][1]);
      *reinterpret_cast<uint32_t*>(dst + (long long)(g + 8) * p.ldo + nt * 8 + 2 * t) = pack_bf16(acc[nt][2], acc[nt][3]);
    }
    zero_acc(acc);
  };

  for (int i = 0; i < n_items; ++i) {
    const int buf = i & 1;
    if (i + 1 < n_items)
      prefetch_kv<256>(p, sm.items[i + 1], h, smem_u32(sm.k[buf ^ 1]), smem_u32(sm.v[buf ^ 1]), sm.kvalid[buf ^ 1]);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const Item it = sm.items[i];
    while (cur_mod < it.mod) { flush_mod(cur_mod); ++cur_mod; }
    if (it.first) { zero_acc(o); m_run[0] = m_run[1] = -INFINITY; l_run[0] = l_run[1] = 0.f; }

    float s[8][4];
    zero_acc(s);
    mma_a_tileT(s, qf, smem_u32(sm.k[buf]), lane);
    // mask + online softmax (rows g and g+8 of this warp's 16)
    float bm[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = nt * 8 + 2 * t + (c & 1);
        const int qr = r0 + g + ((c >> 1) << 3);
        bool ok = sm.kvalid[buf][j] != 0;
        if (p.causal) ok = ok && (it.key0 + j <= qr);
        s[nt][c] = ok ? s[nt][c] * sc : -INFINITY;
        bm[c >> 1] = fmaxf(bm[c >> 1], s[nt][c]);
      }
    }
    float corr[2], mnew[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      bm[r] = fmaxf(bm[r], __shfl_xor_sync(0xffffffffu, bm[r], 1));
      bm[r] = fmaxf(bm[r], __shfl_xor_sync(0xffffffffu, bm[r], 2));
      mnew[r] = fmaxf(m_run[r], bm[r]);
      const float msafe = (mnew[r] == -INFINITY) ? 0.f : mnew[r];
      corr[r] = exp2f(m_run[r] - msafe);   // m_run = -inf -> 0
      m_run[r] = mnew[r];
      mnew[r] = msafe;
    }
    float rs[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float e = exp2f(s[nt][c] - mnew[c >> 1]);
        s[nt][c] = e;
        rs[c >> 1] += e;
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * corr[r] + rs[r];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { o[nt][0] *= corr[0]; o[nt][1] *= corr[0]; o[nt][2] *= corr[1]; o[nt][3] *= corr[1]; }
    uint32_t pf[4][4];
    pack_p(pf, s);
    mma_p_tile(o, pf, smem_u32(sm.v[buf]), lane);

    if (it.last) {
      const float inv_n = p.inv_n ? p.inv_n[(long long)qseq * p.n_mod + it.mod] : 1.f;
      float* lse = p.LSE + (((long long)qseq * p.H + h) * p.E_total + it.ent) * SQ + r0;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float l = l_run[r];
        l += __shfl_xor_sync(0xffffffffu, l, 1);
        l += __shfl_xor_sync(0xffffffffu, l, 2);
        const float w = (l > 0.f) ? inv_n / l : 0.f;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) { acc[nt][2 * r] += o[nt][2 * r] * w; acc[nt][2 * r + 1] += o[nt][2 * r + 1] * w; }
        if (t == 0) lse[g + 8 * r] = (l > 0.f) ? (m_run[r] + log2f(l)) : INFINITY;  // log2 domain, includes scale
      }
    }
    __syncthreads();   // everyone done with buf before it is refilled two iterations later
  }
  while (cur_mod < p.n_mod) { flush_mod(cur_mod); ++cur_mod; }
}

// ----------------------------------------------------------------------------------------------
// backward, part 1 (per query sequence): dQ and DELTA
//   P = exp2(sc*QK^T - LSE),  dP' = dA V^T,  delta' = rowsum(P o dP'),
//   dQ += scale * inv_n * ( (P o (dP' - dref)) K + (dref - delta') * (P K) )   summed over entities and modalities
//   (dref = estimate of delta' from the entity's first key block, so no cancellation happens after bf16 rounding)
// DELTA stores delta' (un-normalised by inv_n) for part 2.
// ----------------------------------------------------------------------------------------------
struct BwdQSmem {
  uint8_t q[SQ * 128];
  uint8_t dA[SQ * 128];
  uint8_t k[2][KB * 128];
  uint8_t v[2][KB * 128];
  uint8_t kvalid[2][KB];
  Item items[kMaxItems];
  int mod_begin[4];
  int n_items;
};

__global__ void __launch_bounds__(256, 1) attn_bwd_dq_kernel(const MmsumAttnArgs p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  BwdQSmem& sm = *reinterpret_cast<BwdQSmem*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int tgt = blockIdx.x % p.R;
  const int h = (blockIdx.x / p.R) % p.H;
  const int biz = blockIdx.x / (p.R * p.H);
  const int qseq = biz * p.R + tgt;
  const long long qrow0 = (long long)qseq * SQ;
  const int r0 = warp * 16;

  if (threadIdx.x == 0) sm.n_items = build_items(p, qseq, sm.items, sm.mod_begin);
  const uint32_t sq = smem_u32(sm.q), sda = smem_u32(sm.dA);
  load_tile_async<256>(sq, reinterpret_cast<const bf16*>(p.Q) + qrow0 * p.ldq + p.q_col + h * HD, p.ldq, SQ, SQ);
  cp_async_commit();
  __syncthreads();
  const int n_items = sm.n_items;
  cp_async_wait<0>();
  __syncthreads();
  uint32_t qf[4][4];
  load_a_frags(qf, sq, r0, lane);

  const float sc = p.scale * kLog2e;
  float dq[8][4], X[8][4], Y[8][4];
  zero_acc(dq);
  float dl[2] = {0.f, 0.f}, dref[2] = {0.f, 0.f};
  uint32_t daf[4][4];
  int cur_mod = -1;
  const bf16* dO = reinterpret_cast<const bf16*>(p.O);

  if (n_items > 0) prefetch_kv<256>(p, sm.items[0], h, smem_u32(sm.k[0]), smem_u32(sm.v[0]), sm.kvalid[0]);
  cp_async_commit();

  for (int i = 0; i < n_items; ++i) {
    const int buf = i & 1;
    const Item it = sm.items[i];
    if (it.mod != cur_mod) {
      // upstream gradient tile of this modality
      __syncthreads();
      load_tile_async<256>(sda, dO + p.mods[it.mod].o_off + qrow0 * p.ldo + h * HD, p.ldo, SQ, SQ);
      cp_async_commit();
      cp_async_wait<0>();
      __syncthreads();
      load_a_frags(daf, sda, r0, lane);
      cur_mod = it.mod;
    }
    if (i + 1 < n_items)
      prefetch_kv<256>(p, sm.items[i + 1], h, smem_u32(sm.k[buf ^ 1]), smem_u32(sm.v[buf ^ 1]), sm.kvalid[buf ^ 1]);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (it.first) { zero_acc(X); zero_acc(Y); dl[0] = dl[1] = 0.f; dref[0] = dref[1] = 0.f; }

    const float* lse = p.LSE + (((long long)qseq * p.H + h) * p.E_total + it.ent) * SQ + r0;
    const float lse_r[2] = {lse[g], lse[g + 8]};
    float s[8][4], dp[8][4];
    zero_acc(s); zero_acc(dp);
    mma_a_tileT(s, qf, smem_u32(sm.k[buf]), lane);
    mma_a_tileT(dp, daf, smem_u32(sm.v[buf]), lane);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = nt * 8 + 2 * t + (c & 1);
        const int qr = r0 + g + ((c >> 1) << 3);
        bool ok = sm.kvalid[buf][j] != 0;
        if (p.causal) ok = ok && (it.key0 + j <= qr);
        const float pr = ok ? exp2f(s[nt][c] * sc - lse_r[c >> 1]) : 0.f;
        s[nt][c] = pr;
        dl[c >> 1] += dp[nt][c] * pr;
      }
    }
    if (it.first) {
      // delta_ref: the softmax-weighted mean of dP' over the FIRST key block.  dS is formed as P o (dP' - delta_ref)
      // before the bf16 rounding (no cancellation after rounding); the exact delta is restored through Y at the end.
      float pa[2] = {0.f, 0.f};
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) { pa[0] += s[nt][0] + s[nt][1]; pa[1] += s[nt][2] + s[nt][3]; }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float a = dl[r], b = pa[r];
        a += __shfl_xor_sync(0xffffffffu, a, 1); a += __shfl_xor_sync(0xffffffffu, a, 2);
        b += __shfl_xor_sync(0xffffffffu, b, 1); b += __shfl_xor_sync(0xffffffffu, b, 2);
        dref[r] = (b > 0.f) ? a / b : 0.f;
      }
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int c = 0; c < 4; ++c) dp[nt][c] = s[nt][c] * (dp[nt][c] - dref[c >> 1]);
    }
    uint32_t pf[4][4], pdf[4][4];
    pack_p(pf, s);
    pack_p(pdf, dp);
    mma_p_tile2(X, pdf, Y, pf, smem_u32(sm.k[buf]), lane);

    if (it.last) {
      const float inv_n = p.inv_n ? p.inv_n[(long long)qseq * p.n_mod + it.mod] : 1.f;
      float* dlt = p.DELTA + (((long long)qseq * p.H + h) * p.E_total + it.ent) * SQ + r0;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float d = dl[r];
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        if (t == 0) dlt[g + 8 * r] = d;
        const float w = p.scale * inv_n;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          dq[nt][2 * r] += w * (X[nt][2 * r] + (dref[r] - d) * Y[nt][2 * r]);
          dq[nt][2 * r + 1] += w * (X[nt][2 * r + 1] + (dref[r] - d) * Y[nt][2 * r + 1]);
        }
      }
    }
    __syncthreads();
  }
  bf16* dQ = reinterpret_cast<bf16*>(p.dQ) + (qrow0 + r0) * p.lddq + p.dq_col + h * HD;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    *reinterpret_cast<uint32_t*>(dQ + (long long)g * p.lddq + nt * 8 + 2 * t) = pack_bf16(dq[nt][0], dq[nt][1]);
    *reinterpret_cast<uint32_t*>(dQ + (long long)(g + 8) * p.lddq + nt * 8 + 2 * t) = pack_bf16(dq[nt][2], dq[nt][3]);
  }
}

// ----------------------------------------------------------------------------------------------
// backward, part 2 (per business x head x entity x 64-key block): dK, dV summed over the consumer targets
//   Pn^T = inv_n * exp2(sc*K Q^T - LSE),  dV += Pn^T dA,  dP'^T = V dA^T,  dS^T = Pn^T o (dP'^T - delta'),
//   dK += scale * dS^T Q
// ----------------------------------------------------------------------------------------------
struct BwdKVSmem {
  uint8_t k[KB * 128];
  uint8_t v[KB * 128];
  uint8_t q[2][64 * 128];
  uint8_t dA[2][64 * 128];
  float lse[2][64];
  float dlt[2][64];
};

__global__ void __launch_bounds__(128, 2) attn_bwd_dkv_kernel(const MmsumAttnArgs p, int blocks_per_bh) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  BwdKVSmem& sm = *reinterpret_cast<BwdKVSmem*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  // decode (biz, h, modality, entity, key block)
  int rem = blockIdx.x % blocks_per_bh;
  const int bh = blockIdx.x / blocks_per_bh;
  const int h = bh % p.H, biz = bh / p.H;
  int m = 0, e = 0, kb = 0;
  for (m = 0; m < p.n_mod; ++m) {
    const int nb = (p.mods[m].Sk + KB - 1) / KB;
    const int cnt = p.mods[m].E * nb;
    if (rem < cnt) { e = rem / nb; kb = rem % nb; break; }
    rem -= cnt;
  }
  const MmsumAttnMod& md = p.mods[m];
  const int ge = md.ent_base + e;
  const int key0 = kb * KB;
  const int nkeys = min(KB, md.Sk - key0);
  const long long kvrow0 = md.kv_row_base + ((long long)biz * md.E + e) * md.Sk + key0;
  const int k0w = warp * 16;  // this warp's 16 keys
  bf16* dKV = reinterpret_cast<bf16*>(p.dKV);

  float dk[8][4], dv[8][4];
  zero_acc(dk); zero_acc(dv);
  const bool ent_ok = (p.ent_valid == nullptr) || (p.ent_valid[(long long)biz * p.E_total + ge] != 0);
  const bool blk_ok = !(p.causal && key0 > SQ - 1);

  if (ent_ok && blk_ok) {
    const bf16* kv = reinterpret_cast<const bf16*>(p.KV);
    const uint32_t sk = smem_u32(sm.k), sv = smem_u32(sm.v);
    load_tile_async<128>(sk, kv + kvrow0 * p.ldkv + p.k_col + h * HD, p.ldkv, KB, nkeys);
    load_tile_async<128>(sv, kv + kvrow0 * p.ldkv + p.v_col + h * HD, p.ldkv, KB, nkeys);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    uint32_t kf[4][4], vf[4][4];
    load_a_frags(kf, sk, k0w, lane);
    load_a_frags(vf, sv, k0w, lane);
    bool kval[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int j = k0w + g + 8 * r;
      kval[r] = (j < nkeys) && (p.key_valid == nullptr || p.key_valid[kvrow0 + j] != 0);
    }
    const float sc = p.scale * kLog2e;
    const bf16* Q = reinterpret_cast<const bf16*>(p.Q);
    const bf16* dO = reinterpret_cast<const bf16*>(p.O);

    // flattened loop over (target, 64-query block); skip the leave-one-out target and causally dead blocks
    const int n_steps = p.R * 2;
    auto step_ok = [&](int s) {
      const int tg = s >> 1, qb = s & 1;
      if (md.loo && tg == e) return false;
      if (p.causal && (qb * 64 + 63 < key0)) return false;
      return true;
    };
    auto issue = [&](int s, int buf) {
      const int tg = s >> 1, qb = s & 1;
      const int qseq = biz * p.R + tg;
      const long long row0 = (long long)qseq * SQ + qb * 64;
      load_tile_async<128>(smem_u32(sm.q[buf]), Q + row0 * p.ldq + p.q_col + h * HD, p.ldq, 64, 64);
      load_tile_async<128>(smem_u32(sm.dA[buf]), dO + md.o_off + row0 * p.ldo + h * HD, p.ldo, 64, 64);
      if (threadIdx.x < 64) {
        const long long li = (((long long)qseq * p.H + h) * p.E_total + ge) * SQ + qb * 64 + threadIdx.x;
        sm.lse[buf][threadIdx.x] = p.LSE[li];
        sm.dlt[buf][threadIdx.x] = p.DELTA[li];
      }
    };
    int s_cur = 0;
    while (s_cur < n_steps && !step_ok(s_cur)) ++s_cur;
    if (s_cur < n_steps) issue(s_cur, 0);
    cp_async_commit();
    int buf = 0;
    while (s_cur < n_steps) {
      int s_next = s_cur + 1;
      while (s_next < n_steps && !step_ok(s_next)) ++s_next;
      if (s_next < n_steps) issue(s_next, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
      __syncthreads();
      const int tg = s_cur >> 1, qb = s_cur & 1;
      const int qseq = biz * p.R + tg;
      const float inv_n = p.inv_n ? p.inv_n[(long long)qseq * p.n_mod + m] : 1.f;
      const uint32_t sq = smem_u32(sm.q[buf]), sda = smem_u32(sm.dA[buf]);

      float st[8][4], dpt[8][4];
      zero_acc(st); zero_acc(dpt);
      mma_a_tileT(st, kf, sq, lane);     // S^T  = K Q^T      [16 keys x 64 queries]
      mma_a_tileT(dpt, vf, sda, lane);   // dP'^T = V dA^T
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int jq = nt * 8 + 2 * t + (c & 1);       // query within the block
          const int r = c >> 1;                          // key row g / g+8
          bool ok = kval[r];
          if (p.causal) ok = ok && (key0 + k0w + g + 8 * r <= qb * 64 + jq);
          const float pr = ok ? inv_n * exp2f(st[nt][c] * sc - sm.lse[buf][jq]) : 0.f;
          st[nt][c] = pr;
          dpt[nt][c] = pr * (dpt[nt][c] - sm.dlt[buf][jq]);
        }
      }
      uint32_t pf[4][4], dsf[4][4];
      pack_p(pf, st);
      pack_p(dsf, dpt);
      mma_p_tile(dv, pf, sda, lane);     // dV += Pn^T dA
      mma_p_tile(dk, dsf, sq, lane);     // dK += dS^T Q
      __syncthreads();
      s_cur = s_next;
      buf ^= 1;
    }
  }
  // write (zeros for skipped / null entities so the buffer never needs a memset)
  for (int r = 0; r < 2; ++r) {
    const int j = k0w + g + 8 * r;
    if (j < nkeys) {
      bf16* dkp = dKV + (kvrow0 + j) * p.lddkv + p.dk_col + h * HD;
      bf16* dvp = dKV + (kvrow0 + j) * p.lddkv + p.dv_col + h * HD;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        *reinterpret_cast<uint32_t*>(dkp + nt * 8 + 2 * t) = pack_bf16(dk[nt][2 * r] * p.scale, dk[nt][2 * r + 1] * p.scale);
        *reinterpret_cast<uint32_t*>(dvp + nt * 8 + 2 * t) = pack_bf16(dv[nt][2 * r], dv[nt][2 * r + 1]);
      }
    }
  }
}

static int validate(const MmsumAttnArgs* a, bool bwd) {
  if (!a || !a->Q || !a->KV || !a->O || !a->LSE) return MMSUM_ERR_INVALID;
  if (a->n_qseq <= 0 || a->H <= 0 || a->R <= 0 || a->n_mod < 1 || a->n_mod > 3) return MMSUM_ERR_INVALID;
  if (a->n_qseq % a->R) return MMSUM_ERR_INVALID;
  if ((a->ldq % 8) || (a->ldkv % 8) || (a->ldo % 8) || (a->q_col % 8) || (a->k_col % 8) || (a->v_col % 8)) return MMSUM_ERR_INVALID;
  int items = 0, ents = 0;
  for (int m = 0; m < a->n_mod; ++m) {
    if (a->mods[m].E <= 0 || a->mods[m].Sk <= 0) return MMSUM_ERR_INVALID;
    items += a->mods[m].E * ((a->mods[m].Sk + KB - 1) / KB);
    ents += a->mods[m].E;
    if (a->mods[m].o_off % 8) return MMSUM_ERR_INVALID;
  }
  if (items > kMaxItems || ents > a->E_total) return MMSUM_ERR_INVALID;
  if (a->causal && (a->n_mod != 1 || a->mods[0].Sk != SQ)) return MMSUM_ERR_INVALID;
  if (bwd) {
    if (!a->DELTA || !a->dQ || !a->dKV) return MMSUM_ERR_INVALID;
    if ((a->lddq % 8) || (a->lddkv % 8) || (a->dq_col % 8) || (a->dk_col % 8) || (a->dv_col % 8)) return MMSUM_ERR_INVALID;
  }
  return 0;
}

}  // namespace mmsum

using namespace mmsum;

extern "C" int mmsum_attn_fwd(const MmsumAttnArgs* a, void* stream_v) {
  if (int rc = validate(a, false)) return rc;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FwdSmem));
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  attn_fwd_kernel<<<a->n_qseq * a->H, 256, sizeof(FwdSmem), stream>>>(*a);
  MMSUM_CHECK_LAUNCH();
  return 0;
}

extern "C" int mmsum_attn_bwd(const MmsumAttnArgs* a, void* stream_v) {
  if (int rc = validate(a, true)) return rc;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_v);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BwdQSmem));
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BwdKVSmem));
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  attn_bwd_dq_kernel<<<a->n_qseq * a->H, 256, sizeof(BwdQSmem), stream>>>(*a);
  MMSUM_CHECK_LAUNCH();
  int per_bh = 0;
  for (int m = 0; m < a->n_mod; ++m) per_bh += a->mods[m].E * ((a->mods[m].Sk + KB - 1) / KB);
  const int n_biz = a->n_qseq / a->R;
  attn_bwd_dkv_kernel<<<n_biz * a->H * per_bh, 128, sizeof(BwdKVSmem), stream>>>(*a, per_bh);
  MMSUM_CHECK_LAUNCH();
  return 0;
}
