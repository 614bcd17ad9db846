// Causal flash attention forward / backward on tcgen05 (sm_100a), head_dim = 128.
// Oracle: F.scaled_dot_product_attention(q, k, v, is_causal=True) as called by HF
// LlamaAttention with _attn_implementation == "sdpa" (SURVEY.md §8 a7).
//
// Both kernels keep every accumulator in TMEM and give each of the 128 threads one TMEM lane
// (= one matrix row), so softmax row statistics never need a cross-thread reduction.
//
// forward   CTA = (128-query tile, head, sequence); loop over 64-key blocks:
//             S = Q K^T            (tcgen05, K-major x K-major)      -> TMEM
//             P = exp2(S - max)    (registers, fp32)                 -> smem bf16, SW128 K-major
//             O_blk = P V          (tcgen05, K-major x MN-major V)   -> TMEM -> registers (+=)
// backward  CTA = (128-key block, kv head, sequence); loop over 64-query blocks, transposed
//           formulation so that P^T / dS^T come out of TMEM already in the layout the next
//           MMAs want:
//             S^T  = K Q^T, dP^T = V dO^T                            -> TMEM
//             P^T  = exp2(S^T - lse), dS^T = P^T (dP^T - delta)*scale -> smem bf16
//             dV  += P^T dO,  dK += dS^T Q   (accumulate in TMEM over the whole loop)
//             dQ^T = K^T dS^T                -> TMEM -> fp32 red.global.add into dq32
#include "host_common.h"
#include "ops.h"
#include "ptx.cuh"

namespace b200w {

namespace {

using bf16 = __nv_bfloat16;
constexpr int DH = 128;
constexpr int ATOM64 = 64 * 128;    // bytes of a [64 rows x 128 B] swizzle-atom column
constexpr int ATOM128 = 128 * 128;  // bytes of a [128 rows x 128 B] one

__device__ __forceinline__ void require_1024_aligned(const void* p) {
  if (smem_u32(p) & 1023u) {
    if (threadIdx.x == 0) printf("b200w: dynamic shared memory base is not 1024-byte aligned\n");
    __trap();
  }
}

// ==========================================================================================
// forward
// ==========================================================================================
constexpr int FWD_BQ = 128, FWD_BKV = 64;
constexpr int FWD_SMEM = 2 * ATOM128 /*Q*/ + 2 * 2 * ATOM64 /*K x2*/ + 2 * 2 * ATOM64 /*V x2*/ +
                         ATOM128 /*P*/ + 256 /*barriers*/;
constexpr int FWD_TMEM_COLS = 256;  // S: [0,64)  O_blk: [64,192)

__global__ void __launch_bounds__(128, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, bf16* __restrict__ out, int ld_out,
                float* __restrict__ lse2, int k_off, int v_off, int B, int S, int H, int Hkv,
                float scale_log2) {
  extern __shared__ __align__(1024) uint8_t smem[];
  require_1024_aligned(smem);
  uint8_t* sQ = smem;                      // 2 atoms (dh halves) x [128 x 128 B]
  uint8_t* sK = sQ + 2 * ATOM128;          // 2 bufs x 2 atoms x [64 x 128 B]
  uint8_t* sV = sK + 2 * 2 * ATOM64;       // 2 bufs x 2 atoms x [64 kv rows x 128 B]
  uint8_t* sP = sV + 2 * 2 * ATOM64;       // [128 x 128 B]
  uint64_t* bar_q = reinterpret_cast<uint64_t*>(sP + ATOM128);
  uint64_t* bar_kv = bar_q + 1;  // [2]
  uint64_t* bar_s = bar_kv + 2;
  uint64_t* bar_o = bar_s + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_o + 1);

  const int nq = S / FWD_BQ;
  const int bh = blockIdx.x % (B * H);
  const int qi = nq - 1 - blockIdx.x / (B * H);  // longest (most key blocks) tiles first
  const int h = bh % H, b = bh / H;
  const int hk = h / (H / Hkv);
  const int tok0 = b * S;                // first token of this sequence
  const int q0 = qi * FWD_BQ;            // first query row inside the sequence
  const int njb = 2 * qi + 2;            // causal: key blocks [0, njb)
  const int tid = threadIdx.x, warp = tid >> 5;

  if (tid == 0) {
    tma_prefetch_desc(&tm_qkv);
    mbar_init(bar_q, 1);
    mbar_init(&bar_kv[0], 1);
    mbar_init(&bar_kv[1], 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, FWD_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base, tmem_O = tmem_base + 64;

  auto load_kv = [&](int j, int buf) {
    mbar_arrive_expect_tx(&bar_kv[buf], 4 * ATOM64);
    const int row = tok0 + j * FWD_BKV;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      tma_load_2d(sK + (buf * 2 + a) * ATOM64, &tm_qkv, &bar_kv[buf], k_off + hk * DH + a * 64, row);
      tma_load_2d(sV + (buf * 2 + a) * ATOM64, &tm_qkv, &bar_kv[buf], v_off + hk * DH + a * 64, row);
    }
  };

  if (tid == 0) {
    mbar_arrive_expect_tx(bar_q, 2 * ATOM128);
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int r = 0; r < 2; ++r)
        tma_load_2d(sQ + a * ATOM128 + r * ATOM64, &tm_qkv, bar_q, h * DH + a * 64,
                    tok0 + q0 + r * 64);
    load_kv(0, 0);
  }

  constexpr uint32_t idesc_s = make_idesc_bf16(128, FWD_BKV, false, false);
  constexpr uint32_t idesc_o = make_idesc_bf16(128, DH, false, true);

  float o[DH];
#pragma unroll
  for (int i = 0; i < DH; ++i) o[i] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;
  const int row_local = tid;             // TMEM lane == query row inside the tile
  const int row_seq = q0 + row_local;    // query position inside the sequence
  const uint32_t lane_base = (warp * 32u) << 16;

  for (int j = 0; j < njb; ++j) {
    const int buf = j & 1;
    if (tid == 0) {
      if (j + 1 < njb) load_kv(j + 1, buf ^ 1);  // buf^1 was released by bar_o of block j-1
      if (j == 0) mbar_wait(bar_q, 0);
      mbar_wait(&bar_kv[buf], (j >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < DH / 16; ++k) {
        const uint64_t da = make_smem_desc(smem_u32(sQ + (k / 4) * ATOM128) + (k % 4) * 32, 16, 1024);
        const uint64_t db =
            make_smem_desc(smem_u32(sK + (buf * 2 + k / 4) * ATOM64) + (k % 4) * 32, 16, 1024);
        tc_mma_bf16(tmem_S, da, db, idesc_s, k != 0);
      }
      tc_commit(bar_s);
    }
    mbar_wait(bar_s, j & 1);
    __syncwarp();
    tc_fence_after();

    uint32_t sr[2][32];
    tmem_ld32(tmem_S + lane_base, sr[0]);
    tmem_ld32(tmem_S + lane_base + 32, sr[1]);
    tmem_ld_wait();

    const int col0 = j * FWD_BKV;
    const bool diag = (col0 + FWD_BKV - 1) > q0;  // block reaches past the first row's diagonal
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < 64; ++c) {
      float s = __uint_as_float(sr[c >> 5][c & 31]) * scale_log2;
      if (diag && (col0 + c > row_seq)) s = -INFINITY;
      sr[c >> 5][c & 31] = __float_as_uint(s);
      mx = fmaxf(mx, s);
    }
    const float m_new = fmaxf(m_run, mx);  // finite from block 0 on (column 0 is always visible)
    const float alpha = exp2f(m_run - m_new);
    float psum = 0.f;
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) {
      float p[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int c = c8 * 8 + e;
        p[e] = exp2f(__uint_as_float(sr[c >> 5][c & 31]) - m_new);
        psum += p[e];
      }
      uint4 u;
      u.x = pack_bf16x2(p[0], p[1]); u.y = pack_bf16x2(p[2], p[3]);
      u.z = pack_bf16x2(p[4], p[5]); u.w = pack_bf16x2(p[6], p[7]);
      *reinterpret_cast<uint4*>(sP + sw128_offset(row_local, c8)) = u;
    }
    l_run = l_run * alpha + psum;
    m_run = m_new;

    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();

    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < FWD_BKV / 16; ++k) {
        const uint64_t da = make_smem_desc(smem_u32(sP) + k * 32, 16, 1024);
        // V as MN-major B: N = dh (2 atoms of 64, ATOM64 apart), K = kv rows (16 rows = 2048 B)
        const uint64_t db =
            make_smem_desc(smem_u32(sV + buf * 2 * ATOM64) + k * 2048, ATOM64, 1024);
        tc_mma_bf16(tmem_O, da, db, idesc_o, k != 0);
      }
      tc_commit(bar_o);
    }
#pragma unroll
    for (int i = 0; i < DH; ++i) o[i] *= alpha;
    mbar_wait(bar_o, j & 1);
    __syncwarp();
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < DH / 32; ++c) {
      uint32_t r[32];
      tmem_ld32(tmem_O + lane_base + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[c * 32 + i] += __uint_as_float(r[i]);
    }
    // The next block's S MMA / P stores / PV MMA are ordered behind this block's reads by the
    // post-softmax __syncthreads of the next iteration plus the bar_o wait above.
    tc_fence_before();
  }

  const float inv_l = 1.f / l_run;
  bf16* orow = out + static_cast<size_t>(tok0 + row_seq) * ld_out + h * DH;
#pragma unroll
  for (int c8 = 0; c8 < DH / 8; ++c8) {
    uint4 u;
    u.x = pack_bf16x2(o[c8 * 8 + 0] * inv_l, o[c8 * 8 + 1] * inv_l);
    u.y = pack_bf16x2(o[c8 * 8 + 2] * inv_l, o[c8 * 8 + 3] * inv_l);
    u.z = pack_bf16x2(o[c8 * 8 + 4] * inv_l, o[c8 * 8 + 5] * inv_l);
    u.w = pack_bf16x2(o[c8 * 8 + 6] * inv_l, o[c8 * 8 + 7] * inv_l);
    reinterpret_cast<uint4*>(orow)[c8] = u;
  }
  lse2[static_cast<size_t>(h) * (static_cast<size_t>(B) * S) + tok0 + row_seq] =
      m_run + log2f(l_run);

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, FWD_TMEM_COLS);
  }
}

// ==========================================================================================
// backward
// ==========================================================================================
constexpr int BWD_BKV = 128, BWD_BQ = 64;
constexpr int BWD_SMEM = 2 * ATOM128 /*K*/ + 2 * ATOM128 /*V*/ + 2 * 2 * ATOM64 /*Q x2*/ +
                         2 * 2 * ATOM64 /*dO x2*/ + ATOM128 /*P^T*/ + ATOM128 /*dS^T*/ +
                         2 * 2 * 64 * 4 /*lse, delta x2*/ + 256;
constexpr int BWD_TMEM_COLS = 512;  // S^T [0,64) dP^T [64,128) dV [128,256) dK [256,384) dQ^T [384,448)

__global__ void __launch_bounds__(128, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                const float* __restrict__ lse2, const float* __restrict__ delta,
                float* __restrict__ dq32, bf16* __restrict__ dqkv, int ld_qkv, int k_off, int v_off,
                int B, int S, int H, int Hkv, float scale, float scale_log2) {
  extern __shared__ __align__(1024) uint8_t smem[];
  require_1024_aligned(smem);
  uint8_t* sK = smem;                       // 2 atoms (dh halves) x [128 kv x 128 B]
  uint8_t* sV = sK + 2 * ATOM128;
  uint8_t* sQ = sV + 2 * ATOM128;           // 2 bufs x 2 atoms x [64 q x 128 B]
  uint8_t* sdO = sQ + 2 * 2 * ATOM64;
  uint8_t* sP = sdO + 2 * 2 * ATOM64;       // P^T  [128 kv x 64 q]
  uint8_t* sdS = sP + ATOM128;              // dS^T [128 kv x 64 q]
  float* sStat = reinterpret_cast<float*>(sdS + ATOM128);  // [2 bufs][lse 64 | delta 64]
  uint64_t* bar_kv = reinterpret_cast<uint64_t*>(sStat + 2 * 128);
  uint64_t* bar_q = bar_kv + 1;  // [2]
  uint64_t* bar_s = bar_q + 2;
  uint64_t* bar_dq = bar_s + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_dq + 1);

  const int nkb = S / BWD_BKV;
  const int G = H / Hkv;
  // earliest key blocks (longest query loops) first
  const int jb = blockIdx.x / (B * Hkv);
  const int bhk = blockIdx.x % (B * Hkv);
  const int hk = bhk % Hkv, b = bhk / Hkv;
  const int tok0 = b * S;
  const int kv0 = jb * BWD_BKV;
  const int nqb = S / BWD_BQ - 2 * jb;  // query blocks [2 jb, S/64) see this key block
  const int n_iter = G * nqb;
  const int tid = threadIdx.x, warp = tid >> 5;
  const size_t Ttot = static_cast<size_t>(B) * S;
  (void)nkb;

  if (tid == 0) {
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
    mbar_init(bar_kv, 1);
    mbar_init(&bar_q[0], 1);
    mbar_init(&bar_q[1], 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_dq, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, BWD_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_ST = tmem_base, tmem_dPT = tmem_base + 64, tmem_dV = tmem_base + 128,
                 tmem_dK = tmem_base + 256, tmem_dQT = tmem_base + 384;

  auto iter_head = [&](int it) { return hk * G + it / nqb; };
  auto iter_qrow = [&](int it) { return (2 * jb + it % nqb) * BWD_BQ; };  // inside the sequence

  auto load_q = [&](int it, int buf) {
    mbar_arrive_expect_tx(&bar_q[buf], 4 * ATOM64);
    const int h = iter_head(it), row = tok0 + iter_qrow(it);
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      tma_load_2d(sQ + (buf * 2 + a) * ATOM64, &tm_qkv, &bar_q[buf], h * DH + a * 64, row);
      tma_load_2d(sdO + (buf * 2 + a) * ATOM64, &tm_do, &bar_q[buf], h * DH + a * 64, row);
    }
  };
  auto load_stats = [&](int it, int buf) {  // all 128 threads
    const int h = iter_head(it);
    const size_t base = static_cast<size_t>(h) * Ttot + tok0 + iter_qrow(it);
    sStat[buf * 128 + tid] = (tid < 64) ? lse2[base + tid] : delta[base + tid - 64];
  };

  if (tid == 0) {
    mbar_arrive_expect_tx(bar_kv, 4 * ATOM128);
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        tma_load_2d(sK + a * ATOM128 + r * ATOM64, &tm_qkv, bar_kv, k_off + hk * DH + a * 64,
                    tok0 + kv0 + r * 64);
        tma_load_2d(sV + a * ATOM128 + r * ATOM64, &tm_qkv, bar_kv, v_off + hk * DH + a * 64,
                    tok0 + kv0 + r * 64);
      }
    load_q(0, 0);
  }
  load_stats(0, 0);
  __syncthreads();

  constexpr uint32_t idesc_st = make_idesc_bf16(128, BWD_BQ, false, false);  // S^T, dP^T
  constexpr uint32_t idesc_dv = make_idesc_bf16(128, DH, false, true);       // dV, dK
  constexpr uint32_t idesc_dq = make_idesc_bf16(128, BWD_BQ, true, true);    // dQ^T

  const int row_local = tid;                 // TMEM lane == key row inside the block
  const int kv_seq = kv0 + row_local;        // key position inside the sequence
  const uint32_t lane_base = (warp * 32u) << 16;

  for (int it = 0; it < n_iter; ++it) {
    const int buf = it & 1;
    const int q_seq0 = iter_qrow(it);
    const int h = iter_head(it);
    if (it + 1 < n_iter) load_stats(it + 1, buf ^ 1);  // visible after this iteration's barrier
    if (tid == 0) {
      if (it + 1 < n_iter) load_q(it + 1, buf ^ 1);  // buf^1 released by bar_dq of it-1
      if (it == 0) mbar_wait(bar_kv, 0);
      mbar_wait(&bar_q[buf], (it >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < DH / 16; ++k) {
        const uint32_t ko = (k % 4) * 32;
        const uint64_t dk = make_smem_desc(smem_u32(sK + (k / 4) * ATOM128) + ko, 16, 1024);
        const uint64_t dq = make_smem_desc(smem_u32(sQ + (buf * 2 + k / 4) * ATOM64) + ko, 16, 1024);
        tc_mma_bf16(tmem_ST, dk, dq, idesc_st, k != 0);
      }
#pragma unroll
      for (int k = 0; k < DH / 16; ++k) {
        const uint32_t ko = (k % 4) * 32;
        const uint64_t dv = make_smem_desc(smem_u32(sV + (k / 4) * ATOM128) + ko, 16, 1024);
        const uint64_t ddo = make_smem_desc(smem_u32(sdO + (buf * 2 + k / 4) * ATOM64) + ko, 16, 1024);
        tc_mma_bf16(tmem_dPT, dv, ddo, idesc_st, k != 0);
      }
      tc_commit(bar_s);
    }
    mbar_wait(bar_s, it & 1);
    __syncwarp();
    tc_fence_after();

    const float* st = sStat + buf * 128;
    const bool diag = (q_seq0 < kv0 + BWD_BKV);  // some (q, kv) pairs of this block are masked
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t s_r[32], dp_r[32];
      tmem_ld32(tmem_ST + lane_base + half * 32, s_r);
      tmem_ld32(tmem_dPT + lane_base + half * 32, dp_r);
      tmem_ld_wait();
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        float p[8], ds[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int c = half * 32 + c8 * 8 + e;  // query column inside the block
          float pv = exp2f(__uint_as_float(s_r[c8 * 8 + e]) * scale_log2 - st[c]);
          if (diag && (q_seq0 + c < kv_seq)) pv = 0.f;
          p[e] = pv;
          ds[e] = pv * (__uint_as_float(dp_r[c8 * 8 + e]) - st[64 + c]) * scale;
        }
        uint4 u, w;
        u.x = pack_bf16x2(p[0], p[1]); u.y = pack_bf16x2(p[2], p[3]);
        u.z = pack_bf16x2(p[4], p[5]); u.w = pack_bf16x2(p[6], p[7]);
        w.x = pack_bf16x2(ds[0], ds[1]); w.y = pack_bf16x2(ds[2], ds[3]);
        w.z = pack_bf16x2(ds[4], ds[5]); w.w = pack_bf16x2(ds[6], ds[7]);
        const uint32_t off = sw128_offset(row_local, half * 4 + c8);
        *reinterpret_cast<uint4*>(sP + off) = u;
        *reinterpret_cast<uint4*>(sdS + off) = w;
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();

    if (tid == 0) {
      tc_fence_after();
      // dV += P^T dO, dK += dS^T Q : A K-major [128 kv x 64 q], B MN-major (N = dh, K = q rows)
#pragma unroll
      for (int k = 0; k < BWD_BQ / 16; ++k) {
        const uint64_t dp = make_smem_desc(smem_u32(sP) + k * 32, 16, 1024);
        const uint64_t ddo = make_smem_desc(smem_u32(sdO + buf * 2 * ATOM64) + k * 2048, ATOM64, 1024);
        tc_mma_bf16(tmem_dV, dp, ddo, idesc_dv, (it | k) != 0);
      }
#pragma unroll
      for (int k = 0; k < BWD_BQ / 16; ++k) {
        const uint64_t dds = make_smem_desc(smem_u32(sdS) + k * 32, 16, 1024);
        const uint64_t dq = make_smem_desc(smem_u32(sQ + buf * 2 * ATOM64) + k * 2048, ATOM64, 1024);
        tc_mma_bf16(tmem_dK, dds, dq, idesc_dv, (it | k) != 0);
      }
      // dQ^T = K^T dS^T : A MN-major (M = dh: 2 atoms ATOM128 apart, K = kv rows),
      //                   B MN-major (N = q: one atom, K = kv rows)
#pragma unroll
      for (int k = 0; k < BWD_BKV / 16; ++k) {
        const uint64_t dk = make_smem_desc(smem_u32(sK) + k * 2048, ATOM128, 1024);
        const uint64_t dds = make_smem_desc(smem_u32(sdS) + k * 2048, ATOM128, 1024);
        tc_mma_bf16(tmem_dQT, dk, dds, idesc_dq, k != 0);
      }
      tc_commit(bar_dq);
    }
    mbar_wait(bar_dq, it & 1);
    __syncwarp();
    tc_fence_after();
    // lane = dh index, column = query: coalesced fp32 reductions into dq32[T, H*128]
    float* dq_col = dq32 + (static_cast<size_t>(tok0 + q_seq0)) * (static_cast<size_t>(H) * DH) +
                    h * DH + row_local;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t r[32];
      tmem_ld32(tmem_dQT + lane_base + half * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 32; ++c)
        atomicAdd(dq_col + static_cast<size_t>(half * 32 + c) * (static_cast<size_t>(H) * DH),
                  __uint_as_float(r[c]));
    }
    tc_fence_before();
  }

  // dV, dK: lane = key row, 128 dh columns each
  bf16* dvrow = dqkv + static_cast<size_t>(tok0 + kv_seq) * ld_qkv + v_off + hk * DH;
  bf16* dkrow = dqkv + static_cast<size_t>(tok0 + kv_seq) * ld_qkv + k_off + hk * DH;
#pragma unroll
  for (int c = 0; c < DH / 32; ++c) {
    uint32_t r[32];
    tmem_ld32(tmem_dV + lane_base + c * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 u;
      u.x = pack_bf16x2(__uint_as_float(r[i * 8 + 0]), __uint_as_float(r[i * 8 + 1]));
      u.y = pack_bf16x2(__uint_as_float(r[i * 8 + 2]), __uint_as_float(r[i * 8 + 3]));
      u.z = pack_bf16x2(__uint_as_float(r[i * 8 + 4]), __uint_as_float(r[i * 8 + 5]));
      u.w = pack_bf16x2(__uint_as_float(r[i * 8 + 6]), __uint_as_float(r[i * 8 + 7]));
      reinterpret_cast<uint4*>(dvrow + c * 32)[i] = u;
    }
    tmem_ld32(tmem_dK + lane_base + c * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 u;
      u.x = pack_bf16x2(__uint_as_float(r[i * 8 + 0]), __uint_as_float(r[i * 8 + 1]));
      u.y = pack_bf16x2(__uint_as_float(r[i * 8 + 2]), __uint_as_float(r[i * 8 + 3]));
      u.z = pack_bf16x2(__uint_as_float(r[i * 8 + 4]), __uint_as_float(r[i * 8 + 5]));
      u.w = pack_bf16x2(__uint_as_float(r[i * 8 + 6]), __uint_as_float(r[i * 8 + 7]));
      reinterpret_cast<uint4*>(dkrow + c * 32)[i] = u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BWD_TMEM_COLS);
  }
}

}  // namespace

void attention_fwd(const void* qkv, int ld_qkv, int k_off, int v_off, void* out, int ld_out,
                   float* lse2, int B, int S, int H, int Hkv, float scale, cudaStream_t s) {
  B200W_CHECK(S % 128 == 0, "sequence length must be a multiple of 128");
  B200W_CHECK(H % Hkv == 0 && ld_out % 8 == 0, "bad head configuration");
  const size_t T = static_cast<size_t>(B) * S;
  CUtensorMap tm = make_tmap_bf16_2d(qkv, T, ld_qkv, ld_qkv, 64, 64);
  static bool attr = false;
  if (!attr) {
    B200W_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    FWD_SMEM));
    attr = true;
  }
  const int grid = (S / FWD_BQ) * B * H;
  const float scale_log2 = scale * 1.4426950408889634f;
  attn_fwd_kernel<<<grid, 128, FWD_SMEM, s>>>(tm, static_cast<bf16*>(out), ld_out, lse2, k_off,
                                              v_off, B, S, H, Hkv, scale_log2);
  B200W_CUDA(cudaGetLastError());
}

void attention_bwd(const void* qkv, int ld_qkv, int k_off, int v_off, const void* out,
                   const void* dout, int ld_out, const float* lse2, float* delta, float* dq32,
                   void* dqkv, int B, int S, int H, int Hkv, float scale, cudaStream_t s) {
  B200W_CHECK(S % 128 == 0, "sequence length must be a multiple of 128");
  B200W_CHECK(H % Hkv == 0, "bad head configuration");
  const size_t T = static_cast<size_t>(B) * S;
  attn_bwd_delta(out, dout, ld_out, delta, static_cast<int>(T), H, s);
  CUtensorMap tm_qkv = make_tmap_bf16_2d(qkv, T, ld_qkv, ld_qkv, 64, 64);
  CUtensorMap tm_do = make_tmap_bf16_2d(dout, T, ld_out, ld_out, 64, 64);
  static bool attr = false;
  if (!attr) {
    B200W_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    BWD_SMEM));
    attr = true;
  }
  const int grid = (S / BWD_BKV) * B * Hkv;
  const float scale_log2 = scale * 1.4426950408889634f;
  attn_bwd_kernel<<<grid, 128, BWD_SMEM, s>>>(tm_qkv, tm_do, lse2, delta, dq32,
                                              static_cast<bf16*>(dqkv), ld_qkv, k_off, v_off, B, S,
                                              H, Hkv, scale, scale_log2);
  B200W_CUDA(cudaGetLastError());
}

}  // namespace b200w
