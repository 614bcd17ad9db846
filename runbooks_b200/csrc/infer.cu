// Batched greedy decode engine for the Server path: one new token per cache
// slot per step — prompt ingestion and generation are the same step, which is what a
// continuous-batching server wants. Families:
//   FALCON  (HF models/falcon/modeling_falcon.py, falcon-7b layout: multi_query, parallel_attn,
//            one input_layernorm per layer, bias-free linears, LayerNorm with bias, exact GeLU,
//            rotate_half RoPE, lm_head tied to word_embeddings)
//   LLAMA   (HF models/llama/modeling_llama.py: RMSNorm, SwiGLU, sequential residual, GQA)
// Decode at batch 32 is HBM-bound on the weights (13.84 GB per step for Falcon-7B):
// the projections run on the same tcgen05 GEMM as training (M = batch rows, TMA zero-fills the
// rest of the 128-row tile without fetching it), attention over the KV cache is a CUDA-core kernel
// that reads each K/V row once per group of 8 query heads (MQA/GQA aware).
#include <math.h>
#include <string.h>

#include <algorithm>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/b200w.h"
#include "ctx_access.h"
#include "host_common.h"
#include "ops.h"
#include "ptx.cuh"

using namespace b200w;
using bf16 = __nv_bfloat16;

namespace {

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ float block_reduce(float v, float* red, bool is_max) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = is_max ? warp_max_f(v) : warp_sum_f(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : (is_max ? -INFINITY : 0.f);
  return is_max ? warp_max_f(r) : warp_sum_f(r);
}

// LayerNorm with bias (oracle: torch.nn.LayerNorm as used by FalconDecoderLayer), fp32 statistics.
// One block per row, the row is read from HBM/L2 once and kept in registers (d <= 256*8*4).
constexpr int LN_MAXP = 4;
__global__ void __launch_bounds__(256)
layernorm_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, const bf16* __restrict__ b,
                 bf16* __restrict__ y, int d, float eps) {
  __shared__ float red[32];
  const bf16* xr = x + static_cast<size_t>(blockIdx.x) * d;
  float v[LN_MAXP][8];
  float s = 0.f;
#pragma unroll
  for (int p = 0; p < LN_MAXP; ++p) {
    const int c = (p * 256 + threadIdx.x) * 8;
    if (c < d) {
      const uint4 u = *reinterpret_cast<const uint4*>(xr + c);
      const uint32_t wv[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16x2(wv[j]);
        v[p][2 * j] = f.x;
        v[p][2 * j + 1] = f.y;
        s += f.x + f.y;
      }
    }
  }
  const float mean = block_reduce(s, red, false) / d;
  float q = 0.f;
#pragma unroll
  for (int p = 0; p < LN_MAXP; ++p) {
    const int c = (p * 256 + threadIdx.x) * 8;
    if (c < d) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float t = v[p][j] - mean;
        q += t * t;
      }
    }
  }
  const float rstd = rsqrtf(block_reduce(q, red, false) / d + eps);
  bf16* yr = y + static_cast<size_t>(blockIdx.x) * d;
#pragma unroll
  for (int p = 0; p < LN_MAXP; ++p) {
    const int c = (p * 256 + threadIdx.x) * 8;
    if (c < d) {
      const uint4 uw = *reinterpret_cast<const uint4*>(w + c);
      const uint4 ub = *reinterpret_cast<const uint4*>(b + c);
      const uint32_t ww[4] = {uw.x, uw.y, uw.z, uw.w}, bb[4] = {ub.x, ub.y, ub.z, ub.w};
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 fw = unpack_bf16x2(ww[j]), fb = unpack_bf16x2(bb[j]);
        o[j] = pack_bf16x2((v[p][2 * j] - mean) * rstd * fw.x + fb.x, (v[p][2 * j + 1] - mean) * rstd * fw.y + fb.y);
      }
      *reinterpret_cast<uint4*>(yr + c) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// exact (erf) GeLU in place — transformers get_activation("gelu")
__global__ void gelu_kernel(bf16* x, size_t n) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float v = __bfloat162float(x[i]);
    x[i] = __float2bfloat16_rn(0.5f * v * (1.f + erff(v * 0.70710678118654752f)));
  }
}

// rotate_half RoPE on the q heads (in place) and on k; k and v are written into the cache at
// [slot][pos]. One thread per (row, head, pair index).
__global__ void rope_append_kernel(bf16* __restrict__ qkv, int ld, const float* __restrict__ inv_freq,
                                   const int32_t* __restrict__ pos, const int32_t* __restrict__ slot,
                                   bf16* __restrict__ kcache, bf16* __restrict__ vcache, int n, int H,
                                   int Hkv, int dh, int max_ctx) {
  const int half = dh / 2;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(n) * (H + Hkv) * half;
  if (idx >= total) return;
  const int i = static_cast<int>(idx % half);
  const int h = static_cast<int>((idx / half) % (H + Hkv));
  const int r = static_cast<int>(idx / (static_cast<long long>(half) * (H + Hkv)));
  const int p = pos[r];
  const float ang = static_cast<float>(p) * inv_freq[i];
  float sn, cs;
  sincosf(ang, &sn, &cs);
  bf16* src = qkv + static_cast<size_t>(r) * ld + h * dh;
  const float x1 = __bfloat162float(src[i]), x2 = __bfloat162float(src[i + half]);
  const bf16 o1 = __float2bfloat16_rn(x1 * cs - x2 * sn), o2 = __float2bfloat16_rn(x2 * cs + x1 * sn);
  if (h < H) {
    src[i] = o1;
    src[i + half] = o2;
  } else {
    const int hk = h - H;
    const size_t off = ((static_cast<size_t>(slot[r]) * max_ctx + p) * Hkv + hk) * dh;
    kcache[off + i] = o1;
    kcache[off + i + half] = o2;
    const bf16* vsrc = qkv + static_cast<size_t>(r) * ld + (H + Hkv + hk) * dh;
    vcache[off + i] = vsrc[i];
    vcache[off + i + half] = vsrc[i + half];
  }
}

// Attention of one new query token per row over its slot's cache [0, pos]. Block = (row, kv head,
// group of GT query heads that share that kv head): every K/V row is read once per block.
constexpr int ATT_GT = 8;
constexpr int ATT_THREADS = 256;
template <int DH>
__global__ void __launch_bounds__(ATT_THREADS)
decode_attn_kernel(const bf16* __restrict__ qkv, int ld, const bf16* __restrict__ kcache,
                   const bf16* __restrict__ vcache, const int32_t* __restrict__ pos,
                   const int32_t* __restrict__ slot, bf16* __restrict__ out, int H, int Hkv,
                   int max_ctx, int sc_stride, float scale) {
  extern __shared__ float sm[];
  const int r = blockIdx.x, hk = blockIdx.y, G = H / Hkv;
  const int g0 = blockIdx.z * ATT_GT;
  const int ng = min(ATT_GT, G - g0);
  const int len = pos[r] + 1;
  float* sq = sm;                       // [ATT_GT][DH]
  float* sc = sq + ATT_GT * DH;         // [ATT_GT][sc_stride]; reused to combine output slices
  float* red = sc + ATT_GT * sc_stride; // [32]
  const int tid = threadIdx.x;
  for (int i = tid; i < ng * DH; i += ATT_THREADS) {
    const int g = i / DH, c = i % DH;
    sq[g * DH + c] = __bfloat162float(qkv[static_cast<size_t>(r) * ld + (hk * G + g0 + g) * DH + c]) * scale;
  }
  __syncthreads();
  const size_t cbase = static_cast<size_t>(slot[r]) * max_ctx;
  // scores: one key position per thread, the K row lives in registers for all ng heads
  for (int t = tid; t < len; t += ATT_THREADS) {
    const uint4* krow = reinterpret_cast<const uint4*>(kcache + ((cbase + t) * Hkv + hk) * DH);
    float acc[ATT_GT];
#pragma unroll
    for (int g = 0; g < ATT_GT; ++g) acc[g] = 0.f;
#pragma unroll
    for (int c8 = 0; c8 < DH / 8; ++c8) {
      const uint4 u = krow[c8];
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 kf = unpack_bf16x2(w[j]);
#pragma unroll
        for (int g = 0; g < ATT_GT; ++g)
          acc[g] += kf.x * sq[g * DH + c8 * 8 + 2 * j] + kf.y * sq[g * DH + c8 * 8 + 2 * j + 1];
      }
    }
#pragma unroll
    for (int g = 0; g < ATT_GT; ++g)
      if (g < ng) sc[g * sc_stride + t] = acc[g];
  }
  __syncthreads();
  // softmax per head
  float inv_sum[ATT_GT];
  for (int g = 0; g < ng; ++g) {
    float m = -INFINITY;
    for (int t = tid; t < len; t += ATT_THREADS) m = fmaxf(m, sc[g * sc_stride + t]);
    m = block_reduce(m, red, true);
    float s = 0.f;
    for (int t = tid; t < len; t += ATT_THREADS) {
      const float e = __expf(sc[g * sc_stride + t] - m);
      sc[g * sc_stride + t] = e;
      s += e;
    }
    inv_sum[g] = 1.f / block_reduce(s, red, false);
  }
  __syncthreads();
  // output: thread = (8 dims via one 16-byte load, position slice); 4 independent positions in
  // flight per thread; V read once for all ng heads
  constexpr int DG = DH / 8;                 // 16-byte groups per row
  constexpr int SLICES = ATT_THREADS / DG;   // 32 (dh 64) or 16 (dh 128) position slices
  const int dg = tid % DG, sl = tid / DG;
  float o[ATT_GT][8];
#pragma unroll
  for (int g = 0; g < ATT_GT; ++g)
#pragma unroll
    for (int e = 0; e < 8; ++e) o[g][e] = 0.f;
  for (int t0 = sl; t0 < len; t0 += 4 * SLICES) {
    uint4 vv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int t = t0 + u * SLICES;
      vv[u] = t < len ? *reinterpret_cast<const uint4*>(vcache + ((cbase + t) * Hkv + hk) * DH + dg * 8)
                      : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int t = t0 + u * SLICES;
      if (t < len) {
        const uint32_t w[4] = {vv[u].x, vv[u].y, vv[u].z, vv[u].w};
        float vf[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack_bf16x2(w[j]);
          vf[2 * j] = f.x;
          vf[2 * j + 1] = f.y;
        }
#pragma unroll
        for (int g = 0; g < ATT_GT; ++g) {
          const float p = sc[g * sc_stride + t];
#pragma unroll
          for (int e = 0; e < 8; ++e) o[g][e] += p * vf[e];
        }
      }
    }
  }
  __syncthreads();  // scores are dead: reuse their space to combine the slices
  float* so = sc;   // [SLICES][ATT_GT][DH]  (sc_stride >= SLICES * DH / ... checked on the host)
#pragma unroll
  for (int g = 0; g < ATT_GT; ++g)
    if (g < ng) {
#pragma unroll
      for (int e = 0; e < 8; ++e) so[(sl * ATT_GT + g) * DH + dg * 8 + e] = o[g][e];
    }
  __syncthreads();
  for (int i = tid; i < ng * DH; i += ATT_THREADS) {
    const int g = i / DH, dim = i % DH;
    float a = 0.f;
    for (int s2 = 0; s2 < SLICES; ++s2) a += so[(s2 * ATT_GT + g) * DH + dim];
    out[static_cast<size_t>(r) * (H * DH) + (hk * G + g0 + g) * DH + dim] = __float2bfloat16_rn(a * inv_sum[g]);
  }
}

// greedy token: first index of the row maximum (torch.argmax tie-breaking)
__global__ void argmax_kernel(const bf16* __restrict__ logits, int V, int32_t* __restrict__ out) {
  __shared__ float sv[32];
  __shared__ int si[32];
  const bf16* row = logits + static_cast<size_t>(blockIdx.x) * V;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    const float v = __bfloat162float(row[i]);
    if (v > best) { best = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sv[warp] = best; si[warp] = bi; }
  __syncthreads();
  if (warp == 0) {
    best = lane < (blockDim.x >> 5) ? sv[lane] : -INFINITY;
    bi = lane < (blockDim.x >> 5) ? si[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) out[blockIdx.x] = bi;
  }
}

struct IParam {
  std::string name;
  int64_t rows, cols;
  size_t off;
};

struct Infer {
  b200w_infer_arch a{};
  int max_batch = 0;
  std::vector<IParam> params;
  std::unordered_map<std::string, int> index;
  size_t n_elems = 0;
  bf16* w = nullptr;
  struct L { size_t ln1_w, ln1_b, ln2_w, wqkv, wo, w1, w2; };
  std::vector<L> lp;
  size_t p_embed = 0, p_lnf_w = 0, p_lnf_b = 0, p_lm = 0;
  bf16 *h = nullptr, *h2 = nullptr, *nrm = nullptr, *qkv = nullptr, *att = nullptr, *mid = nullptr,
       *act = nullptr, *logits = nullptr;
  bf16 *kc = nullptr, *vc = nullptr;  // [L][max_batch][max_ctx][Hkv*dh]
  float* inv_freq = nullptr;
  float* ws = nullptr;          // split-K workspace [max_batch, max N] (kept zeroed)
  unsigned* counters = nullptr;
  int32_t *tok = nullptr, *pos = nullptr, *slot = nullptr, *next = nullptr;
  int32_t* pin = nullptr;  // pinned host staging: tok | pos | slot | next, max_batch each
  std::unordered_map<int, cudaGraphExec_t> graphs;  // whole decode step per row count
  std::unordered_map<int, int> warm;
  std::vector<void*> allocs;
  int64_t bytes = 0;
  template <typename T>
  T* alloc(size_t n) {
    void* p = nullptr;
    const size_t b = ((n * sizeof(T) + 255) / 256) * 256;
    if (cudaMalloc(&p, b) != cudaSuccess) {
      cudaGetLastError();
      throw std::bad_alloc();
    }
    allocs.push_back(p);
    bytes += b;
    return static_cast<T*>(p);
  }
  ~Infer() {
    for (auto& g : graphs) cudaGraphExecDestroy(g.second);
    if (pin) cudaFreeHost(pin);
    for (void* p : allocs) cudaFree(p);
  }
};

void add(Infer* m, const std::string& name, int64_t r, int64_t c, size_t* off) {
  *off = m->n_elems;
  m->index[name] = static_cast<int>(m->params.size());
  m->params.push_back({name, r, c, m->n_elems});
  m->n_elems += static_cast<size_t>(r) * c;
}

void build(Infer* m) {
  const auto& a = m->a;
  const int d = a.hidden_size, f = a.intermediate_size, L = a.num_layers;
  const int qd = a.num_heads * a.head_dim, kd = a.num_kv_heads * a.head_dim;
  m->lp.resize(L);
  size_t dummy;
  if (a.family == B200W_FAMILY_FALCON) {
    add(m, "transformer.word_embeddings.weight", a.vocab_size, d, &m->p_embed);
    for (int l = 0; l < L; ++l) {
      const std::string p = "transformer.h." + std::to_string(l) + ".";
      add(m, p + "input_layernorm.weight", 1, d, &m->lp[l].ln1_w);
      add(m, p + "input_layernorm.bias", 1, d, &m->lp[l].ln1_b);
      add(m, p + "self_attention.query_key_value.weight", qd + 2 * kd, d, &m->lp[l].wqkv);
      add(m, p + "self_attention.dense.weight", d, qd, &m->lp[l].wo);
      add(m, p + "mlp.dense_h_to_4h.weight", f, d, &m->lp[l].w1);
      add(m, p + "mlp.dense_4h_to_h.weight", d, f, &m->lp[l].w2);
    }
    add(m, "transformer.ln_f.weight", 1, d, &m->p_lnf_w);
    add(m, "transformer.ln_f.bias", 1, d, &m->p_lnf_b);
  } else {
    add(m, "model.embed_tokens.weight", a.vocab_size, d, &m->p_embed);
    for (int l = 0; l < L; ++l) {
      const std::string p = "model.layers." + std::to_string(l) + ".";
      add(m, p + "input_layernorm.weight", 1, d, &m->lp[l].ln1_w);
      add(m, p + "post_attention_layernorm.weight", 1, d, &m->lp[l].ln2_w);
      add(m, p + "self_attn.q_proj.weight", qd, d, &m->lp[l].wqkv);
      add(m, p + "self_attn.k_proj.weight", kd, d, &dummy);
      add(m, p + "self_attn.v_proj.weight", kd, d, &dummy);
      add(m, p + "self_attn.o_proj.weight", d, qd, &m->lp[l].wo);
      add(m, p + "mlp.gate_proj.weight", f, d, &m->lp[l].w1);
      add(m, p + "mlp.up_proj.weight", f, d, &dummy);
      add(m, p + "mlp.down_proj.weight", d, f, &m->lp[l].w2);
    }
    add(m, "model.norm.weight", 1, d, &m->p_lnf_w);
  }
  if (a.tie_embeddings) m->p_lm = m->p_embed;
  else add(m, "lm_head.weight", a.vocab_size, d, &m->p_lm);
}

template <typename F>
int iguard(b200w_ctx* ctx, F&& f) {
  if (!ctx) return B200W_ERR_INVALID;
  try {
    B200W_CUDA(cudaSetDevice(ctx_device(ctx)));
    f();
    return B200W_OK;
  } catch (const std::bad_alloc&) {
    ctx_set_error(ctx, "device memory exhausted");
    return B200W_ERR_OOM;
  } catch (const Error& e) {
    ctx_set_error(ctx, e.what());
    return std::string(e.what()).rfind("check failed", 0) == 0 ? B200W_ERR_INVALID : B200W_ERR_CUDA;
  } catch (const std::exception& e) {
    ctx_set_error(ctx, e.what());
    return B200W_ERR_INVALID;
  }
}

Infer* model(b200w_ctx* ctx) {
  Infer* m = static_cast<Infer*>(ctx_infer_slot(ctx));
  B200W_CHECK(m != nullptr, "inference model not initialised");
  return m;
}

void infer_destroy(void* p) { delete static_cast<Infer*>(p); }

}  // namespace

extern "C" {

int b200w_infer_init(b200w_ctx* ctx, const b200w_infer_arch* arch, int max_batch) {
  return iguard(ctx, [&] {
    B200W_CHECK(arch != nullptr && max_batch >= 1 && max_batch <= 128, "bad arch / max_batch (1..128)");
    B200W_CHECK(ctx_infer_slot(ctx) == nullptr, "inference model already initialised");
    B200W_CHECK(arch->family == B200W_FAMILY_LLAMA || arch->family == B200W_FAMILY_FALCON, "unknown family");
    B200W_CHECK(arch->head_dim == 64 || arch->head_dim == 128, "head_dim must be 64 or 128");
    B200W_CHECK(arch->num_heads % arch->num_kv_heads == 0, "heads must be a multiple of kv heads");
    B200W_CHECK(arch->hidden_size % 8 == 0 && arch->intermediate_size % 8 == 0 && arch->vocab_size % 8 == 0,
                "sizes must be multiples of 8");
    B200W_CHECK(arch->hidden_size <= 256 * 8 * LN_MAXP, "hidden_size too large for the decode LayerNorm");
    B200W_CHECK(arch->max_ctx >= 1 && arch->max_ctx <= 8192, "max_ctx must be in 1..8192");
    auto m = std::make_unique<Infer>();
    m->a = *arch;
    m->max_batch = max_batch;
    build(m.get());
    const auto& a = m->a;
    const size_t B = max_batch, d = a.hidden_size, f = a.intermediate_size;
    const size_t qd = a.num_heads * a.head_dim, kd = a.num_kv_heads * a.head_dim;
    m->w = m->alloc<bf16>(m->n_elems);
    m->h = m->alloc<bf16>(B * d);
    m->h2 = m->alloc<bf16>(B * d);
    m->nrm = m->alloc<bf16>(B * d);
    m->qkv = m->alloc<bf16>(B * (qd + 2 * kd));
    m->att = m->alloc<bf16>(B * qd);
    const size_t fmid = a.family == B200W_FAMILY_LLAMA ? 2 * f : f;
    m->mid = m->alloc<bf16>(B * fmid);
    m->act = m->alloc<bf16>(B * f);
    m->logits = m->alloc<bf16>(B * a.vocab_size);
    const size_t cache = static_cast<size_t>(a.num_layers) * B * a.max_ctx * kd;
    m->kc = m->alloc<bf16>(cache);
    m->vc = m->alloc<bf16>(cache);
    m->tok = m->alloc<int32_t>(B);
    m->pos = m->alloc<int32_t>(B);
    m->slot = m->alloc<int32_t>(B);
    m->next = m->alloc<int32_t>(B);
    B200W_CUDA(cudaMallocHost(reinterpret_cast<void**>(&m->pin), 4 * B * sizeof(int32_t)));
    m->inv_freq = m->alloc<float>(a.head_dim / 2);
    const size_t max_n = std::max<size_t>({static_cast<size_t>(a.vocab_size), fmid, qd + 2 * kd, d});
    m->ws = m->alloc<float>(B * max_n);
    m->counters = m->alloc<unsigned>((max_n + 127) / 128);
    B200W_CUDA(cudaMemset(m->ws, 0, B * max_n * sizeof(float)));
    B200W_CUDA(cudaMemset(m->counters, 0, ((max_n + 127) / 128) * sizeof(unsigned)));
    std::vector<float> inv(a.head_dim / 2);
    for (int i = 0; i < a.head_dim / 2; ++i)
      inv[i] = static_cast<float>(1.0 / pow(static_cast<double>(a.rope_theta), 2.0 * i / a.head_dim));
    B200W_CUDA(cudaMemcpy(m->inv_freq, inv.data(), inv.size() * 4, cudaMemcpyHostToDevice));
    ctx_set_infer(ctx, m.release(), infer_destroy);
  });
}

int b200w_infer_param_count(b200w_ctx* ctx, int64_t* n_tensors, int64_t* n_elements) {
  return iguard(ctx, [&] {
    Infer* m = model(ctx);
    if (n_tensors) *n_tensors = static_cast<int64_t>(m->params.size());
    if (n_elements) *n_elements = static_cast<int64_t>(m->n_elems);
  });
}

int b200w_infer_param_info(b200w_ctx* ctx, int64_t index, char* name, size_t name_cap, int64_t* rows,
                           int64_t* cols) {
  return iguard(ctx, [&] {
    Infer* m = model(ctx);
    B200W_CHECK(index >= 0 && index < (int64_t)m->params.size(), "bad index");
    const IParam& p = m->params[index];
    if (name && name_cap) {
      strncpy(name, p.name.c_str(), name_cap - 1);
      name[name_cap - 1] = 0;
    }
    if (rows) *rows = p.rows;
    if (cols) *cols = p.cols;
  });
}

int b200w_infer_load_tensor(b200w_ctx* ctx, const char* name, const void* host, b200w_dtype dtype,
                            int64_t n_elements) {
  return iguard(ctx, [&] {
    Infer* m = model(ctx);
    B200W_CHECK(name && host, "NULL name / buffer");
    auto it = m->index.find(name);
    if (it == m->index.end()) throw Error(std::string("check failed: unknown parameter ") + name);
    const IParam& p = m->params[it->second];
    B200W_CHECK(n_elements == p.rows * p.cols, "element count does not match the parameter shape");
    cudaStream_t s = ctx_stream(ctx);
    const size_t n = static_cast<size_t>(n_elements);
    if (dtype == B200W_BF16) {
      B200W_CUDA(cudaMemcpyAsync(m->w + p.off, host, n * 2, cudaMemcpyHostToDevice, s));
      B200W_CUDA(cudaStreamSynchronize(s));
    } else if (dtype == B200W_F32) {
      void* tmp = nullptr;
      B200W_CUDA(cudaMalloc(&tmp, n * 4));
      B200W_CUDA(cudaMemcpyAsync(tmp, host, n * 4, cudaMemcpyHostToDevice, s));
      cast_f32_to_bf16(static_cast<float*>(tmp), m->w + p.off, n, s);
      B200W_CUDA(cudaStreamSynchronize(s));
      cudaFree(tmp);
    } else {
      throw Error("check failed: dtype must be bf16 or f32");
    }
  });
}

int b200w_infer_init_random(b200w_ctx* ctx, uint64_t seed, float std) {
  return iguard(ctx, [&] {
    Infer* m = model(ctx);
    // host-side generation is fine for a one-off benchmark initialisation, but 7B elements would
    // take minutes: fill on the device through the training engine's generator instead
    ctx_fill_normal(ctx, m->w, m->n_elems, seed, std);
    for (const IParam& p : m->params) {
      const bool norm_w = p.name.find("layernorm.weight") != std::string::npos ||
                          p.name.find("ln_f.weight") != std::string::npos ||
                          p.name.find("norm.weight") != std::string::npos;
      const bool norm_b = p.name.find("layernorm.bias") != std::string::npos ||
                          p.name.find("ln_f.bias") != std::string::npos;
      if (norm_w || norm_b) ctx_fill_const(ctx, m->w + p.off, static_cast<size_t>(p.rows * p.cols), norm_w ? 1.f : 0.f);
    }
    B200W_CUDA(cudaStreamSynchronize(ctx_stream(ctx)));
  });
}

int b200w_infer_step(b200w_ctx* ctx, const int32_t* tokens, const int32_t* positions,
                     const int32_t* slots, int n, int32_t* next_tokens, float* logits_out) {
  return iguard(ctx, [&] {
    Infer* m = model(ctx);
    const auto& a = m->a;
    B200W_CHECK(tokens && positions && slots && n >= 1 && n <= m->max_batch, "bad batch");
    for (int i = 0; i < n; ++i) {
      B200W_CHECK(positions[i] >= 0 && positions[i] < a.max_ctx, "position outside the KV cache");
      B200W_CHECK(slots[i] >= 0 && slots[i] < m->max_batch, "bad cache slot");
      B200W_CHECK(tokens[i] >= 0 && tokens[i] < a.vocab_size, "token id outside the vocabulary");
    }
    cudaStream_t s = ctx_stream(ctx);
    int64_t& nl = ctx_launches(ctx);
    const int B = m->max_batch;
    memcpy(m->pin, tokens, n * 4);
    memcpy(m->pin + B, positions, n * 4);
    memcpy(m->pin + 2 * B, slots, n * 4);
    const int d = a.hidden_size, f = a.intermediate_size, H = a.num_heads, Hkv = a.num_kv_heads,
              dh = a.head_dim, V = a.vocab_size;
    const int qd = H * dh, kd = Hkv * dh, qkvd = qd + 2 * kd;
    const float scale = 1.f / sqrtf(static_cast<float>(dh));
    const bool falcon = a.family == B200W_FAMILY_FALCON;
    const size_t layer_cache = static_cast<size_t>(m->max_batch) * a.max_ctx * kd;
    const int G = H / Hkv;
    const dim3 agrid(n, Hkv, (G + ATT_GT - 1) / ATT_GT);
    const int slices = ATT_THREADS / (dh / 8);   // staging area: slices * ATT_GT * dh floats
    const int sc_stride = std::max(a.max_ctx, slices * dh);
    const size_t att_smem = (static_cast<size_t>(ATT_GT) * dh + static_cast<size_t>(ATT_GT) * sc_stride + 32) * 4;
    static PerDeviceOnce once;
    once.run([&] {
      B200W_CUDA(cudaFuncSetAttribute(decode_attn_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      B200W_CUDA(cudaFuncSetAttribute(decode_attn_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    });
    B200W_CHECK(att_smem <= 200 * 1024, "max_ctx too large for the decode attention kernel");

    int64_t step_launches = 0;
    // Everything from the H2D of the three index vectors to the D2H of the argmax, on stream s.
    // Run eagerly the first time a row count is seen (first-use attribute calls), captured into a
    // CUDA graph the second time, replayed from then on: ~330 launches become one.
    auto enqueue = [&]() {
    int64_t& nl = step_launches;
    B200W_CUDA(cudaMemcpyAsync(m->tok, m->pin, n * 4, cudaMemcpyHostToDevice, s));
    B200W_CUDA(cudaMemcpyAsync(m->pos, m->pin + B, n * 4, cudaMemcpyHostToDevice, s));
    B200W_CUDA(cudaMemcpyAsync(m->slot, m->pin + 2 * B, n * 4, cudaMemcpyHostToDevice, s));
    bf16* h = m->h;
    bf16* h2 = m->h2;
    embed_fwd(m->tok, m->w + m->p_embed, nullptr, h, n, d, V, 1, 0, s); ++nl;
    auto gemm = [&](const bf16* A, int K, size_t woff, int N, bf16* D, const bf16* C, int act = 0) {
      gemm_decode(A, m->w + woff, D, C, m->ws, m->counters, n, N, K, N, act, s); ++nl;
    };
    for (int l = 0; l < a.num_layers; ++l) {
      const auto& p = m->lp[l];
      bf16* kc = m->kc + l * layer_cache;
      bf16* vc = m->vc + l * layer_cache;
      if (falcon) layernorm_kernel<<<n, 256, 0, s>>>(h, m->w + p.ln1_w, m->w + p.ln1_b, m->nrm, d, a.norm_eps);
      else rmsnorm_fwd(h, m->w + p.ln1_w, m->nrm, nullptr, n, d, a.norm_eps, s);
      ++nl;
      gemm(m->nrm, d, p.wqkv, qkvd, m->qkv, nullptr);
      const long long rp = static_cast<long long>(n) * (H + Hkv) * (dh / 2);
      rope_append_kernel<<<static_cast<int>((rp + 255) / 256), 256, 0, s>>>(
          m->qkv, qkvd, m->inv_freq, m->pos, m->slot, kc, vc, n, H, Hkv, dh, a.max_ctx); ++nl;
      if (dh == 64)
        decode_attn_kernel<64><<<agrid, ATT_THREADS, att_smem, s>>>(m->qkv, qkvd, kc, vc, m->pos, m->slot,
                                                                    m->att, H, Hkv, a.max_ctx, sc_stride, scale);
      else
        decode_attn_kernel<128><<<agrid, ATT_THREADS, att_smem, s>>>(m->qkv, qkvd, kc, vc, m->pos, m->slot,
                                                                     m->att, H, Hkv, a.max_ctx, sc_stride, scale);
      ++nl;
      if (falcon) {
        // parallel residual: h' = h + dense(attn) + W2 gelu(W1 ln(h))   (modeling_falcon.py
        // FalconDecoderLayer.forward, parallel_attn branch)
        gemm(m->att, qd, p.wo, d, h2, h);
        gemm(m->nrm, d, p.w1, f, m->mid, nullptr, /*act=*/1);  // exact GeLU in the GEMM epilogue
        gemm(m->mid, f, p.w2, d, h, h2);
      } else {
        gemm(m->att, qd, p.wo, d, h2, h);
        rmsnorm_fwd(h2, m->w + p.ln2_w, m->nrm, nullptr, n, d, a.norm_eps, s); ++nl;
        gemm(m->nrm, d, p.w1, 2 * f, m->mid, nullptr);
        swiglu_fwd(m->mid, m->act, n, f, s); ++nl;
        gemm(m->act, f, p.w2, d, h, h2);
      }
    }
    if (falcon) layernorm_kernel<<<n, 256, 0, s>>>(h, m->w + m->p_lnf_w, m->w + m->p_lnf_b, m->nrm, d, a.norm_eps);
    else rmsnorm_fwd(h, m->w + m->p_lnf_w, m->nrm, nullptr, n, d, a.norm_eps, s);
    ++nl;
    gemm(m->nrm, d, m->p_lm, V, m->logits, nullptr);
    argmax_kernel<<<n, 1024, 0, s>>>(m->logits, V, m->next); ++nl;
    B200W_CUDA(cudaGetLastError());
    B200W_CUDA(cudaMemcpyAsync(m->pin + 3 * B, m->next, n * 4, cudaMemcpyDeviceToHost, s));
    };

    auto git = m->graphs.find(n);
    if (git != m->graphs.end()) {
      B200W_CUDA(cudaGraphLaunch(git->second, s));
      step_launches = m->warm[n];
    } else if (m->warm.count(n)) {
      cudaGraph_t graph = nullptr;
      B200W_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
      try {
        enqueue();
      } catch (...) {
        cudaStreamEndCapture(s, &graph);
        if (graph) cudaGraphDestroy(graph);
        throw;
      }
      B200W_CUDA(cudaStreamEndCapture(s, &graph));
      cudaGraphExec_t exec = nullptr;
      B200W_CUDA(cudaGraphInstantiate(&exec, graph, 0));
      cudaGraphDestroy(graph);
      m->graphs[n] = exec;
      B200W_CUDA(cudaGraphLaunch(exec, s));
    } else {
      enqueue();
      m->warm[n] = static_cast<int>(step_launches);
    }
    nl += step_launches;
    B200W_CUDA(cudaStreamSynchronize(s));
    if (next_tokens) memcpy(next_tokens, m->pin + 3 * B, n * 4);
    if (logits_out) {
      void* tmp = nullptr;
      B200W_CUDA(cudaMalloc(&tmp, static_cast<size_t>(n) * V * 4));
      cast_bf16_to_f32(m->logits, static_cast<float*>(tmp), static_cast<size_t>(n) * V, s);
      B200W_CUDA(cudaStreamSynchronize(s));
      B200W_CUDA(cudaMemcpy(logits_out, tmp, static_cast<size_t>(n) * V * 4, cudaMemcpyDeviceToHost));
      cudaFree(tmp);
    }
    B200W_CUDA(cudaStreamSynchronize(s));
  });
}

int64_t b200w_infer_device_bytes(b200w_ctx* ctx) {
  Infer* m = ctx ? static_cast<Infer*>(ctx_infer_slot(ctx)) : nullptr;
  return m ? m->bytes : 0;
}

}  // extern "C"
