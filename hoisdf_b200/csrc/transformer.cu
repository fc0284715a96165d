// hoisdf_encoder_fwd -- a whole post-norm transformer ENCODER stack (upstream common/nets/transformer.py:175-202 over
// TransformerEncoderLayer.forward_post :279-302; pos_embed is all zeros upstream, main/model.py:541-543) behind ONE C entry
// point, batch-major rows (row = b * S + token), d_model 256, 4 heads of 64:
//   per layer  qkv = x W_in^T + b         FP16x3 tcgen05 GEMM (hoisdf_linear_h3_fwd), fp32 (rows, 768)
//              att = softmax(q k^T / 8) v  tcgen05 flash attention, result in split-half format (hoisdf_attention_split_fwd)
//              x1  = LN1(x + att W_o^T)    GEMM + fused residual LayerNorm (+ split-half copy)
//              out = LN2(x1 + relu(x1 W_1^T) W_2^T)  two GEMMs (hidden kept in split-half format) + fused residual LayerNorm,
//                    which also emits inter_norm(out) -- the shared LayerNorm of :196-197 -- and both split-half copies
// The same launches, in the same order, as hoisdf_b200/nets/transformer.py:TransformerEncoder.forward_bm (bit-identical).
// The caller owns every buffer (workspace: hoisdf_encoder_workspace_bytes); nothing is allocated, nothing is read back.
#include <cstring>

#include "common.cuh"

namespace hoisdf {
namespace {

constexpr int64_t kAlignT = 256;
inline int64_t upT(int64_t x) { return (x + kAlignT - 1) / kAlignT * kAlignT; }

struct EncLayout { int64_t xs[2], xf[2], qkv, att, y, x1, x1s, h, attn_ws, attn_bytes, total; };

EncLayout enc_layout(int64_t batch, int64_t seq, int64_t d_ff, int64_t heads) {
  EncLayout L{};
  int64_t o = 0;
  auto take = [&](int64_t bytes) { const int64_t at = o; o += upT(bytes); return at; };
  const int64_t n = batch * seq, d = heads * 64;
  for (int i = 0; i < 2; ++i) L.xs[i] = take(n * 2 * d * 2);       // split-half (rows, 2, d)
  for (int i = 0; i < 2; ++i) L.xf[i] = take(n * d * 4);
  L.qkv = take(n * 3 * d * 4);
  L.att = take(n * 2 * d * 2);
  L.y = take(n * d * 4);
  L.x1 = take(n * d * 4);
  L.x1s = take(n * 2 * d * 2);
  L.h = take(n * 2 * d_ff * 2);
  L.attn_bytes = hoisdf_attention_workspace_bytes(batch, heads, seq, seq);
  L.attn_ws = take(L.attn_bytes);
  L.total = o;
  return L;
}

__global__ void add_rows_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) y[i] = a[i] + b[i];
}

struct DecLayout { int64_t tin[2], add, xs, qk, kv, att, y, t1, q, mkv, att2, t2, t2s, h, attn_ws, attn_bytes, total; };

DecLayout dec_layout(int64_t batch, int64_t lq, int64_t seq, int64_t d_ff, int64_t heads) {
  DecLayout L{};
  int64_t o = 0;
  auto take = [&](int64_t bytes) { const int64_t at = o; o += upT(bytes); return at; };
  const int64_t n = batch * lq, d = heads * 64;
  for (int i = 0; i < 2; ++i) L.tin[i] = take(n * d * 4);      // layer input / output (fp32), ping-pong
  L.add = take(n * d * 4);
  L.xs = take(n * 2 * d * 2);                                   // split-half scratch for the (rows, d) operands
  L.qk = take(n * 2 * d * 4);
  L.kv = take(n * 2 * d * 4);
  L.att = take(n * d * 4);
  L.y = take(n * d * 4);
  L.t1 = take(n * d * 4);
  L.q = take(n * d * 4);
  L.mkv = take(batch * seq * 2 * d * 4);
  L.att2 = take(n * d * 4);
  L.t2 = take(n * d * 4);
  L.t2s = take(n * 2 * d * 2);
  L.h = take(n * 2 * d_ff * 2);
  L.attn_bytes = hoisdf_attention_workspace_bytes(batch, heads, lq, seq);
  L.attn_ws = take(L.attn_bytes);
  L.total = o;
  return L;
}

int h3_linear(const uint16_t* x_hi, const uint16_t* x_lo, int64_t ldx, const hoisdf_h3_linear& w, int64_t m, int64_t n,
              int64_t k, int act, float* y, int64_t ldy, uint16_t* y_hi, uint16_t* y_lo, int64_t ldyh, void* stream) {
  hoisdf_linear_h3_args l;
  std::memset(&l, 0, sizeof(l));
  l.x_hi = x_hi; l.x_lo = x_lo; l.ldx = ldx;
  l.w_a = w.a; l.w_b = w.b; l.w_c = w.c; l.ldw = w.ld; l.bias = w.bias;
  l.y = y; l.ldy = ldy; l.y_hi = y_hi; l.y_lo = y_lo; l.ldyh = ldyh;
  l.m = m; l.n = n; l.k = k; l.act = act; l.chunk_kb = 0; l.single_pass = 0; l.w_scale = w.scale > 0.f ? w.scale : 1.f;
  return hoisdf_linear_h3_fwd(&l, stream);
}

}  // namespace
}  // namespace hoisdf

using namespace hoisdf;

HOISDF_API int64_t hoisdf_encoder_workspace_bytes(int64_t batch, int64_t seq, int64_t d_ff, int32_t heads) {
  if (batch <= 0 || seq <= 32 || d_ff <= 0 || heads <= 0) return 0;
  return enc_layout(batch, seq, d_ff, heads).total;
}

HOISDF_API int hoisdf_encoder_fwd(const hoisdf_encoder_args* a, void* stream) {
  if (a == nullptr || a->layers == nullptr || a->x == nullptr || a->out == nullptr || a->workspace == nullptr)
    return HOISDF_E_NULL;
  if (a->num_layers <= 0 || a->batch <= 0 || a->heads * 64 != 256 || a->d_ff <= 0 || (a->d_ff & 31)) return HOISDF_E_SHAPE;
  if (a->seq <= 32) return HOISDF_E_UNSUPPORTED;          // short sequences take the SIMT attention kernel (caller's path)
  const bool want_inter = a->inter != nullptr;
  if (want_inter && (a->inter_g == nullptr || a->inter_b == nullptr)) return HOISDF_E_NULL;
  if ((a->out_hi == nullptr) != (a->out_lo == nullptr) || (a->inter_hi == nullptr) != (a->inter_lo == nullptr))
    return HOISDF_E_NULL;
  if (!aligned16(a->workspace)) return HOISDF_E_ALIGN;
  const int64_t B = a->batch, S = a->seq, n = B * S, d = 256, H = a->heads;
  const EncLayout L = enc_layout(B, S, a->d_ff, H);
  if (L.total > a->workspace_bytes) return HOISDF_E_WORKSPACE;
  char* ws = static_cast<char*>(a->workspace);
  auto h16 = [&](int64_t off) { return reinterpret_cast<uint16_t*>(ws + off); };
  auto f32 = [&](int64_t off) { return reinterpret_cast<float*>(ws + off); };
  const int64_t lds = 2 * d, ldh = 2 * a->d_ff;          // row pitch of the (rows, 2, cols) split-half buffers
  int st;
  // layer 0 reads the caller's fp32 tokens: split them once
  const float* x = a->x;
  uint16_t* xs = h16(L.xs[0]);
  st = hoisdf_split_rows(x, n, d, d, d, xs, xs + d, lds, stream);
  if (st != HOISDF_OK) return st;
  for (int li = 0; li < a->num_layers; ++li) {
    const hoisdf_encoder_layer& ly = a->layers[li];
    const bool last = li == a->num_layers - 1;
    float* qkv = f32(L.qkv);
    st = h3_linear(xs, xs + d, lds, ly.qkv, n, 3 * d, d, HOISDF_ACT_NONE, qkv, 3 * d, nullptr, nullptr, 0, stream);
    if (st != HOISDF_OK) return st;
    uint16_t* att = h16(L.att);
    st = hoisdf_attention_split_fwd(qkv, 3 * d, qkv + d, qkv + 2 * d, 3 * d, att, att + d, lds, B, H, S, S, S, ws + L.attn_ws,
                                    L.attn_bytes, stream);
    if (st != HOISDF_OK) return st;
    float* y = f32(L.y);
    st = h3_linear(att, att + d, lds, ly.out, n, d, d, HOISDF_ACT_NONE, y, d, nullptr, nullptr, 0, stream);
    if (st != HOISDF_OK) return st;
    float* x1 = f32(L.x1);
    uint16_t* x1s = h16(L.x1s);
    st = hoisdf_add_layernorm_split_fwd(y, x, ly.norm1_g, ly.norm1_b, x1, nullptr, nullptr, nullptr, n, d, x1s, x1s + d, lds,
                                        nullptr, nullptr, 0, stream);
    if (st != HOISDF_OK) return st;
    uint16_t* hh = h16(L.h);
    st = h3_linear(x1s, x1s + d, lds, ly.lin1, n, a->d_ff, d, HOISDF_ACT_RELU, nullptr, 0, hh, hh + a->d_ff, ldh, stream);
    if (st != HOISDF_OK) return st;
    st = h3_linear(hh, hh + a->d_ff, ldh, ly.lin2, n, d, a->d_ff, HOISDF_ACT_NONE, y, d, nullptr, nullptr, 0, stream);
    if (st != HOISDF_OK) return st;
    // the layer's output: fp32 + split-half (next layer's operand / the caller's `out_hi`), and inter_norm of it
    float* xo = last ? a->out : f32(L.xf[li & 1]);
    uint16_t* xo_hi = (last && a->out_hi != nullptr) ? a->out_hi : h16(L.xs[(li + 1) & 1]);
    uint16_t* xo_lo = (last && a->out_hi != nullptr) ? a->out_lo : xo_hi + d;
    const int64_t ldxo = (last && a->out_hi != nullptr) ? a->ld_out : lds;
    float* it = want_inter ? a->inter + static_cast<int64_t>(li) * n * d : nullptr;
    uint16_t* it_hi = (want_inter && a->inter_hi != nullptr) ? a->inter_hi + static_cast<int64_t>(li) * n * a->ld_inter : nullptr;
    uint16_t* it_lo = (want_inter && a->inter_hi != nullptr) ? a->inter_lo + static_cast<int64_t>(li) * n * a->ld_inter : nullptr;
    st = hoisdf_add_layernorm_split_fwd(y, x1, ly.norm2_g, ly.norm2_b, xo, want_inter ? a->inter_g : nullptr,
                                        want_inter ? a->inter_b : nullptr, it, n, d, xo_hi, xo_lo, ldxo, it_hi, it_lo,
                                        it_hi != nullptr ? a->ld_inter : 0, stream);
    if (st != HOISDF_OK) return st;
    x = xo;
    xs = xo_hi;
    if (!(last && a->out_hi != nullptr) && xo_lo != xs + d) return HOISDF_E_SHAPE;
  }
  return launch_status();
}

// ---------------------------------------------------------------------------------------------------------------------
// hoisdf_decoder_fwd -- the post-norm transformer DECODER stack over a handful of learned queries (upstream
// common/nets/transformer.py:214-252 over TransformerDecoderLayer.forward_post :366-395; tgt = 0, query_pos = the query
// embedding, pos = 0): masked self-attention among the queries, cross-attention to the encoder memory with the keys
// >= kv_valid blocked (the only memory_mask upstream builds, common/utils/misc.py:42-47), ReLU feed-forward, three residual
// LayerNorms, and `norm(out_l)` of every layer as the result.  Same launches, same order as
// hoisdf_b200/nets/transformer.py:TransformerDecoderLayer.forward_bm (bit-identical).
HOISDF_API int64_t hoisdf_decoder_workspace_bytes(int64_t batch, int64_t queries, int64_t seq, int64_t d_ff, int32_t heads) {
  if (batch <= 0 || queries <= 0 || seq <= 0 || d_ff <= 0 || heads <= 0) return 0;
  return dec_layout(batch, queries, seq, d_ff, heads).total;
}

HOISDF_API int hoisdf_decoder_fwd(const hoisdf_decoder_args* a, void* stream) {
  if (a == nullptr || a->layers == nullptr || a->query_pos == nullptr || a->memory_hi == nullptr || a->memory_lo == nullptr ||
      a->norm_g == nullptr || a->norm_b == nullptr || a->hs == nullptr || a->workspace == nullptr)
    return HOISDF_E_NULL;
  if (a->num_layers <= 0 || a->batch <= 0 || a->queries <= 0 || a->seq <= 0 || a->heads * 64 != 256 || a->d_ff <= 0 ||
      (a->d_ff & 31) || a->kv_valid <= 0 || a->kv_valid > a->seq)
    return HOISDF_E_SHAPE;
  if (!aligned16(a->workspace)) return HOISDF_E_ALIGN;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t B = a->batch, Lq = a->queries, S = a->seq, n = B * Lq, d = 256, H = a->heads;
  const DecLayout L = dec_layout(B, Lq, S, a->d_ff, H);
  if (L.total > a->workspace_bytes) return HOISDF_E_WORKSPACE;
  char* ws = static_cast<char*>(a->workspace);
  auto h16 = [&](int64_t off) { return reinterpret_cast<uint16_t*>(ws + off); };
  auto f32 = [&](int64_t off) { return reinterpret_cast<float*>(ws + off); };
  const int64_t lds = 2 * d, ldh = 2 * a->d_ff;
  const unsigned add_blocks = static_cast<unsigned>(ceil_div(n * d, 256));
  int st;
  // fp32 (rows, d) -> split-half scratch -> GEMM
  auto lin = [&](const float* x, const hoisdf_h3_linear& w, int64_t nn, int act, float* y, int64_t ldy) -> int {
    uint16_t* xs = h16(L.xs);
    int e = hoisdf_split_rows(x, n, d, d, d, xs, xs + d, lds, stream);
    if (e != HOISDF_OK) return e;
    return h3_linear(xs, xs + d, lds, w, n, nn, d, act, y, ldy, nullptr, nullptr, 0, stream);
  };
  float* t = f32(L.tin[0]);
  if (cudaMemsetAsync(t, 0, n * d * 4, s) != cudaSuccess) return HOISDF_E_SHAPE;        // tgt = zeros (transformer.py:150)
  for (int li = 0; li < a->num_layers; ++li) {
    const hoisdf_decoder_layer& ly = a->layers[li];
    float* add = f32(L.add);
    // ---- masked self-attention among the queries: q = k = tgt + query_pos, v = tgt
    add_rows_kernel<<<add_blocks, 256, 0, s>>>(t, a->query_pos, add, n * d);
    float* qk = f32(L.qk);
    if ((st = lin(add, ly.sa_qk, 2 * d, HOISDF_ACT_NONE, qk, 2 * d)) != HOISDF_OK) return st;
    float* kv = f32(L.kv);
    if ((st = lin(t, ly.sa_v, d, HOISDF_ACT_NONE, kv + d, 2 * d)) != HOISDF_OK) return st;
    if (cudaMemcpy2DAsync(kv, 2 * d * 4, qk + d, 2 * d * 4, d * 4, n, cudaMemcpyDeviceToDevice, s) != cudaSuccess)
      return HOISDF_E_SHAPE;
    float* att = f32(L.att);
    st = hoisdf_attention_fwd(qk, 2 * d, kv, kv + d, 2 * d, att, d, B, H, Lq, Lq, Lq, a->tgt_mask, nullptr, 0, stream);
    if (st != HOISDF_OK) return st;
    float* y = f32(L.y);
    if ((st = lin(att, ly.sa_out, d, HOISDF_ACT_NONE, y, d)) != HOISDF_OK) return st;
    float* t1 = f32(L.t1);
    st = hoisdf_add_layernorm_fwd(y, t, ly.norm1_g, ly.norm1_b, t1, nullptr, nullptr, nullptr, n, d, stream);
    if (st != HOISDF_OK) return st;
    // ---- cross-attention to the memory (keys >= kv_valid blocked)
    add_rows_kernel<<<add_blocks, 256, 0, s>>>(t1, a->query_pos, add, n * d);
    float* q = f32(L.q);
    if ((st = lin(add, ly.ca_q, d, HOISDF_ACT_NONE, q, d)) != HOISDF_OK) return st;
    float* mkv = f32(L.mkv);
    st = h3_linear(a->memory_hi, a->memory_lo, a->ld_memory, ly.ca_kv, B * S, 2 * d, d, HOISDF_ACT_NONE, mkv, 2 * d, nullptr,
                   nullptr, 0, stream);
    if (st != HOISDF_OK) return st;
    float* att2 = f32(L.att2);
    st = hoisdf_attention_fwd(q, d, mkv, mkv + d, 2 * d, att2, d, B, H, Lq, S, a->kv_valid, nullptr, ws + L.attn_ws, L.attn_bytes,
                              stream);
    if (st != HOISDF_OK) return st;
    if ((st = lin(att2, ly.ca_out, d, HOISDF_ACT_NONE, y, d)) != HOISDF_OK) return st;
    float* t2 = f32(L.t2);
    uint16_t* t2s = h16(L.t2s);
    st = hoisdf_add_layernorm_split_fwd(y, t1, ly.norm2_g, ly.norm2_b, t2, nullptr, nullptr, nullptr, n, d, t2s, t2s + d, lds,
                                        nullptr, nullptr, 0, stream);
    if (st != HOISDF_OK) return st;
    // ---- feed-forward, third LayerNorm, and the stack's norm of the layer output
    uint16_t* hh = h16(L.h);
    st = h3_linear(t2s, t2s + d, lds, ly.lin1, n, a->d_ff, d, HOISDF_ACT_RELU, nullptr, 0, hh, hh + a->d_ff, ldh, stream);
    if (st != HOISDF_OK) return st;
    st = h3_linear(hh, hh + a->d_ff, ldh, ly.lin2, n, d, a->d_ff, HOISDF_ACT_NONE, y, d, nullptr, nullptr, 0, stream);
    if (st != HOISDF_OK) return st;
    float* out = f32(L.tin[(li + 1) & 1]);
    st = hoisdf_add_layernorm_fwd(y, t2, ly.norm3_g, ly.norm3_b, out, a->norm_g, a->norm_b, a->hs + static_cast<int64_t>(li) * n * d,
                                  n, d, stream);
    if (st != HOISDF_OK) return st;
    t = out;
  }
  return launch_status();
}
