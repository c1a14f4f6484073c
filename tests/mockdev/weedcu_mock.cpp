// weedcu_mock.cpp — TEST INFRASTRUCTURE ONLY. Never shipped, never loaded by weed_b200/.
//
// A host-memory stand-in for libweedcu.so that implements every entry point of include/weedcu.h by
// forwarding to the C oracle (oracle/weed_oracle.c). It exists so that the HOST logic of the product
// (weed_b200/host: views, broadcasting, the autograd graph, module composition, optimiser plumbing)
// can be exercised by `pytest -m "not gpu"` in a container without a GPU: tests/mockdev/Makefile
// links a second copy of the host library + harness against this file. "Device memory" is malloc'd
// host memory, streams and events are no-ops, every "launch" runs synchronously.
// The GPU tests (-m gpu) and bench.py use the real CUDA library; nothing here is a fallback.
#include "weedcu.h"
extern "C" {
#include "weed_oracle.h"
}

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

static uint64_t g_launches = 0, g_mallocs = 0, g_frees = 0;
#define VIEW(v) reinterpret_cast<const wo_view *>(v)
#define MAT(m) reinterpret_cast<const wo_mat *>(m)
static const bool g_trace = getenv("WEEDCU_MOCK_TRACE") != nullptr; // print one line per kernel call
static inline void trace_call(const char *what) {
  if (!g_trace) return;
  const char *p = strchr(what, '(');
  static auto last = std::chrono::steady_clock::now();
  const auto now = std::chrono::steady_clock::now();
  fprintf(stderr, "[mock] %.*s +%.1fus\n", p ? (int)(p - what) : (int)strlen(what), what,
          std::chrono::duration<double, std::micro>(now - last).count()); // host time since the previous kernel call
  last = now;
}
// WEEDCU_MOCK_NOCOMPUTE: skip the oracle call (results are garbage) — isolates the HOST cost of a step
static const bool g_nocompute = getenv("WEEDCU_MOCK_NOCOMPUTE") != nullptr;
#define RUN(expr) (++g_launches, trace_call(#expr), g_nocompute ? 0 : ((expr) == 0 ? 0 : WEEDCU_EINVAL))

extern "C" {
int weedcu_device_count(int *count) { if (!count) return WEEDCU_EINVAL; *count = 1; return 0; }
int weedcu_set_device(int) { return 0; }
int weedcu_get_device(int *device) { if (!device) return WEEDCU_EINVAL; *device = 0; return 0; }
int weedcu_device_info(int, char *name, int name_len, uint64_t *total_mem, int *sm_count, int *cc_major, int *cc_minor) {
  if (name && name_len > 0) { strncpy(name, "mock (host memory, oracle-backed)", (size_t)name_len - 1); name[name_len - 1] = 0; }
  if (total_mem) *total_mem = 1ull << 34;
  if (sm_count) *sm_count = 1;
  if (cc_major) *cc_major = 0;
  if (cc_minor) *cc_minor = 0;
  return 0;
}
const char *weedcu_error_string(int code) { return code == 0 ? "ok" : (code == WEEDCU_EINVAL ? "weedcu(mock): invalid argument" : "weedcu(mock): error"); }
void *weedcu_default_stream(void) { return (void *)0x1; }
int weedcu_set_default_stream(void *) { return 0; }
int weedcu_stream_create(void **stream) { if (!stream) return WEEDCU_EINVAL; *stream = (void *)0x1; return 0; }
int weedcu_stream_create_priority(void **stream, int) { return weedcu_stream_create(stream); }
int weedcu_stream_destroy(void *) { return 0; }
int weedcu_stream_sync(void *) { return 0; }
int weedcu_stream_wait_event(void *, void *) { return 0; }
int weedcu_event_create(void **event) { if (!event) return WEEDCU_EINVAL; *event = new double(0); return 0; }
int weedcu_event_destroy(void *event) { delete (double *)event; return 0; }
int weedcu_event_record(void *event, void *) {
  *(double *)event = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
  return 0;
}
int weedcu_event_sync(void *) { return 0; }
int weedcu_event_elapsed_ms(void *start, void *stop, float *ms) { if (!ms) return WEEDCU_EINVAL; *ms = (float)(*(double *)stop - *(double *)start); return 0; }
int weedcu_malloc(void **ptr, size_t bytes, void *) { if (!ptr) return WEEDCU_EINVAL; *ptr = malloc(bytes ? bytes : 16); ++g_mallocs; return *ptr ? 0 : 2; }
int weedcu_free(void *ptr, void *) { if (ptr) ++g_frees; free(ptr); return 0; }
int weedcu_pool_trim(void) { return 0; }
int weedcu_mem_info(uint64_t *f, uint64_t *t) { if (f) *f = 1ull << 33; if (t) *t = 1ull << 34; return 0; }
int weedcu_host_alloc(void **ptr, size_t bytes) { if (!ptr) return WEEDCU_EINVAL; *ptr = malloc(bytes ? bytes : 16); return *ptr ? 0 : 2; }
int weedcu_host_free(void *ptr) { free(ptr); return 0; }
int weedcu_memcpy_h2d(void *dst, const void *src, size_t bytes, void *) { memcpy(dst, src, bytes); return 0; }
int weedcu_memcpy_d2h(void *dst, const void *src, size_t bytes, void *) { memcpy(dst, src, bytes); return 0; }
int weedcu_memcpy_d2d(void *dst, const void *src, size_t bytes, void *) { memmove(dst, src, bytes); return 0; }
int weedcu_launch_count(uint64_t *count) { if (!count) return WEEDCU_EINVAL; *count = g_launches; return 0; }
int weedcu_host_stats(double *a, uint64_t *b, double *c, uint64_t *d) { if (a) *a = 0; if (b) *b = g_mallocs; if (c) *c = 0; if (d) *d = g_frees; return 0; }
int weedcu_prof_enable(int) { return 0; }
int weedcu_gemm_set_dynamic(int) { return 0; }
int weedcu_gemm_set_mode(int) { return 0; }
int weedcu_set_pdl(int) { return 0; }
int weedcu_prof_read(int, double *t, uint64_t *n, double *w) { if (t) *t = 0; if (n) *n = 0; if (w) *w = 0; return 0; }

int weedcu_fill_real(float *p, uint64_t n, float value, void *) { return RUN(wo_fill_real(p, n, value)); }
int weedcu_fill_int(int32_t *p, uint64_t n, int32_t value, void *) { for (uint64_t i = 0; i < n; ++i) p[i] = value; ++g_launches; return 0; }
int weedcu_binary_real(int op, const float *a, const weedcu_view *av, const float *b, const weedcu_view *bv, float *out, const weedcu_view *ov, void *) {
  return RUN(wo_binary_real(op, a, VIEW(av), b, VIEW(bv), out, VIEW(ov)));
}
int weedcu_inplace_real(int op, float *a, const weedcu_view *av, const float *b, const weedcu_view *bv, void *) {
  return RUN(wo_inplace_real(op, a, VIEW(av), b, VIEW(bv)));
}
int weedcu_copy_real(float *dst, const weedcu_view *dv, const float *src, const weedcu_view *sv, void *) { return RUN(wo_copy_real(dst, VIEW(dv), src, VIEW(sv))); }
int weedcu_unary_real(int op, float param, const float *a, const weedcu_view *av, float *out, const weedcu_view *ov, void *) {
  return RUN(wo_unary_real(op, param, a, VIEW(av), out, VIEW(ov)));
}
int weedcu_unary_grad_real(int op, float *din, const weedcu_view *dinv, const float *in, const weedcu_view *inv, const float *dout, const weedcu_view *doutv, int accumulate, void *) {
  return RUN(wo_unary_grad_real(op, din, VIEW(dinv), in, VIEW(inv), dout, VIEW(doutv), accumulate));
}
int weedcu_reduce_real(const float *a, const weedcu_view *av, int axis, float *out, int index_order, void *) { return RUN(wo_reduce_real(a, VIEW(av), axis, out, index_order)); }
int weedcu_reduce_grad_real(float *din, const weedcu_view *dinv, const float *dout, const weedcu_view *doutv, int axis, int index_order, void *) {
  return RUN(wo_reduce_grad_real(din, VIEW(dinv), dout, VIEW(doutv), axis, index_order));
}
int weedcu_clamp_real(const float *a, const weedcu_view *av, float lo, float hi, float *out, const weedcu_view *ov, void *) {
  return RUN(wo_clamp_real(a, VIEW(av), lo, hi, out, VIEW(ov)));
}
int weedcu_clamp_grad_real(float *din, const weedcu_view *dinv, const float *in, const weedcu_view *inv, const float *dout, const weedcu_view *doutv, float lo, float hi, void *) {
  return RUN(wo_clamp_grad_real(din, VIEW(dinv), in, VIEW(inv), dout, VIEW(doutv), lo, hi));
}
int weedcu_extremum_real(int is_min, const float *a, const weedcu_view *av, float *out, void *) { return RUN(wo_extremum_real(is_min, a, VIEW(av), out)); }
int weedcu_match_grad_full_real(float *din, const weedcu_view *dinv, const float *in, const weedcu_view *inv, const float *dout, const weedcu_view *doutv, const float *extremum, void *) {
  return RUN(wo_match_grad_full_real(din, VIEW(dinv), in, VIEW(inv), dout, VIEW(doutv), extremum));
}
int weedcu_extremum_axis_real(int is_min, const float *a, const weedcu_view *av, int axis, float *out, int index_order, void *) {
  return RUN(wo_extremum_axis_real(is_min, a, VIEW(av), axis, out, index_order));
}
int weedcu_match_grad_real(float *din, const weedcu_view *dinv, const float *in, const weedcu_view *inv, const float *dout, const weedcu_view *doutv, const float *reduced, int axis,
                           int index_order, void *) {
  return RUN(wo_match_grad_real(din, VIEW(dinv), in, VIEW(inv), dout, VIEW(doutv), reduced, axis, index_order));
}
int weedcu_sum_real(const float *a, const weedcu_view *av, float scale, float *out, void *) { return RUN(wo_sum_real(a, VIEW(av), scale, out)); }
int weedcu_softmax_real(int log_mode, const float *a, const weedcu_view *av, int axis, float *out, const weedcu_view *ov, void *) {
  return RUN(wo_softmax_real(log_mode, a, VIEW(av), axis, out, VIEW(ov)));
}
int weedcu_softmax_grad_real(int log_mode, float *din, const weedcu_view *dinv, const float *out, const weedcu_view *ov, const float *dout, const weedcu_view *doutv, int axis, void *) {
  return RUN(wo_softmax_grad_real(log_mode, din, VIEW(dinv), out, VIEW(ov), dout, VIEW(doutv), axis));
}
int weedcu_attn_softmax_real(const float *scores, float *out, uint32_t batch, uint32_t Tq, uint32_t Tk, float divisor, float mask_val, int causal, int batch_fastest, void *) {
  return RUN(wo_attn_softmax_real(scores, out, batch, Tq, Tk, divisor, mask_val, causal, batch_fastest));
}
int weedcu_attention_fwd(const float *q, const float *k, const float *v, float *out, uint32_t B, uint32_t T, uint32_t H, uint32_t hd, float divisor, float mask_val, int causal, void *) {
  if ((T % 8u) || T < 64u || (hd != 64u && T > 1024u) || hd < 16u || (hd % 8u)) return WEEDCU_ENOSUP; // same envelope as the device entry
  return RUN(wo_attention_fwd(q, k, v, out, B, T, H, hd, divisor, mask_val, (causal && T > 1) ? 1 : 0));
}
int weedcu_attention_fwd_bf16out(const float *q, const float *k, const float *v, float *out, uint16_t *out_bf16, uint32_t B, uint32_t T, uint32_t H, uint32_t hd, float divisor,
                                 float mask_val, int causal, void *stream) {
  if (out_bf16 && (B % 4u)) return WEEDCU_ENOSUP;
  const int rc = weedcu_attention_fwd(q, k, v, out, B, T, H, hd, divisor, mask_val, causal, stream);
  if (rc == 0 && out_bf16 && !g_nocompute)
    for (uint64_t i = 0; i < (uint64_t)B * T * H * hd; ++i) out_bf16[i] = wo_f32_to_bf16(out[i]);
  return rc;
}
int weedcu_attention_fwd_bf16in(const uint16_t *q, const uint16_t *k, const uint16_t *v, float *out, uint16_t *out_bf16, uint32_t B, uint32_t T, uint32_t H, uint32_t hd,
                                float divisor, float mask_val, int causal, void *stream) {
  if (!q || !k || !v) return WEEDCU_EINVAL;
  if (B % 8u) return WEEDCU_ENOSUP;
  const size_t n = (size_t)B * T * H * hd;
  std::vector<float> wq(n), wk(n), wv(n);
  for (size_t i = 0; i < n; ++i) {
    uint32_t a = (uint32_t)q[i] << 16, b = (uint32_t)k[i] << 16, c = (uint32_t)v[i] << 16;
    memcpy(&wq[i], &a, 4);
    memcpy(&wk[i], &b, 4);
    memcpy(&wv[i], &c, 4);
  }
  return weedcu_attention_fwd_bf16out(wq.data(), wk.data(), wv.data(), out, out_bf16, B, T, H, hd, divisor, mask_val, causal, stream);
}
int weedcu_attention_decode(const float *q, const float *k, const float *v, float *k_cache, float *v_cache, float *out, uint32_t B, uint32_t T_new, uint32_t H, uint32_t hd,
                            uint32_t S, uint32_t cache_len, float divisor, float mask_val, int causal, void *) {
  if (hd > 64u) return WEEDCU_ENOSUP;
  if ((uint64_t)cache_len + T_new > S) return WEEDCU_EINVAL;
  return RUN(wo_attention_decode(q, k, v, k_cache, v_cache, out, B, T_new, H, hd, S, cache_len, divisor, mask_val, causal));
}
int weedcu_matmul_skinny(const float *a, const weedcu_mat *am, const float *b, const weedcu_mat *bm, float *c, const weedcu_mat *cm, uint32_t M, uint32_t K, uint32_t N,
                         const float *bias, int accumulate, void *) {
  if (M > 16u) return WEEDCU_ENOSUP;
  return RUN(wo_matmul_skinny(a, MAT(am), b, MAT(bm), c, MAT(cm), M, K, N, bias, accumulate));
}
int weedcu_matmul_skinny_residual(const float *a, const weedcu_mat *am, const float *b, const weedcu_mat *bm, float *c, const weedcu_mat *cm, uint32_t M, uint32_t K, uint32_t N,
                                  const float *bias, const float *residual, void *stream) {
  if (!residual) return WEEDCU_EINVAL;
  const int rc = weedcu_matmul_skinny(a, am, b, bm, c, cm, M, K, N, bias, 0, stream);
  if (rc || g_nocompute) return rc;
  for (uint32_t m = 0; m < M; ++m)
    for (uint32_t n = 0; n < N; ++n) {
      const uint64_t o = cm->offset + (uint64_t)m * cm->s0 + (uint64_t)n * cm->s1;
      c[o] = c[o] + residual[o];
    }
  return 0;
}
int weedcu_matmul_skinny_grouped(const float *a, const weedcu_mat *am, uint32_t groups, const float *const *b, const weedcu_mat *bm, float *const *c, const weedcu_mat *cm, uint32_t M,
                                 uint32_t K, uint32_t N, const float *const *bias, void *stream) {
  if (!groups || groups > 3u || !b || !c) return WEEDCU_EINVAL;
  for (uint32_t g = 0; g < groups; ++g) {
    const int rc = weedcu_matmul_skinny(a, am, b[g], bm, c[g], cm, M, K, N, bias ? bias[g] : nullptr, 0, stream);
    if (rc) return rc;
  }
  return 0;
}
int weedcu_cross_entropy_fwd(const float *logits, uint64_t offset, uint32_t rows, uint32_t V, uint32_t rs, uint32_t vs, const int32_t *targets, float *lse, float *loss, void *) {
  return RUN(wo_cross_entropy_fwd(logits, offset, rows, V, rs, vs, targets, lse, loss));
}
int weedcu_cross_entropy_bwd(const float *logits, uint64_t offset, uint32_t rows, uint32_t V, uint32_t rs, uint32_t vs, const int32_t *targets, const float *lse, const float *dloss,
                             float *dlogits, uint64_t d_offset, int accumulate, void *) {
  return RUN(wo_cross_entropy_bwd(logits, offset, rows, V, rs, vs, targets, lse, dloss, dlogits, d_offset, accumulate));
}
int weedcu_layernorm_fwd_stats(const float *x, uint32_t rows, uint32_t F, const float *stats, uint32_t tiles, uint32_t tile_cols, const float *gamma,
                               const float *beta, float eps, float *y, float *mean, float *rstd, uint16_t *y_bf16, void *stream) {
  if (!stats || !tiles || (uint64_t)tiles * tile_cols < F || (rows % 4u)) return rows % 4u ? WEEDCU_ENOSUP : WEEDCU_EINVAL;
  return weedcu_layernorm_fwd_bf16(x, rows, F, gamma, beta, eps, y, mean, rstd, y_bf16, stream); // the partials only save a pass
}
int weedcu_layernorm_fwd_bf16(const float *x, uint32_t rows, uint32_t F, const float *gamma, const float *beta, float eps, float *y, float *mean, float *rstd, uint16_t *y_bf16,
                              void *) {
  if (y_bf16 && ((rows % 8u) || rows <= 256u)) return WEEDCU_ENOSUP;
  return RUN(wo_layernorm_fwd_bf16(x, rows, F, gamma, beta, eps, y, mean, rstd, y_bf16));
}
int weedcu_gelu_fwd_bf16(const float *x, float *y, uint16_t *y_bf16, uint64_t n, void *) {
  if (n % 4u) return WEEDCU_ENOSUP;
  if (!y) { // bf16 operand copy only
    std::vector<float> tmp((size_t)n);
    return RUN(wo_gelu_fwd_bf16(x, tmp.data(), y_bf16, n));
  }
  return RUN(wo_gelu_fwd_bf16(x, y, y_bf16, n));
}
int weedcu_gelu_grad_pack(float *din, const float *in, const float *dout, uint32_t rows, uint32_t cols, int accumulate, uint16_t *din_bf16, float *colsum, void *) {
  if (rows % 8u) return WEEDCU_ENOSUP;
  if (!din) { // operand copy + column sums only (the fp32 values are materialised on demand by the host)
    if (accumulate) return WEEDCU_EINVAL;
    std::vector<float> tmp((size_t)rows * cols);
    return RUN(wo_gelu_grad_pack(tmp.data(), in, dout, rows, cols, 0, din_bf16, colsum));
  }
  return RUN(wo_gelu_grad_pack(din, in, dout, rows, cols, accumulate, din_bf16, colsum));
}
int weedcu_cross_entropy_bwd_pack(const float *logits, uint64_t offset, uint32_t rows, uint32_t V, const int32_t *targets, const float *lse, const float *dloss, float *dlogits,
                                  uint64_t d_offset, int accumulate, uint16_t *dlogits_bf16, float *colsum, void *) {
  if (rows % 8u) return WEEDCU_ENOSUP;
  if (!dlogits) {
    if (accumulate) return WEEDCU_EINVAL;
    std::vector<float> tmp((size_t)rows * V);
    return RUN(wo_cross_entropy_bwd_pack(logits, offset, rows, V, targets, lse, dloss, tmp.data(), 0, 0, dlogits_bf16, colsum));
  }
  return RUN(wo_cross_entropy_bwd_pack(logits, offset, rows, V, targets, lse, dloss, dlogits, d_offset, accumulate, dlogits_bf16, colsum));
}
int weedcu_layernorm_fwd(const float *x, uint32_t rows, uint32_t F, const float *gamma, const float *beta, float eps, float *y, float *mean, float *rstd, void *) {
  return RUN(wo_layernorm_fwd(x, rows, F, gamma, beta, eps, y, mean, rstd));
}
int weedcu_layernorm_bwd(const float *x, const float *dy, uint32_t rows, uint32_t F, const float *gamma, const float *mean, const float *rstd, float *dx, float *dgamma, float *dbeta,
                         int grad_mode, int accumulate, void *) {
  return RUN(wo_layernorm_bwd(x, dy, rows, F, gamma, mean, rstd, dx, dgamma, dbeta, grad_mode, accumulate));
}
int weedcu_layernorm_bwd_from(const float *x, const float *dy, uint32_t rows, uint32_t F, const float *gamma, const float *mean, const float *rstd, const float *dx_in, float *dx,
                              float *dgamma, float *dbeta, int grad_mode, void *) {
  if (dx_in && dx_in != dx) memcpy(dx, dx_in, sizeof(float) * (size_t)rows * F); // then accumulate in place, as the oracle does
  return RUN(wo_layernorm_bwd(x, dy, rows, F, gamma, mean, rstd, dx, dgamma, dbeta, grad_mode, dx_in ? 1 : 0));
}
int weedcu_embedding_gather(const int32_t *idx, uint64_t idx_off, uint32_t idx_stride, uint32_t n, const float *W, uint64_t w_off, uint32_t w_s0, uint32_t w_s1, uint32_t D, float *out,
                            uint64_t o_off, uint32_t o_s0, uint32_t o_s1, void *) {
  return RUN(wo_embedding_gather(idx, idx_off, idx_stride, n, W, w_off, w_s0, w_s1, D, out, o_off, o_s0, o_s1));
}
int weedcu_embedding_scatter_add(float *dW, uint64_t w_off, uint32_t w_s0, uint32_t w_s1, const int32_t *idx, uint64_t idx_off, uint32_t idx_stride, uint32_t n, uint32_t D,
                                 const float *dout, uint64_t o_off, uint32_t o_s0, uint32_t o_s1, void *) {
  return RUN(wo_embedding_scatter_add(dW, w_off, w_s0, w_s1, idx, idx_off, idx_stride, n, D, dout, o_off, o_s0, o_s1));
}
int weedcu_triu_fill_real(float *a, const weedcu_view *av, float val, uint32_t diagonal, void *) { return RUN(wo_triu_fill_real(a, VIEW(av), val, diagonal)); }
int weedcu_argmax_rows(const float *x, uint64_t offset, uint32_t rows, uint32_t V, uint32_t rs, uint32_t vs, int32_t *out, void *) { return RUN(wo_argmax_rows(x, offset, rows, V, rs, vs, out)); }
int weedcu_sgd_step(float *p, const float *g, uint64_t n, float lr, float gscale, void *) { return RUN(wo_sgd_step(p, g, n, lr, gscale)); }
int weedcu_adam_step(float *p, const float *g, float *m, float *v, uint64_t n, float lr, float beta1, float beta2, float eps, float bc1, float bc2, float gscale, void *) {
  return RUN(wo_adam_step(p, g, m, v, n, lr, beta1, beta2, eps, bc1, bc2, gscale));
}
int weedcu_adam_step_multi_shadow(uint32_t count, float *const *p, const float *const *g, float *const *m, float *const *v, const uint64_t *n, uint16_t *const *shadow, float lr,
                                  float beta1, float beta2, float eps, float bc1, float bc2, float gscale, void *) {
  ++g_launches;
  for (uint32_t t = 0; t < count; ++t) {
    std::vector<float> zeros;
    if (!g[t]) zeros.assign(n[t], 0.0f); // NULL gradient = all zeros
    if (wo_adam_step(p[t], g[t] ? g[t] : zeros.data(), m[t], v[t], n[t], lr, beta1, beta2, eps, bc1, bc2, gscale) != 0) return WEEDCU_EINVAL;
    if (shadow && shadow[t])
      for (uint64_t i = 0; i < n[t]; ++i) shadow[t][i] = wo_f32_to_bf16(p[t][i]);
  }
  return 0;
}
int weedcu_adam_step_multi_zero(uint32_t count, float *const *p, const float *const *g, float *const *m, float *const *v, const uint64_t *n, uint16_t *const *shadow,
                                const uint8_t *zero_grad, float lr, float beta1, float beta2, float eps, float bc1, float bc2, float gscale, void *stream) {
  const int rc = weedcu_adam_step_multi_shadow(count, p, g, m, v, n, shadow, lr, beta1, beta2, eps, bc1, bc2, gscale, stream);
  if (rc == 0 && zero_grad)
    for (uint32_t t = 0; t < count; ++t)
      if (zero_grad[t] && g[t]) memset(const_cast<float *>(g[t]), 0, sizeof(float) * n[t]);
  return rc;
}
int weedcu_adam_step_multi(uint32_t count, float *const *p, const float *const *g, float *const *m, float *const *v, const uint64_t *n, float lr, float beta1, float beta2, float eps,
                           float bc1, float bc2, float gscale, void *stream) {
  return weedcu_adam_step_multi_shadow(count, p, g, m, v, n, nullptr, lr, beta1, beta2, eps, bc1, bc2, gscale, stream);
}
int weedcu_matmul_real(const float *a, const weedcu_mat *am, const float *b, const weedcu_mat *bm, float *c, const weedcu_mat *cm, uint32_t M, uint32_t K, uint32_t N, uint32_t batch,
                       int accumulate, int precision, void *) {
  if (precision == WEEDCU_GEMM_BF16) return RUN(wo_matmul_bf16_model(a, MAT(am), b, MAT(bm), c, MAT(cm), M, K, N, batch, accumulate));
  return RUN(wo_matmul_real(a, MAT(am), b, MAT(bm), c, MAT(cm), M, K, N, batch, accumulate));
}
// bf16 operands: the same rounding model as wo_matmul_bf16_model, split the way the product does it
// (pack once with wo_f32_to_bf16, multiply the widened values in fp32)
static inline float bf16_widen(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, sizeof(f));
  return f;
}
int weedcu_pack_bf16(const float *src, uint64_t offset, uint32_t s0, uint32_t s1, uint32_t rows, uint32_t cols, uint16_t *dst, int dst_major, void *) {
  if (!src || !dst || !rows || !cols) return WEEDCU_EINVAL;
  trace_call("pack_bf16(");
  if (g_nocompute) return 0;
  const uint64_t ld = ((uint64_t)(dst_major ? rows : cols) + 7U) & ~(uint64_t)7U;
  for (uint32_t r = 0; r < rows; ++r)
    for (uint32_t c = 0; c < cols; ++c)
      dst[dst_major ? (r + c * ld) : (c + r * ld)] = wo_f32_to_bf16(src[offset + (uint64_t)r * s0 + (uint64_t)c * s1]);
  return 0;
}
int weedcu_pack_bf16_colsum(const float *src, uint64_t offset, uint32_t s0, uint32_t s1, uint32_t rows, uint32_t cols, uint16_t *dst, int dst_major, float *colsum, int accumulate,
                            void *) {
  if (!src || !dst || !colsum || !rows || !cols) return WEEDCU_EINVAL;
  const uint32_t n_fast = dst_major ? rows : cols, n_slow = dst_major ? cols : rows;
  const uint64_t s_fast = dst_major ? s0 : s1, ss = dst_major ? s1 : s0;
  if (s_fast != 1 || (n_fast % 8u) || (ss % 4u)) return WEEDCU_ENOSUP; // same envelope as the device entry
  trace_call("pack_bf16_colsum(");
  ++g_launches;
  if (g_nocompute) return 0;
  const uint64_t ld = ((uint64_t)n_fast + 7U) & ~(uint64_t)7U;
  for (uint32_t j = 0; j < n_slow; ++j) {
    float sum = 0.0f;
    for (uint32_t i = 0; i < n_fast; ++i) {
      const float x = src[offset + (uint64_t)j * ss + i];
      sum += x;
      dst[(uint64_t)j * ld + i] = wo_f32_to_bf16(x);
    }
    colsum[j] = accumulate ? colsum[j] + sum : sum;
  }
  return 0;
}
int weedcu_gemm_bf16(const uint16_t *a, int a_major, uint64_t lda, const uint16_t *b, int b_major, uint64_t ldb, float *c, uint64_t ldc, uint32_t M, uint32_t N, uint32_t K,
                     int accumulate, const float *col_bias, void *) {
  if (!a || !b || !c || !M || !N || !K) return WEEDCU_EINVAL;
  trace_call("gemm_bf16(");
  ++g_launches;
  if (g_nocompute) return 0;
  for (uint32_t m = 0; m < M; ++m)
    for (uint32_t n = 0; n < N; ++n) {
      double sum = 0.0; // same accumulation as wo_matmul_bf16_model
      for (uint32_t k = 0; k < K; ++k)
        sum += (double)bf16_widen(a[a_major ? (m + k * lda) : (k + m * lda)]) * (double)bf16_widen(b[b_major ? (n + k * ldb) : (k + n * ldb)]);
      float *o = &c[m + (uint64_t)n * ldc];
      *o = accumulate ? (float)(*o + sum) : (float)sum;
      if (col_bias) *o = *o + col_bias[n];
    }
  return 0;
}
int weedcu_gemm_bf16_residual(const uint16_t *a, int a_major, uint64_t lda, const uint16_t *b, int b_major, uint64_t ldb, float *c, uint64_t ldc, uint32_t M, uint32_t N, uint32_t K,
                              const float *col_bias, const float *residual, uint64_t ldr, void *stream) {
  if (!residual) return WEEDCU_EINVAL;
  const int rc = weedcu_gemm_bf16(a, a_major, lda, b, b_major, ldb, c, ldc, M, N, K, 0, col_bias, stream);
  if (rc || g_nocompute) return rc;
  for (uint32_t n = 0; n < N; ++n)
    for (uint32_t m = 0; m < M; ++m) c[m + (uint64_t)n * ldc] = c[m + (uint64_t)n * ldc] + residual[m + (uint64_t)n * ldr];
  return 0;
}
int weedcu_gemm_bf16_grouped(const uint16_t *a, int a_major, uint64_t lda, uint32_t groups, const uint16_t *const *b, int b_major, uint64_t ldb, float *const *c, uint64_t ldc,
                             uint32_t M, uint32_t N, uint32_t K, int accumulate, const float *const *col_bias, void *stream) {
  if (!groups || groups > 3 || !b || !c) return WEEDCU_EINVAL;
  for (uint32_t g = 0; g < groups; ++g) {
    const int rc = weedcu_gemm_bf16(a, a_major, lda, b[g], b_major, ldb, c[g], ldc, M, N, K, accumulate, col_bias ? col_bias[g] : nullptr, stream);
    if (rc) return rc;
  }
  return 0;
}
int weedcu_cross_entropy_fwd_stats(const float *stats, uint32_t tiles, uint32_t rows, uint32_t V, const uint16_t *a, int a_major, uint64_t lda, const uint16_t *b,
                                   int b_major, uint64_t ldb, uint32_t K, const float *col_bias, const int32_t *targets, float *lse, float *loss, void *) {
  if (!stats || !tiles || !a || !b || !targets || !lse || !loss) return WEEDCU_EINVAL;
  trace_call("cross_entropy_fwd_stats(");
  ++g_launches;
  if (g_nocompute) return 0;
  double total = 0.0;
  for (uint32_t r = 0; r < rows; ++r) {
    double M = -1.0 / 0.0, S = 0.0;
    for (uint32_t t = 0; t < tiles; ++t) M = stats[2 * ((uint64_t)t * rows + r)] > M ? stats[2 * ((uint64_t)t * rows + r)] : M;
    for (uint32_t t = 0; t < tiles; ++t) S += (double)stats[2 * ((uint64_t)t * rows + r) + 1] * exp((double)stats[2 * ((uint64_t)t * rows + r)] - M);
    const uint32_t tg = (uint32_t)targets[r];
    if (tg >= V) return WEEDCU_EINVAL;
    double xt = 0.0;
    for (uint32_t k = 0; k < K; ++k)
      xt += (double)bf16_widen(a[a_major ? (r + k * lda) : (k + r * lda)]) * (double)bf16_widen(b[b_major ? (tg + k * ldb) : (k + tg * ldb)]);
    float x32 = (float)xt;
    if (col_bias) x32 = x32 + col_bias[tg];
    const float l = (float)(M + log(S));
    lse[r] = l;
    total += (double)(x32 - l);
  }
  *loss = (float)(-total / rows);
  return 0;
}
int weedcu_cross_entropy_fwd_bf16in(const uint16_t *logits_bf16, uint32_t rows, uint32_t V, const uint16_t *a, int a_major, uint64_t lda, const uint16_t *b, int b_major,
                                    uint64_t ldb, uint32_t K, const float *col_bias, const int32_t *targets, float *lse, float *loss, void *stream) {
  if (!logits_bf16) return WEEDCU_EINVAL;
  if (rows % 8u) return WEEDCU_ENOSUP;
  std::vector<float> stats(2 * (size_t)rows); // one partial per row over the whole vocabulary
  for (uint32_t r = 0; r < rows; ++r) {
    double M = -1.0 / 0.0, S = 0.0;
    for (uint32_t v = 0; v < V; ++v) M = bf16_widen(logits_bf16[r + (uint64_t)v * rows]) > M ? bf16_widen(logits_bf16[r + (uint64_t)v * rows]) : M;
    for (uint32_t v = 0; v < V; ++v) S += exp((double)bf16_widen(logits_bf16[r + (uint64_t)v * rows]) - M);
    stats[2 * (size_t)r] = (float)M;
    stats[2 * (size_t)r + 1] = (float)S;
  }
  return weedcu_cross_entropy_fwd_stats(stats.data(), 1, rows, V, a, a_major, lda, b, b_major, ldb, K, col_bias, targets, lse, loss, stream);
}
int weedcu_gelu_grad_pack_bf16dy(float *din, const float *in, const uint16_t *dout_bf16, uint32_t rows, uint32_t cols, int accumulate, uint16_t *din_bf16,
                                 float *colsum, void *stream) {
  std::vector<float> wide((size_t)rows * cols);
  for (size_t i = 0; i < wide.size(); ++i) wide[i] = bf16_widen(dout_bf16[i]);
  return weedcu_gelu_grad_pack(din, in, wide.data(), rows, cols, accumulate, din_bf16, colsum, stream);
}
int weedcu_cross_entropy_bwd_pack_bf16in(const uint16_t *logits_bf16, uint32_t rows, uint32_t V, const int32_t *targets, const float *lse, const float *dloss,
                                         float *dlogits, uint64_t d_offset, int accumulate, uint16_t *dlogits_bf16, float *colsum, void *stream) {
  if (!logits_bf16) return WEEDCU_EINVAL;
  std::vector<float> wide((size_t)rows * V);
  for (size_t i = 0; i < wide.size(); ++i) wide[i] = bf16_widen(logits_bf16[i]);
  return weedcu_cross_entropy_bwd_pack(wide.data(), 0, rows, V, targets, lse, dloss, dlogits, d_offset, accumulate, dlogits_bf16, colsum, stream);
}
// extended epilogue: the plain product, then every extra output from the fp32 values (column tiles of 256)
int weedcu_gemm_bf16_ex(const uint16_t *a, int a_major, uint64_t lda, const uint16_t *b, int b_major, uint64_t ldb, float *c, uint64_t ldc, uint16_t *c_bf16,
                        uint64_t ldc_bf16, uint32_t M, uint32_t N, uint32_t K, const weedcu_gemm_epilogue *epi, void *stream) {
  if (!a || !b || (!c && !c_bf16)) return WEEDCU_EINVAL;
  if (c_bf16 && (ldc_bf16 % 8)) return WEEDCU_ENOSUP;
  if (c && (ldc % 4)) return WEEDCU_ENOSUP;
  std::vector<float> tmp;
  float *out = c;
  uint64_t ld = ldc;
  if (!out) {
    tmp.resize((size_t)M * N);
    out = tmp.data();
    ld = M;
  }
  int rc = weedcu_gemm_bf16(a, a_major, lda, b, b_major, ldb, out, ld, M, N, K, 0, epi ? epi->col_bias : nullptr, stream);
  if (rc || g_nocompute) return rc;
  if (epi && epi->residual)
    for (uint32_t n = 0; n < N; ++n)
      for (uint32_t m = 0; m < M; ++m) out[m + (uint64_t)n * ld] = out[m + (uint64_t)n * ld] + epi->residual[m + (uint64_t)n * epi->ldr];
  if (c_bf16) {
    wo_view v;
    memset(&v, 0, sizeof(v));
    v.rank = 1;
    v.shape[0] = M;
    v.stride[0] = 1;
    std::vector<float> col(M);
    for (uint32_t n = 0; n < N; ++n) {
      const float *src = out + (uint64_t)n * ld;
      if (epi && epi->activation) {
        wo_unary_real(7 /* GELU */, 0.0f, src, &v, col.data(), &v);
        src = col.data();
      }
      for (uint32_t m = 0; m < M; ++m) c_bf16[m + (uint64_t)n * ldc_bf16] = wo_f32_to_bf16(src[m]);
    }
  }
  if (epi && epi->row_stats) {
    const uint32_t cols = 256, tiles = (N + cols - 1) / cols;
    if (!epi->stats || tiles > epi->stats_capacity_tiles) return WEEDCU_EINVAL;
    for (uint32_t t = 0; t < tiles; ++t) {
      const uint32_t n0 = t * cols, n1 = (n0 + cols < N) ? n0 + cols : N;
      for (uint32_t m = 0; m < M; ++m) {
        double a0 = 0.0, a1 = 0.0;
        if (epi->row_stats == 1) {
          for (uint32_t n = n0; n < n1; ++n) a0 += out[m + (uint64_t)n * ld];
          a0 /= (double)(n1 - n0);
          for (uint32_t n = n0; n < n1; ++n) a1 += (out[m + (uint64_t)n * ld] - a0) * (out[m + (uint64_t)n * ld] - a0);
        } else {
          a0 = -1.0 / 0.0;
          for (uint32_t n = n0; n < n1; ++n) a0 = out[m + (uint64_t)n * ld] > a0 ? out[m + (uint64_t)n * ld] : a0;
          for (uint32_t n = n0; n < n1; ++n) a1 += exp((double)out[m + (uint64_t)n * ld] - a0);
        }
        epi->stats[2 * ((uint64_t)t * M + m)] = (float)a0;
        epi->stats[2 * ((uint64_t)t * M + m) + 1] = (float)a1;
      }
    }
    if (epi->stats_tiles) *epi->stats_tiles = tiles;
    if (epi->stats_tile_cols) *epi->stats_tile_cols = tiles == 1 ? N : cols;
  }
  return 0;
}
int weedcu_gemm_bf16_grouped_bf16out(const uint16_t *a, int a_major, uint64_t lda, uint32_t groups, const uint16_t *const *b, int b_major, uint64_t ldb,
                                     uint16_t *const *c_bf16, uint64_t ldc_bf16, uint32_t M, uint32_t N, uint32_t K, const float *const *col_bias, void *stream) {
  if (!groups || groups > 3 || !b || !c_bf16) return WEEDCU_EINVAL;
  for (uint32_t g = 0; g < groups; ++g) {
    weedcu_gemm_epilogue e;
    memset(&e, 0, sizeof(e));
    e.col_bias = col_bias ? col_bias[g] : nullptr;
    const int rc = weedcu_gemm_bf16_ex(a, a_major, lda, b[g], b_major, ldb, nullptr, 0, c_bf16[g], ldc_bf16, M, N, K, &e, stream);
    if (rc) return rc;
  }
  return 0;
}
int weedcu_gemm_workspace_bytes(uint32_t, uint32_t, uint32_t, uint32_t, int, uint64_t *bytes) { if (bytes) *bytes = 0; return 0; }
// collectives: single-process identity (world size 1); the gloo world_size-2 tests patch these from Python
int weedcu_nccl_load(const char *) { return 0; }
int weedcu_nccl_unique_id(void *id128) { if (id128) memset(id128, 0, 128); return 0; }
int weedcu_nccl_init(const void *, int, int, void **comm) { if (comm) *comm = (void *)0x2; return 0; }
int weedcu_nccl_destroy(void *) { return 0; }
int weedcu_nccl_group_start(void) { return 0; }
int weedcu_nccl_group_end(void) { return 0; }
typedef int (*mock_allreduce_hook)(float *buf, uint64_t n);
static mock_allreduce_hook g_allreduce_hook = nullptr, g_bcast_hook = nullptr;
void weedcu_mock_set_collective_hooks(mock_allreduce_hook allreduce, mock_allreduce_hook bcast) { g_allreduce_hook = allreduce; g_bcast_hook = bcast; }
int weedcu_multi_copy(uint32_t count, const float *const *src, float *const *dst, const uint64_t *n, void *) {
  ++g_launches;
  for (uint32_t t = 0; t < count; ++t) memmove(dst[t], src[t], sizeof(float) * (size_t)n[t]);
  return 0;
}
int weedcu_nccl_allreduce_sum(void *, float *buf, uint64_t n, void *) { ++g_launches; return g_allreduce_hook ? g_allreduce_hook(buf, n) : 0; }
int weedcu_nccl_broadcast(void *, float *buf, uint64_t n, int, void *) { ++g_launches; return g_bcast_hook ? g_bcast_hook(buf, n) : 0; }
}
