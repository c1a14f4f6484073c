/*
 * weedcu.h — C-ABI of the B200 (sm_100a) compute layer for Weed's GPU device.
 *
 * This is the drop-in boundary: the entry points a `WEED_ENABLE_CUDA` build of Weed
 * binds where the reference's OpenCL build calls `GpuDevice::RequestKernel(OCLAPI, ...)`
 * (reference: include/devices/gpu_device.hpp:30-287, include/common/oclapi.hpp:18-121).
 * Plain pointers and sizes only; no C++ or torch types cross this line. Every function
 * returns 0 on success, a positive cudaError_t / ncclResult_t (+1000) value on a driver
 * failure, or a negative WEEDCU_E* code on bad arguments. Nothing here throws and nothing
 * here falls back to the host: if the device path cannot run, the call fails.
 *
 * Data model (reference: include/tensors/base_tensor.hpp:25-142):
 *   a tensor is a VIEW (offset, shape[], stride[]) over a flat device buffer of `real`
 *   (= float, WEED_FPPOW=5, include/common/weed_types.hpp:91-95). Layout is column-major:
 *   flat element index i resolves as  offset + sum_d ((i / prod(shape[<d])) % shape[d]) * stride[d]
 *   (BaseTensor::get_storage_index, base_tensor.hpp:123-142). stride 0 = broadcast.
 *   Shapes/strides are `tcapint` = uint32 (WEED_TCAPPOW=5, weed_types.hpp:54-57); byte
 *   offsets are formed in 64-bit on the device.
 *
 * All launches are asynchronous on `stream` (a cudaStream_t passed as void*; NULL selects
 * the library's own compute stream, weedcu_default_stream()).
 */
#ifndef WEEDCU_H
#define WEEDCU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WEEDCU_MAX_RANK 8

#define WEEDCU_OK 0
#define WEEDCU_EINVAL (-1)   /* bad argument (rank, null pointer, shape mismatch) */
#define WEEDCU_ENOSUP (-2)   /* valid request this build cannot serve (e.g. TMA alignment) */
#define WEEDCU_ENCCL  (-3)   /* NCCL library not loaded */

typedef struct weedcu_view {
  uint64_t offset;                   /* BaseTensor::offset */
  int32_t rank;                      /* shape.size() */
  uint32_t shape[WEEDCU_MAX_RANK];   /* BaseTensor::shape  */
  uint32_t stride[WEEDCU_MAX_RANK];  /* BaseTensor::stride */
} weedcu_view;

/* ------------------------------------------------------------------ runtime / memory
 * Replaces OCLEngine device discovery (include/common/oclengine.hpp:249-395) and
 * GpuDevice::MakeBuffer / LockSync / clFinish (src/devices/gpu_device.cpp:34-76,388-447). */
int weedcu_device_count(int *count);
int weedcu_set_device(int device);
int weedcu_get_device(int *device);
int weedcu_device_info(int device, char *name, int name_len, uint64_t *total_mem, int *sm_count,
                       int *cc_major, int *cc_minor);
const char *weedcu_error_string(int code);
void *weedcu_default_stream(void);          /* per-device compute stream, created on demand */
int weedcu_set_default_stream(void *stream);/* adopt an external stream (e.g. torch's) */
int weedcu_stream_create(void **stream);
/* high != 0: the highest stream priority of the device — blocks of kernels queued on it are scheduled ahead of the
 * compute stream's as SMs free up (the communication stream of the data-parallel gradient all-reduce) */
int weedcu_stream_create_priority(void **stream, int high);
int weedcu_stream_destroy(void *stream);
int weedcu_stream_sync(void *stream);       /* GpuDevice::clFinish */
int weedcu_stream_wait_event(void *stream, void *event);
int weedcu_event_create(void **event);
int weedcu_event_destroy(void *event);
int weedcu_event_record(void *event, void *stream);
int weedcu_event_sync(void *event);
int weedcu_event_elapsed_ms(void *start, void *stop, float *ms);
/* Stream-ordered pool allocation (cudaMallocAsync with an unbounded release threshold):
 * freeing while kernels are in flight is safe, like QueueItem holding BufferPtrs
 * (include/devices/queue_item.hpp:27-47). */
int weedcu_malloc(void **ptr, size_t bytes, void *stream);
int weedcu_free(void *ptr, void *stream);
/* weedcu_malloc/free keep freed blocks on per-stream free lists (a training step re-requests the
 * same sizes every iteration); this returns every cached block to the driver. */
int weedcu_pool_trim(void);
int weedcu_mem_info(uint64_t *free_bytes, uint64_t *total_bytes);
int weedcu_host_alloc(void **ptr, size_t bytes); /* pinned staging memory */
int weedcu_host_free(void *ptr);
int weedcu_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream);
int weedcu_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream);
int weedcu_memcpy_d2d(void *dst, const void *src, size_t bytes, void *stream);
int weedcu_launch_count(uint64_t *count);   /* kernels launched by this library so far */
/* Programmatic dependent launch between consecutive kernels of this library on one stream (on by default; env WEEDCU_PDL=0).
 * on = 0 / 1 sets it for the launches that follow; on < 0 only queries. Returns the previous setting. No reference
 * counterpart: the reference's FIFO waits on the host per launch (src/devices/gpu_device.cpp:296-305). */
int weedcu_set_pdl(int on);
int weedcu_host_stats(double *malloc_ms, uint64_t *mallocs, double *free_ms, uint64_t *frees);

/* ------------------------------------------------------------------ F1 fills
 * GpuDevice::ClearRealBuffer / FillOnesReal / FillValueReal (src/devices/gpu_device.cpp:314-386);
 * kernels clear_buffer_real / fill_ones_real / fill_value_real (src/common/qengine.cl:99-134). */
int weedcu_fill_real(float *p, uint64_t n, float value, void *stream);
int weedcu_fill_int(int32_t *p, uint64_t n, int32_t value, void *stream);

/* ------------------------------------------------------------------ E1-E3 elementwise
 * Weed::add / mul (src/ops/commuting.cpp:53-58,87-92,187-192), sub (src/ops/sub.cpp:48-53),
 * div (src/ops/div.cpp:48-53): out[i] = a[i] (op) b[i] over the flat col-major index of the
 * common broadcast shape; every operand resolves i through its own view. The reference's
 * OpenCL launch only forwards stride[0] (SURVEY §2.3 defect 1); this entry is N-D correct
 * like the CPU path. All three views must have equal rank and shape. */
enum { WEEDCU_ADD = 0, WEEDCU_MUL = 1, WEEDCU_SUB = 2, WEEDCU_DIV = 3 };
int weedcu_binary_real(int op, const float *a, const weedcu_view *av, const float *b,
                       const weedcu_view *bv, float *out, const weedcu_view *ov, void *stream);
/* Weed::add_in_place / sub_in_place (src/ops/in_place.cpp:49-54,79-84): a[i] (+/-)= b[i],
 * i over a's broadcast size. */
int weedcu_inplace_real(int op, float *a, const weedcu_view *av, const float *b,
                        const weedcu_view *bv, void *stream);
/* Weed::copy_broadcast (src/ops/copy_broadcast.cpp:44-49): dst[i] = src[i]. */
int weedcu_copy_real(float *dst, const weedcu_view *dv, const float *src, const weedcu_view *sv,
                     void *stream);

/* ------------------------------------------------------------------ U1-U3 unary + grads
 * relu/sigmoid/tanh (src/ops/real_unary.cpp:77-83,132-138,188-194), abs (src/ops/abs.cpp:88-94),
 * pow/exp/log (src/ops/pow.cpp:52-77; `param` = p, log(b), 1/log(b) respectively),
 * gelu = fused Tensor::gelu (src/tensors/tensor.cpp:841-851, tanh approximation). */
enum {
  WEEDCU_RELU = 0, WEEDCU_SIGMOID = 1, WEEDCU_TANH = 2, WEEDCU_ABS = 3, WEEDCU_POW = 4,
  WEEDCU_EXP = 5, WEEDCU_LOG = 6, WEEDCU_GELU = 7, WEEDCU_SIN = 8, WEEDCU_COS = 9
};
int weedcu_unary_real(int op, float param, const float *a, const weedcu_view *av, float *out,
                      const weedcu_view *ov, void *stream);
/* din[i] += f'(.) * dout[i]. `in` is the forward INPUT for relu/abs/gelu/sin/cos and the forward
 * OUTPUT for sigmoid/tanh (src/ops/real_unary.cpp:47-76; src/tensors/tensor.cpp:867-935).
 * accumulate == 0 stores instead of adding (din is known to be all zeros and is not read): the
 * host's lazily zero-filled gradients use it to skip the fill and the read (16 -> 12 B/elem). */
int weedcu_unary_grad_real(int op, float *din, const weedcu_view *dinv, const float *in,
                           const weedcu_view *inv, const float *dout, const weedcu_view *doutv,
                           int accumulate, void *stream);

/* Tensor::gelu forward on a dense tensor (tensor.cpp:841-851, as weedcu_unary_real(WEEDCU_GELU)) that also
 * writes y_bf16[i] = bf16(y[i]): the GEMM operand of the Linear that follows (ff2). n % 4 == 0,
 * 16-byte aligned x / y, 8-byte aligned y_bf16; otherwise WEEDCU_ENOSUP. y may be NULL (bf16 copy only:
 * 10 -> 6 B/elem; weedcu_unary_real(WEEDCU_GELU) gives the fp32 values when something needs them). */
int weedcu_gelu_fwd_bf16(const float *x, float *y, uint16_t *y_bf16, uint64_t n, void *stream);
/* gelu_grad (tensor.cpp:841-851 backward; as weedcu_unary_grad_real(WEEDCU_GELU)) on a dense [rows, cols]
 * matrix with rows contiguous, fused with the preparation of the Linear backward that consumes din:
 * din (+)= dout * gelu'(in), din_bf16[i] = bf16(din[i]), colsum[c] = sum_r din[r, c] (stored; fixed order).
 * WEEDCU_ENOSUP unless rows % 8 == 0 and all buffers are 16-byte aligned.
 * din may be NULL when accumulate == 0 (bf16 copy and column sums only; weedcu_unary_grad_real gives the fp32 values). */
int weedcu_gelu_grad_pack(float *din, const float *in, const float *dout, uint32_t rows, uint32_t cols,
                          int accumulate, uint16_t *din_bf16, float *colsum, void *stream);
/* weedcu_gelu_grad_pack with dout given as its bf16 copy ([rows, cols], rows contiguous): the case where the product that
 * formed dout (the dA of the Linear behind the activation, matmul backward tensor.cpp:1105-1136) wrote only that copy
 * because this node is its only reader. Same outputs. */
int weedcu_gelu_grad_pack_bf16dy(float *din, const float *in, const uint16_t *dout_bf16, uint32_t rows, uint32_t cols,
                                 int accumulate, uint16_t *din_bf16, float *colsum, void *stream);

/* ------------------------------------------------------------------ R1-R2 reductions
 * Weed::reduce (src/ops/reduce.cpp:17-38,60-66): out[o] = sum_j a[base(o) + j*stride[axis]];
 * `a` must be contiguous (Tensor::sum makes it so, src/tensors/tensor.cpp:626), out is a dense
 * buffer of prod(shape)/shape[axis] elements.
 * index_order 0: o enumerates the non-axis coordinates column-major, i.e. the layout the
 *   output tensor built by Tensor::sum (tensor.cpp:629-645) is read with (intended semantics).
 * index_order 1: reproduces the reference CPU loop bit-for-bit: o is decomposed LAST dimension
 *   fastest (REDUCE_HEAD, reduce.cpp:17-31), which permutes the output whenever two or more
 *   non-axis dims exceed 1 (see DESIGN.md "Reference defects"). */
int weedcu_reduce_real(const float *a, const weedcu_view *av, int axis, float *out,
                       int index_order, void *stream);
/* Weed::reduce_grad (src/ops/reduce.cpp:84-113): din[i] += dout[o(i)], dout broadcast along axis.
 * index_order as above (REDUCE_GRAD_HEAD, reduce.cpp:84-99). */
int weedcu_reduce_grad_real(float *din, const weedcu_view *dinv, const float *dout,
                            const weedcu_view *doutv, int axis, int index_order, void *stream);
/* Weed::clamp / clamp_grad (reference src/ops/clamp.cpp:67-73,55-60; OpenCL kernels `clamp_real` / `clamp_grad_real`,
 * src/common/qengine.cl): out = min(max(a, lo), hi); din += dout where lo < in < hi. Views as weedcu_unary_real. */
int weedcu_clamp_real(const float *a, const weedcu_view *av, float lo, float hi, float *out, const weedcu_view *ov, void *stream);
int weedcu_clamp_grad_real(float *din, const weedcu_view *dinv, const float *in, const weedcu_view *inv, const float *dout,
                           const weedcu_view *doutv, float lo, float hi, void *stream);
/* Weed::max / Weed::min over the whole tensor (reference src/ops/real_extremum.cpp:88-106,151-163: the reference copies the
 * buffer to the host for this): *out = extremum of the view's elements, two-pass device reduction. is_min: 0 max, 1 min. */
int weedcu_extremum_real(int is_min, const float *a, const weedcu_view *av, float *out, void *stream);
/* their backward (real_extremum.cpp:50-57; OpenCL `match_grad_real`): din += dout where in == *extremum (device scalar);
 * dout is the scalar gradient broadcast over din's shape (all strides 0) or any view of the same shape. */
int weedcu_match_grad_full_real(float *din, const weedcu_view *dinv, const float *in, const weedcu_view *inv, const float *dout,
                                const weedcu_view *doutv, const float *extremum, void *stream);
/* Weed::max / Weed::min along one axis (reference src/ops/reduce.cpp:40-58,239-248,362-386) with the same output layout and
 * index_order switch as weedcu_reduce_real, and Weed::match_grad (:84-101,263-277,433-449): din += dout[o] where
 * in == reduced[o], o = the output element the input element was reduced into; `reduced` is read through doutv too. */
int weedcu_extremum_axis_real(int is_min, const float *a, const weedcu_view *av, int axis, float *out, int index_order, void *stream);
int weedcu_match_grad_real(float *din, const weedcu_view *dinv, const float *in, const weedcu_view *inv, const float *dout,
                           const weedcu_view *doutv, const float *reduced, int axis, int index_order, void *stream);
/* Weed::sum / mean (src/ops/sum.cpp:74-98): *out = scale * sum_i a[i]. Device-side (the
 * reference copies the buffer to the host, sum.cpp:52-67). Deterministic two-pass tree. */
int weedcu_sum_real(const float *a, const weedcu_view *av, float scale, float *out, void *stream);

/* ------------------------------------------------------------------ A2 KV-cache decode
 * MultiHeadAttention::forward with use_kv_cache and kv_quant_bits = 0
 * (src/modules/multihead_attention.cpp:169-199, 278-287, 313-345), for q, k, v, out = [B, T_new, H*hd]
 * column-major (the W_q / W_k / W_v outputs; b fastest) and float caches [B, H, S, hd] column-major
 * (element (b,h,s,j) at b + B*(h + H*(s + S*j)), zero-initialised by the caller):
 *   1. k_cache[b,h,cache_len+t,j] += k[b,t,h+H*j], same for v      (add_in_place into the slot, :282-283)
 *   2. x[b,h,t,s] = q.k_cache / divisor, s < cache_len + T_new; when causal and T_new > 1, + mask_val
 *      where t + 1 <= s — the [T_q, T_k] triu_fill(diagonal 1) mask of :322-328, whose query index
 *      counts from this call's first token (reference behaviour, reproduced)
 *   3. softmax over s, out[b,t,h+H*j] = sum_s p * v_cache[b,h,s,j]
 * (feature c = h + H*j: the column-major reshape to {B, T, H, hd} of :155-157 makes H the fast extent)
 * The caches are read in place (the reference copies both contiguous on every call).
 * WEEDCU_ENOSUP for hd > 64 (callers compose the generic ops); WEEDCU_EINVAL when cache_len + T_new > S. */
int weedcu_attention_decode(const float *q, const float *k, const float *v, float *k_cache, float *v_cache,
                            float *out, uint32_t B, uint32_t T_new, uint32_t H, uint32_t hd, uint32_t S,
                            uint32_t cache_len, float divisor, float mask_val, int causal, void *stream);

/* ------------------------------------------------------------------ S1-S2 softmax family
 * Weed::softmax / softmax_grad (src/ops/softmax.cpp:85-138), logsoftmax / logsoftmax_grad
 * (src/ops/logsoftmax.cpp:87-149). Rows run along `axis`; all views share one shape.
 * grad: din += out*(dout - sum(dout*out))   |   din += dout - exp(out)*sum(dout). */
int weedcu_softmax_real(int log_mode, const float *a, const weedcu_view *av, int axis, float *out,
                        const weedcu_view *ov, void *stream);
int weedcu_softmax_grad_real(int log_mode, float *din, const weedcu_view *dinv, const float *out,
                             const weedcu_view *ov, const float *dout, const weedcu_view *doutv,
                             int axis, void *stream);
/* Fused causal attention probabilities: out = softmax(scores/divisor + triu_mask(mask_val), key
 * axis) — the div + triu_fill + add + softmax chain of MultiHeadAttention::forward
 * (src/modules/multihead_attention.cpp:319-334) in one pass (out may alias scores).
 * batch_fastest 1: scores[batch, Tq, Tk] column-major (strides 1, batch, batch*Tq), the layout
 *                  the reference's [B,H,Tq,Tk] scores tensor has;
 * batch_fastest 0: one contiguous [Tq, Tk] column-major matrix per batch (strides 1, Tq, Tq*Tk),
 *                  the layout the fused attention path keeps its per-head products in.
 * Masked where q + 1 <= k (triu_fill diagonal 1, src/ops/triu_fill.cpp:48-56) when causal != 0. */
int weedcu_attn_softmax_real(const float *scores, float *out, uint32_t batch, uint32_t Tq,
                             uint32_t Tk, float divisor, float mask_val, int causal,
                             int batch_fastest, void *stream);
/* Fused attention core on bf16 tensor cores for q, k, v, out = [B, T, H*hd] column-major (b fastest;
 * the layout Linear::forward leaves them in; feature c belongs to head h = c % H, component
 * j = c / H — the column-major reshape to {B, T, H, hd} of multihead_attention.cpp:155-157 makes H
 * the fast extent): per (b, h)  S = Q K^T / divisor (+ mask_val where
 * q + 1 <= k when causal), P = softmax_k(S), O = P V — the chain of MultiHeadAttention::forward,
 * src/modules/multihead_attention.cpp:289-345, without its transposing copies. Q, K, V and P are
 * rounded to bf16, accumulation is fp32; like the reference's batched matmul (tensor.cpp:1253-1271)
 * the result carries no autograd edge. head_dim 64 runs a flash-style kernel (scores and
 * probabilities never leave the SM: tcgen05 + TMEM, online softmax); other head sizes store S and
 * P (bf16) in a workspace. Returns WEEDCU_ENOSUP for shapes outside T % 8 == 0, T >= 64,
 * hd % 8 == 0, hd >= 16 (and T <= 1024 unless hd == 64); callers then compose the generic ops. */
int weedcu_attention_fwd(const float *q, const float *k, const float *v, float *out, uint32_t B,
                         uint32_t T, uint32_t H, uint32_t hd, float divisor, float mask_val,
                         int causal, void *stream);
/* The same entry that also writes out_bf16[i] = bf16(out[i]) (RNE) at the same linear index: the A operand
 * of the W_o product that follows (no pack pass). out_bf16 == NULL is weedcu_attention_fwd;
 * WEEDCU_ENOSUP additionally when B % 4 != 0 or out is not 16-byte aligned. */
int weedcu_attention_fwd_bf16out(const float *q, const float *k, const float *v, float *out, uint16_t *out_bf16,
                                 uint32_t B, uint32_t T, uint32_t H, uint32_t hd, float divisor, float mask_val,
                                 int causal, void *stream);
/* weedcu_attention_fwd_bf16out whose q, k, v arrive as bf16 [B, T, H*hd] column-major (the operand copies the grouped
 * W_q / W_k / W_v product wrote, weedcu_gemm_bf16_grouped_bf16out): the head relayout reads 2 B/elem instead of 4 and the
 * fp32 projections are never written. Same rounding points as the fp32-input entry (q, k, v are rounded to bf16 there
 * too). WEEDCU_ENOSUP unless B % 8 == 0 and the pointers are 16-byte aligned. */
int weedcu_attention_fwd_bf16in(const uint16_t *q_bf16, const uint16_t *k_bf16, const uint16_t *v_bf16, float *out, uint16_t *out_bf16,
                                uint32_t B, uint32_t T, uint32_t H, uint32_t hd, float divisor, float mask_val, int causal, void *stream);
/* Fused cross-entropy over logits[rows, V] (row stride rs, vocab stride vs):
 * cross_entropy_loss (include/autograd/cross_entropy_loss.hpp:21-34) = -mean_rows lsm[row, target].
 * fwd writes per-row log-sum-exp (lse[rows]) and the scalar loss; bwd does
 * dlogits += (softmax - onehot) * (dloss / rows). Targets are int32, gathered exactly. */
int weedcu_cross_entropy_fwd(const float *logits, uint64_t offset, uint32_t rows, uint32_t V,
                             uint32_t rs, uint32_t vs, const int32_t *targets, float *lse,
                             float *loss, void *stream);
int weedcu_cross_entropy_bwd(const float *logits, uint64_t offset, uint32_t rows, uint32_t V,
                             uint32_t rs, uint32_t vs, const int32_t *targets, const float *lse,
                             const float *dloss, float *dlogits, uint64_t d_offset, int accumulate,
                             void *stream); /* accumulate == 0: dlogits = ... (not read) */

/* cross_entropy backward fused with the preparation of the Linear backward that consumes dlogits
 * (Tensor::matmul_backward packs dY for its two tensor-core GEMMs and the bias node sums its columns,
 * tensor.cpp:1105-1136,1361-1400): one pass writes dlogits (fp32, exactly as weedcu_cross_entropy_bwd
 * with rs = 1, vs = rows), dlogits_bf16[v*rows + r] = bf16(dlogits) and colsum[v] = sum_r dlogits[r,v]
 * (stored, not accumulated; deterministic order). 10-14 B/elem instead of 8-12 + 6 + 4.
 * WEEDCU_ENOSUP unless rows % 8 == 0 and all buffers are 16-byte aligned.
 * dlogits may be NULL when accumulate == 0: only the bf16 operand copy and the column sums are written
 * (10 -> 6 B/elem); the caller then owns producing the fp32 values on demand (weedcu_cross_entropy_bwd). */
int weedcu_cross_entropy_bwd_pack(const float *logits, uint64_t offset, uint32_t rows, uint32_t V,
                                  const int32_t *targets, const float *lse, const float *dloss,
                                  float *dlogits, uint64_t d_offset, int accumulate, uint16_t *dlogits_bf16,
                                  float *colsum, void *stream);
/* cross_entropy_loss forward (include/autograd/cross_entropy_loss.hpp:21-34) from the log-sum-exp partials the LM head's
 * GEMM epilogue left (weedcu_gemm_bf16_ex, row_stats 2; stats[t][row] = (max, sum exp) of column tile t) — the logits are not
 * read at all. The target logit of each row is recomputed as the same bf16 x bf16 -> fp32 dot product (+ col_bias) from
 * the GEMM's operands a [rows, K] / b [V, K] (majors and leading dimensions as in weedcu_gemm_bf16). lse[row], *loss as
 * weedcu_cross_entropy_fwd. */
int weedcu_cross_entropy_fwd_stats(const float *stats, uint32_t tiles, uint32_t rows, uint32_t V, const uint16_t *a, int a_major, uint64_t lda,
                                   const uint16_t *b, int b_major, uint64_t ldb, uint32_t K, const float *col_bias, const int32_t *targets,
                                   float *lse, float *loss, void *stream);
/* cross_entropy_loss forward reading the bf16 copy of the logits ([rows, V] dense, rows contiguous, rows % 8 == 0; what
 * weedcu_gemm_bf16_ex leaves when the fp32 logits are not written): 2 B/elem instead of 4. The target logit of each row is
 * recomputed exactly from the product's operands as in weedcu_cross_entropy_fwd_stats. */
int weedcu_cross_entropy_fwd_bf16in(const uint16_t *logits_bf16, uint32_t rows, uint32_t V, const uint16_t *a, int a_major, uint64_t lda,
                                    const uint16_t *b, int b_major, uint64_t ldb, uint32_t K, const float *col_bias, const int32_t *targets,
                                    float *lse, float *loss, void *stream);
/* weedcu_cross_entropy_bwd_pack reading the bf16 copy of the logits ([rows, V] dense, rows contiguous) instead of fp32. */
int weedcu_cross_entropy_bwd_pack_bf16in(const uint16_t *logits_bf16, uint32_t rows, uint32_t V, const int32_t *targets, const float *lse,
                                         const float *dloss, float *dlogits, uint64_t d_offset, int accumulate, uint16_t *dlogits_bf16,
                                         float *colsum, void *stream);

/* ------------------------------------------------------------------ L1 LayerNorm (fused)
 * LayerNorm::forward (src/modules/layernorm.cpp:29-42): x[rows, F] with row stride 1 and
 * feature stride `rows` (last axis is slowest in col-major). y = (x-mean)/sqrt(var+eps)*gamma+beta,
 * biased variance. Saves mean[rows], rstd[rows] for the backward. */
int weedcu_layernorm_fwd(const float *x, uint32_t rows, uint32_t F, const float *gamma,
                         const float *beta, float eps, float *y, float *mean, float *rstd,
                         void *stream);
/* The same forward that also writes y_bf16[i] = bf16(y[i]) (RNE) at the same linear index: for rows % 8 == 0
 * that is exactly the tensor-core GEMM operand of the Linear that consumes y, so no fp32 -> bf16 pack
 * pass follows (10 B/elem instead of 8 + 6). WEEDCU_ENOSUP (nothing done) for rows % 8 != 0, rows <= 256
 * or unaligned buffers. y_bf16 == NULL is weedcu_layernorm_fwd. */
int weedcu_layernorm_fwd_bf16(const float *x, uint32_t rows, uint32_t F, const float *gamma,
                              const float *beta, float eps, float *y, float *mean, float *rstd,
                              uint16_t *y_bf16, void *stream);
/* weedcu_layernorm_fwd_bf16 for an x whose row statistics already exist as per-column-tile partials (mean_t, M2_t) left by
 * the epilogue of the GEMM that produced x (weedcu_gemm_bf16_ex, row_stats 1; tile t covers features
 * [t * tile_cols, min(F, (t + 1) * tile_cols))): one pass over x instead of two. Same result as the two-pass kernel up to
 * fp32 rounding of the merge (LayerNorm::forward, src/modules/layernorm.cpp:29-42). mean / rstd / y_bf16 may be NULL.
 * WEEDCU_ENOSUP unless rows % 4 == 0 and the pointers are 16-byte aligned. */
int weedcu_layernorm_fwd_stats(const float *x, uint32_t rows, uint32_t F, const float *stats, uint32_t tiles, uint32_t tile_cols,
                               const float *gamma, const float *beta, float eps, float *y, float *mean, float *rstd, uint16_t *y_bf16,
                               void *stream);
/* dx += ..., dgamma[F] += sum_rows dy*xhat, dbeta[F] += sum_rows dy.
 * grad_mode 0 reproduces what the reference's autograd chain computes: its div node omits dout on
 *   the denominator branch (src/tensors/tensor.cpp:1506-1521), so the variance path contributes
 *   only  xc * (-rstd^3/F) * sum_f(xc)  (rounding-level):  dxc = g*rstd + that;  dx += dxc - mean_f(dxc).
 * grad_mode 1 is the analytic LayerNorm gradient: dx += rstd*(g - mean_f(g) - xhat*mean_f(g*xhat)).
 * (g = dy*gamma, xc = x-mean, xhat = xc*rstd)
 * accumulate == 0: dx is stored, not added to (dgamma / dbeta always accumulate). */
int weedcu_layernorm_bwd(const float *x, const float *dy, uint32_t rows, uint32_t F,
                         const float *gamma, const float *mean, const float *rstd, float *dx,
                         float *dgamma, float *dbeta, int grad_mode, int accumulate, void *stream);
/* weedcu_layernorm_bwd with the old values of dx read from a second buffer: dx = dx_in + (LayerNorm input gradient);
 * dx_in == dx accumulates in place, dx_in == NULL stores. Lets a gradient that shares another tensor's buffer
 * copy-on-write be accumulated into without being copied first (12 + 4 B/elem either way). */
int weedcu_layernorm_bwd_from(const float *x, const float *dy, uint32_t rows, uint32_t F, const float *gamma,
                              const float *mean, const float *rstd, const float *dx_in, float *dx, float *dgamma,
                              float *dbeta, int grad_mode, void *stream);

/* ------------------------------------------------------------------ M1-M2 embedding, mask
 * Weed::embedding_gather / embedding_scatter_add (src/ops/embedding.cpp:56-110):
 * out[i + d*o_s1] = W[w_off + tok_i*w_s0 + d*w_s1]; dW[...] += dout[...] (atomic on duplicates).
 * Indices are int32 `symint`, used exactly. */
int weedcu_embedding_gather(const int32_t *idx, uint64_t idx_off, uint32_t idx_stride, uint32_t n,
                            const float *W, uint64_t w_off, uint32_t w_s0, uint32_t w_s1,
                            uint32_t D, float *out, uint64_t o_off, uint32_t o_s0, uint32_t o_s1,
                            void *stream);
int weedcu_embedding_scatter_add(float *dW, uint64_t w_off, uint32_t w_s0, uint32_t w_s1,
                                 const int32_t *idx, uint64_t idx_off, uint32_t idx_stride,
                                 uint32_t n, uint32_t D, const float *dout, uint64_t o_off,
                                 uint32_t o_s0, uint32_t o_s1, void *stream);
/* Weed::triu_fill (src/ops/triu_fill.cpp:41-59): a[i,j] = val where i + diagonal <= j. */
int weedcu_triu_fill_real(float *a, const weedcu_view *av, float val, uint32_t diagonal,
                          void *stream);
/* argmax over the last axis of logits[rows,V] (greedy decode; the reference only has axis max,
 * src/ops/reduce.cpp:40-49). Ties resolve to the lowest index. */
int weedcu_argmax_rows(const float *x, uint64_t offset, uint32_t rows, uint32_t V, uint32_t rs,
                       uint32_t vs, int32_t *out, void *stream);

/* ------------------------------------------------------------------ O1-O3 optimisers
 * sgd_step (include/autograd/sgd.hpp:23-37): p -= lr * (gscale*g).
 * adam_step (include/autograd/adam.hpp:70-106):
 *   m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g*g ; p -= lr*m / (bc1*(sqrt(v/bc2)+eps)).
 * gscale folds the 1/world_size of data-parallel gradient averaging into the read of g. */
int weedcu_sgd_step(float *p, const float *g, uint64_t n, float lr, float gscale, void *stream);
int weedcu_adam_step(float *p, const float *g, float *m, float *v, uint64_t n, float lr,
                     float beta1, float beta2, float eps, float bc1, float bc2, float gscale,
                     void *stream);
/* The same update for `count` parameters in ONE launch (adam_step walks a parameter list,
 * adam.hpp:70-106; a 12-layer transformer has ~150 of them, most of them tiny). Host arrays of
 * device pointers and sizes; they are copied before the call returns. g[i] == NULL stands for an
 * all-zero gradient (a parameter nothing back-propagated into: no fill, no read). */
int weedcu_adam_step_multi(uint32_t count, float *const *p, const float *const *g, float *const *m,
                           float *const *v, const uint64_t *n, float lr, float beta1, float beta2,
                           float eps, float bc1, float bc2, float gscale, void *stream);

/* The multi-tensor update that also refreshes the bf16 GEMM operand of a weight: shadow[i] (or the
 * whole array) may be NULL; otherwise it receives bf16(p) (RNE) at the same linear index, so the
 * next forward pass needs no fp32 -> bf16 pack of the weights (+2 B/param instead of a 6 B/param pass). */
int weedcu_adam_step_multi_shadow(uint32_t count, float *const *p, const float *const *g, float *const *m,
                                  float *const *v, const uint64_t *n, uint16_t *const *shadow, float lr,
                                  float beta1, float beta2, float eps, float bc1, float bc2, float gscale,
                                  void *stream);
/* weedcu_adam_step_multi_shadow that also performs zero_grad (include/autograd/zero_grad.hpp:21-25) for the parameters
 * with zero_grad[t] != 0: their gradient is overwritten with zeros right after it is read (+4 B/param on an existing
 * pass instead of one fill launch per gradient in the next step). g[t] is written in that case. zero_grad may be NULL. */
int weedcu_adam_step_multi_zero(uint32_t count, float *const *p, const float *const *g, float *const *m, float *const *v,
                                const uint64_t *n, uint16_t *const *shadow, const uint8_t *zero_grad, float lr, float beta1, float beta2,
                                float eps, float bc1, float bc2, float gscale, void *stream);

/* ------------------------------------------------------------------ G1-G4 matmul
 * Weed::matmul (src/ops/matmul.cpp:242-279; dims :95-122): C[M,N] (+)= A[M,K] * B[K,N], every
 * operand an arbitrary (offset, s0, s1) view. `batch` > 1 runs independent products with
 * per-operand batch strides (the host loop of Tensor::matmul, src/tensors/tensor.cpp:1259-1269).
 * accumulate != 0 adds into C (fuses the tmp + add_in_place of make_matmul_node,
 * tensor.cpp:1361-1400).
 * precision: WEEDCU_GEMM_FP32  — FFMA, fp32 in / fp32 accumulate (FpMath parity path)
 *            WEEDCU_GEMM_BF16  — operands rounded to bf16, tcgen05.mma kind::f16 with fp32
 *                                accumulators in TMEM, operands staged by TMA
 *            WEEDCU_GEMM_TF32X3 — 3xTF32 split on tcgen05 kind::tf32 (fp32-accurate to ~1e-6) */
enum { WEEDCU_GEMM_FP32 = 0, WEEDCU_GEMM_BF16 = 1, WEEDCU_GEMM_TF32X3 = 2 };
typedef struct weedcu_mat {
  uint64_t offset;      /* element offset of [0,0] of batch 0 */
  uint32_t s0, s1;      /* element strides of the row / column index */
  uint64_t batch_stride;
} weedcu_mat;
int weedcu_matmul_real(const float *a, const weedcu_mat *am, const float *b, const weedcu_mat *bm,
                       float *c, const weedcu_mat *cm, uint32_t M, uint32_t K, uint32_t N,
                       uint32_t batch, int accumulate, int precision, void *stream);
/* Weed::matmul for M <= 16 rows (a decode step: Linear::forward on B tokens), fp32 FFMA, any operand
 * strides, optional dense bias[N] added to every row (Linear's `y + bias`), optional C +=. The weight
 * matrix is read once: HBM-bound on 4*K*N bytes. WEEDCU_ENOSUP for M > 16. */
int weedcu_matmul_skinny(const float *a, const weedcu_mat *am, const float *b, const weedcu_mat *bm, float *c,
                         const weedcu_mat *cm, uint32_t M, uint32_t K, uint32_t N, const float *bias,
                         int accumulate, void *stream);
/* `groups` (1..3) skinny products A * B_g (+ bias_g) -> C_g that share A, M, K, N and the strides of bm / cm (whose
 * offsets are added to every b[g] / c[g]) as ONE launch: the W_q / W_k / W_v projections of a decode step
 * (multihead_attention.cpp:151-153). Same arithmetic as `groups` weedcu_matmul_skinny calls; C is stored. */
int weedcu_matmul_skinny_grouped(const float *a, const weedcu_mat *am, uint32_t groups, const float *const *b,
                                 const weedcu_mat *bm, float *const *c, const weedcu_mat *cm, uint32_t M, uint32_t K,
                                 uint32_t N, const float *const *bias, void *stream);
/* weedcu_matmul_skinny with a residual laid out like C (same cm, including its offset): C = (A * B + bias) + residual,
 * stored — the `x + Linear(...)` of a transformer block on a decode step. bias may be NULL. */
int weedcu_matmul_skinny_residual(const float *a, const weedcu_mat *am, const float *b, const weedcu_mat *bm, float *c,
                                  const weedcu_mat *cm, uint32_t M, uint32_t K, uint32_t N, const float *bias,
                                  const float *residual, void *stream);
/* bf16 tensor-core GEMM on operands already held in bf16 (raw uint16 bit patterns).
 * a_major / b_major: 0 = K contiguous, 1 = M (resp. N) contiguous; lda/ldb are the strides
 * (in elements) of the non-contiguous index. C is fp32, column-major with leading dim ldc.
 * col_bias (optional, may be NULL): [N] fp32 added to every row in the epilogue — the bias add of
 * Linear::forward (src/modules/linear.cpp) without a second pass over C. */
int weedcu_gemm_bf16(const uint16_t *a, int a_major, uint64_t lda, const uint16_t *b, int b_major,
                     uint64_t ldb, float *c, uint64_t ldc, uint32_t M, uint32_t N, uint32_t K,
                     int accumulate, const float *col_bias, void *stream);
/* `groups` (1..3) products A * B_g -> C_g (+ col_bias_g) that share the A operand and all dimensions
 * (the W_q / W_k / W_v projections of one activation, multihead_attention.cpp:151-153) as ONE launch
 * of the same kernel: bit-identical to `groups` weedcu_gemm_bf16 calls, one prologue instead of three.
 * b, c, col_bias: host arrays of `groups` device pointers (col_bias or its entries may be NULL). */
int weedcu_gemm_bf16_grouped(const uint16_t *a, int a_major, uint64_t lda, uint32_t groups,
                             const uint16_t *const *b, int b_major, uint64_t ldb, float *const *c, uint64_t ldc,
                             uint32_t M, uint32_t N, uint32_t K, int accumulate, const float *const *col_bias,
                             void *stream);
/* weedcu_gemm_bf16 with a residual: C = (A * B + col_bias) + residual, residual [M, N] fp32 column-major with leading
 * dimension ldr, added in the epilogue in that order — the `x + Linear(...)` of a transformer block
 * (src/modules/transformer_encoder_layer.cpp:63-125: Linear::forward, then Tensor::add) as one kernel: bit-identical to
 * weedcu_gemm_bf16 followed by weedcu_binary_real(ADD), without writing and re-reading the Linear output. C is stored
 * (no accumulate); WEEDCU_ENOSUP when C does not meet the TMA store rules. col_bias may be NULL. */
int weedcu_gemm_bf16_residual(const uint16_t *a, int a_major, uint64_t lda, const uint16_t *b, int b_major,
                              uint64_t ldb, float *c, uint64_t ldc, uint32_t M, uint32_t N, uint32_t K,
                              const float *col_bias, const float *residual, uint64_t ldr, void *stream);
/* Extended epilogue of the tensor-core GEMM: everything the consumers of a Linear output read is left by the product's own
 * epilogue, from the accumulator registers, instead of by extra passes over C.
 *   c        fp32 output [M, N] column-major (leading dimension ldc), or NULL when only the bf16 copy is wanted
 *   c_bf16   bf16 copy of the output [M, N] column-major (leading dimension ldc_bf16, a multiple of 8), or NULL: the GEMM
 *            operand copy (Bf16 shadow) of C that the next Linear / the attention relayout reads
 *   epi->col_bias, epi->residual / ldr: as weedcu_gemm_bf16 / weedcu_gemm_bf16_residual (Linear::forward's bias add,
 *            src/modules/linear.cpp:86-100; the `x + Linear(...)` of transformer_encoder_layer.cpp:63-125)
 *   epi->activation 1: c_bf16 = bf16(gelu(value)) while c keeps the pre-activation (Tensor::gelu, src/tensors/tensor.cpp:841-851,
 *            of a value that is rounded to bf16 right after: hardware tanh, 2^-11 relative)
 *   epi->row_stats 1: stats[t][m] = (mean, M2 = sum (x - mean)^2) of row m over the columns of column tile t — the partials of
 *            LayerNorm::forward's two means (src/modules/layernorm.cpp:29-42), merged by weedcu_layernorm_fwd_stats
 *   epi->row_stats 2: stats[t][m] = (max, sum exp(x - max)) of row m over tile t — the log-sum-exp partials of
 *            cross_entropy_loss (include/autograd/cross_entropy_loss.hpp:21-34), merged by weedcu_cross_entropy_fwd_stats
 *   *epi->stats_tiles / *epi->stats_tile_cols (host, written before the call returns): number of column tiles and their
 *            width; partial t covers columns [t * cols, min(N, (t + 1) * cols)) (empty when t * cols >= N: it then holds
 *            (0, 0) resp. (-inf, 0)). stats holds stats_capacity_tiles * M * 2 floats (>= 2 * ceil(N / 128) is always enough: the
 *            kernel leaves one partial per half of a >= 128-wide column tile).
 * C is stored (no accumulate, no split-K). WEEDCU_ENOSUP when the outputs do not meet the TMA store rules. */
typedef struct weedcu_gemm_epilogue {
  const float *col_bias;
  const float *residual;
  uint64_t ldr;
  int activation;
  int row_stats;
  float *stats;
  uint32_t stats_capacity_tiles;
  uint32_t *stats_tiles;
  uint32_t *stats_tile_cols;
} weedcu_gemm_epilogue;
int weedcu_gemm_bf16_ex(const uint16_t *a, int a_major, uint64_t lda, const uint16_t *b, int b_major, uint64_t ldb, float *c,
                        uint64_t ldc, uint16_t *c_bf16, uint64_t ldc_bf16, uint32_t M, uint32_t N, uint32_t K,
                        const weedcu_gemm_epilogue *epi, void *stream);
/* weedcu_gemm_bf16_grouped whose outputs leave the kernel as bf16 only (c_bf16[g]: [M, N] column-major, leading dimension
 * ldc_bf16): the W_q / W_k / W_v projections, whose only reader is the attention core's head relayout
 * (src/modules/multihead_attention.cpp:151-157). */
int weedcu_gemm_bf16_grouped_bf16out(const uint16_t *a, int a_major, uint64_t lda, uint32_t groups, const uint16_t *const *b,
                                     int b_major, uint64_t ldb, uint16_t *const *c_bf16, uint64_t ldc_bf16, uint32_t M, uint32_t N,
                                     uint32_t K, const float *const *col_bias, void *stream);
/* Tile family of the bf16 tensor-core GEMM: 0 = cost model over single-CTA 128 x N tiles and CTA-pair
 * (tcgen05 cta_group::2, two SMs of a TPC on one 256 x N tile) kernels (default; env WEEDCU_GEMM_MODE),
 * 1 = single-CTA tiles only, 2 = CTA pairs wherever the operands allow,
 * pair * 1000000 + BLOCK_N * 1000 + splits = one forced configuration (tuning sweeps). Results of the
 * families differ only by fp32 summation order across split-K slices. */
int weedcu_gemm_set_mode(int mode);
/* Dynamic tile scheduling in the CTA-pair GEMM kernel (env WEEDCU_GEMM_DYNAMIC): work units are drawn from a per-launch
 * counter instead of being strided over the clusters, so that a cluster whose SMs are busy with another stream's kernel
 * (the data-parallel all-reduce) does not hold a fixed share of the tiles. Same results: a unit is computed the same way by
 * whichever cluster takes it. */
int weedcu_gemm_set_dynamic(int on);
/* strided fp32 -> packed bf16 (round-to-nearest-even); dst is a dense [rows, cols] matrix whose
 * contiguous index is chosen by dst_major (0: cols contiguous, 1: rows contiguous). */
int weedcu_pack_bf16(const float *src, uint64_t offset, uint32_t s0, uint32_t s1, uint32_t rows,
                     uint32_t cols, uint16_t *dst, int dst_major, void *stream);
/* weedcu_pack_bf16 that also writes, for every index of the NON-contiguous dst dimension, the fp32 sum
 * over the contiguous one: colsum[j] (+)= sum_i src[i, j]. Packing dY [rows = B*T, cols = N] for the
 * backward GEMMs yields the bias gradient (the reduce node of `y + bias`, tensor.cpp:1105-1136) in
 * the same pass. WEEDCU_ENOSUP unless the source is contiguous along the packed dimension. */
int weedcu_pack_bf16_colsum(const float *src, uint64_t offset, uint32_t s0, uint32_t s1, uint32_t rows,
                            uint32_t cols, uint16_t *dst, int dst_major, float *colsum, int accumulate,
                            void *stream);
int weedcu_gemm_workspace_bytes(uint32_t M, uint32_t K, uint32_t N, uint32_t batch,
                                int precision, uint64_t *bytes);

/* ------------------------------------------------------------------ per-kernel-class timing
 * Measurement support for bench.py's `roofline` object: when enabled, every launch of the classes
 * below is bracketed by CUDA events on its own stream; weedcu_prof_read() synchronises and returns
 * the summed device time, launch count and algorithmic work (FLOP for the GEMM classes, bytes for
 * the rest, SURVEY §8d figures). Off by default: the normal hot path records nothing. */
enum {
  WEEDCU_PROF_GEMM_TC = 1, WEEDCU_PROF_GEMM_F32 = 2, WEEDCU_PROF_PACK = 3, WEEDCU_PROF_ELEMENTWISE = 4,
  WEEDCU_PROF_SOFTMAX = 5, WEEDCU_PROF_LAYERNORM = 6, WEEDCU_PROF_CROSS_ENTROPY = 7,
  WEEDCU_PROF_OPTIMIZER = 8, WEEDCU_PROF_REDUCE = 9, WEEDCU_PROF_EMBEDDING = 10, WEEDCU_PROF_FILL = 11,
  WEEDCU_PROF_NCCL = 12, WEEDCU_PROF_ATTENTION = 13, WEEDCU_PROF_NUM_CLASSES = 14
};
int weedcu_prof_enable(int on);
int weedcu_prof_read(int cls, double *total_ms, uint64_t *launches, double *work);

/* ------------------------------------------------------------------ data-parallel collectives
 * No reference counterpart (Weed has no gradient exchange). NCCL over NVLink,
 * one process per GPU. The unique id is produced by rank 0 and shipped by the caller
 * (bench.py uses torch.distributed for that plumbing). */
int weedcu_nccl_load(const char *libnccl_path);
int weedcu_nccl_unique_id(void *id128);      /* writes 128 bytes */
int weedcu_nccl_init(const void *id128, int rank, int world, void **comm);
int weedcu_nccl_destroy(void *comm);
/* ncclGroupStart / ncclGroupEnd: the per-parameter all-reduces of one step are issued as one group
 * (NCCL fuses them into few kernels instead of one launch per parameter). */
int weedcu_nccl_group_start(void);
int weedcu_nccl_group_end(void);
int weedcu_nccl_allreduce_sum(void *comm, float *buf, uint64_t n, void *stream);
/* dst[t][0 .. n[t]) = src[t][0 .. n[t]) for `count` dense fp32 runs in one or a few launches (host arrays of device
 * pointers). Used to gather the small gradients of a bucket into one all-reduce message and to scatter the sum back. */
int weedcu_multi_copy(uint32_t count, const float *const *src, float *const *dst, const uint64_t *n, void *stream);
int weedcu_nccl_broadcast(void *comm, float *buf, uint64_t n, int root, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* WEEDCU_H */
