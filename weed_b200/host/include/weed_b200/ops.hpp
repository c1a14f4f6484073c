// ops.hpp — Weed's tensor-op entry points (reference include/ops/*.hpp), same names and argument
// order. Each validates on the host with the reference's exception types, then, for GPU tensors,
// issues ONE typed call into the extern "C" layer (include/weedcu.h). The reference packs a
// 12-slot `tcapint` vector for RequestKernel (e.g. src/ops/commuting.cpp:37-50) — an OpenCL
// artefact this backend does not keep. There is no CPU compute path here: ops on CPU-tagged
// tensors throw std::domain_error (upstream Weed's own cpu_* functions serve that device).
#pragma once
#include "weed_b200/core.hpp"

namespace Weed {
void validate_all_same_device(const std::vector<const BaseTensor *> &t, const std::string cls);

void add(const Tensor &a, const Tensor &b, Tensor &out);
void mul(const Tensor &a, const Tensor &b, Tensor &out);
void sub(const Tensor &a, const Tensor &b, Tensor &out);
void div(const Tensor &a, const Tensor &b, Tensor &out);
void add_in_place(Tensor &a, const Tensor &b);
void sub_in_place(Tensor &a, const Tensor &b);
void copy_broadcast(Tensor &a, const Tensor &b);

void relu(const Tensor &a, Tensor &out);
void relu_grad(Tensor &din, const Tensor &in, const Tensor &dout);
void sigmoid(const Tensor &a, Tensor &out);
void sigmoid_grad(Tensor &din, const Tensor &in, const Tensor &dout);
void tanh(const Tensor &a, Tensor &out);
void tanh_grad(Tensor &din, const Tensor &in, const Tensor &dout);
void sin(const Tensor &a, Tensor &out);
void sin_grad(Tensor &din, const Tensor &in, const Tensor &dout);
void cos(const Tensor &a, Tensor &out);
void cos_grad(Tensor &din, const Tensor &in, const Tensor &dout);
void abs(const Tensor &a, Tensor &out);
void abs_grad(Tensor &din, const Tensor &in, const Tensor &dout);
void pow(const Tensor &a, const real1 &p, Tensor &out);
void exp(const Tensor &a, const real1 &b, Tensor &out);
void log(const Tensor &a, const real1 &b, Tensor &out);
// fused additions (no reference op of that name; behind Tensor::gelu, tensor.cpp:841-851)
void gelu(const Tensor &a, Tensor &out);
void gelu_grad(Tensor &din, const Tensor &in, const Tensor &dout);

void sum(const Tensor &a, Tensor &out);
void mean(const Tensor &a, Tensor &out);
void reduce(const tcapint &index, const Tensor &a, Tensor &out);
void reduce_grad(const tcapint &index, Tensor &din, const Tensor &a, const Tensor &dout);

// reference include/ops/clamp.hpp, real_extremum.hpp, reduce.hpp:27-36
void clamp(const Tensor &a, const real1 &l, const real1 &h, Tensor &out);
void clamp_grad(Tensor &din, const Tensor &in, const Tensor &dout, const real1 &l, const real1 &h);
void max(const Tensor &a, Tensor &out);
void max_grad(Tensor &din, const Tensor &in, const Tensor &dout, const Tensor &out);
void min(const Tensor &a, Tensor &out);
void min_grad(Tensor &din, const Tensor &in, const Tensor &dout, const Tensor &out);
void max(const tcapint &index, const Tensor &a, Tensor &out);
void min(const tcapint &index, const Tensor &a, Tensor &out);
void match_grad(const tcapint &index, Tensor &din, const Tensor &in, const Tensor &dout, const Tensor &out);

void softmax(const tcapint &index, const Tensor &a, Tensor &out);
void softmax_grad(const tcapint &index, Tensor &din, const Tensor &out, const Tensor &dout);
void logsoftmax(const tcapint &index, const Tensor &a, Tensor &out);
void logsoftmax_grad(const tcapint &index, Tensor &din, const Tensor &out, const Tensor &dout);

void matmul(const Tensor &a, const Tensor &b, Tensor &out);
// C (+)= A*B with the accumulate folded into the GEMM epilogue (make_matmul_node's tmp + add_in_place)
void matmul_accumulate(const Tensor &a, const Tensor &b, Tensor &out);
// out = a b + bias (bias: dense [N], broadcast over rows) in one tensor-core GEMM; false (nothing
// computed) when that kernel does not apply to these operands
bool matmul_bias(const Tensor &a, const Tensor &b, const Tensor &bias, Tensor &out, const Tensor *residual = nullptr);
// h = a w + bias (fp32, kept for the GELU backward) and the bf16 GEMM operand copy of y = gelu(h) from ONE tensor-core GEMM
// (weedcu_gemm_bf16_ex, activation 1); y's fp32 values are deferred. false: nothing was done.
bool matmul_bias_gelu(const Tensor &a, const Tensor &w, const Tensor &bias, Tensor &h, Tensor &y);
// LM head: out = a w + bias leaves the kernel as its bf16 copy + per-row log-sum-exp partials only (fp32 values deferred:
// recomputed by the plain GEMM if something reads them). false: nothing was done.
bool matmul_bias_lse(const Tensor &a, const Tensor &w, const Tensor &bias, Tensor &out);
// cross_entropy_loss forward for logits produced by matmul_bias_lse: no pass over the logits. false: nothing was done.
bool cross_entropy_fwd_from_stats(const Tensor &logits, const SymbolTensor &targets, Tensor &lse, Tensor &loss, tcapint rows, tcapint V);
// Packs dy [rows, N] into its bf16 GEMM shadow and adds its column sums into `sums` (dense [N]) in the
// same pass — Linear's bias gradient without a separate reduction. false: nothing was done.
bool pack_with_column_sums(const Tensor &dy, Tensor &sums);
// `batch` independent products over 3-D views [batch, M, K] x [batch, K, N] -> [batch, M, N]
// (the host loop of Tensor::matmul, tensor.cpp:1259-1269, as one strided-batched launch)
void matmul_batched(const Tensor &a3, const Tensor &b3, Tensor &out3);

void embedding_gather(const SymbolTensor &indices, const Tensor &weight, Tensor &out);
void embedding_scatter_add(Tensor &dW, const SymbolTensor &indices, const Tensor &dout);
// Greedy decoding (no reference counterpart: Weed has axis max but no arg-max, SURVEY §7 hard part 9):
// logits [B, T, V] -> the arg-max token of the LAST position of every sequence as a device
// SymbolTensor [B, 1] that can be fed straight back into Sequential::forward. Lowest index wins ties.
SymbolTensorPtr argmax_last_token(const Tensor &logits);
// outs[g] = a * ws[g] + biases[g] for up to three weights of one shape as ONE grouped tensor-core launch
// (the W_q / W_k / W_v projections); false (nothing computed) when the bf16 operand path does not apply.
// bf16_only: the caller reads the outputs only through their bf16 operand copies (the attention core): the fp32 values are
// not written unless something else reads them
bool matmul_bias_grouped(const Tensor &a, const std::vector<const Tensor *> &ws, const std::vector<const Tensor *> &biases,
                         const std::vector<Tensor *> &outs, bool bf16_only = false);
// MultiHeadAttention::forward's fused core on bf16 tensor cores (weedcu_attention_fwd_bf16out / _bf16in); returns the
// weedcu code (WEEDCU_ENOSUP: not eligible, nothing launched)
int attention_forward_bf16(const Tensor &q, const Tensor &k, const Tensor &v, Tensor &out, uint16_t *out_bf16, tcapint B, tcapint T, tcapint H, tcapint hd,
                           real1 divisor, real1 mask_val, int causal);
bool matmul_skinny_grouped(const Tensor &a, const std::vector<const Tensor *> &ws, const std::vector<const Tensor *> &biases,
                           const std::vector<Tensor *> &outs);
// Producer-side bf16 operand shadows. A kernel that is about to overwrite the dense tensor `out` may
// also emit bf16(out) at the same linear index; when the Linear that consumes `out` next views it as
// [rows, cols] (rows contiguous, rows % 8 == 0) that copy IS its tensor-core GEMM operand and the pack
// pass disappears. begin_ returns the buffer to fill (nullptr: bf16 path off or layout not eligible);
// end_ is called after the producing kernel has been issued through out's write accessor.
struct OutputShadow {
  GpuRealStorage *storage = nullptr;
  size_t index = 0U;
  uint16_t *ptr = nullptr;
};
OutputShadow begin_output_shadow(const Tensor &out, tcapint cols);
void end_output_shadow(const OutputShadow &os);
// LayerNorm::forward's fused kernel / Tensor::gelu's forward with the shadow emitted when eligible
void layernorm_forward(const Tensor &x, tcapint rows, tcapint features, const Tensor &gamma, const Tensor &beta, real1 eps, Tensor &y, Tensor &mean,
                       Tensor &rstd);
// Fused cross-entropy backward that also leaves dlogits' bf16 GEMM shadow and its column sums on the
// gradient's storage (include/weedcu.h: weedcu_cross_entropy_bwd_pack), so that the Linear node that
// consumes dlogits next neither re-reads it to pack nor to sum. False (nothing done) when the bf16
// operand path is off or the layout is not the dense [rows, V] one.
bool cross_entropy_bwd_pack(const Tensor &logits, const SymbolTensor &targets, const Tensor &lse, const Tensor &dloss, Tensor &dlogits,
                            tcapint rows, tcapint V);
void triu_fill(Tensor &a, const complex &val, const tcapint diagonal = 1);
} // namespace Weed
