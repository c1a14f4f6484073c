"""Seeded test cases shared by the CPU oracle tests and the GPU parity tests.

Every case is `fn(be, rng) -> {name: ndarray}`; `be` is a backend from backends.py, so the SAME
inputs and views go through oracle/liboracle.so (CPU restatement of the reference) and through
weed_b200/libweedcu.so (the sm_100a kernels behind include/weedcu.h).

Tolerances follow BASELINE.json north_star: relative-to-max error <= 1e-5 per fp32 op, exact (0.0)
for copies / fills / integer gathers, a stated looser bound for the bf16 tensor-core path.
"""
import ctypes as C

import numpy as np

from weed_b200._lib import Mat, contiguous_stride, make_view

U64, U32, I32 = C.c_uint64, C.c_uint32, C.c_int
F32 = np.float32
ADD, MUL, SUB, DIV = 0, 1, 2, 3
RELU, SIGMOID, TANH, ABS, POW, EXP, LOG, GELU, SIN, COS = range(10)

TOL = 1e-5


def rel_err(y, ref):
    y = np.asarray(y, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    denom = max(np.max(np.abs(ref)) if ref.size else 0.0, 1e-30)
    return float(np.max(np.abs(y - ref)) / denom) if ref.size else 0.0


def cview(shape, offset=0):
    return make_view(shape, contiguous_stride(shape), offset)


def span(shape, stride, offset=0):
    return offset + 1 + sum((s - 1) * t for s, t in zip(shape, stride))


def uni(rng, n, lo=-1.0, hi=1.0):
    return rng.uniform(lo, hi, size=n).astype(F32)


CASES = []


def case(name, tol=TOL):
    def deco(fn):
        CASES.append((name, fn, tol))
        return fn
    return deco


# --------------------------------------------------------------------------------------- binary
def _binary(be, rng, op, shape, sa, sb, so, oa=0, ob=0, oo=0, positive_b=False):
    a = uni(rng, span(shape, sa, oa))
    b = uni(rng, span(shape, sb, ob), 0.5 if positive_b else -1.0, 1.5 if positive_b else 1.0)
    out = np.full(span(shape, so, oo), 7.0, F32)
    ha, hb, ho = be.buf(a), be.buf(b), be.buf(out)
    be.call("binary_real", I32(op), ha, make_view(shape, sa, oa), hb, make_view(shape, sb, ob), ho,
            make_view(shape, so, oo))
    return {"out": ho.get()}


for _op, _nm in ((ADD, "add"), (MUL, "mul"), (SUB, "sub"), (DIV, "div")):
    @case(f"binary_{_nm}_contig_1003")
    def _c(be, rng, op=_op):
        s = [1003]
        return _binary(be, rng, op, s, [1], [1], [1], positive_b=True)

    @case(f"binary_{_nm}_bias_broadcast")
    def _c(be, rng, op=_op):  # y[M,N] + bias[N] after match_shape: stride [0,1]
        s = [64, 48]
        return _binary(be, rng, op, s, [1, 64], [0, 1], [1, 64], positive_b=True)

    @case(f"binary_{_nm}_scalar")
    def _c(be, rng, op=_op):
        s = [8, 16, 12]
        return _binary(be, rng, op, s, [1, 8, 128], [0, 0, 0], [1, 8, 128], positive_b=True)


@case("binary_add_mask_4d")
def _c(be, rng):  # scores[B,H,Tq,Tk] + mask[Tq,Tk] broadcast (multihead_attention.cpp:322-328)
    s = [2, 3, 8, 8]
    return _binary(be, rng, ADD, s, contiguous_stride(s), [0, 0, 1, 8], contiguous_stride(s))


@case("binary_add_transposed_view")
def _c(be, rng):  # contiguous(transpose(x,1,2)) = zeros + x  (tensor.hpp:319-331)
    s = [4, 6, 5, 8]          # logical [B,H,T,hd] view of a [B,T,H,hd] buffer
    sa = [1, 4 * 5, 4, 4 * 5 * 6]
    return _binary(be, rng, ADD, s, sa, contiguous_stride(s), contiguous_stride(s))


@case("binary_mul_offset_views_rank5")
def _c(be, rng):
    s = [3, 2, 4, 2, 5]
    st = contiguous_stride(s)
    return _binary(be, rng, MUL, s, st, st, st, oa=5, ob=3, oo=2)


@case("binary_add_large_vec")
def _c(be, rng):
    s = [4096, 33]
    st = contiguous_stride(s)
    return _binary(be, rng, ADD, s, st, st, st)


@case("inplace_add_broadcast")
def _c(be, rng):  # grad accumulation with a broadcast b (in_place.cpp:27-35)
    s = [32, 20]
    a, b = uni(rng, 640), uni(rng, 20)
    ha, hb = be.buf(a), be.buf(b)
    be.call("inplace_real", I32(ADD), ha, cview(s), hb, make_view(s, [0, 1]))
    return {"a": ha.get()}


@case("inplace_sub_contig")
def _c(be, rng):
    s = [1001]
    a, b = uni(rng, 1001), uni(rng, 1001)
    ha, hb = be.buf(a), be.buf(b)
    be.call("inplace_real", I32(SUB), ha, cview(s), hb, cview(s))
    return {"a": ha.get()}


@case("inplace_add_kv_slot")
def _c(be, rng):  # k_cache slice += K (multihead_attention.cpp:278-283): strided destination window
    full = [2, 3, 16, 4]
    st = contiguous_stride(full)
    s = [2, 3, 5, 4]
    a, b = np.zeros(int(np.prod(full)), F32), uni(rng, int(np.prod(s)))
    ha, hb = be.buf(a), be.buf(b)
    be.call("inplace_real", I32(ADD), ha, make_view(s, st, 7 * st[2]), hb, cview(s))
    return {"a": ha.get()}


@case("copy_broadcast_materialize", tol=0.0)
def _c(be, rng):  # Tensor::materialize_broadcast (tensor.cpp:334-353)
    s = [16, 12, 3]
    src = uni(rng, 12)
    dst = np.zeros(int(np.prod(s)), F32)
    hs, hd = be.buf(src), be.buf(dst)
    be.call("copy_real", hd, cview(s), hs, make_view(s, [0, 1, 0]))
    return {"dst": hd.get()}


@case("copy_transpose", tol=0.0)
def _c(be, rng):
    s = [40, 24]
    src = uni(rng, 960)
    dst = np.zeros(960, F32)
    hs, hd = be.buf(src), be.buf(dst)
    be.call("copy_real", hd, cview(s), hs, make_view(s, [24, 1]))
    return {"dst": hd.get()}


@case("fill_real", tol=0.0)
def _c(be, rng):
    buf = np.zeros(1031, F32)
    h = be.buf(buf)
    be.call("fill_real", h, U64(1031), F32(1.0))
    return {"buf": h.get()}


# --------------------------------------------------------------------------------------- unary
_UNARY = [(RELU, "relu", 0.0, -1, 1), (SIGMOID, "sigmoid", 0.0, -6, 6), (TANH, "tanh", 0.0, -4, 4),
          (ABS, "abs", 0.0, -1, 1), (POW, "pow0.5", 0.5, 0.01, 4), (POW, "pow2", 2.0, -2, 2),
          (EXP, "exp_e", 1.0, -5, 5), (LOG, "log_e", 1.0, 0.01, 9), (GELU, "gelu", 0.0, -5, 5),
          (SIN, "sin", 0.0, -3, 3), (COS, "cos", 0.0, -3, 3)]
for _op, _nm, _p, _lo, _hi in _UNARY:
    @case(f"unary_{_nm}")
    def _c(be, rng, op=_op, p=_p, lo=_lo, hi=_hi):
        s = [257, 9]
        a = uni(rng, 257 * 9, lo, hi)
        ha, ho = be.buf(a), be.buf(np.zeros_like(a))
        be.call("unary_real", I32(op), F32(p), ha, cview(s), ho, cview(s))
        return {"out": ho.get()}

for _op, _nm, _lo, _hi in [(RELU, "relu", -1, 1), (SIGMOID, "sigmoid", 0.01, 0.99), (TANH, "tanh", -0.99, 0.99),
                           (ABS, "abs", -1, 1), (GELU, "gelu", -4, 4), (SIN, "sin", -3, 3), (COS, "cos", -3, 3)]:
    @case(f"unary_grad_store_{_nm}")
    def _c(be, rng, op=_op, lo=_lo, hi=_hi):  # accumulate = 0: din is overwritten, never read
        s = [257, 5]
        n = 257 * 5
        din, x, dout = uni(rng, n), uni(rng, n, lo, hi), uni(rng, n)
        hd, hx, hg = be.buf(din), be.buf(x), be.buf(dout)
        be.call("unary_grad_real", I32(op), hd, cview(s), hx, cview(s), hg, cview(s), I32(0))
        return {"din": hd.get()}

    @case(f"unary_grad_{_nm}")
    def _c(be, rng, op=_op, lo=_lo, hi=_hi):
        s = [130, 7]
        n = 130 * 7
        din, x, dout = uni(rng, n), uni(rng, n, lo, hi), uni(rng, n)
        if op in (RELU, ABS):
            x[::11] = 0.0
        hd, hx, hg = be.buf(din), be.buf(x), be.buf(dout)
        be.call("unary_grad_real", I32(op), hd, cview(s), hx, cview(s), hg, cview(s), I32(1))
        return {"din": hd.get()}


@case("unary_grad_broadcast_dout")
def _c(be, rng):  # dout is the all-ones seed broadcast from one element (tensor.cpp:377)
    s = [64, 5]
    din, x, dout = uni(rng, 320), uni(rng, 320), np.ones(1, F32)
    hd, hx, hg = be.buf(din), be.buf(x), be.buf(dout)
    be.call("unary_grad_real", I32(TANH), hd, cview(s), hx, cview(s), hg, make_view(s, [0, 0]), I32(1))
    return {"din": hd.get()}


# --------------------------------------------------------------------------------------- reduce
def _reduce(be, rng, shape, axis, order):
    a = uni(rng, int(np.prod(shape)))
    n_out = int(np.prod(shape)) // shape[axis]
    ha, ho = be.buf(a), be.buf(np.zeros(n_out, F32))
    be.call("reduce_real", ha, cview(shape), I32(axis), ho, I32(order))
    return {"out": ho.get()}


for _shape, _axis in [([300, 70], 0), ([300, 70], 1), ([33, 1], 0), ([6, 10, 96], 2), ([6, 10, 96], 1),
                      ([6, 10, 96], 0), ([1, 40, 64], 2), ([64, 3, 5, 7], 2), ([2048, 40], 1),
                      # contiguous axis: short (thread per output, 128-bit / scalar) and medium (warp per output)
                      ([8, 3000], 0), ([7, 513], 0), ([200, 640], 0)]:
    for _order in (0, 1):
        @case(f"reduce_{'x'.join(map(str, _shape))}_axis{_axis}_order{_order}")
        def _c(be, rng, shape=_shape, axis=_axis, order=_order):
            return _reduce(be, rng, shape, axis, order)

for _shape, _axis in [([30, 17], 1), ([30, 17], 0), ([4, 6, 16], 2), ([1, 12, 16], 2), ([4, 6, 16], 1)]:
    for _order in (0, 1):
        @case(f"reduce_grad_{'x'.join(map(str, _shape))}_axis{_axis}_order{_order}")
        def _c(be, rng, shape=_shape, axis=_axis, order=_order):
            n = int(np.prod(shape))
            oshape = list(shape)
            oshape[axis] = 1
            ost = contiguous_stride(oshape)      # dout built by Tensor::sum, then match_shape'd
            din, dout = uni(rng, n), uni(rng, n // shape[axis])
            hd, hg = be.buf(din), be.buf(dout)
            be.call("reduce_grad_real", hd, cview(shape), hg, make_view(shape, ost), I32(axis), I32(order))
            return {"din": hd.get()}


# ---- round 2: clamp / max / min / match_grad (reference clamp.cpp, real_extremum.cpp, reduce.cpp:40-113)
def _quantised(rng, n, levels=23):
    """values on a coarse grid: the extremum is hit by several elements, which is what match_grad is about"""
    return (rng.integers(-levels, levels + 1, size=n).astype(F32) / F32(4.0)).astype(F32)


for _shape in ([1000], [37, 29], [6, 10, 16]):
    @case(f"clamp_{'x'.join(map(str, _shape))}", tol=0.0)
    def _c(be, rng, shape=_shape):
        n = int(np.prod(shape))
        a = uni(rng, n, -2, 2)
        ha, ho = be.buf(a), be.buf(np.zeros(n, F32))
        be.call("clamp_real", ha, cview(shape), F32(-0.75), F32(0.5), ho, cview(shape))
        din, dout = uni(rng, n), uni(rng, n)
        hd, hg = be.buf(din), be.buf(dout)
        be.call("clamp_grad_real", hd, cview(shape), ha, cview(shape), hg, cview(shape), F32(-0.75), F32(0.5))
        return {"out": ho.get(), "din": hd.get()}


@case("clamp_strided_views", tol=0.0)
def _c(be, rng):
    s = [20, 12]
    a, out = uni(rng, 20 * 40 + 7, -2, 2), np.zeros(12 * 25, F32)
    ha, ho = be.buf(a), be.buf(out)
    be.call("clamp_real", ha, make_view(s, [1, 40], 7), F32(-1.0), F32(1.0), ho, make_view(s, [12, 1], 0))  # transposed output
    return {"out": ho.get()}


for _ismin in (0, 1):
    for _shape in ([100003], [300, 70], [5, 7, 11]):
        @case(f"{'min' if _ismin else 'max'}_full_{'x'.join(map(str, _shape))}", tol=0.0)
        def _c(be, rng, shape=_shape, ismin=_ismin):
            n = int(np.prod(shape))
            a = _quantised(rng, n)
            ha, ho = be.buf(a), be.buf(np.zeros(1, F32))
            be.call("extremum_real", I32(ismin), ha, cview(shape), ho)
            din, g = uni(rng, n), uni(rng, 1)
            hd, hg = be.buf(din), be.buf(g)
            be.call("match_grad_full_real", hd, cview(shape), ha, cview(shape), hg, make_view(shape, [0] * len(shape)), ho)
            return {"out": ho.get(), "din": hd.get()}

    @case(f"{'min' if _ismin else 'max'}_full_strided_view", tol=0.0)
    def _c(be, rng, ismin=_ismin):
        a = _quantised(rng, 50 * 64 + 3)
        ha, ho = be.buf(a), be.buf(np.zeros(1, F32))
        be.call("extremum_real", I32(ismin), ha, make_view([50, 30], [1, 64], 3), ho)
        return {"out": ho.get()}

    for _shape, _axis in [([300, 70], 0), ([300, 70], 1), ([6, 10, 96], 2), ([6, 10, 96], 1), ([6, 10, 96], 0), ([1, 40, 64], 2)]:
        for _order in (0, 1):
            @case(f"{'min' if _ismin else 'max'}_axis_{'x'.join(map(str, _shape))}_axis{_axis}_order{_order}", tol=0.0)
            def _c(be, rng, shape=_shape, axis=_axis, order=_order, ismin=_ismin):
                n = int(np.prod(shape))
                a = _quantised(rng, n)
                n_out = n // shape[axis]
                ha, ho = be.buf(a), be.buf(np.zeros(n_out, F32))
                be.call("extremum_axis_real", I32(ismin), ha, cview(shape), I32(axis), ho, I32(order))
                oshape = list(shape)
                oshape[axis] = 1
                ost = contiguous_stride(oshape)      # out / dout as Tensor::max builds them, then match_shape'd
                din, dout = uni(rng, n), uni(rng, n_out)
                hd, hg = be.buf(din), be.buf(dout)
                be.call("match_grad_real", hd, cview(shape), ha, cview(shape), hg, make_view(shape, ost), ho, I32(axis), I32(order))
                return {"out": ho.get(), "din": hd.get()}


@case("sum_linear_100003", tol=2e-5)
def _c(be, rng):
    # positive inputs: the tolerance is relative to |sum|, so keep the sum well conditioned (the
    # serial fp32 oracle loop and the device tree differ by rounding order only)
    a = uni(rng, 100003, 0.0, 1.0)
    ha, ho = be.buf(a), be.buf(np.zeros(1, F32))
    be.call("sum_real", ha, cview([100003]), F32(1.0), ho)
    return {"out": ho.get()}


@case("mean_strided_view", tol=2e-5)
def _c(be, rng):
    s = [50, 30]
    a = uni(rng, 50 * 64)
    ha, ho = be.buf(a), be.buf(np.zeros(1, F32))
    be.call("sum_real", ha, make_view(s, [1, 64], 3), F32(1.0 / 1500), ho)
    return {"out": ho.get()}


# --------------------------------------------------------------------------------------- softmax
def _softmax(be, rng, log_mode, shape, axis, stride=None, lo=-10, hi=10):
    st = stride or contiguous_stride(shape)
    a = uni(rng, span(shape, st), lo, hi)
    ha, ho = be.buf(a), be.buf(np.zeros(int(np.prod(shape)), F32))
    be.call("softmax_real", I32(log_mode), ha, make_view(shape, st), I32(axis), ho, cview(shape))
    return {"out": ho.get()}


def _softmax_bwd(be, rng, log_mode, shape, axis):
    n = int(np.prod(shape))
    x = uni(rng, n, -3, 3)
    hx, hy = be.buf(x), be.buf(np.zeros(n, F32))
    be.call("softmax_real", I32(log_mode), hx, cview(shape), I32(axis), hy, cview(shape))
    din, dout = uni(rng, n), uni(rng, n)
    hd, hg = be.buf(din), be.buf(dout)
    be.call("softmax_grad_real", I32(log_mode), hd, cview(shape), hy, cview(shape), hg, cview(shape), I32(axis))
    return {"din": hd.get()}


for _lm, _nm in ((0, "softmax"), (1, "logsoftmax")):
    for _shape, _axis in [([3], 0), ([37, 11], 0), ([37, 11], 1), ([2, 3, 40, 33], 3), ([70, 600], 1),
                          ([5, 9, 4], 1), ([64, 2000], 1)]:
        @case(f"{_nm}_{'x'.join(map(str, _shape))}_axis{_axis}")
        def _c(be, rng, lm=_lm, shape=_shape, axis=_axis):
            return _softmax(be, rng, lm, shape, axis)

        @case(f"{_nm}_bwd_{'x'.join(map(str, _shape))}_axis{_axis}", tol=2e-5)
        def _c(be, rng, lm=_lm, shape=_shape, axis=_axis):
            return _softmax_bwd(be, rng, lm, shape, axis)

    @case(f"{_nm}_adversarial_1000")
    def _c(be, rng, lm=_lm):  # test_softmax_forward_numerical_stability (reference tests.cpp:1134-1145)
        a = np.array([1000, 1001, 1002], F32)
        ha, ho = be.buf(a), be.buf(np.zeros(3, F32))
        be.call("softmax_real", I32(lm), ha, cview([3]), I32(0), ho, cview([3]))
        return {"out": ho.get()}

    @case(f"{_nm}_long_rows_unstaged")
    def _c(be, rng, lm=_lm):  # 40 x 3000 x 4 B > staging budget: online two-pass path
        return _softmax(be, rng, lm, [40, 3000], 1)

    # >= 2^22 elements: the two streaming passes (row statistics split over column ranges, 128-bit apply);
    # an odd row count takes the scalar apply, outer > 1 the per-slab indexing
    for _shape, _axis in [([2048, 2100], 1), ([2050, 2050], 1), ([1024, 70, 64], 1), ([64, 40, 1700], 2)]:
        @case(f"{_nm}_two_pass_{'x'.join(map(str, _shape))}_axis{_axis}")
        def _c(be, rng, lm=_lm, shape=_shape, axis=_axis):
            return _softmax(be, rng, lm, shape, axis)

    @case(f"{_nm}_noncontiguous_input")
    def _c(be, rng, lm=_lm):
        return _softmax(be, rng, lm, [12, 20], 1, stride=[24, 1])


@case("attn_softmax_causal")
def _c(be, rng):
    batch, tq, tk = 6, 24, 24
    s = uni(rng, batch * tq * tk, -4, 4)
    hs, ho = be.buf(s), be.buf(np.zeros_like(s))
    be.call("attn_softmax_real", hs, ho, U32(batch), U32(tq), U32(tk), F32(np.sqrt(F32(16.0))),
            F32(-1.701411835e38), I32(1), I32(1))
    return {"out": ho.get()}


@case("attn_softmax_causal_per_batch_inplace")
def _c(be, rng):  # layout of the fused attention path; probabilities overwrite the scores
    batch, tq, tk = 5, 40, 40
    s = uni(rng, batch * tq * tk, -4, 4)
    hs = be.buf(s)
    be.call("attn_softmax_real", hs, hs, U32(batch), U32(tq), U32(tk), F32(8.0), F32(-1.701411835e38), I32(1), I32(0))
    return {"out": hs.get()}


@case("attn_softmax_decode_row")
def _c(be, rng):  # Tq = 1: no mask is applied (multihead_attention.cpp:322)
    batch, tq, tk = 12, 1, 40
    s = uni(rng, batch * tq * tk, -4, 4)
    hs, ho = be.buf(s), be.buf(np.zeros_like(s))
    be.call("attn_softmax_real", hs, ho, U32(batch), U32(tq), U32(tk), F32(8.0), F32(-1.701411835e38), I32(1), I32(1))
    return {"out": ho.get()}


def _ce(be, rng, rows, V, acc=1):
    logits = uni(rng, rows * V, -5, 5)
    tg = rng.integers(0, V, size=rows).astype(np.int32)
    hl, ht = be.buf(logits), be.buf(tg)
    hlse, hloss = be.buf(np.zeros(rows, F32)), be.buf(np.zeros(1, F32))
    be.call("cross_entropy_fwd", hl, U64(0), U32(rows), U32(V), U32(1), U32(rows), ht, hlse, hloss)
    dl = uni(rng, rows * V)
    hdl, hg = be.buf(dl), be.buf(np.ones(1, F32))
    be.call("cross_entropy_bwd", hl, U64(0), U32(rows), U32(V), U32(1), U32(rows), ht, hlse, hg, hdl, U64(0), I32(acc))
    return {"lse": hlse.get(), "loss": hloss.get(), "dlogits": hdl.get()}


def _ce_pack(be, rng, rows, V, acc):
    logits = uni(rng, rows * V, -5, 5)
    tg = rng.integers(0, V, size=rows).astype(np.int32)
    hl, ht = be.buf(logits), be.buf(tg)
    hlse, hloss = be.buf(np.zeros(rows, F32)), be.buf(np.zeros(1, F32))
    be.call("cross_entropy_fwd", hl, U64(0), U32(rows), U32(V), U32(1), U32(rows), ht, hlse, hloss)
    hdl, hg = be.buf(uni(rng, rows * V)), be.buf(np.full(1, 0.7, F32))
    hsh, hcs = be.buf(np.zeros(rows * V, np.uint16)), be.buf(np.full(V, 3.0, F32))
    be.call("cross_entropy_bwd_pack", hl, U64(0), U32(rows), U32(V), ht, hlse, hg, hdl, U64(0), I32(acc), hsh, hcs)
    d = hdl.get()
    # the bf16 copy must be the RNE rounding of THIS backend's dlogits, bit for bit (comparing the
    # copies of two backends would let a one-ulp fp32 difference flip a rounding)
    u = d.view(np.uint32).astype(np.uint64)
    want = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)
    ok = np.array([1.0 if np.array_equal(hsh.get(), want) else 0.0], F32)
    return {"dlogits": d, "shadow_is_rne_of_dlogits": ok, "colsum": hcs.get()}


for _rows, _V, _acc in [(48, 1000, 1), (1032, 77, 0), (2048, 40, 1)]:
    @case(f"cross_entropy_bwd_pack_{_rows}x{_V}_acc{_acc}", tol=2e-5)
    def _c(be, rng, rows=_rows, V=_V, acc=_acc):
        out = _ce_pack(be, rng, rows, V, acc)
        return out


@case("cross_entropy_48x1000", tol=2e-5)
def _c(be, rng):
    return _ce(be, rng, 48, 1000)


@case("cross_entropy_7x13", tol=2e-5)
def _c(be, rng):
    return _ce(be, rng, 7, 13)


@case("cross_entropy_300x2000_store", tol=2e-5)
def _c(be, rng):  # accumulate = 0: dlogits is overwritten (vocab split over several blocks)
    return _ce(be, rng, 300, 2000, acc=0)


for _rows, _V in [(1028, 300), (2048, 1000), (516, 50257)]:
    @case(f"cross_entropy_{_rows}x{_V}_vec4_rows", tol=2e-5)
    def _c(be, rng, rows=_rows, V=_V):  # forward with 4 adjacent rows per thread (128-row tiles, ragged last tile / vocabulary slice)
        return _ce(be, rng, rows, V)


@case("cross_entropy_33x50257", tol=2e-5)
def _c(be, rng):  # GPT-2 vocabulary, ragged row tile
    return _ce(be, rng, 33, 50257)


# --------------------------------------------------------------------------------------- layernorm
def _ln(be, rng, rows, F, grad_mode=0, acc=1):
    x = uni(rng, rows * F, -2, 2)
    gamma, beta = uni(rng, F, 0.5, 1.5), uni(rng, F)
    hx, hg, hb = be.buf(x), be.buf(gamma), be.buf(beta)
    hy, hm, hr = be.buf(np.zeros_like(x)), be.buf(np.zeros(rows, F32)), be.buf(np.zeros(rows, F32))
    eps = F32(np.finfo(np.float32).eps / 4)  # FP_NORM_EPSILON (weed_types.hpp:213-214)
    be.call("layernorm_fwd", hx, U32(rows), U32(F), hg, hb, eps, hy, hm, hr)
    dy, dx = uni(rng, rows * F), uni(rng, rows * F)
    dg, db = uni(rng, F), uni(rng, F)
    hdy, hdx, hdg, hdb = be.buf(dy), be.buf(dx), be.buf(dg), be.buf(db)
    be.call("layernorm_bwd", hx, hdy, U32(rows), U32(F), hg, hm, hr, hdx, hdg, hdb, I32(grad_mode), I32(acc))
    return {"y": hy.get(), "mean": hm.get(), "rstd": hr.get(), "dx": hdx.get(), "dgamma": hdg.get(),
            "dbeta": hdb.get()}


for _rows, _F in [(40, 24), (5, 8), (300, 64), (5000, 16), (9600, 12)]:
    for _gm in (0, 1):
        @case(f"layernorm_{_rows}x{_F}_gradmode{_gm}", tol=3e-5)
        def _c(be, rng, rows=_rows, F=_F, gm=_gm):
            return _ln(be, rng, rows, F, gm)


@case("layernorm_8192x768_store", tol=3e-5)
def _c(be, rng):  # GPT-2 shape; accumulate = 0 (dx overwritten); persistent blocks loop over row tiles
    return _ln(be, rng, 8192, 768, 0, acc=0)


for _rows, _F, _why in [(8, 768, "decode_step_single_launch"), (130, 1000, "single_launch_ragged"), (1026, 72, "odd_rows_scalar_apply"),
                        (516, 1100, "features_beyond_register_tile"), (4100, 24, "vector_apply_ragged_chunk")]:
    @case(f"layernorm_{_rows}x{_F}_{_why}", tol=3e-5)
    def _c(be, rng, rows=_rows, F=_F):
        return _ln(be, rng, rows, F, 1)


def _rne_ok(values, shadow_u16):
    u = values.view(np.uint32).astype(np.uint64)
    want = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)
    return np.array([1.0 if np.array_equal(shadow_u16, want) else 0.0], F32)


@case("layernorm_fwd_bf16_1032x72", tol=3e-5)
def _c(be, rng):  # forward that also emits the bf16 GEMM operand copy of y
    rows, F = 1032, 72
    x = uni(rng, rows * F, -2, 2)
    gamma, beta = uni(rng, F, 0.5, 1.5), uni(rng, F)
    hx, hg, hb = be.buf(x), be.buf(gamma), be.buf(beta)
    hy, hm, hr = be.buf(np.zeros_like(x)), be.buf(np.zeros(rows, F32)), be.buf(np.zeros(rows, F32))
    hs = be.buf(np.zeros(rows * F, np.uint16))
    be.call("layernorm_fwd_bf16", hx, U32(rows), U32(F), hg, hb, F32(np.finfo(np.float32).eps / 4), hy, hm, hr, hs)
    y = hy.get()
    return {"y": y, "mean": hm.get(), "rstd": hr.get(), "shadow_is_rne_of_y": _rne_ok(y, hs.get())}


@case("gelu_fwd_bf16_4100")
def _c(be, rng):
    n = 4100
    x = uni(rng, n, -4, 4)
    hx, hy, hs = be.buf(x), be.buf(np.zeros(n, F32)), be.buf(np.zeros(n, np.uint16))
    be.call("gelu_fwd_bf16", hx, hy, hs, U64(n))
    y = hy.get()
    return {"y": y, "shadow_is_rne_of_y": _rne_ok(y, hs.get())}


for _rows, _cols, _acc in [(1032, 50, 0), (2048, 24, 1)]:
    @case(f"gelu_grad_pack_{_rows}x{_cols}_acc{_acc}", tol=2e-5)
    def _c(be, rng, rows=_rows, cols=_cols, acc=_acc):
        n = rows * cols
        hd, hx, hg = be.buf(uni(rng, n)), be.buf(uni(rng, n, -4, 4)), be.buf(uni(rng, n))
        hs, hc = be.buf(np.zeros(n, np.uint16)), be.buf(np.full(cols, 3.0, F32))
        be.call("gelu_grad_pack", hd, hx, hg, U32(rows), U32(cols), I32(acc), hs, hc)
        d = hd.get()
        return {"din": d, "shadow_is_rne_of_din": _rne_ok(d, hs.get()), "colsum": hc.get()}


for _rows, _F in [(8200, 768), (4096, 520), (1028, 1024), (2052, 96)]:
    @case(f"layernorm_{_rows}x{_F}_large_ragged", tol=3e-5)
    def _c(be, rng, rows=_rows, F=_F):  # large inputs with a ragged last row tile / feature group
        return _ln(be, rng, rows, F, 1)


@case("layernorm_fwd_bf16_8200x768", tol=3e-5)
def _c(be, rng):
    rows, F = 8200, 768
    x = uni(rng, rows * F, -2, 2)
    gamma, beta = uni(rng, F, 0.5, 1.5), uni(rng, F)
    hx, hg, hb = be.buf(x), be.buf(gamma), be.buf(beta)
    hy, hm, hr = be.buf(np.zeros_like(x)), be.buf(np.zeros(rows, F32)), be.buf(np.zeros(rows, F32))
    hs = be.buf(np.zeros(rows * F, np.uint16))
    be.call("layernorm_fwd_bf16", hx, U32(rows), U32(F), hg, hb, F32(np.finfo(np.float32).eps / 4), hy, hm, hr, hs)
    y = hy.get()
    return {"y": y, "mean": hm.get(), "rstd": hr.get(), "shadow_is_rne_of_y": _rne_ok(y, hs.get())}


@case("layernorm_20000x40_multi_tile", tol=3e-5)
def _c(be, rng):  # more row tiles than blocks: the per-block column sums span several tiles
    return _ln(be, rng, 20000, 40, 1)


# --------------------------------------------------------------------------------------- embedding etc.
@case("embedding_gather", tol=0.0)
def _c(be, rng):
    V, D, n = 50, 12, 37
    W = uni(rng, V * D)
    idx = rng.integers(0, V, size=n).astype(np.int32)
    hw, hi, ho = be.buf(W), be.buf(idx), be.buf(np.zeros(n * D, F32))
    be.call("embedding_gather", hi, U64(0), U32(1), U32(n), hw, U64(0), U32(1), U32(V), U32(D), ho, U64(0),
            U32(1), U32(n))
    return {"out": ho.get()}


@case("embedding_scatter_add_duplicates")
def _c(be, rng):
    V, D, n = 9, 6, 64  # many duplicate tokens
    dW = uni(rng, V * D)
    idx = rng.integers(0, V, size=n).astype(np.int32)
    dout = uni(rng, n * D)
    hw, hi, hd = be.buf(dW), be.buf(idx), be.buf(dout)
    be.call("embedding_scatter_add", hw, U64(0), U32(1), U32(V), hi, U64(0), U32(1), U32(n), U32(D), hd, U64(0),
            U32(1), U32(n))
    return {"dW": hw.get()}


@case("triu_fill", tol=0.0)
def _c(be, rng):
    a = np.zeros(20 * 20, F32)
    ha = be.buf(a)
    be.call("triu_fill_real", ha, cview([20, 20]), F32(-1.701411835e38), U32(1))
    return {"a": ha.get()}


@case("argmax_rows", tol=0.0)
def _c(be, rng):
    rows, V = 45, 777
    x = uni(rng, rows * V)
    x[3 + 5 * rows] = x[3 + 9 * rows] = 5.0  # tie: lowest index wins
    hx, ho = be.buf(x), be.buf(np.zeros(rows, np.int32))
    be.call("argmax_rows", hx, U64(0), U32(rows), U32(V), U32(1), U32(rows), ho)
    return {"idx": ho.get()}


@case("argmax_rows_decode_shape_ties_across_slices", tol=0.0)
def _c(be, rng):  # 8 rows x GPT-2 vocabulary: the scan is split over vocabulary slices; ties in different slices
    rows, V = 8, 50257
    x = uni(rng, rows * V)
    x[2 + 40000 * rows] = x[2 + 123 * rows] = x[2 + 50256 * rows] = 7.0
    x[5 + 50256 * rows] = 9.0
    hx, ho = be.buf(x), be.buf(np.zeros(rows, np.int32))
    be.call("argmax_rows", hx, U64(0), U32(rows), U32(V), U32(1), U32(rows), ho)
    return {"idx": ho.get()}


# --------------------------------------------------------------------------------------- optimisers
@case("sgd_step")
def _c(be, rng):
    n = 4099
    p, g = uni(rng, n), uni(rng, n)
    hp, hg = be.buf(p), be.buf(g)
    be.call("sgd_step", hp, hg, U64(n), F32(0.01), F32(1.0))
    return {"p": hp.get()}


@case("adam_step_3_iterations")
def _c(be, rng):
    n = 2051
    p, m, v = uni(rng, n), np.zeros(n, F32), np.zeros(n, F32)
    hp, hm, hv = be.buf(p), be.buf(m), be.buf(v)
    b1, b2 = F32(0.9), F32(0.999)
    for t in range(1, 4):
        g = uni(rng, n)
        hg = be.buf(g)
        bc1 = F32(1.0) - F32(np.power(b1, F32(t), dtype=F32))
        bc2 = F32(1.0) - F32(np.power(b2, F32(t), dtype=F32))
        be.call("adam_step", hp, hg, hm, hv, U64(n), F32(0.001), b1, b2, F32(1e-8), bc1, bc2, F32(1.0))
    return {"p": hp.get(), "m": hm.get(), "v": hv.get()}


@case("adam_step_gscale")
def _c(be, rng):
    n = 512
    p, g, m, v = uni(rng, n), uni(rng, n), uni(rng, n, 0, 0.1), uni(rng, n, 0, 0.1)
    hp, hg, hm, hv = be.buf(p), be.buf(g), be.buf(m), be.buf(v)
    be.call("adam_step", hp, hg, hm, hv, U64(n), F32(0.001), F32(0.9), F32(0.999), F32(1e-8), F32(0.1),
            F32(0.001), F32(0.125))
    return {"p": hp.get(), "m": hm.get(), "v": hv.get()}


# --------------------------------------------------------------------------------------- matmul
def _mat(offset, s0, s1, bs=0):
    m = Mat()
    m.offset, m.s0, m.s1, m.batch_stride = offset, s0, s1, bs
    return m


def _matmul(be, rng, M, K, N, a_layout, b_layout, batch=1, accumulate=0, precision=0, fn="matmul_real"):
    """layouts: 'col' = (1, rows), 'row' = (cols, 1), 'batchfast' = (batch, batch*rows) (tensor.cpp:1242-1262)"""
    def lay(rows, cols, kind):
        if kind == "col":
            return 1, rows, rows * cols
        if kind == "row":
            return cols, 1, rows * cols
        return batch, batch * rows, 1  # batch index fastest
    as0, as1, abs_ = lay(M, K, a_layout)
    bs0, bs1, bbs = lay(K, N, b_layout)
    a = uni(rng, batch * M * K + 3)
    b = uni(rng, batch * K * N + 5)
    c = uni(rng, batch * M * N)
    ha, hb, hc = be.buf(a), be.buf(b), be.buf(c)
    args = [ha, _mat(3 if a_layout != "batchfast" else 0, as0, as1, abs_),
            hb, _mat(5 if b_layout != "batchfast" else 0, bs0, bs1, bbs),
            hc, _mat(0, 1, M, M * N), U32(M), U32(K), U32(N), U32(batch), I32(accumulate)]
    if fn == "matmul_real":
        args.append(I32(precision))
    be.call(fn, *args)
    return {"c": hc.get()}


for _M, _K, _N, _al, _bl, _batch, _acc in [
        (2, 3, 2, "col", "col", 1, 0),          # test_matmul_gradient_sum_loss shapes
        (130, 70, 150, "col", "col", 1, 0),      # forward: A MN-major, B K-major
        (130, 70, 150, "col", "row", 1, 1),      # dA: B^T view, accumulate
        (70, 130, 150, "row", "col", 1, 1),      # dB: A^T view, accumulate
        (257, 33, 129, "row", "row", 1, 0),
        (64, 13, 26, "col", "col", 1, 0),        # heart_attack layer 1 tile
        (500, 26, 1, "col", "col", 1, 0),        # heart_attack head: thin kernel
        (1, 96, 300, "col", "col", 1, 0),        # decode GEMV
        (40, 16, 40, "batchfast", "batchfast", 6, 0),  # attention QK^T slices
        (128, 128, 128, "col", "col", 3, 0),
]:
    @case(f"matmul_f32_{_M}x{_K}x{_N}_{_al}_{_bl}_b{_batch}_acc{_acc}", tol=2e-5)
    def _c(be, rng, M=_M, K=_K, N=_N, al=_al, bl=_bl, batch=_batch, acc=_acc):
        return _matmul(be, rng, M, K, N, al, bl, batch, acc, 0)


# Few outputs, long reduction: the weight-gradient products of the small-feature configs (C2's
# 13x65536x26, C4's 512x32768x1), which the product runs split-K (gemm_f32.cu). The oracle sums K
# terms in order in fp32, the kernel in per-slice partials, so the bound scales with sqrt(K)*eps.
for _M, _K, _N, _al, _bl, _acc in [
        (13, 65536, 26, "row", "col", 0),       # dW = X^T dY: both operands K-contiguous
        (13, 65536, 26, "row", "col", 1),
        (26, 32768, 1, "row", "col", 1),
        (512, 4096, 1, "col", "col", 0),        # A K-strided
        (70, 2048, 150, "col", "row", 0),       # B K-strided, ragged tiles
        (64, 2051, 64, "row", "row", 0),        # ragged K
]:
    @case(f"matmul_f32_splitk_{_M}x{_K}x{_N}_{_al}_{_bl}_acc{_acc}", tol=1e-4)
    def _c(be, rng, M=_M, K=_K, N=_N, al=_al, bl=_bl, acc=_acc):
        return _matmul(be, rng, M, K, N, al, bl, 1, acc, 0)


# --------------------------------------------------------------------------------------- decode
def _decode(be, rng, B, H, hd, S, steps, causal=1):
    """KV-cache attention over a sequence of calls (cache append + attention): `steps` = T_new of
    every call; returns every call's output and the final caches."""
    BH = B * H
    kc, vc = np.zeros(BH * S * hd, F32), np.zeros(BH * S * hd, F32)
    hkc, hvc = be.buf(kc), be.buf(vc)
    div, mask = np.float32(np.sqrt(hd)), np.float32(-1.701411835e38)
    out, cache_len = {}, 0
    for i, T in enumerate(steps):
        n = B * T * H * hd
        q, k, v = uni(rng, n), uni(rng, n), uni(rng, n)
        hq, hk, hv, ho = be.buf(q), be.buf(k), be.buf(v), be.buf(np.zeros(n, F32))
        be.call("attention_decode", hq, hk, hv, hkc, hvc, ho, U32(B), U32(T), U32(H), U32(hd), U32(S), U32(cache_len), div, mask, I32(causal))
        out[f"out{i}"] = ho.get()
        cache_len += T
    out["k_cache"], out["v_cache"] = hkc.get(), hvc.get()
    return out


for _B, _H, _hd, _S, _steps in [
        (3, 4, 4, 16, [1, 1, 1, 1, 1]),          # token by token from an empty cache
        (2, 2, 16, 40, [6, 1, 1, 1]),            # prefill then decode (the reference throws there, D10)
        (8, 12, 64, 300, [128, 1, 1]),           # GPT-2 heads: 96 (b,h) pairs = 3 warps of lanes, many key chunks
        (5, 3, 8, 24, [4, 4, 4]),                # chunks with T_new > 1 after the first: [T_q, T_k] triu mask quirk
        (1, 1, 32, 64, [33, 1]),                 # one head, partial warp
        (4, 40, 64, 20, [1, 2, 1])]:             # BH = 160 > 128
    @case(f"attention_decode_B{_B}_H{_H}_hd{_hd}_S{_S}_steps{'_'.join(map(str, _steps))}", tol=2e-5)
    def _c(be, rng, B=_B, H=_H, hd=_hd, S=_S, steps=_steps):
        return _decode(be, rng, B, H, hd, S, steps)


for _M, _K, _N, _al, _bl, _bias, _acc in [
        (8, 768, 96, "col", "col", 1, 0),        # a decode step of Linear(768, .) on 8 tokens
        (1, 300, 50, "col", "col", 1, 0),
        (16, 130, 33, "col", "col", 0, 1),
        (3, 64, 10, "row", "row", 1, 0),
        (5, 1000, 7, "col", "row", 0, 0)]:
    @case(f"matmul_skinny_{_M}x{_K}x{_N}_{_al}_{_bl}_bias{_bias}_acc{_acc}", tol=2e-5)
    def _c(be, rng, M=_M, K=_K, N=_N, al=_al, bl=_bl, bias=_bias, acc=_acc):
        lay = lambda r, c, kind: (1, r) if kind == "col" else (c, 1)
        as0, as1 = lay(M, K, al)
        bs0, bs1 = lay(K, N, bl)
        a, b, c = uni(rng, M * K + 3), uni(rng, K * N + 5), uni(rng, M * N)
        bv = uni(rng, N)
        ha, hb, hc, hbias = be.buf(a), be.buf(b), be.buf(c), be.buf(bv)
        be.call("matmul_skinny", ha, _mat(3, as0, as1, 0), hb, _mat(5, bs0, bs1, 0), hc, _mat(0, 1, M, 0), U32(M), U32(K), U32(N),
                hbias if bias else None, I32(acc))
        return {"c": hc.get()}


# bf16 tensor-core path: checked against the bf16-rounding model (operands RNE-rounded to bf16,
# exact products, wide accumulation). Only the fp32 accumulation order differs -> 1e-4; the error
# against the un-rounded fp32 product is bounded separately in test_kernels_gpu.py (<= 2e-2).
BF16_CASES = [
    (128, 64, 128, "col", "col", 1, 0),
    (256, 192, 256, "col", "col", 1, 0),
    (256, 192, 256, "col", "row", 1, 1),
    (192, 256, 320, "row", "col", 1, 1),
    (200, 100, 136, "row", "row", 1, 0),
    (130, 70, 150, "col", "col", 1, 0),
    (384, 96, 128, "col", "col", 3, 0),
    (128, 64, 128, "batchfast", "batchfast", 4, 0),
    (1024, 512, 768, "col", "col", 1, 0),
    # few output tiles, long K: split-K slices meeting in C by TMA reduce-add (weight-gradient shapes)
    (256, 2048, 256, "row", "col", 1, 0),
    (256, 2048, 256, "row", "col", 1, 1),
    (130, 1100, 200, "col", "col", 1, 0),
    (384, 1536, 128, "col", "row", 2, 1),
]
