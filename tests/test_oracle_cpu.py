"""CPU-only: pins the C oracle (oracle/weed_oracle.c) before anything trusts it.

1. Known-answer vectors transcribed from the reference's own unit tests (tests/golden/ref_unit_vectors.json).
2. Outputs of the UNMODIFIED reference CPU build for seeded inputs: committed fixtures
   (tests/golden/ref_pins.npz, made by tests/golden/make_golden.py) and, when oracle/_ref/ is present,
   the live compiled reference through harness/weed_harness.cpp.
3. The reference defects the restatement reproduces or deliberately does not (reduce index order).
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

import cases
import refpins
from backends import OracleBackend
from cases import F32, I32, U32, U64, cview
from weed_b200._lib import Mat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libweed_ref_harness.so")


@pytest.fixture(scope="module")
def O():
    return OracleBackend()


def seed_of(name):
    return sum(ord(ch) * (i + 1) for i, ch in enumerate(name)) % (2**31)


def _mat(off, s0, s1):
    m = Mat()
    m.offset, m.s0, m.s1, m.batch_stride = off, s0, s1, 0
    return m


VECTORS = json.load(open(os.path.join(GOLDEN, "ref_unit_vectors.json")))["vectors"]


@pytest.mark.parametrize("v", VECTORS, ids=[v["name"] for v in VECTORS])
def test_reference_unit_test_vectors(O, v):
    op = v["op"]
    f = lambda a: np.asarray(a, F32)  # noqa: E731
    if op in ("sum", "mean"):
        x = f(v["x"])
        ho = O.buf(np.zeros(1, F32))
        O.call("sum_real", O.buf(x), cview(v["shape"]), F32(1.0 / x.size if op == "mean" else 1.0), ho)
        got = ho.get()
    elif op == "sum_axis":
        x = f(v["x"])
        ho = O.buf(np.zeros(x.size // v["shape"][v["axis"]], F32))
        O.call("reduce_real", O.buf(x), cview(v["shape"]), I32(v["axis"]), ho, I32(1))
        got = ho.get()
    elif op == "mul_scalar_tensor":
        b = f(v["b"])
        ho = O.buf(np.zeros(b.size, F32))
        sh = [b.size]
        O.call("binary_real", I32(1), O.buf(f(v["a"])), cases.make_view(sh, [0]), O.buf(b), cview(sh), ho, cview(sh))
        got = ho.get()
    elif op in ("matmul", "matmul_dA", "matmul_dB"):
        (M, K), (_, N) = v["a_shape"], v["b_shape"]
        if op == "matmul":
            hc = O.buf(np.zeros(M * N, F32))
            O.call("matmul_real", O.buf(f(v["a"])), _mat(0, 1, M), O.buf(f(v["b"])), _mat(0, 1, K), hc, _mat(0, 1, M),
                   U32(M), U32(K), U32(N), U32(1), I32(0))
            got = hc.get()
        elif op == "matmul_dA":  # loss = sum(C): dC = ones; dA = dC * B^T
            hd = O.buf(np.zeros(M * K, F32))
            O.call("matmul_real", O.buf(np.ones(M * N, F32)), _mat(0, 1, M), O.buf(f(v["b"])), _mat(0, K, 1), hd,
                   _mat(0, 1, M), U32(M), U32(N), U32(K), U32(1), I32(1))
            got = hd.get()
        else:  # dB = A^T * dC
            hd = O.buf(np.zeros(K * N, F32))
            O.call("matmul_real", O.buf(f(v["a"])), _mat(0, M, 1), O.buf(np.ones(M * N, F32)), _mat(0, 1, M), hd,
                   _mat(0, 1, K), U32(K), U32(M), U32(N), U32(1), I32(1))
            got = hd.get()
    elif op in ("softmax", "logsoftmax", "softmax_sum", "softmax_rowsum_axis1"):
        x = f(v["x"])
        axis = v.get("axis", 1)
        hy = O.buf(np.zeros(x.size, F32))
        O.call("softmax_real", I32(1 if op == "logsoftmax" else 0), O.buf(x), cview(v["shape"]), I32(axis), hy, cview(v["shape"]))
        y = hy.get()
        if op == "softmax_sum":
            got = np.array([y.sum()], F32)
        elif op == "softmax_rowsum_axis1":
            got = np.array([y[0] + y[2] + y[4], y[1] + y[3] + y[5]], F32)  # col-major rows (tests.cpp:1214-1218)
        else:
            got = y
    elif op in ("softmax_grad_pick", "logsoftmax_grad_pick"):
        x = f(v["x"])
        lm = 1 if op.startswith("log") else 0
        hy = O.buf(np.zeros(x.size, F32))
        O.call("softmax_real", I32(lm), O.buf(x), cview(v["shape"]), I32(0), hy, cview(v["shape"]))
        dout = np.zeros(x.size, F32)
        dout[v["pick"]] = 1.0
        hd = O.buf(np.zeros(x.size, F32))
        O.call("softmax_grad_real", I32(lm), hd, cview(v["shape"]), hy, cview(v["shape"]), O.buf(dout), cview(v["shape"]), I32(0))
        got = hd.get()
    elif op in ("relu", "sigmoid", "tanh"):
        x = f(v["x"])
        hy = O.buf(np.zeros(x.size, F32))
        O.call("unary_real", I32({"relu": 0, "sigmoid": 1, "tanh": 2}[op]), F32(0), O.buf(x), cview(v["shape"]), hy, cview(v["shape"]))
        got = hy.get()
    elif op in ("relu_grad", "sigmoid_grad", "tanh_grad"):
        src = f(v.get("x", v.get("y")))
        hd = O.buf(np.zeros(src.size, F32))
        sh = [src.size]
        O.call("unary_grad_real", I32({"relu_grad": 0, "sigmoid_grad": 1, "tanh_grad": 2}[op]), hd, cview(sh), O.buf(src), cview(sh),
               O.buf(np.ones(src.size, F32)), cview(sh), I32(1))
        got = hd.get()
    else:
        raise AssertionError(f"unhandled vector op {op}")
    exp = f(v["expect"])
    if v["tol"] == 0:
        assert np.array_equal(got, exp), f"{v['name']} ({v['ref']}): {got} != {exp}"
    else:
        assert np.max(np.abs(got - exp)) <= v["tol"], f"{v['name']} ({v['ref']}): {got} vs {exp}"


def _run_oracle(orc, O, inp, ref_out):
    import inspect
    return orc(O, inp, ref_out) if len(inspect.signature(orc).parameters) == 3 else orc(O, inp)


@pytest.mark.parametrize("name,fn,tol", refpins.PINS, ids=[p[0] for p in refpins.PINS])
def test_oracle_matches_reference_fixtures(O, name, fn, tol):
    """Oracle vs stored outputs of the compiled reference (always available)."""
    path = os.path.join(GOLDEN, "ref_pins.npz")
    assert os.path.exists(path), "tests/golden/ref_pins.npz missing: run tests/golden/make_golden.py"
    z = np.load(path)
    inp, _ref, orc = fn(np.random.default_rng(seed_of(name)))
    for k, v in inp.items():  # the fixture was generated from these exact inputs
        assert np.array_equal(z[f"{name}/in/{k}"], v), f"{name}: seeded input {k} drifted from the fixture"
    ref_out = {k.split("/out/")[1]: z[k] for k in z.files if k.startswith(f"{name}/out/")}
    got = _run_oracle(orc, O, inp, ref_out)
    assert got.keys() == ref_out.keys()
    for k in got:
        err = cases.rel_err(got[k], ref_out[k])
        assert err <= tol, f"{name}:{k} oracle vs reference fixture rel-to-max {err:.3e} > {tol:.1e}"


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("name,fn,tol", refpins.PINS, ids=[p[0] for p in refpins.PINS])
def test_oracle_matches_live_reference(O, name, fn, tol):
    """Oracle vs the compiled reference run right now (build container / GPU box with oracle/_ref)."""
    from weed_b200.harness import Harness
    R = Harness.reference()
    inp, ref, orc = fn(np.random.default_rng(seed_of(name)))
    ref_out = ref(R, inp)
    R.reset()
    got = _run_oracle(orc, O, inp, ref_out)
    for k in got:
        err = cases.rel_err(got[k], ref_out[k])
        assert err <= tol, f"{name}:{k} oracle vs live reference rel-to-max {err:.3e} > {tol:.1e}"


def test_reduce_reference_order_is_a_permutation_of_intended_order(O):
    """Documents the reference defect the oracle models with index_order: for rank >= 3
    with two non-axis dims > 1 the CPU loop stores sums in row-major output order."""
    shape, axis = [2, 3, 4], 2
    x = np.arange(24, dtype=F32)
    out0, out1 = O.buf(np.zeros(6, F32)), O.buf(np.zeros(6, F32))
    O.call("reduce_real", O.buf(x), cview(shape), I32(axis), out0, I32(0))
    O.call("reduce_real", O.buf(x), cview(shape), I32(axis), out1, I32(1))
    intended = x.reshape(4, 3, 2).sum(axis=0).ravel()  # numpy C-order of reversed dims == col-major
    assert np.array_equal(out0.get(), intended)
    assert np.array_equal(out1.get(), np.array([36, 44, 52, 40, 48, 56], F32))  # measured on the reference build
    assert sorted(out0.get()) == sorted(out1.get()) and not np.array_equal(out0.get(), out1.get())


def test_all_gpu_cases_run_on_the_oracle(O):
    """Every case the GPU parity suite uses must at least execute and stay finite on the oracle."""
    for name, fn, _tol in cases.CASES:
        out = fn(O, np.random.default_rng(seed_of(name)))
        for k, v in out.items():
            if v.dtype.kind == "f":
                assert np.all(np.isfinite(v)), f"{name}:{k}"


def test_cross_entropy_backward_is_softmax_minus_onehot(O):
    """The oracle's CE backward (analytic; the reference's own is identically zero, see refpins)."""
    rows, V = 6, 9
    rng = np.random.default_rng(5)
    logits = rng.uniform(-3, 3, size=rows * V).astype(F32)
    tg = rng.integers(0, V, size=rows).astype(np.int32)
    hl, ht = O.buf(logits), O.buf(tg)
    hlse, hloss = O.buf(np.zeros(rows, F32)), O.buf(np.zeros(1, F32))
    O.call("cross_entropy_fwd", hl, U64(0), U32(rows), U32(V), U32(1), U32(rows), ht, hlse, hloss)
    hd = O.buf(np.zeros(rows * V, F32))
    O.call("cross_entropy_bwd", hl, U64(0), U32(rows), U32(V), U32(1), U32(rows), ht, hlse, O.buf(np.ones(1, F32)), hd, U64(0), I32(1))
    x = logits.astype(np.float64).reshape(V, rows).T  # [rows, V]
    p = np.exp(x - x.max(1, keepdims=True))
    p /= p.sum(1, keepdims=True)
    oh = np.eye(V)[tg]
    want = ((p - oh) / rows).T.ravel()
    assert cases.rel_err(hd.get(), want) <= 1e-5
    assert abs(hloss.get()[0] - (-np.mean(np.log(p[np.arange(rows), tg])))) <= 1e-5
