"""GPU parity: every weedcu_* kernel (through the C-ABI, include/weedcu.h) against the CPU oracle
(oracle/liboracle.so, a restatement of the reference's CPU path) on identical seeded inputs."""
import numpy as np
import pytest

import cases
from backends import GpuBackend, OracleBackend

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    return GpuBackend()


@pytest.fixture(scope="module")
def oracle():
    return OracleBackend()


@pytest.mark.parametrize("name,fn,tol", cases.CASES, ids=[c[0] for c in cases.CASES])
def test_kernel_matches_oracle(gpu, oracle, name, fn, tol):
    seed = abs(hash(name)) % (2**31)
    seed = sum(ord(ch) * (i + 1) for i, ch in enumerate(name)) % (2**31)  # stable across processes
    ref = fn(oracle, np.random.default_rng(seed))
    got = fn(gpu, np.random.default_rng(seed))
    gpu.sync()
    assert ref.keys() == got.keys()
    for k in ref:
        if ref[k].dtype.kind in "iu" or tol == 0.0:
            assert np.array_equal(ref[k], got[k]), f"{name}:{k} not bit-exact"
        else:
            err = cases.rel_err(got[k], ref[k])
            assert np.all(np.isfinite(got[k])), f"{name}:{k} has non-finite values"
            assert err <= tol, f"{name}:{k} rel-to-max error {err:.3e} > {tol:.1e}"


@pytest.mark.parametrize("M,K,N,al,bl,batch,acc", cases.BF16_CASES,
                         ids=[f"bf16_{c[0]}x{c[1]}x{c[2]}_{c[3]}_{c[4]}_b{c[5]}_acc{c[6]}" for c in cases.BF16_CASES])
def test_matmul_bf16_tensor_core(gpu, oracle, M, K, N, al, bl, batch, acc):
    """tcgen05 path. (1) vs the bf16-rounding model: only accumulation order differs -> 1e-4.
    (2) vs the exact fp32 product: the stated bf16 bound, 2e-2 relative-to-max."""
    seed = 7000 + M + 3 * K + 5 * N + batch
    model = cases._matmul(oracle, np.random.default_rng(seed), M, K, N, al, bl, batch, acc, fn="matmul_bf16_model")
    exact = cases._matmul(oracle, np.random.default_rng(seed), M, K, N, al, bl, batch, acc, 0)
    got = cases._matmul(gpu, np.random.default_rng(seed), M, K, N, al, bl, batch, acc, 1)
    gpu.sync()
    e_model = cases.rel_err(got["c"], model["c"])
    e_exact = cases.rel_err(got["c"], exact["c"])
    assert e_model <= 1e-4, f"vs bf16 model: {e_model:.3e}"
    assert e_exact <= 2e-2, f"vs fp32: {e_exact:.3e}"


def test_launch_counter_counts(gpu):
    import ctypes as C
    n0, n1 = C.c_uint64(0), C.c_uint64(0)
    gpu.lib.weedcu_launch_count(C.byref(n0))
    h = gpu.buf(np.zeros(64, np.float32))
    gpu.call("fill_real", h, C.c_uint64(64), np.float32(2.0))
    gpu.lib.weedcu_launch_count(C.byref(n1))
    assert n1.value == n0.value + 1
    assert np.all(h.get() == 2.0)
