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


ATTN_CASES = [  # B, T, H, hd, causal
    (2, 64, 2, 16, 1), (3, 128, 4, 64, 1), (1, 72, 3, 32, 1), (2, 200, 2, 64, 0), (8, 256, 2, 64, 1),
    # head_dim 64 runs the flash kernel (scores in TMEM): full GPT-2 sequence, ragged last tiles, T > 1024
    (1, 1024, 2, 64, 1), (1, 1160, 1, 64, 1), (2, 328, 1, 64, 0), (1, 64, 1, 64, 1),
]


@pytest.mark.parametrize("B,T,H,hd,causal", ATTN_CASES, ids=[f"attn_B{c[0]}_T{c[1]}_H{c[2]}_hd{c[3]}_causal{c[4]}" for c in ATTN_CASES])
def test_attention_fwd_bf16(gpu, oracle, B, T, H, hd, causal):
    """Fused attention core (heads relayout + bf16 pack, tcgen05 QK^T, softmax -> bf16 P, tcgen05 PV).
    (1) vs the oracle model with the same bf16 rounding points: accumulation order and one-ulp flips
    of rounded probabilities (the flash kernel rounds p = 2^(t - m) against the RUNNING row maximum,
    the model against the final one, so mantissas differ by a non-power-of-two factor before the
    rounding) -> 4e-3 relative-to-max, i.e. two bf16 ulps (2^-8);
    (2) vs the exact fp32 chain (reference arithmetic, multihead_attention.cpp:289-345): the bf16
    bound, 2e-2 relative-to-max."""
    import ctypes as C
    rng = np.random.default_rng(9000 + B + T + H + hd)
    n = B * T * H * hd
    q, k, v = (rng.uniform(-1, 1, n).astype(np.float32) for _ in range(3))
    div, mask = np.float32(np.sqrt(hd)), np.float32(-1.701411835e38)

    def run(be):
        hq, hk, hv, ho = be.buf(q), be.buf(k), be.buf(v), be.buf(np.zeros(n, np.float32))
        be.call("attention_fwd", hq, hk, hv, ho, C.c_uint32(B), C.c_uint32(T), C.c_uint32(H), C.c_uint32(hd), div, mask, C.c_int(causal))
        return ho.get()

    got, model = run(gpu), run(oracle)
    gpu.sync()
    # exact fp32 chain in numpy: x[b, t, c] lives at b + B*t + B*T*c
    def heads(x):
        return x.reshape(hd, H, T, B).transpose(3, 1, 2, 0).astype(np.float64)  # feature c = h + H*j -> [B, H, T, hd]
    Q, K, V = heads(q), heads(k), heads(v)
    S = Q @ K.transpose(0, 1, 3, 2) / float(div)
    if causal:
        S = S + np.triu(np.full((T, T), float(mask)), 1)
    S = S - S.max(-1, keepdims=True)
    Pm = np.exp(S)
    Pm /= Pm.sum(-1, keepdims=True)
    exact = (Pm @ V).transpose(3, 1, 2, 0).reshape(-1).astype(np.float32)  # back to [B, T, H*hd] col-major (c = h + H*j)
    assert np.all(np.isfinite(got))
    assert cases.rel_err(got, model) <= 4e-3, f"vs bf16 model: {cases.rel_err(got, model):.3e}"
    assert cases.rel_err(got, exact) <= 2e-2, f"vs fp32 chain: {cases.rel_err(got, exact):.3e}"


def test_attention_fwd_unsupported_shapes_say_so(gpu):
    """Outside the tensor-map envelope the entry returns WEEDCU_ENOSUP (callers compose the generic
    ops) instead of computing something else."""
    import ctypes as C
    h = gpu.buf(np.zeros(4 * 60 * 16, np.float32))
    fn = gpu.lib.weedcu_attention_fwd
    fn.restype = C.c_int
    rc = fn(C.c_void_p(h.ptr), C.c_void_p(h.ptr), C.c_void_p(h.ptr), C.c_void_p(h.ptr), C.c_uint32(4), C.c_uint32(60), C.c_uint32(1), C.c_uint32(16),
            C.c_float(4.0), C.c_float(-1e38), C.c_int(1), C.c_void_p(gpu.stream))
    assert rc == -2


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (300, 200, 96), (8192, 768, 256), (257, 130, 72)])
def test_gemm_bf16_col_bias_epilogue(gpu, oracle, M, N, K):
    """weedcu_pack_bf16 + weedcu_gemm_bf16 with the column bias added in the epilogue (TMA-store and
    direct-store paths, ragged tiles) against the bf16-rounding model + a broadcast add."""
    import ctypes as C
    rng = np.random.default_rng(M + N + K)
    a = rng.uniform(-1, 1, M * K).astype(np.float32)   # [M, K] M contiguous
    b = rng.uniform(-1, 1, K * N).astype(np.float32)   # [K, N] K contiguous
    bias = rng.uniform(-2, 2, N).astype(np.float32)
    r8 = lambda x: (x + 7) // 8 * 8
    ha, hb, hbias = gpu.buf(a), gpu.buf(b), gpu.buf(bias)
    pa, pb = gpu.buf(np.zeros(r8(M) * K + 8, np.uint16)), gpu.buf(np.zeros(r8(K) * N + 8, np.uint16))
    hc = gpu.buf(np.zeros(M * N, np.float32))
    U32, U64, I32 = C.c_uint32, C.c_uint64, C.c_int
    gpu.call("pack_bf16", ha, U64(0), U32(1), U32(M), U32(M), U32(K), pa, I32(1))   # rows = m (stride 1), cols = k
    gpu.call("pack_bf16", hb, U64(0), U32(K), U32(1), U32(N), U32(K), pb, I32(0))   # rows = n (stride K), cols = k
    gpu.call("gemm_bf16", pa, I32(1), U64(r8(M)), pb, I32(0), U64(r8(K)), hc, U64(M), U32(M), U32(N), U32(K), I32(0), hbias)
    got = hc.get().reshape(N, M).T
    gpu.sync()
    bf = lambda x: (lambda u: ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32).view(np.float32))(x.view(np.uint32).astype(np.uint64))
    A = bf(a).reshape(K, M).T.astype(np.float64)
    B = bf(b).reshape(N, K).T.astype(np.float64)
    want = (A @ B + bias[None, :].astype(np.float64)).astype(np.float32)
    assert cases.rel_err(got, want) <= 1e-4


def _bf16_bits(x):
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    return (((u + 0x7FFF + ((u >> 16) & 1)) >> 16) & 0xFFFF).astype(np.uint16)


def _bf16_widen(bits):
    return (bits.astype(np.uint32) << 16).view(np.float32)


def _operand(rng, mn, k, major):
    """bf16 [mn, k] operand in the layout weedcu_gemm_bf16 takes: major 1 = mn contiguous, 0 = k contiguous;
    returns (device array with the padded leading dimension, ld, exact fp64 values)."""
    r8 = lambda x: (x + 7) // 8 * 8
    bits = _bf16_bits(rng.uniform(-1, 1, (mn, k)))
    ld = r8(mn) if major else r8(k)
    dev = np.zeros((k, ld) if major else (mn, ld), np.uint16)
    if major:
        dev[:, :mn] = bits.T
    else:
        dev[:, :k] = bits
    return dev.reshape(-1), ld, _bf16_widen(bits).astype(np.float64)


PAIR_CASES = [  # M, N, K, a_major, b_major, mode (pair * 1e6 + BLOCK_N * 1e3 + splits), accumulate, bias
    (256, 256, 64, 1, 0, 1256001, 0, 0), (256, 256, 64, 0, 0, 1256001, 0, 0), (256, 256, 64, 1, 1, 1256001, 0, 0),
    (1024, 512, 768, 1, 0, 1256001, 0, 1), (1024, 512, 768, 1, 1, 1256001, 1, 0), (1024, 512, 768, 0, 0, 1256001, 0, 0),
    (1024, 384, 768, 1, 0, 1192001, 0, 1), (1024, 384, 768, 0, 0, 1192001, 1, 0),
    (1024, 384, 512, 1, 0, 1128001, 0, 1), (1024, 384, 512, 1, 1, 1128001, 0, 0), (1024, 384, 512, 0, 0, 1128001, 1, 0),
    (768, 1000, 2048, 0, 0, 1256004, 0, 0), (768, 1000, 2048, 0, 0, 1256004, 1, 0), (768, 520, 1024, 1, 1, 1128002, 0, 0),
    (300, 200, 96, 1, 0, 1256001, 0, 1), (900, 330, 200, 1, 1, 1256001, 0, 0), (388, 520, 136, 0, 0, 1192001, 0, 1),
    (8192, 768, 768, 1, 0, 1256001, 0, 1), (8192, 3072, 768, 1, 0, 1192001, 0, 1), (8192, 768, 3072, 1, 1, 1128001, 1, 0),
    (4104, 1544, 264, 1, 0, 0, 0, 1), (4104, 1544, 264, 1, 1, 2, 1, 0), (768, 3072, 8192, 0, 0, 2, 1, 0),
]


@pytest.mark.parametrize("M,N,K,am,bm,mode,acc,bias", PAIR_CASES, ids=[f"{c[0]}x{c[1]}x{c[2]}_a{c[3]}b{c[4]}_mode{c[5]}_acc{c[6]}_bias{c[7]}" for c in PAIR_CASES])
def test_gemm_bf16_cta_pair_matches_model(gpu, M, N, K, am, bm, mode, acc, bias):
    """The cta_group::2 kernel (two SMs on one 256 x BLOCK_N tile, leader-issued MMA, remote barrier
    arrivals) against the exact product of the bf16-rounded operands: every operand majorness, all three
    tile widths, split-K slices meeting by reduce-add, C +=, the bias epilogue, ragged edges where the
    peer CTA's half of a tile is partly or wholly out of range."""
    import ctypes as C
    rng = np.random.default_rng(M * 7 + N * 3 + K + mode)
    U32, U64, I32 = C.c_uint32, C.c_uint64, C.c_int
    a_dev, lda, A = _operand(rng, M, K, am)
    b_dev, ldb, B = _operand(rng, N, K, bm)
    c0 = rng.uniform(-1, 1, M * N).astype(np.float32) if acc else np.full(M * N, 7.0, np.float32)
    hb = rng.uniform(-2, 2, N).astype(np.float32)
    pa, pb, hc, hbias = gpu.buf(a_dev), gpu.buf(b_dev), gpu.buf(c0), gpu.buf(hb)
    assert gpu.lib.weedcu_gemm_set_mode(C.c_int(mode)) == 0
    try:
        gpu.call("gemm_bf16", pa, I32(am), U64(lda), pb, I32(bm), U64(ldb), hc, U64(M), U32(M), U32(N), U32(K), I32(acc),
                 hbias if bias else C.c_void_p(0))
        got = hc.get().reshape(N, M).T
        gpu.sync()
    finally:
        gpu.lib.weedcu_gemm_set_mode(C.c_int(0))
    want = A @ B.T
    if bias:
        want = want + hb[None, :].astype(np.float64)
    if acc:
        want = want + c0.reshape(N, M).T.astype(np.float64)
    assert np.all(np.isfinite(got))
    assert cases.rel_err(got, want.astype(np.float32)) <= 1e-4


@pytest.mark.parametrize("M,N,K,am,bm,mode,splits_hint", [(8192, 768, 768, 1, 0, 1192001, 1), (4096, 1024, 512, 1, 1, 1256001, 1), (2048, 520, 4096, 0, 0, 1128002, 2),
                                                          (1100, 1000, 320, 1, 0, 1128001, 1)])
def test_gemm_bf16_cta_pair_dynamic_tile_scheduler(gpu, M, N, K, am, bm, mode, splits_hint):
    """weedcu_gemm_set_dynamic(1): the clusters of the CTA-pair kernel draw their work units from a per-launch counter through
    the shared-memory unit ring instead of striding over them. Bit-identical to the static schedule (a unit is computed the
    same way by whichever cluster takes it), also over many consecutive launches (counter slots are re-armed by the launch
    that used them) and with more units than clusters (several rounds through the 8-entry ring)."""
    import ctypes as C
    rng = np.random.default_rng(M + N + K)
    U32, U64, I32 = C.c_uint32, C.c_uint64, C.c_int
    a_dev, lda, A = _operand(rng, M, K, am)
    b_dev, ldb, B = _operand(rng, N, K, bm)
    pa, pb = gpu.buf(a_dev), gpu.buf(b_dev)
    out = {}
    assert gpu.lib.weedcu_gemm_set_mode(C.c_int(mode)) == 0
    try:
        for dyn in (0, 1):
            assert gpu.lib.weedcu_gemm_set_dynamic(C.c_int(dyn)) == 0
            hc = gpu.buf(np.zeros(M * N, np.float32))
            for _ in range(300 if dyn else 1):  # more launches than counter slots
                gpu.call("gemm_bf16", pa, I32(am), U64(lda), pb, I32(bm), U64(ldb), hc, U64(M), U32(M), U32(N), U32(K), I32(0), C.c_void_p(0))
            gpu.sync()
            out[dyn] = hc.get()
    finally:
        gpu.lib.weedcu_gemm_set_dynamic(C.c_int(0))
        gpu.lib.weedcu_gemm_set_mode(C.c_int(0))
    if splits_hint == 1:
        assert np.array_equal(out[0], out[1])
    else:  # split-K slices meet by reduce-add: the order of the adds is not fixed in either schedule
        assert cases.rel_err(out[1], out[0]) <= 1e-6
    assert cases.rel_err(out[1].reshape(N, M).T, (A @ B.T).astype(np.float32)) <= 1e-4


@pytest.mark.parametrize("mode", [256001, 192001, 1256001, 1128001])
@pytest.mark.parametrize("M,N,K,groups", [(8192, 768, 768, 3), (300, 200, 96, 2), (1024, 130, 256, 3), (256, 64, 64, 1)])
def test_gemm_bf16_grouped_equals_separate_launches(gpu, M, N, K, groups, mode):
    """weedcu_gemm_bf16_grouped (the W_q / W_k / W_v projections as one launch) is bit-identical to
    `groups` weedcu_gemm_bf16 calls: same tiles, same k order, only the tile loop is shared."""
    import ctypes as C
    rng = np.random.default_rng(M + N + K + groups)
    r8 = lambda x: (x + 7) // 8 * 8
    U32, U64, I32 = C.c_uint32, C.c_uint64, C.c_int
    gpu.lib.weedcu_gemm_set_mode(C.c_int(mode))  # one fixed tile configuration (single-CTA or CTA-pair) for both sides
    a = rng.uniform(-1, 1, M * K).astype(np.float32)
    ha = gpu.buf(a)
    pa = gpu.buf(np.zeros(r8(M) * K + 8, np.uint16))
    gpu.call("pack_bf16", ha, U64(0), U32(1), U32(M), U32(M), U32(K), pa, I32(1))
    pbs, biases, sep, grp = [], [], [], []
    for g in range(groups):
        b = rng.uniform(-1, 1, K * N).astype(np.float32)
        hb = gpu.buf(b)
        pb = gpu.buf(np.zeros(r8(K) * N + 8, np.uint16))
        gpu.call("pack_bf16", hb, U64(0), U32(K), U32(1), U32(N), U32(K), pb, I32(0))
        pbs.append(pb)
        biases.append(gpu.buf(rng.uniform(-2, 2, N).astype(np.float32)))
        sep.append(gpu.buf(np.zeros(M * N, np.float32)))
        grp.append(gpu.buf(np.full(M * N, 5.0, np.float32)))
        gpu.call("gemm_bf16", pa, I32(1), U64(r8(M)), pb, I32(0), U64(r8(K)), sep[g], U64(M), U32(M), U32(N), U32(K), I32(0), biases[g])
    PtrArr = C.c_void_p * groups
    gpu.call("gemm_bf16_grouped", pa, I32(1), U64(r8(M)), U32(groups), PtrArr(*[p.ptr for p in pbs]), I32(0), U64(r8(K)),
             PtrArr(*[c.ptr for c in grp]), U64(M), U32(M), U32(N), U32(K), I32(0), PtrArr(*[b.ptr for b in biases]))
    gpu.lib.weedcu_gemm_set_mode(C.c_int(0))
    for g in range(groups):
        assert np.array_equal(sep[g].get(), grp[g].get()), f"group {g}"


@pytest.mark.parametrize("M,K,N,groups", [(8, 768, 768, 3), (8, 768, 96, 2), (3, 70, 37, 3), (16, 130, 33, 1)])
def test_matmul_skinny_grouped_equals_separate_launches(gpu, M, K, N, groups):
    """weedcu_matmul_skinny_grouped (the W_q / W_k / W_v projections of a decode step as one launch, blockIdx.y = product)
    is bit-identical to `groups` weedcu_matmul_skinny calls."""
    import ctypes as C
    rng = np.random.default_rng(M + K + N + groups)
    U32 = C.c_uint32
    a = rng.uniform(-1, 1, M * K + 4).astype(np.float32)
    ha = gpu.buf(a)
    am, bm, cm = cases._mat(4, 1, M, 0), cases._mat(8, 1, K, 0), cases._mat(0, 1, M, 0)
    hbs, hbias, sep, grp = [], [], [], []
    for g in range(groups):
        hbs.append(gpu.buf(rng.uniform(-1, 1, K * N + 8).astype(np.float32)))
        hbias.append(gpu.buf(rng.uniform(-2, 2, N).astype(np.float32)))
        sep.append(gpu.buf(np.zeros(M * N, np.float32)))
        grp.append(gpu.buf(np.full(M * N, 5.0, np.float32)))
        gpu.call("matmul_skinny", ha, am, hbs[g], bm, sep[g], cm, U32(M), U32(K), U32(N), hbias[g], C.c_int(0))
    PtrArr = C.c_void_p * groups
    gpu.call("matmul_skinny_grouped", ha, am, U32(groups), PtrArr(*[b.ptr for b in hbs]), bm, PtrArr(*[c.ptr for c in grp]), cm,
             U32(M), U32(K), U32(N), PtrArr(*[b.ptr for b in hbias]))
    for g in range(groups):
        assert np.array_equal(sep[g].get(), grp[g].get()), f"group {g}"


@pytest.mark.parametrize("mode", [0, 256001, 1256001, 1192001, 1128002])
@pytest.mark.parametrize("M,N,K", [(8192, 768, 768), (1024, 520, 264), (388, 200, 96)])
def test_gemm_bf16_residual_equals_gemm_then_add(gpu, M, N, K, mode):
    """weedcu_gemm_bf16_residual (the `x + Linear(...)` of a transformer block with the add in the GEMM epilogue) is
    bit-identical to weedcu_gemm_bf16 followed by the fp32 add, in the single-CTA and the CTA-pair kernels, with the
    bias, ragged tiles and split-K slices (only the first slice adds bias and residual)."""
    import ctypes as C
    rng = np.random.default_rng(M + N + K)
    U32, U64, I32 = C.c_uint32, C.c_uint64, C.c_int
    a_dev, lda, _ = _operand(rng, M, K, 1)
    b_dev, ldb, _ = _operand(rng, N, K, 0)
    res = rng.uniform(-3, 3, M * N).astype(np.float32)
    bias = rng.uniform(-2, 2, N).astype(np.float32)
    pa, pb, hres, hbias = gpu.buf(a_dev), gpu.buf(b_dev), gpu.buf(res), gpu.buf(bias)
    plain, fused = gpu.buf(np.zeros(M * N, np.float32)), gpu.buf(np.full(M * N, 9.0, np.float32))
    gpu.lib.weedcu_gemm_set_mode(C.c_int(mode))
    try:
        gpu.call("gemm_bf16", pa, I32(1), U64(lda), pb, I32(0), U64(ldb), plain, U64(M), U32(M), U32(N), U32(K), I32(0), hbias)
        gpu.call("gemm_bf16_residual", pa, I32(1), U64(lda), pb, I32(0), U64(ldb), fused, U64(M), U32(M), U32(N), U32(K), hbias, hres, U64(M))
        want = plain.get() + res
        got = fused.get()
    finally:
        gpu.lib.weedcu_gemm_set_mode(C.c_int(0))
    if mode % 100 > 1:  # split-K: the slice that adds the residual is reduce-added in a different order than C + residual
        assert cases.rel_err(got, want) <= 1e-6
    else:
        assert np.array_equal(got, want)


@pytest.mark.parametrize("M,K,N,bias", [(8, 768, 768, 1), (8, 3072, 768, 1), (5, 70, 37, 0), (16, 130, 33, 1)])
def test_matmul_skinny_residual_equals_skinny_then_add(gpu, M, K, N, bias):
    """weedcu_matmul_skinny_residual (x + Linear(...) of a decode step in one launch) == weedcu_matmul_skinny + fp32 add."""
    import ctypes as C
    rng = np.random.default_rng(M + K + N)
    U32 = C.c_uint32
    ha = gpu.buf(rng.uniform(-1, 1, M * K).astype(np.float32))
    hb = gpu.buf(rng.uniform(-1, 1, K * N).astype(np.float32))
    res = rng.uniform(-3, 3, M * N).astype(np.float32)
    hres, hbias = gpu.buf(res), gpu.buf(rng.uniform(-2, 2, N).astype(np.float32))
    plain, fused = gpu.buf(np.zeros(M * N, np.float32)), gpu.buf(np.full(M * N, 9.0, np.float32))
    am, bm, cm = cases._mat(0, 1, M, 0), cases._mat(0, 1, K, 0), cases._mat(0, 1, M, 0)
    gpu.call("matmul_skinny", ha, am, hb, bm, plain, cm, U32(M), U32(K), U32(N), hbias if bias else None, C.c_int(0))
    gpu.call("matmul_skinny_residual", ha, am, hb, bm, fused, cm, U32(M), U32(K), U32(N), hbias if bias else None, hres)
    assert np.array_equal(fused.get(), plain.get() + res)


EX_SHAPES = [(8192, 768, 768), (1024, 520, 264), (392, 200, 96), (2048, 5003, 128)]


def _merge_ln(stats, tiles, cols, N):
    """Chan merge of per-tile (mean, M2) partials -> (mean, variance) per row, in fp64"""
    n, mean, m2 = 0.0, 0.0, 0.0
    for t in range(tiles):
        cnt = min(N, (t + 1) * cols) - t * cols
        if cnt <= 0:
            continue
        mt, m2t = stats[t, :, 0].astype(np.float64), stats[t, :, 1].astype(np.float64)
        tot = n + cnt
        delta = mt - mean
        mean = mean + delta * cnt / tot
        m2 = m2 + m2t + delta * delta * n * cnt / tot
        n = tot
    return mean, m2 / N


@pytest.mark.parametrize("mode", [0, 256001, 1256001, 1192001, 1128001])
@pytest.mark.parametrize("M,N,K", EX_SHAPES)
def test_gemm_bf16_ex_epilogue(gpu, M, N, K, mode):
    """weedcu_gemm_bf16_ex: the extended epilogue of the tensor-core GEMM in every tile family.
    (a) fp32 C + bias + residual + LayerNorm partials: C bit-identical to weedcu_gemm_bf16_residual, merged (mean, var) == numpy on C.
    (b) bf16-only output + log-sum-exp partials: bf16 copy == RNE(bf16) of the plain product, merged lse == numpy.
    (c) fp32 C + bf16 copy of gelu(C): C bit-identical, copy within 2 bf16 ulps of bf16(gelu(C)) (hardware tanh)."""
    import ctypes as C
    from weed_b200._lib import GemmEpilogue
    rng = np.random.default_rng(M + N + K + 11)
    U32, U64, I32 = C.c_uint32, C.c_uint64, C.c_int
    a_dev, lda, _ = _operand(rng, M, K, 1)
    b_dev, ldb, _ = _operand(rng, N, K, 0)
    res = rng.uniform(-3, 3, M * N).astype(np.float32)
    bias = rng.uniform(-2, 2, N).astype(np.float32)
    pa, pb, hres, hbias = gpu.buf(a_dev), gpu.buf(b_dev), gpu.buf(res), gpu.buf(bias)
    plain, plain_res = gpu.buf(np.zeros(M * N, np.float32)), gpu.buf(np.zeros(M * N, np.float32))
    cap = 2 * ((N + 127) // 128)
    r8 = (M + 7) // 8 * 8
    if M % 8:
        pytest.skip("bf16 outputs need a leading dimension that is a multiple of 8 == M for a dense operand copy")
    gpu.lib.weedcu_gemm_set_mode(C.c_int(mode))
    try:
        gpu.call("gemm_bf16", pa, I32(1), U64(lda), pb, I32(0), U64(ldb), plain, U64(M), U32(M), U32(N), U32(K), I32(0), hbias)
        gpu.call("gemm_bf16_residual", pa, I32(1), U64(lda), pb, I32(0), U64(ldb), plain_res, U64(M), U32(M), U32(N), U32(K), hbias, hres, U64(M))
        want, want_res = plain.get().reshape(N, M), plain_res.get().reshape(N, M)

        def run(c, c16, **kw):
            tiles, cols = U32(0), U32(0)
            stats = gpu.buf(np.full(cap * M * 2, np.nan, np.float32))
            e = GemmEpilogue(col_bias=hbias.ptr, residual=kw.get("residual", 0), ldr=M, activation=kw.get("act", 0), row_stats=kw.get("stats", 0),
                             stats=stats.ptr, stats_capacity_tiles=cap, stats_tiles=C.pointer(tiles), stats_tile_cols=C.pointer(cols))
            gpu.call("gemm_bf16_ex", pa, I32(1), U64(lda), pb, I32(0), U64(ldb), c, U64(M), c16, U64(r8), U32(M), U32(N), U32(K), e)
            return stats.get().reshape(cap, M, 2), tiles.value, cols.value

        # (a)
        c = gpu.buf(np.full(M * N, 9.0, np.float32))
        st, tiles, cols = run(c, None, residual=hres.ptr, stats=1)
        assert np.array_equal(c.get().reshape(N, M), want_res)
        assert tiles >= (N + cols - 1) // cols and tiles <= cap
        mean, var = _merge_ln(st, tiles, cols, N)
        w64 = want_res.astype(np.float64)
        assert np.max(np.abs(mean - w64.mean(0))) <= 1e-6 * np.abs(w64).max()
        assert np.max(np.abs(var - w64.var(0)) / w64.var(0)) <= 1e-5
        # (b)
        c16 = gpu.buf(np.full(r8 * N, 0x7FC0, np.uint16))
        st, tiles, cols = run(None, c16, stats=2)
        assert np.array_equal(c16.get().reshape(N, r8)[:, :M], _bf16_bits(want))
        mx = np.max(st[:tiles, :, 0], axis=0).astype(np.float64)
        ssum = np.sum(st[:tiles, :, 1].astype(np.float64) * np.exp(st[:tiles, :, 0].astype(np.float64) - mx[None, :]), axis=0)
        lse = mx + np.log(ssum)
        w64 = want.astype(np.float64)
        ref_lse = w64.max(0) + np.log(np.exp(w64 - w64.max(0)[None, :]).sum(0))
        assert np.max(np.abs(lse - ref_lse)) <= 2e-6 * np.abs(ref_lse).max()
        # (c)
        c = gpu.buf(np.full(M * N, 9.0, np.float32))
        c16 = gpu.buf(np.full(r8 * N, 0x7FC0, np.uint16))
        run(c, c16, act=1)
        assert np.array_equal(c.get().reshape(N, M), want)
        x = want.astype(np.float64)
        g = 0.5 * x * (1.0 + np.tanh(0.7978845608028654 * (x + 0.044715 * x ** 3)))
        got = _bf16_widen(c16.get().reshape(N, r8)[:, :M]).astype(np.float64)
        assert np.max(np.abs(got - g) / np.maximum(np.abs(g), 0.25)) <= 2 ** -7   # 2 ulps of bf16 (ulp = 2^-8 relative), floor at 0.25
    finally:
        gpu.lib.weedcu_gemm_set_mode(C.c_int(0))


@pytest.mark.parametrize("mode", [0, 256001, 1256001, 1128001])
@pytest.mark.parametrize("M,N,K,groups", [(8192, 768, 768, 3), (1024, 136, 256, 2)])
def test_gemm_bf16_grouped_bf16out(gpu, M, N, K, groups, mode):
    """weedcu_gemm_bf16_grouped_bf16out == RNE(bf16) of weedcu_gemm_bf16_grouped's fp32 outputs, bit for bit."""
    import ctypes as C
    rng = np.random.default_rng(M + N + K + groups + 5)
    U32, U64, I32 = C.c_uint32, C.c_uint64, C.c_int
    a_dev, lda, _ = _operand(rng, M, K, 1)
    pa = gpu.buf(a_dev)
    pbs, biases, f32, b16 = [], [], [], []
    ldb = None
    for g in range(groups):
        b_dev, ldb, _ = _operand(rng, N, K, 0)
        pbs.append(gpu.buf(b_dev))
        biases.append(gpu.buf(rng.uniform(-2, 2, N).astype(np.float32)))
        f32.append(gpu.buf(np.zeros(M * N, np.float32)))
        b16.append(gpu.buf(np.full(M * N, 0x7FC0, np.uint16)))
    PtrArr = C.c_void_p * groups
    gpu.lib.weedcu_gemm_set_mode(C.c_int(mode))
    try:
        gpu.call("gemm_bf16_grouped", pa, I32(1), U64(lda), U32(groups), PtrArr(*[p.ptr for p in pbs]), I32(0), U64(ldb),
                 PtrArr(*[c.ptr for c in f32]), U64(M), U32(M), U32(N), U32(K), I32(0), PtrArr(*[b.ptr for b in biases]))
        gpu.call("gemm_bf16_grouped_bf16out", pa, I32(1), U64(lda), U32(groups), PtrArr(*[p.ptr for p in pbs]), I32(0), U64(ldb),
                 PtrArr(*[c.ptr for c in b16]), U64(M), U32(M), U32(N), U32(K), PtrArr(*[b.ptr for b in biases]))
        for g in range(groups):
            assert np.array_equal(b16[g].get(), _bf16_bits(f32[g].get())), f"group {g}"
    finally:
        gpu.lib.weedcu_gemm_set_mode(C.c_int(0))


@pytest.mark.parametrize("rows,V,K", [(1024, 5003, 64), (256, 1000, 136), (8192, 2048, 768), (264, 1000, 100), (2048, 70, 64)])
def test_cross_entropy_on_bf16_logits(gpu, rows, V, K):
    """weedcu_cross_entropy_fwd_bf16in / _bwd_pack_bf16in: the loss path when the LM head's epilogue wrote only the bf16 copy
    of the logits. lse from the bf16 logits, the target logit recomputed exactly from the product's bf16 operands
    (+ bias), dlogits = (exp(l - lse) - onehot) * dloss / rows with its RNE bf16 copy and column sums — against numpy."""
    import ctypes as C
    rng = np.random.default_rng(rows + V + K)
    U32, U64, I32 = C.c_uint32, C.c_uint64, C.c_int
    a_dev, lda, A = _operand(rng, rows, K, 1)
    b_dev, ldb, B = _operand(rng, V, K, 0)
    bias = rng.uniform(-2, 2, V).astype(np.float32)
    exact = A @ B.T + bias[None, :].astype(np.float64)                 # [rows, V]
    l16_bits = _bf16_bits(exact.astype(np.float32))
    l16 = _bf16_widen(l16_bits).astype(np.float64)
    tg = rng.integers(0, V, size=rows).astype(np.int32)
    pa, pb, hbias, ht = gpu.buf(a_dev), gpu.buf(b_dev), gpu.buf(bias), gpu.buf(tg)
    hl16 = gpu.buf(np.ascontiguousarray(l16_bits.T).reshape(-1))       # column-major [rows, V]: rows contiguous
    hlse, hloss = gpu.buf(np.zeros(rows, np.float32)), gpu.buf(np.zeros(1, np.float32))
    gpu.call("cross_entropy_fwd_bf16in", hl16, U32(rows), U32(V), pa, I32(1), U64(lda), pb, I32(0), U64(ldb), U32(K), hbias, ht, hlse, hloss)
    m = l16.max(1)
    ref_lse = m + np.log(np.exp(l16 - m[:, None]).sum(1))
    got_lse = hlse.get().astype(np.float64)
    assert np.max(np.abs(got_lse - ref_lse)) <= 2e-6 * np.abs(ref_lse).max()
    ref_loss = np.mean(ref_lse - exact[np.arange(rows), tg])
    assert abs(float(hloss.get()[0]) - ref_loss) <= 2e-6 * abs(ref_loss)
    for acc in (0, 1):
        d0 = rng.uniform(-1, 1, rows * V).astype(np.float32)
        hd, hg = gpu.buf(d0), gpu.buf(np.full(1, 0.7, np.float32))
        hsh, hcs = gpu.buf(np.zeros(rows * V, np.uint16)), gpu.buf(np.full(V, 3.0, np.float32))
        gpu.call("cross_entropy_bwd_pack_bf16in", hl16, U32(rows), U32(V), ht, hlse, hg, hd, U64(0), I32(acc), hsh, hcs)
        d = hd.get()
        onehot = np.zeros((rows, V))
        onehot[np.arange(rows), tg] = 1.0
        want = (np.exp(l16 - got_lse[:, None]) - onehot) * (0.7 / rows)
        if acc:
            want = want + d0.reshape(V, rows).T
        assert cases.rel_err(d.reshape(V, rows).T, want) <= 2e-5
        assert np.array_equal(hsh.get(), _bf16_bits(d))
        assert cases.rel_err(hcs.get(), d.reshape(V, rows).sum(1)) <= 2e-5


@pytest.mark.parametrize("rows,cols,acc", [(1032, 50, 0), (2048, 24, 1), (8192, 96, 0)])
def test_gelu_grad_pack_on_bf16_dout(gpu, oracle, rows, cols, acc):
    """weedcu_gelu_grad_pack_bf16dy: dout arrives as the bf16 copy the product behind it wrote. Bit-identical to
    weedcu_gelu_grad_pack on the widened values, and equal to the oracle's gelu_grad_pack on them within 2e-5."""
    import ctypes as C
    U32, I32 = C.c_uint32, C.c_int
    rng = np.random.default_rng(rows * 7 + cols)
    n = rows * cols
    d0, x = rng.uniform(-1, 1, n).astype(np.float32), rng.uniform(-4, 4, n).astype(np.float32)
    g_bits = _bf16_bits(rng.uniform(-1, 1, n).astype(np.float32))
    g_wide = _bf16_widen(g_bits).astype(np.float32)
    out = {}
    for be, name in ((gpu, "bf16dy"), (gpu, "wide"), (oracle, "oracle")):
        hd, hx = be.buf(d0.copy()), be.buf(x)
        hs, hc = be.buf(np.zeros(n, np.uint16)), be.buf(np.full(cols, 3.0, np.float32))
        if name == "bf16dy":
            be.call("gelu_grad_pack_bf16dy", hd, hx, be.buf(g_bits), U32(rows), U32(cols), I32(acc), hs, hc)
        else:
            be.call("gelu_grad_pack", hd, hx, be.buf(g_wide), U32(rows), U32(cols), I32(acc), hs, hc)
        be.sync()
        out[name] = (hd.get(), hs.get(), hc.get())
    for k in range(3):
        assert np.array_equal(out["bf16dy"][k], out["wide"][k])
    assert cases.rel_err(out["bf16dy"][0], out["oracle"][0]) <= 2e-5
    assert cases.rel_err(out["bf16dy"][2], out["oracle"][2]) <= 2e-5
    if not acc:
        # no fp32 output asked for (the production call): the tanh.approx path; the bf16 copy may differ from the exact one by
        # one bf16 ulp (2^-7 of the value), the column sums by the approximation's 2^-11
        hs, hc = gpu.buf(np.zeros(n, np.uint16)), gpu.buf(np.zeros(cols, np.float32))
        gpu.call("gelu_grad_pack_bf16dy", None, gpu.buf(x), gpu.buf(g_bits), U32(rows), U32(cols), I32(0), hs, hc)
        gpu.sync()
        exact = _bf16_widen(out["oracle"][1]).astype(np.float64)
        assert np.max(np.abs(_bf16_widen(hs.get()).astype(np.float64) - exact)) <= 2.0 ** -7 * np.max(np.abs(exact)) + 1e-30
        assert cases.rel_err(hc.get(), out["oracle"][2]) <= 2e-3
