"""Kernel microbenchmarks (config C3 of BASELINE.json: op sweep on one B200).

Times each weedcu_* kernel with CUDA events on the launching stream after warm-up, with working
sets rotated through enough distinct buffers to exceed the 126 MB L2, and reports achieved
algorithmic GB/s (SURVEY §8d per-element byte counts) or TFLOP/s next to MEASURED_PEAKS.json.
Usage: python tools/microbench.py [--group ew|gemm|all] [--out gpurun_out/microbench.json]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from weed_b200 import weedcu, check, Mat  # noqa: E402
from weed_b200._lib import contiguous_view, make_view  # noqa: E402

U64, U32, I32, F = C.c_uint64, C.c_uint32, C.c_int, C.c_float
lib = None
STREAM = None


def P(t):
    return C.c_void_p(t.data_ptr())


def call(name, *args):
    fn = getattr(lib, "weedcu_" + name)
    fn.restype = C.c_int
    conv = [C.byref(a) if isinstance(a, C.Structure) else a for a in args]
    check(fn(*conv, C.c_void_p(STREAM)), name)


_BLOCK = None


def timeit(fn, nrot, iters=20, warmup=5):
    """Device time per launch. A long blocker kernel is queued first so that the host (ctypes call +
    tensor-map encodes, ~10-20 us per launch from Python) runs ahead of the GPU: without it a kernel
    shorter than the host's issue time measures the host, not the kernel."""
    global _BLOCK
    for i in range(warmup):
        fn(i % nrot)
    torch.cuda.synchronize()
    if _BLOCK is None:
        _BLOCK = torch.randn(6144, 6144, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.mm(_BLOCK, _BLOCK)  # ~10 ms of fp32 work on the current (timing) stream
    e0.record()
    for i in range(iters):
        fn(i % nrot)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters  # ms


def rot_count(bytes_per_set):
    return max(2, int(np.ceil(400e6 / bytes_per_set)))


def bufs(n, count, dtype=torch.float32, fill=None):
    out = []
    for _ in range(count):
        t = torch.empty(n, dtype=dtype, device="cuda")
        if fill is None:
            t.uniform_(-1, 1) if dtype == torch.float32 else t.zero_()
        else:
            t.fill_(fill)
        out.append(t)
    return out


def mat(offset, s0, s1, bs=0):
    m = Mat()
    m.offset, m.s0, m.s1, m.batch_stride = offset, s0, s1, bs
    return m


def bench_ew(results, peaks):
    hbm = peaks.get("hbm_gbs", 6650.0)

    def rec(name, ms, bytes_alg, note=""):
        gbs = bytes_alg / ms / 1e6
        results.append({"kernel": name, "ms": round(ms, 4), "alg_GBps": round(gbs, 1),
                        "frac_of_hbm": round(gbs / hbm, 3), "note": note})
        print(f"{name:42s} {ms:9.4f} ms  {gbs:8.1f} GB/s  {gbs / hbm:5.2f} of HBM  {note}", flush=True)

    # --- binary add, same shape (12 B/elem)
    for logn in (20, 24, 26):
        n = 1 << logn
        k = rot_count(12 * n)
        A, B, O = bufs(n, k), bufs(n, k), bufs(n, k)
        v = contiguous_view([n])
        ms = timeit(lambda i: call("binary_real", I32(0), P(A[i]), v, P(B[i]), v, P(O[i]), v), k)
        rec(f"add_same_shape_2^{logn}", ms, 12 * n)
        del A, B, O
    # --- bias add [M,N] + [0,1]-strided (8 B/elem)
    M, N = 8192, 3072
    k = rot_count(8 * M * N)
    A, O = bufs(M * N, k), bufs(M * N, k)
    bias = bufs(N, 1)[0]
    va, vb = contiguous_view([M, N]), make_view([M, N], [0, 1])
    ms = timeit(lambda i: call("binary_real", I32(0), P(A[i]), va, P(bias), vb, P(O[i]), va), k)
    rec("bias_add_8192x3072", ms, 8 * M * N)
    # --- scalar mul (8 B/elem)
    sc = bufs(1, 1)[0]
    vs = make_view([M, N], [0, 0])
    ms = timeit(lambda i: call("binary_real", I32(1), P(A[i]), va, P(sc), vs, P(O[i]), va), k)
    rec("scalar_mul_8192x3072", ms, 8 * M * N)
    # --- in-place add (12 B/elem)
    ms = timeit(lambda i: call("inplace_real", I32(0), P(O[i]), va, P(A[i]), va), k)
    rec("inplace_add_8192x3072", ms, 12 * M * N)
    # --- gelu fwd (8 B/elem) and grad (16 B/elem)
    ms = timeit(lambda i: call("unary_real", I32(7), F(0), P(A[i]), va, P(O[i]), va), k)
    rec("gelu_fwd_8192x3072", ms, 8 * M * N)
    ms = timeit(lambda i: call("unary_real", I32(0), F(0), P(A[i]), va, P(O[i]), va), k)
    rec("relu_fwd_8192x3072", ms, 8 * M * N)
    G = bufs(M * N, k)
    ms = timeit(lambda i: call("unary_grad_real", I32(7), P(O[i]), va, P(A[i]), va, P(G[i]), va, I32(1)), k)
    rec("gelu_grad_8192x3072", ms, 16 * M * N)
    # --- transposed copy: contiguous(transpose(x,1,2)) for [B,T,H,hd] -> [B,H,T,hd]
    Bq, T, H, hd = 8, 1024, 12, 64
    vsrc = make_view([Bq, H, T, hd], [1, Bq * T, Bq, Bq * T * H])
    vdst = contiguous_view([Bq, H, T, hd])
    n = Bq * T * H * hd
    k2 = rot_count(8 * n)
    S, D = bufs(n, k2), bufs(n, k2)
    ms = timeit(lambda i: call("copy_real", P(D[i]), vdst, P(S[i]), vsrc), k2)
    rec("head_transpose_copy_8x12x1024x64", ms, 8 * n)
    del A, O, G, S, D
    # --- fill (4 B/elem)
    n = 1 << 26
    Z = bufs(n, 2)
    ms = timeit(lambda i: call("fill_real", P(Z[i]), U64(n), F(0.0)), 2)
    rec("fill_zero_2^26", ms, 4 * n)
    del Z
    # --- Adam (28 B/param), SGD (12 B/param)
    n = 1 << 25
    k = rot_count(28 * n)
    Pp, G, Mm, V = bufs(n, k), bufs(n, k), bufs(n, k, fill=0.0), bufs(n, k, fill=0.0)
    ms = timeit(lambda i: call("adam_step", P(Pp[i]), P(G[i]), P(Mm[i]), P(V[i]), U64(n), F(1e-3), F(0.9),
                               F(0.999), F(1e-8), F(0.1), F(0.001), F(1.0)), k)
    rec("adam_step_2^25", ms, 28 * n)
    ms = timeit(lambda i: call("sgd_step", P(Pp[i]), P(G[i]), U64(n), F(1e-3), F(1.0)), k)
    rec("sgd_step_2^25", ms, 12 * n)
    del Pp, G, Mm, V
    # --- softmax [8,12,1024,1024] last axis (8 B/elem fwd, 16 B/elem bwd)
    shape = [8, 12, 1024, 1024]
    n = int(np.prod(shape))
    v4 = contiguous_view(shape)
    X, Y = bufs(n, 2), bufs(n, 2)
    ms = timeit(lambda i: call("softmax_real", I32(0), P(X[i]), v4, I32(3), P(Y[i]), v4), 2, iters=10)
    rec("softmax_fwd_8x12x1024x1024", ms, 8 * n)
    ms = timeit(lambda i: call("attn_softmax_real", P(X[i]), P(Y[i]), U32(96), U32(1024), U32(1024), F(8.0),
                               F(-1.7e38), I32(1), I32(1)), 2, iters=10)
    rec("attn_softmax_fused_96x1024x1024", ms, 8 * n)
    DI = bufs(n, 2)
    ms = timeit(lambda i: call("softmax_grad_real", I32(0), P(DI[i]), v4, P(Y[i]), v4, P(X[i]), v4, I32(3)), 2, iters=10)
    rec("softmax_bwd_8x12x1024x1024", ms, 16 * n)
    del X, Y, DI
    # --- softmax sweep [8192, L]
    for L in (128, 1024, 4096, 16384):
        n = 8192 * L
        k = rot_count(8 * n)
        X, Y = bufs(n, k), bufs(n, k)
        v2 = contiguous_view([8192, L])
        ms = timeit(lambda i: call("softmax_real", I32(0), P(X[i]), v2, I32(1), P(Y[i]), v2), k)
        rec(f"softmax_fwd_8192x{L}", ms, 8 * n)
        ms = timeit(lambda i: call("softmax_real", I32(1), P(X[i]), v2, I32(1), P(Y[i]), v2), k)
        rec(f"logsoftmax_fwd_8192x{L}", ms, 8 * n)
        del X, Y
    # --- LayerNorm [8192, F]
    for Fd in (768, 1024, 4096):
        rows = 8192
        n = rows * Fd
        k = rot_count(8 * n)
        X, Y, DY, DX = bufs(n, k), bufs(n, k), bufs(n, k), bufs(n, k)
        g, b = bufs(Fd, 1, fill=1.0)[0], bufs(Fd, 1, fill=0.0)[0]
        mu, rs = bufs(rows, 1)[0], bufs(rows, 1)[0]
        dg, db = bufs(Fd, 1, fill=0.0)[0], bufs(Fd, 1, fill=0.0)[0]
        ms = timeit(lambda i: call("layernorm_fwd", P(X[i]), U32(rows), U32(Fd), P(g), P(b), F(3e-8), P(Y[i]),
                                   P(mu), P(rs)), k)
        rec(f"layernorm_fwd_8192x{Fd}", ms, 8 * n)
        ms = timeit(lambda i: call("layernorm_bwd", P(X[i]), P(DY[i]), U32(rows), U32(Fd), P(g), P(mu), P(rs),
                                   P(DX[i]), P(dg), P(db), I32(0), I32(1)), k)
        rec(f"layernorm_bwd_8192x{Fd}", ms, 16 * n)
        ms = timeit(lambda i: call("layernorm_bwd", P(X[i]), P(DY[i]), U32(rows), U32(Fd), P(g), P(mu), P(rs),
                                   P(DX[i]), P(dg), P(db), I32(0), I32(0)), k)
        rec(f"layernorm_bwd_store_8192x{Fd}", ms, 12 * n)
        del X, Y, DY, DX
    # --- axis reductions
    rows, Fd = 8192, 768
    n = rows * Fd
    k = rot_count(4 * n)
    X = bufs(n, k)
    out = bufs(rows, 1)[0]
    v2 = contiguous_view([rows, Fd])
    ms = timeit(lambda i: call("reduce_real", P(X[i]), v2, I32(1), P(out), I32(0)), k)
    rec("reduce_axis1_8192x768 (LayerNorm mean)", ms, 4 * n)
    ms = timeit(lambda i: call("reduce_real", P(X[i]), v2, I32(0), P(out), I32(0)), k)
    rec("reduce_axis0_8192x768 (bias grad)", ms, 4 * n)
    del X
    n = 1 << 26
    X = bufs(n, 2)
    ms = timeit(lambda i: call("sum_real", P(X[i]), contiguous_view([n]), F(1.0), P(out)), 2)
    rec("sum_full_2^26", ms, 4 * n)
    del X
    # --- cross entropy [8192, 50257]
    rows, V = 8192, 50257
    n = rows * V
    L1 = bufs(n, 1)[0]
    DL = bufs(n, 1, fill=0.0)[0]
    tg = torch.randint(0, V, (rows,), dtype=torch.int32, device="cuda")
    lse, loss, one = bufs(rows, 1)[0], bufs(1, 1)[0], bufs(1, 1, fill=1.0)[0]
    ms = timeit(lambda i: call("cross_entropy_fwd", P(L1), U64(0), U32(rows), U32(V), U32(1), U32(rows), P(tg),
                               P(lse), P(loss)), 1, iters=5, warmup=2)
    rec("cross_entropy_fwd_8192x50257", ms, 4 * n)
    ms = timeit(lambda i: call("cross_entropy_bwd", P(L1), U64(0), U32(rows), U32(V), U32(1), U32(rows), P(tg),
                               P(lse), P(one), P(DL), U64(0), I32(1)), 1, iters=5, warmup=2)
    rec("cross_entropy_bwd_8192x50257", ms, 12 * n)
    ms = timeit(lambda i: call("cross_entropy_bwd", P(L1), U64(0), U32(rows), U32(V), U32(1), U32(rows), P(tg),
                               P(lse), P(one), P(DL), U64(0), I32(0)), 1, iters=5, warmup=2)
    rec("cross_entropy_bwd_store_8192x50257", ms, 8 * n)
    del L1, DL
    # --- embedding
    V, D, ntok = 50257, 768, 8192
    W = bufs(V * D, 1)[0]
    idx = torch.randint(0, V, (ntok,), dtype=torch.int32, device="cuda")
    O = bufs(ntok * D, 1)[0]
    ms = timeit(lambda i: call("embedding_gather", P(idx), U64(0), U32(1), U32(ntok), P(W), U64(0), U32(1), U32(V),
                               U32(D), P(O), U64(0), U32(1), U32(ntok)), 1)
    rec("embedding_gather_8192x768", ms, 8 * ntok * D)
    ms = timeit(lambda i: call("embedding_scatter_add", P(W), U64(0), U32(1), U32(V), P(idx), U64(0), U32(1),
                               U32(ntok), U32(D), P(O), U64(0), U32(1), U32(ntok)), 1)
    rec("embedding_scatter_8192x768", ms, 12 * ntok * D)


def bench_attn(results, peaks):
    """Fused attention core at the C5 shape (B=8, H=12, T=1024, hd=64): heads_pack + flash kernel + unheads.
    FLOP = the causal half of 4*B*H*T*T*hd (the key tiles the kernel actually visits are slightly more)."""
    peak = peaks.get("bf16_tflops", 1590.0)
    hbm = peaks.get("hbm_gbs", 6650.0)
    for (B, T, H, hd, causal) in ((8, 1024, 12, 64, 1), (8, 1024, 12, 64, 0), (2, 4096, 12, 64, 1)):
        n = B * T * H * hd
        q, k, v, o = bufs(n, 2), bufs(n, 2), bufs(n, 2), bufs(n, 2)
        ms = timeit(lambda i: call("attention_fwd", P(q[i]), P(k[i]), P(v[i]), P(o[i]), U32(B), U32(T), U32(H), U32(hd),
                                   F(8.0), F(-1.701411835e38), I32(causal)), 2, iters=10, warmup=3)
        flops = 4.0 * B * H * T * T * hd * (0.5 if causal else 1.0)
        tf = flops / ms / 1e9
        results.append({"kernel": f"attention_fwd_B{B}_T{T}_H{H}_hd{hd}_causal{causal}", "ms": round(ms, 4), "TFLOPs": round(tf, 1),
                        "frac_of_bf16_peak": round(tf / peak, 3), "note": "heads_pack + flash (tcgen05, S in TMEM) + unheads"})
        print(f"attention_fwd_B{B}_T{T}_H{H}_hd{hd}_causal{causal:<12d} {ms:9.4f} ms  {tf:8.1f} TFLOP/s  {tf / peak:5.3f} of bf16 peak", flush=True)
        del q, k, v, o
    # broadcast positional-encoding gradient: [B, T*d] summed over the contiguous batch axis (4 B/elem)
    Bq, n_out = 8, 1024 * 768
    X = bufs(Bq * n_out, 8)
    out = bufs(n_out, 1)[0]
    v2 = contiguous_view([Bq, n_out])
    ms = timeit(lambda i: call("reduce_real", P(X[i]), v2, I32(0), P(out), I32(0)), 8)
    gbs = 4.0 * Bq * n_out / ms / 1e6
    results.append({"kernel": "reduce_axis0_8x786432 (pos-enc grad)", "ms": round(ms, 4), "alg_GBps": round(gbs, 1), "frac_of_hbm": round(gbs / hbm, 3)})
    print(f"{'reduce_axis0_8x786432 (pos-enc grad)':42s} {ms:9.4f} ms  {gbs:8.1f} GB/s  {gbs / hbm:5.2f} of HBM", flush=True)


def bench_gemm(results, peaks, which):
    peak = peaks.get("bf16_tflops", 1590.0)

    def rec(name, ms, flops, note=""):
        tf = flops / ms / 1e9
        results.append({"kernel": name, "ms": round(ms, 4), "TFLOPs": round(tf, 1),
                        "frac_of_bf16_peak": round(tf / peak, 3), "note": note})
        print(f"{name:50s} {ms:9.4f} ms  {tf:8.1f} TFLOP/s  {tf / peak:5.3f} of bf16 peak  {note}", flush=True)

    if which in ("gemm", "all", "f32"):
        for n in (1024, 2048, 4096):
            A, B, Cc = bufs(n * n, 2), bufs(n * n, 2), bufs(n * n, 2)
            ms = timeit(lambda i: call("matmul_real", P(A[i]), mat(0, 1, n), P(B[i]), mat(0, 1, n), P(Cc[i]),
                                       mat(0, 1, n), U32(n), U32(n), U32(n), U32(1), I32(0), I32(0)), 2, iters=5, warmup=2)
            rec(f"matmul_fp32_ffma_{n}^3", ms, 2.0 * n ** 3, "fp32 FFMA parity path")
            del A, B, Cc
    if which in ("gemm", "all", "tc"):
        shapes = [(4096, 4096, 4096), (8192, 8192, 8192), (8192, 768, 768), (8192, 3072, 768), (8192, 768, 3072),
                  (8192, 50257, 768), (768, 3072, 8192)]
        for (M, N, K) in shapes:
            for (am, bm) in ((1, 0), (1, 1), (0, 0)):
                lda = (M if am else K)
                ldb = (N if bm else K)
                if lda % 8 or ldb % 8:
                    continue
                a = torch.randn(M * K, device="cuda").to(torch.bfloat16)
                b = torch.randn(N * K, device="cuda").to(torch.bfloat16)
                c = torch.zeros(M * N, device="cuda")
                ms = timeit(lambda i: call("gemm_bf16", P(a), I32(am), U64(lda), P(b), I32(bm), U64(ldb), P(c), U64(M),
                                           U32(M), U32(N), U32(K), I32(0), C.c_void_p(0)), 1, iters=10, warmup=3)
                rec(f"gemm_bf16_tcgen05_M{M}_N{N}_K{K}_a{'MN' if am else 'K'}_b{'MN' if bm else 'K'}", ms,
                    2.0 * M * N * K)
                del a, b, c
        # end-to-end fp32-storage matmul with on-the-fly bf16 packing
        for (M, N, K) in ((8192, 3072, 768), (8192, 768, 3072)):
            A, B, Cc = bufs(M * K, 1), bufs(K * N, 1), bufs(M * N, 1)
            ms = timeit(lambda i: call("matmul_real", P(A[0]), mat(0, 1, M), P(B[0]), mat(0, 1, K), P(Cc[0]), mat(0, 1, M),
                                       U32(M), U32(K), U32(N), U32(1), I32(0), I32(1)), 1, iters=10, warmup=3)
            rec(f"matmul_real_bf16_incl_pack_M{M}_N{N}_K{K}", ms, 2.0 * M * N * K, "pack fp32->bf16 + tcgen05")


def bench_tune(results, peaks):
    """Tile-configuration sweep over the GEMM shapes of the C5 step: every (family, BLOCK_N, splits)
    forced through weedcu_gemm_set_mode, next to what the cost model picks (mode 0). The table is
    what gemm_tc.cu's tile_cost_us() is calibrated against."""
    peak = peaks.get("bf16_tflops", 1590.0)
    r8 = lambda x: (x + 7) // 8 * 8
    # (label, M, N, K, a_major, b_major, accumulate, groups)
    shapes = [("qkv_fwd_grouped", 8192, 768, 768, 1, 0, 0, 3), ("wo_fwd", 8192, 768, 768, 1, 0, 0, 1), ("ff1_fwd", 8192, 3072, 768, 1, 0, 0, 1),
              ("ff2_fwd", 8192, 768, 3072, 1, 0, 0, 1), ("head_fwd", 8192, 50257, 768, 1, 0, 0, 1),
              ("wo_dA", 8192, 768, 768, 1, 1, 0, 1), ("ff1_dA", 8192, 768, 3072, 1, 1, 0, 1), ("ff2_dA", 8192, 3072, 768, 1, 1, 0, 1),
              ("head_dA", 8192, 768, 50257, 1, 1, 0, 1),
              ("wo_dB", 768, 768, 8192, 0, 0, 1, 1), ("ff1_dB", 768, 3072, 8192, 0, 0, 1, 1), ("ff2_dB", 3072, 768, 8192, 0, 0, 1, 1),
              ("head_dB", 768, 50257, 8192, 0, 0, 1, 1), ("square_4096", 4096, 4096, 4096, 1, 0, 0, 1), ("square_8192", 8192, 8192, 8192, 1, 0, 0, 1)]
    for (label, M, N, K, am, bm, acc, groups) in shapes:
        lda, ldb = r8(M if am else K), r8(N if bm else K)
        a = torch.randn(lda * (K if am else M), device="cuda").to(torch.bfloat16)
        bs = [torch.randn(ldb * (K if bm else N), device="cuda").to(torch.bfloat16) for _ in range(groups)]
        cs = [torch.zeros(M * N, device="cuda") for _ in range(groups)]
        PtrArr = C.c_void_p * groups
        bp, cp = PtrArr(*[t.data_ptr() for t in bs]), PtrArr(*[t.data_ptr() for t in cs])

        def run(i):
            call("gemm_bf16_grouped", P(a), I32(am), U64(lda), U32(groups), bp, I32(bm), U64(ldb), cp, U64(M), U32(M), U32(N), U32(K), I32(acc),
                 C.c_void_p(0))

        flops = 2.0 * M * N * K * groups
        tiles128 = ((M + 127) // 128) * ((N + 127) // 128) * groups
        split_list = (1,) if tiles128 > 4 * 148 else (1, 2, 3, 4, 6, 8)
        modes = [0, 1, 2]
        for pair in (0, 1):
            for bn in (256, 192, 128):
                if pair and bn == 192 and bm:
                    continue
                if bn > 128 and N <= bn - 64:
                    continue
                for sp in split_list:
                    if sp > 1 and sp * 4 > (K + 63) // 64:
                        continue
                    modes.append(pair * 1000000 + bn * 1000 + sp)
                    if not pair and sp == 1:
                        modes.append(bn * 1000 + 100 + sp)      # variant 1: direct stores, no staging
        best = None
        for mode in modes:
            lib.weedcu_gemm_set_mode(C.c_int(mode))
            try:
                ms = timeit(run, 1, iters=8, warmup=2)
            except Exception as e:  # an unsupported forced configuration says so
                print(f"tune_{label:18s} mode {mode:8d}  unsupported ({e})", flush=True)
                continue
            finally:
                lib.weedcu_gemm_set_mode(C.c_int(0))
            tf = flops / ms / 1e9
            results.append({"kernel": f"tune_{label}", "mode": mode, "ms": round(ms, 4), "TFLOPs": round(tf, 1), "frac_of_bf16_peak": round(tf / peak, 3)})
            print(f"tune_{label:18s} mode {mode:8d}  {ms * 1000:9.2f} us  {tf:8.1f} TFLOP/s  {tf / peak:5.3f}", flush=True)
            if mode >= 1000 and (best is None or ms < best[1]):
                best = (mode, ms)
        if best:
            print(f"tune_{label:18s} BEST forced mode {best[0]} at {best[1] * 1000:.2f} us", flush=True)
        del a, bs, cs


def main():
    global lib, STREAM
    ap = argparse.ArgumentParser()
    ap.add_argument("--group", default="all")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "microbench.json"))
    args = ap.parse_args()
    lib = weedcu()
    ts = torch.cuda.Stream()  # non-legacy stream: events, allocations and kernels all ordered on it
    torch.cuda.set_stream(ts)
    STREAM = ts.cuda_stream
    assert STREAM != 0
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    results = []
    if args.group in ("ew", "all"):
        bench_ew(results, peaks)
    if args.group in ("attn", "all"):
        bench_attn(results, peaks)
    if args.group == "tune":
        bench_tune(results, peaks)
    if args.group in ("gemm", "all", "f32", "tc"):
        bench_gemm(results, peaks, args.group)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump({"peaks": peaks, "results": results}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
