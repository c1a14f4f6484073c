"""The other named configurations of BASELINE.json next to the C5 training step: C2 (tabular MLP, 64 K rows,
Adam), C3 (op sweep: matmul forward / forward+backward 256..16384 at fp32 and bf16, elementwise / softmax / LayerNorm /
reductions / optimisers at the SURVEY §8(d) sizes), C4 (scaled-up binary-addition transformer) and the decode half of C5.
bench.py calls run_all() after its timed region and puts the result under "extra": {"configs": ...} of its one JSON line.
Every entry carries its own roofline fraction (HBM-bound: algorithmic bytes of SURVEY §8(d) / time against the measured
copy bandwidth; GEMM: 2MNK / time against the measured bf16 peak, also for the fp32 FFMA path so that both precisions
share one denominator) and the SM clocks sampled while it ran.

Can also be run on its own:  python tools/config_sweep.py [--only c2,c3,c4,decode]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

U64, U32, I32, F = C.c_uint64, C.c_uint32, C.c_int, C.c_float


class Timer:
    """CUDA events on the library's own stream (the stream the kernels are launched on)."""

    def __init__(self, lib, stream, check):
        self.lib, self.stream, self.check = lib, stream, check

    def ev(self):
        e = C.c_void_p()
        self.check(self.lib.weedcu_event_create(C.byref(e)))
        return e

    def time(self, fn, iters, warmup):
        for i in range(warmup):
            fn(i)
        e0, e1 = self.ev(), self.ev()
        self.check(self.lib.weedcu_stream_sync(C.c_void_p(self.stream)))
        self.check(self.lib.weedcu_event_record(e0, C.c_void_p(self.stream)))
        for i in range(iters):
            fn(i)
        self.check(self.lib.weedcu_event_record(e1, C.c_void_p(self.stream)))
        self.check(self.lib.weedcu_event_sync(e1))
        ms = C.c_float()
        self.check(self.lib.weedcu_event_elapsed_ms(e0, e1, C.byref(ms)))
        return ms.value / iters


class LossEnds:
    """Drops each step's handles (and with them its autograd graph and activations) as the step ends, reading the loss on
    the first and the last of `total` calls only, so that the steps in between run without a host sync."""

    def __init__(self, P, total):
        self.P, self.total, self.calls, self.first, self.last = P, total, 0, None, None

    def step_done(self, loss, *others):
        self.calls += 1
        if self.calls == 1:
            self.first = float(np.sum(self.P.read(loss)))
        if self.calls == self.total:
            self.last = float(np.sum(self.P.read(loss)))
        for h in (loss,) + others:
            self.P.free(h)


# ----------------------------------------------------------------------------------- C2
def run_c2(P, timer, peaks):
    """examples/heart_attack.cpp scaled to 65536 x 13: Linear(13,26)-Tanh-Linear(26,1), bci_with_logits_loss,
    Adam lr 1e-3, 20 fixed steps. fp32 (the products are 13- and 26-deep: nothing for the tensor cores)."""
    rows = 65536
    rng = np.random.default_rng(1003)
    x = rng.uniform(-1, 1, size=(rows, 13)).astype(np.float32)
    y = (rng.uniform(size=rows) > 0.5).astype(np.float32)
    mark = P.mark()
    P.config("matmul_precision", 0)
    m = P.module("sequential", P.module("linear", 13, 26, 1), P.module("tanh"), P.module("linear", 26, 1, 1))
    P.init_params(m, 2000)
    opt = P.adam(m, 1e-3)
    xt, yt = P.tensor(np.ascontiguousarray(x.T).ravel(), [rows, 13]), P.tensor(y, [rows, 1])
    seen = LossEnds(P, 3 + 20)

    def step(_i):
        pred = P.forward(m, xt)
        loss = P.op("bci_with_logits_loss", [pred, yt])
        P.backward(loss)
        P.adam_step(opt, m)
        P.zero_grad(m)
        seen.step_done(loss, pred)

    ms = timer.time(step, 20, 3)
    first, last = seen.first, seen.last
    P.release_since(mark)
    # algorithmic bytes of one step: x read twice (forward, dW1), hidden [rows, 26] ~10 passes, output column ~12 passes
    alg_bytes = 4.0 * rows * (2 * 13 + 10 * 26 + 12)
    hbm = peaks.get("hbm_gbs", 6650.0)
    return {"workload": "C2 tabular MLP 65536x13, Linear(13,26)-Tanh-Linear(26,1), bci_with_logits_loss, Adam, 20 steps", "ms_per_step": ms,
            "rows_per_s": rows / (ms / 1000.0), "loss_first": first, "loss_last": last, "dtype": "f32",
            "roofline": {"bound": "hbm", "achieved": alg_bytes / ms / 1e6, "peak": hbm, "unit": "GB/s", "frac": alg_bytes / ms / 1e6 / hbm,
                         "note": "launch-bound: ~60 small kernels per step on 1.7 M-element tensors"}}


# ----------------------------------------------------------------------------------- C3
def run_c3(lib, stream, check, peaks, quick=False):
    import torch
    import microbench as mb
    from weed_b200._lib import contiguous_view
    mb.lib, mb.STREAM = lib, stream
    ts = torch.cuda.ExternalStream(stream)
    torch.cuda.set_stream(ts)
    P, call, bufs, mat, timeit, rot = mb.P, mb.call, mb.bufs, mb.mat, mb.timeit, mb.rot_count
    hbm, tpeak = peaks.get("hbm_gbs", 6650.0), peaks.get("bf16_tflops", 1590.0)
    out = []

    def rec_bw(op, size, ms, nbytes):
        out.append({"op": op, "size": size, "ms": round(ms, 5), "achieved": round(nbytes / ms / 1e6, 1), "unit": "GB/s", "bound": "hbm",
                    "frac": round(nbytes / ms / 1e6 / hbm, 3)})

    def rec_tc(op, size, ms, flop, dtype):
        out.append({"op": op, "size": size, "dtype": dtype, "ms": round(ms, 5), "achieved": round(flop / ms / 1e9, 1), "unit": "TFLOP/s",
                    "bound": "tensor", "frac": round(flop / ms / 1e9 / tpeak, 3)})

    # ---- matmul forward and forward + backward (C = A B; dA = dC B^T; dB = A^T dC), fp32 storage, both precisions
    sizes = (256, 512, 1024, 2048, 4096, 8192, 16384)
    for n in sizes:
        A, B, Cc, dA, dB = (bufs(n * n, 1)[0] for _ in range(5))
        for prec, name in ((0, "f32"), (1, "bf16")):
            if prec == 0 and n > 8192 and quick:
                continue
            it = 2 if (prec == 0 and n >= 8192) else (3 if n >= 8192 else 10)

            def fwd(_i):
                call("matmul_real", P(A), mat(0, 1, n), P(B), mat(0, 1, n), P(Cc), mat(0, 1, n), U32(n), U32(n), U32(n), U32(1), I32(0), I32(prec))

            def fwd_bwd(_i):
                fwd(_i)
                call("matmul_real", P(Cc), mat(0, 1, n), P(B), mat(0, n, 1), P(dA), mat(0, 1, n), U32(n), U32(n), U32(n), U32(1), I32(0), I32(prec))
                call("matmul_real", P(A), mat(0, n, 1), P(Cc), mat(0, 1, n), P(dB), mat(0, 1, n), U32(n), U32(n), U32(n), U32(1), I32(0), I32(prec))

            rec_tc("matmul_fwd", f"{n}^3", timeit(fwd, 1, iters=it, warmup=1), 2.0 * n ** 3, name)
            rec_tc("matmul_fwd_bwd", f"{n}^3", timeit(fwd_bwd, 1, iters=it, warmup=1), 6.0 * n ** 3, name)
        del A, B, Cc, dA, dB
        torch.cuda.empty_cache()

    # ---- elementwise / unary / optimisers at n = 2^20 .. 2^28
    for logn in (20, 24, 28):
        n = 1 << logn
        k = 2 if logn == 28 else rot(12 * n)
        A, B, O = bufs(n, k), bufs(n, k), bufs(n, k)
        v = contiguous_view([n])
        size = f"2^{logn}"
        for opn, code in (("add", 0), ("mul", 1), ("div", 3)):
            rec_bw(f"{opn}_same_shape", size, timeit(lambda i: call("binary_real", I32(code), P(A[i]), v, P(B[i]), v, P(O[i]), v), k), 12.0 * n)
        rec_bw("inplace_add", size, timeit(lambda i: call("inplace_real", I32(0), P(O[i]), v, P(A[i]), v), k), 12.0 * n)
        for opn, code in (("relu", 0), ("sigmoid", 1), ("gelu", 7)):
            rec_bw(f"{opn}_fwd", size, timeit(lambda i: call("unary_real", I32(code), F(0), P(A[i]), v, P(O[i]), v), k), 8.0 * n)
            rec_bw(f"{opn}_grad", size, timeit(lambda i: call("unary_grad_real", I32(code), P(O[i]), v, P(A[i]), v, P(B[i]), v, I32(1)), k), 16.0 * n)
        rec_bw("sum_full", size, timeit(lambda i: call("sum_real", P(A[i]), v, F(1.0), P(O[0])), k), 4.0 * n)
        rec_bw("mean_full", size, timeit(lambda i: call("sum_real", P(A[i]), v, F(1.0 / n), P(O[0])), k), 4.0 * n)
        rec_bw("sgd_step", size, timeit(lambda i: call("sgd_step", P(A[i]), P(B[i]), U64(n), F(1e-3), F(1.0)), k), 12.0 * n)
        if logn < 28:
            Mm, V = bufs(n, k, fill=0.0), bufs(n, k, fill=0.0)
            rec_bw("adam_step", size, timeit(lambda i: call("adam_step", P(A[i]), P(B[i]), P(Mm[i]), P(V[i]), U64(n), F(1e-3), F(0.9), F(0.999), F(1e-8), F(0.1),
                                                            F(0.001), F(1.0)), k), 28.0 * n)
            del Mm, V
        del A, B, O
        torch.cuda.empty_cache()
    # bias / scalar broadcast at the C5 activation shape
    M, N = 8192, 3072
    k = rot(8 * M * N)
    A, O = bufs(M * N, k), bufs(M * N, k)
    bias, sc = bufs(N, 1)[0], bufs(1, 1)[0]
    from weed_b200._lib import make_view
    va, vb, vs = contiguous_view([M, N]), make_view([M, N], [0, 1]), make_view([M, N], [0, 0])
    rec_bw("add_bias_broadcast", "8192x3072", timeit(lambda i: call("binary_real", I32(0), P(A[i]), va, P(bias), vb, P(O[i]), va), k), 8.0 * M * N)
    rec_bw("mul_scalar", "8192x3072", timeit(lambda i: call("binary_real", I32(1), P(A[i]), va, P(sc), vs, P(O[i]), va), k), 8.0 * M * N)
    del A, O
    # ---- softmax / logsoftmax forward + backward [8192, L], axis -1
    for L in (128, 1024, 4096, 16384):
        n = 8192 * L
        k = rot(8 * n)
        X, Y, G = bufs(n, k), bufs(n, k), bufs(n, k)
        v2 = contiguous_view([8192, L])
        for lm, nm in ((0, "softmax"), (1, "logsoftmax")):
            rec_bw(f"{nm}_fwd", f"8192x{L}", timeit(lambda i: call("softmax_real", I32(lm), P(X[i]), v2, I32(1), P(Y[i]), v2), k), 8.0 * n)
            rec_bw(f"{nm}_bwd", f"8192x{L}", timeit(lambda i: call("softmax_grad_real", I32(lm), P(G[i]), v2, P(Y[i]), v2, P(X[i]), v2, I32(1)), k), 16.0 * n)
        del X, Y, G
        torch.cuda.empty_cache()
    # ---- LayerNorm forward + backward [8192, F]
    for Fd in (768, 1024, 4096):
        rows = 8192
        n = rows * Fd
        k = rot(8 * n)
        X, Y, DY, DX = bufs(n, k), bufs(n, k), bufs(n, k), bufs(n, k)
        g, b = bufs(Fd, 1, fill=1.0)[0], bufs(Fd, 1, fill=0.0)[0]
        mu, rs, dg, db = bufs(rows, 1)[0], bufs(rows, 1)[0], bufs(Fd, 1, fill=0.0)[0], bufs(Fd, 1, fill=0.0)[0]
        rec_bw("layernorm_fwd", f"8192x{Fd}", timeit(lambda i: call("layernorm_fwd", P(X[i]), U32(rows), U32(Fd), P(g), P(b), F(3e-8), P(Y[i]), P(mu), P(rs)), k), 8.0 * n)
        rec_bw("layernorm_bwd", f"8192x{Fd}", timeit(lambda i: call("layernorm_bwd", P(X[i]), P(DY[i]), U32(rows), U32(Fd), P(g), P(mu), P(rs), P(DX[i]), P(dg), P(db),
                                                                   I32(0), I32(1)), k), 16.0 * n)
        del X, Y, DY, DX
        torch.cuda.empty_cache()
    torch.cuda.synchronize()
    below = [e for e in out if e["bound"] == "hbm" and e["frac"] < 0.70 and (e["size"].startswith("2^2") and int(e["size"][2:]) >= 24 or "x" in e["size"])]
    return {"workload": "C3 op sweep (SURVEY 8d sizes)", "entries": out, "hbm_peak_gbs": hbm, "bf16_peak_tflops": tpeak,
            "hbm_bound_entries_below_0.70_at_streaming_sizes": [f"{e['op']}@{e['size']}={e['frac']}" for e in below]}


# ----------------------------------------------------------------------------------- C4
def run_c4(P, timer, peaks):
    """examples/binary_addition_transformer.cpp scaled up: Embedding(512,512) - LearnedPositionalEncoding -
    TransformerEncoderLayer(512, 8 heads, d_ff 2048) - Linear(512,1), T 128, batch 256, bci_with_logits_loss on the last
    T/2 positions, Adam, 10 fixed steps, bf16 tensor-core GEMMs."""
    B, T, V, d, Hh, dff = 256, 128, 512, 512, 8, 2048
    rng = np.random.default_rng(40)
    tokens = rng.integers(0, V, size=(B, T)).astype(np.int32)
    tlen = T // 2
    target = (rng.uniform(size=(B, tlen)) > 0.5).astype(np.float32)
    mark = P.mark()
    P.config("matmul_precision", 1)
    model = P.module("sequential", P.module("embedding", V, d), P.module("posenc", T, d), P.module("encoder", d, Hh, dff), P.module("linear", d, 1, 1))
    rs = np.random.default_rng(41)
    for i in range(P.param_count(model)):
        n = P.param_size(model, i)
        if n in (d, 1, dff):
            continue
        P.param_set(model, i, rs.uniform(-0.05, 0.05, size=n).astype(np.float32))
    opt = P.adam(model, 1e-4)
    tok = P.symbol(np.ascontiguousarray(tokens.T).ravel(), [B, T])
    tgt = P.tensor(np.ascontiguousarray(target.T).ravel(), [B, tlen])
    seen = LossEnds(P, 3 + 10)

    def step(_i):
        logits = P.forward_symbol(model, tok)
        P.squeeze(logits, 2)
        pred = P.op("slice", [logits], ints=[1, T - tlen, tlen])
        loss = P.op("bci_with_logits_loss", [pred, tgt])
        P.backward(loss)
        P.adam_step(opt, model)
        P.zero_grad(model)
        P.module_set(model, "reset_cache", 1)
        seen.step_done(loss, pred, logits)

    ms = timer.time(step, 10, 3)
    first, last = seen.first / (B * tlen), seen.last / (B * tlen)
    P.release_since(mark)
    toks = B * T
    flop = 2.0 * toks * d * d * 4 + 2.0 * toks * d * dff * 2 + 4.0 * B * Hh * T * T * (d // Hh)   # forward
    flop += 2.0 * (2.0 * toks * d * d + 2.0 * 2.0 * toks * d * dff)                                 # dA + dB of W_o, ff1, ff2
    tpeak = peaks.get("bf16_tflops", 1590.0)
    return {"workload": "C4 scaled binary-addition transformer: vocab 512, d 512, 8 heads, d_ff 2048, T 128, batch 256, Adam, 10 steps",
            "ms_per_step": ms, "samples_per_s": B / (ms / 1000.0), "loss_per_position_first": first, "loss_per_position_last": last, "dtype": "bf16",
            "roofline": {"bound": "tensor", "achieved": flop / ms / 1e9, "peak": tpeak, "unit": "TFLOP/s", "frac": flop / ms / 1e9 / tpeak,
                         "note": "whole step (GEMM FLOP / step time): includes every bandwidth-bound kernel of the step"}}


# ----------------------------------------------------------------------------------- C5 decode
def run_decode(P, lib, check, peaks, new=64, prompt=128, batch=8):
    """Greedy KV-cache decode at the GPT-2-small shape: prompt 128 (one prefill call), `new` single-token steps with the
    arg-max fed back on the device, fp32 weights (tools/decode_bench.py has the long form)."""
    import bench
    cfg = dict(bench.FULL, B=batch)
    V, d = cfg["V"], cfg["d"]
    mark = P.mark()
    P.config("matmul_precision", 0)
    encs = [P.module("encoder", d, cfg["H"], cfg["dff"]) for _ in range(cfg["L"])]
    mods = [P.module("embedding", V, d), P.module("posenc", cfg["T"], d)] + encs + [P.module("layernorm", d), P.module("linear", d, V, 1)]
    model = P.module("sequential", *mods)
    rng = np.random.default_rng(2000)
    for i in range(P.param_count(model)):
        n = P.param_size(model, i)
        if n == d:
            continue
        lim = 0.02 if n >= V * d else float(np.sqrt(6.0 / (d + n // d)))
        P.param_set(model, i, rng.uniform(-lim, lim, size=n).astype(np.float32))
    for e in encs:
        P.module_set(e, "kv_quant_bits", 0)
        P.module_set(e, "use_kv_cache", 1)
        P.module_set(e, "max_kv_seq_len", prompt + new + 8)
    P.module_set(model, "train", 0)
    stream = P.stream()
    pr = rng.integers(0, V, size=batch * prompt).astype(np.int32)

    def ev():
        e = C.c_void_p()
        check(lib.weedcu_event_create(C.byref(e)))
        return e

    def run(n_new):
        P.module_set(model, "reset_cache", 1)
        e0, e1, e2 = ev(), ev(), ev()
        P.sync()
        check(lib.weedcu_event_record(e0, C.c_void_p(stream)))
        lg = P.forward_symbol(model, P.symbol(pr, [batch, prompt]))
        tok = P.argmax_last(lg)
        P.free(lg)
        check(lib.weedcu_event_record(e1, C.c_void_p(stream)))
        n0, n1 = C.c_uint64(), C.c_uint64()
        lib.weedcu_launch_count(C.byref(n0))
        toks = [tok]
        for _ in range(n_new):
            lg = P.forward_symbol(model, tok)
            tok = P.argmax_last(lg)
            P.free(lg)
            toks.append(tok)
        check(lib.weedcu_event_record(e2, C.c_void_p(stream)))
        lib.weedcu_launch_count(C.byref(n1))
        check(lib.weedcu_event_sync(e2))
        a, b = C.c_float(), C.c_float()
        check(lib.weedcu_event_elapsed_ms(e0, e1, C.byref(a)))
        check(lib.weedcu_event_elapsed_ms(e1, e2, C.byref(b)))
        for t in toks:
            P.free(t)
        return a.value, b.value, (n1.value - n0.value) / max(n_new, 1)

    run(4)
    pre, gen, launches = run(new)
    pre2, gen2, _ = run(new)
    gen, pre = min(gen, gen2), min(pre, pre2)
    P.release_since(mark)
    n_w = cfg["L"] * (4 * d * d + 2 * d * cfg["dff"]) + d * V
    kv = 2 * cfg["L"] * batch * d * (prompt + new / 2)
    step_bytes = 4.0 * (n_w + kv)
    ms_step = gen / new
    hbm = peaks.get("hbm_gbs", 6650.0)
    return {"workload": f"C5 decode: batch {batch}, prompt {prompt}, {new} greedy tokens, float KV cache, fp32 weights", "tokens_per_s": batch * new / (gen / 1000.0),
            "ms_per_step": ms_step, "prefill_ms": pre, "launches_per_step": launches, "dtype": "f32",
            "roofline": {"bound": "hbm", "achieved": step_bytes / ms_step / 1e6, "peak": hbm, "unit": "GB/s", "frac": step_bytes / ms_step / 1e6 / hbm,
                         "algorithmic_bytes_per_step": step_bytes}}


def run_all(P, lib, check, peaks, sampler_factory=None, only=None, quick=False):
    """Every configuration, each with the SM clocks sampled while it ran (sampler_factory() -> object with start() / stop())."""
    timer = Timer(lib, P.stream(), check)
    jobs = {"c2": lambda: run_c2(P, timer, peaks), "c3": lambda: run_c3(lib, P.stream(), check, peaks, quick), "c4": lambda: run_c4(P, timer, peaks),
            "decode": lambda: run_decode(P, lib, check, peaks)}
    out = {}
    for name, job in jobs.items():
        if only and name not in only:
            continue
        s = sampler_factory() if sampler_factory else None
        if s:
            s.start()
        t0 = time.perf_counter()
        try:
            out[name] = job()
        except Exception as e:  # one failing configuration must not take the headline line down with it
            out[name] = {"error": f"{type(e).__name__}: {e}"}
        out[name]["wall_s"] = round(time.perf_counter() - t0, 2)
        if s:
            out[name]["clocks"] = s.stop()
        P.config("matmul_precision", 1)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    import torch
    import bench
    from weed_b200 import weedcu, check
    from weed_b200.harness import Harness
    assert torch.cuda.is_available()
    lib = weedcu()
    check(lib.weedcu_set_device(C.c_int(0)))
    P = Harness.product()
    P.config("fused", 1)
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    only = [s for s in args.only.split(",") if s] or None
    res = run_all(P, lib, check, peaks, lambda: bench.ClockSampler(0), only, args.quick)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
