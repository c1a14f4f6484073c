#!/usr/bin/env python
"""Greedy KV-cache decode throughput at the GPT-2-small shape (config C5's decode half):
Embedding(50257,768) + LearnedPositionalEncoding + 12 x TransformerEncoderLayer(768, 12 heads, 3072)
(float KV cache, kv_quant_bits = 0) + LayerNorm + Linear(768, 50257); batch 8, prompt 128 tokens
(one prefill call), then `--new` single-token steps with the arg-max token fed back on the device.

    python tools/decode_bench.py [--batch 8] [--prompt 128] [--new 128] [--precision fp32|bf16] [--layers 12]

Prints one JSON line: tokens/s of the generation phase (device time, CUDA events on the launching
stream), prefill ms, launches per step and per-kernel-class device time of a step.
Positions: LearnedPositionalEncoding::forward adds positions 0..T-1 of the CURRENT call
(learned_positional_encoding.cpp:49-61), so every incrementally fed token gets position 0 — reference
behaviour, reproduced.
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--prompt", type=int, default=128)
    ap.add_argument("--new", type=int, default=128)
    ap.add_argument("--layers", type=int, default=12)
    ap.add_argument("--vocab", type=int, default=50257)
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--fused", type=int, default=1)
    args = ap.parse_args()

    import torch
    import bench
    from weed_b200 import weedcu, check
    from weed_b200.harness import Harness
    assert torch.cuda.is_available(), "decode_bench needs a CUDA device (there is no CPU fallback)"
    lib = weedcu()
    check(lib.weedcu_set_device(C.c_int(0)))
    P = Harness.product()
    P.config("fused", args.fused)
    P.config("ref_index_quirks", 0)
    P.config("matmul_precision", 1 if args.precision == "bf16" else 0)
    cfg = dict(bench.FULL, B=args.batch, L=args.layers, V=args.vocab)
    V, d = cfg["V"], cfg["d"]
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
        P.module_set(e, "max_kv_seq_len", args.prompt + args.new + 8)
    P.module_set(model, "train", 0)
    stream = P.stream()
    B = args.batch
    prompt = rng.integers(0, V, size=B * args.prompt).astype(np.int32)

    def ev():
        e = C.c_void_p()
        check(lib.weedcu_event_create(C.byref(e)))
        return e

    def elapsed(e0, e1):
        ms = C.c_float()
        check(lib.weedcu_event_sync(e1))
        check(lib.weedcu_event_elapsed_ms(e0, e1, C.byref(ms)))
        return ms.value

    def run(n_new, profile=False):
        P.module_set(model, "reset_cache", 1)
        e0, e1, e2 = ev(), ev(), ev()
        P.sync()
        check(lib.weedcu_event_record(e0, C.c_void_p(stream)))
        lg = P.forward_symbol(model, P.symbol(prompt, [B, args.prompt]))
        tok = P.argmax_last(lg)
        P.free(lg)
        check(lib.weedcu_event_record(e1, C.c_void_p(stream)))
        n0, n1 = C.c_uint64(), C.c_uint64()
        lib.weedcu_launch_count(C.byref(n0))
        if profile:
            lib.weedcu_prof_enable(C.c_int(1))
        toks = [tok]
        for _ in range(n_new):
            lg = P.forward_symbol(model, tok)
            tok = P.argmax_last(lg)
            P.free(lg)
            toks.append(tok)
        check(lib.weedcu_event_record(e2, C.c_void_p(stream)))
        lib.weedcu_launch_count(C.byref(n1))
        t_prefill, t_gen = elapsed(e0, e1), elapsed(e1, e2)
        if profile:
            P.sync()
            lib.weedcu_prof_enable(C.c_int(0))
        seq = np.stack([P.read_symbol(t, B) for t in toks], axis=1)
        for t in toks:
            P.free(t)
        return t_prefill, t_gen, (n1.value - n0.value) / max(n_new, 1), seq

    run(8)  # warm-up (allocator pools, module caches)
    t_prefill, t_gen, launches, seq = run(args.new)
    t_prefill2, t_gen2, _, seq2 = run(args.new)
    # fp32: bit-reproducible. bf16: the prefill's tensor-core GEMMs may meet their split-K slices by TMA reduce-add in a
    # different order from run to run (fp32 adds of >= 3 slices), and with random-init weights the top logits are near ties
    deterministic = bool(np.array_equal(seq, seq2))
    assert deterministic or args.precision == "bf16", "greedy decode is not deterministic"
    t_gen = min(t_gen, t_gen2)
    names = {1: "gemm_bf16_tcgen05", 2: "gemm_f32_ffma", 3: "pack_bf16", 4: "elementwise", 5: "softmax", 6: "layernorm", 7: "cross_entropy",
             8: "optimizer", 9: "reduce", 10: "embedding", 11: "fill", 13: "attention"}
    run(16, profile=True)
    breakdown = {}
    for cls, nm in names.items():
        t, n, w = C.c_double(), C.c_uint64(), C.c_double()
        lib.weedcu_prof_read(C.c_int(cls), C.byref(t), C.byref(n), C.byref(w))
        if n.value:
            breakdown[nm] = {"ms_per_step": t.value / 16, "launches_per_step": n.value / 16, "bytes_or_flop_per_step": w.value / 16}
    # algorithmic HBM bytes of one generation step: every fp32 weight once + the visible KV cache once
    n_w = cfg["L"] * (4 * d * d + 2 * d * cfg["dff"]) + d * V
    kv = 2 * cfg["L"] * B * d * (args.prompt + args.new / 2)
    step_bytes = 4.0 * (n_w + kv)
    ms_step = t_gen / args.new
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    hbm = peaks.get("hbm_gbs", 6650.0)
    print(json.dumps({"metric": "greedy decode tokens/s (GPT-2-small shape, KV cache)", "value": B * args.new / (t_gen / 1000.0), "unit": "tokens/s",
                      "batch": B, "prompt": args.prompt, "new_tokens": args.new, "layers": cfg["L"], "precision": args.precision,
                      "fused": args.fused, "prefill_ms": min(t_prefill, t_prefill2), "ms_per_step": ms_step, "launches_per_step": launches,
                      "roofline": {"bound": "hbm", "achieved": step_bytes / ms_step / 1e6, "peak": hbm, "unit": "GB/s",
                                   "frac": step_bytes / ms_step / 1e6 / hbm, "algorithmic_bytes_per_step": step_bytes},
                      "kernel_breakdown": breakdown, "first_tokens": seq[0, :8].tolist(), "deterministic": deterministic}))


if __name__ == "__main__":
    main()
