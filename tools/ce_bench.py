"""Device time of the loss path on bf16 logits at the C5 shape (8192 x 50257, K 768): forward (partials from the bf16
logits + finish with the recomputed target logit + mean) and backward (dlogits bf16 operand copy + column sums).
Usage: python tools/ce_bench.py  (one B200)"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import microbench as mb  # noqa: E402
from weed_b200 import weedcu  # noqa: E402

U64, U32, I32 = C.c_uint64, C.c_uint32, C.c_int


def main():
    mb.lib = weedcu()
    st = torch.cuda.Stream()
    mb.STREAM = st.cuda_stream
    torch.cuda.set_stream(st)
    P = mb.P
    rows, V, K = 8192, 50257, 768
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = float(peaks.get("hbm_gbs", 6462.7))
    nrot = 2
    logits = [(torch.randn(rows * V, device="cuda") * 2.0).to(torch.bfloat16).view(torch.int16) for _ in range(nrot)]
    shadow = [torch.empty(rows * V, dtype=torch.int16, device="cuda") for _ in range(nrot)]
    a = (torch.randn(rows * K, device="cuda") * 0.5).to(torch.bfloat16).view(torch.int16)
    b = (torch.randn(V * K, device="cuda") * 0.05).to(torch.bfloat16).view(torch.int16)
    bias = torch.randn(V, device="cuda")
    tg = torch.randint(0, V, (rows,), dtype=torch.int32, device="cuda")
    lse, loss, g = torch.zeros(rows, device="cuda"), torch.zeros(1, device="cuda"), torch.ones(1, device="cuda")
    colsum = torch.zeros(V, device="cuda")

    def fwd(i):
        mb.call("cross_entropy_fwd_bf16in", P(logits[i]), U32(rows), U32(V), P(a), I32(1), U64(rows), P(b), I32(0), U64(K), U32(K), P(bias), P(tg), P(lse),
                P(loss))

    def bwd(i):
        mb.call("cross_entropy_bwd_pack_bf16in", P(logits[i]), U32(rows), U32(V), P(tg), P(lse), P(g), None, U64(0), I32(0), P(shadow[i]), P(colsum))

    out = {}
    for name, fn, bytes_ in (("ce_fwd_bf16in", fwd, 2.0 * rows * V), ("ce_bwd_pack_bf16in", bwd, 4.0 * rows * V)):
        ms = mb.timeit(fn, nrot, iters=10, warmup=3)
        out[name] = {"us": round(ms * 1e3, 1), "GB/s": round(bytes_ / ms / 1e6, 1), "frac_hbm": round(bytes_ / ms / 1e6 / hbm, 3)}
        print(name, out[name], flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
