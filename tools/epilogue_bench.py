"""Times the extended GEMM epilogue variants at the C5 shapes against the plain products (round 2).
Usage: python tools/epilogue_bench.py  (one B200)"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import microbench as mb  # noqa: E402
from weed_b200 import weedcu, check  # noqa: E402
from weed_b200._lib import GemmEpilogue  # noqa: E402

U64, U32, I32 = C.c_uint64, C.c_uint32, C.c_int


def main():
    mb.lib = weedcu()
    st = torch.cuda.Stream()
    mb.STREAM = st.cuda_stream
    torch.cuda.set_stream(st)
    P = mb.P
    r8 = lambda x: (x + 7) // 8 * 8
    mode = int(os.environ.get("WEEDCU_GEMM_MODE", "0"))
    for (M, N, K, what) in [(8192, 768, 768, "residual"), (8192, 768, 3072, "residual"), (8192, 3072, 768, "gelu"), (8192, 50257, 768, "lmhead"),
                            (8192, 768, 768, "qkv")]:
        G = 3 if what == "qkv" else 1
        nrot = max(2, int(np.ceil(300e6 / (M * N * 4 * G))))
        if what == "lmhead":
            nrot = 2
        a = [torch.zeros(r8(M) * K + 8, dtype=torch.int16, device="cuda") for _ in range(nrot)]
        b = [[torch.zeros(r8(K) * N + 8, dtype=torch.int16, device="cuda") for _ in range(G)] for _ in range(nrot)]
        for t in a:
            t.copy_((torch.randn(t.numel(), device="cuda") * 0.05).to(torch.bfloat16).view(torch.int16))
        for bb in b:
            for t in bb:
                t.copy_((torch.randn(t.numel(), device="cuda") * 0.05).to(torch.bfloat16).view(torch.int16))
        c = [[torch.empty(M * N, dtype=torch.float32, device="cuda") for _ in range(G)] for _ in range(nrot)]
        c16 = [[torch.empty(M * N, dtype=torch.int16, device="cuda") for _ in range(G)] for _ in range(nrot)]
        res = [torch.randn(M * N, device="cuda") for _ in range(nrot)] if what == "residual" else None
        bias = torch.randn(N, device="cuda")
        cap = 2 * ((N + 127) // 128)
        stats = torch.empty(cap * M * 2, device="cuda")
        tiles, cols = U32(0), U32(0)
        flops = 2.0 * M * N * K * G

        def plain(i):
            if what == "residual":
                mb.call("gemm_bf16_residual", P(a[i]), I32(1), U64(r8(M)), P(b[i][0]), I32(0), U64(r8(K)), P(c[i][0]), U64(M), U32(M), U32(N), U32(K), P(bias), P(res[i]), U64(M))
            elif what == "qkv":
                PtrArr = C.c_void_p * G
                mb.call("gemm_bf16_grouped", P(a[i]), I32(1), U64(r8(M)), U32(G), PtrArr(*[t.data_ptr() for t in b[i]]), I32(0), U64(r8(K)),
                        PtrArr(*[t.data_ptr() for t in c[i]]), U64(M), U32(M), U32(N), U32(K), I32(0), PtrArr(*[bias.data_ptr()] * G))
            else:
                mb.call("gemm_bf16", P(a[i]), I32(1), U64(r8(M)), P(b[i][0]), I32(0), U64(r8(K)), P(c[i][0]), U64(M), U32(M), U32(N), U32(K), I32(0), P(bias))

        def ex(i, cc, cc16, **kw):
            e = GemmEpilogue(col_bias=bias.data_ptr(), residual=kw.get("residual", 0), ldr=M, activation=kw.get("act", 0), row_stats=kw.get("stats", 0),
                             stats=stats.data_ptr(), stats_capacity_tiles=cap, stats_tiles=C.pointer(tiles), stats_tile_cols=C.pointer(cols))
            mb.call("gemm_bf16_ex", P(a[i]), I32(1), U64(r8(M)), P(b[i][0]), I32(0), U64(r8(K)), cc, U64(M), cc16, U64(M), U32(M), U32(N), U32(K), e)

        variants = {"plain": plain}
        if what == "residual":
            variants["ex residual+ln_stats"] = lambda i: ex(i, P(c[i][0]), None, residual=res[i].data_ptr(), stats=1)
            variants["ex residual (no stats)"] = lambda i: ex(i, P(c[i][0]), None, residual=res[i].data_ptr())
        elif what == "gelu":
            variants["ex f32 + bf16 gelu (dual)"] = lambda i: ex(i, P(c[i][0]), P(c16[i][0]), act=1)
            variants["ex f32 + bf16 copy (dual, no act)"] = lambda i: ex(i, P(c[i][0]), P(c16[i][0]))
            variants["ex bf16-only gelu"] = lambda i: ex(i, None, P(c16[i][0]), act=1)
            variants["ex bf16-only"] = lambda i: ex(i, None, P(c16[i][0]))
        elif what == "lmhead":
            variants["ex bf16-only + lse stats"] = lambda i: ex(i, None, P(c16[i][0]), stats=2)
            variants["ex bf16-only"] = lambda i: ex(i, None, P(c16[i][0]))
            variants["ex f32 + lse stats"] = lambda i: ex(i, P(c[i][0]), None, stats=2)
        elif what == "qkv":
            PtrArr = C.c_void_p * G
            variants["grouped bf16-only"] = lambda i: mb.call("gemm_bf16_grouped_bf16out", P(a[i]), I32(1), U64(r8(M)), U32(G), PtrArr(*[t.data_ptr() for t in b[i]]), I32(0),
                                                              U64(r8(K)), PtrArr(*[t.data_ptr() for t in c16[i]]), U64(M), U32(M), U32(N), U32(K),
                                                              PtrArr(*[bias.data_ptr()] * G))
        for name, fn in variants.items():
            ms = mb.timeit(fn, nrot, iters=20 if what != "lmhead" else 8, warmup=3)
            print(f"{what:9s} M{M} N{N} K{K} x{G}  {name:36s} {ms * 1000:8.1f} us  {flops / ms / 1e9:7.0f} TFLOP/s  mode={mode}", flush=True)
        del a, b, c, c16, res
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
