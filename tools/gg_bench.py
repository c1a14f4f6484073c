"""Device time of the fused GELU-backward + bf16 pack + column sums at 8192 x 3072 in its variants."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import microbench as mb  # noqa: E402
from weed_b200 import weedcu  # noqa: E402

U32, I32 = C.c_uint32, C.c_int


def main():
    mb.lib = weedcu()
    st = torch.cuda.Stream()
    mb.STREAM = st.cuda_stream
    torch.cuda.set_stream(st)
    P = mb.P
    rows, cols = 8192, 3072
    n = rows * cols
    nrot = 4
    x = [torch.randn(n, device="cuda") for _ in range(nrot)]
    g = [torch.randn(n, device="cuda") for _ in range(nrot)]
    g16 = [t.to(torch.bfloat16).view(torch.int16) for t in g]
    d = [torch.zeros(n, device="cuda") for _ in range(nrot)]
    sh = [torch.empty(n, dtype=torch.int16, device="cuda") for _ in range(nrot)]
    cs = torch.zeros(cols, device="cuda")
    variants = {
        "f32 dy, fp32 + bf16 out (accurate tanh)": lambda i: mb.call("gelu_grad_pack", P(d[i]), P(x[i]), P(g[i]), U32(rows), U32(cols), I32(0), P(sh[i]), P(cs)),
        "f32 dy, bf16 out only": lambda i: mb.call("gelu_grad_pack", None, P(x[i]), P(g[i]), U32(rows), U32(cols), I32(0), P(sh[i]), P(cs)),
        "bf16 dy, bf16 out only": lambda i: mb.call("gelu_grad_pack_bf16dy", None, P(x[i]), P(g16[i]), U32(rows), U32(cols), I32(0), P(sh[i]), P(cs)),
    }
    for name, fn in variants.items():
        ms = mb.timeit(fn, nrot, iters=20, warmup=3)
        print(f"{name:45s} {ms * 1e3:8.1f} us", flush=True)


if __name__ == "__main__":
    main()
