"""One launch of each extended-epilogue variant at a reduced LM-head shape, for `ncu --set full --import-source on`."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import microbench as mb  # noqa: E402
from weed_b200 import weedcu  # noqa: E402
from weed_b200._lib import GemmEpilogue  # noqa: E402

U64, U32, I32 = C.c_uint64, C.c_uint32, C.c_int
mb.lib = weedcu()
st = torch.cuda.Stream()
mb.STREAM = st.cuda_stream
torch.cuda.set_stream(st)
P = mb.P
M, N, K = 8192, 9472, 768   # 37 column tiles of 256: one wave of 2 x 74 ... a few waves, short capture
a = (torch.randn(M * K + 8, device="cuda") * 0.05).to(torch.bfloat16).view(torch.int16)
b = (torch.randn(K * N + 8, device="cuda") * 0.05).to(torch.bfloat16).view(torch.int16)
c = torch.empty(M * N, device="cuda")
c16 = torch.empty(M * N, dtype=torch.int16, device="cuda")
bias = torch.randn(N, device="cuda")
cap = 2 * ((N + 127) // 128)
stats = torch.empty(cap * M * 2, device="cuda")
tiles, cols = U32(0), U32(0)


def ex(cc, cc16, **kw):
    e = GemmEpilogue(col_bias=bias.data_ptr(), residual=0, ldr=M, activation=kw.get("act", 0), row_stats=kw.get("stats", 0), stats=stats.data_ptr(),
                     stats_capacity_tiles=cap, stats_tiles=C.pointer(tiles), stats_tile_cols=C.pointer(cols))
    mb.call("gemm_bf16_ex", P(a), I32(1), U64(M), P(b), I32(0), U64(K), cc, U64(M), cc16, U64(M), U32(M), U32(N), U32(K), e)


for _ in range(2):
    mb.call("gemm_bf16", P(a), I32(1), U64(M), P(b), I32(0), U64(K), P(c), U64(M), U32(M), U32(N), U32(K), I32(0), P(bias))   # launch 0/1: plain
ex(P(c), None, stats=2)      # launch 2: fp32 + lse partials
ex(None, P(c16))             # launch 3: bf16 only
ex(P(c), None, stats=1)      # launch 4: fp32 + LayerNorm partials
torch.cuda.synchronize()
print("done")
