"""One bf16 GEMM shape under one tile configuration, a few launches — the target of `ncu --set full
--import-source on` captures (tools/gpu_runs/gpu_round20.sh).  python tools/gemm_probe.py M N K a_major b_major acc mode"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from weed_b200 import weedcu, check  # noqa: E402

M, N, K, am, bm, acc, mode = [int(x) for x in sys.argv[1:8]]
lib = weedcu()
r8 = lambda x: (x + 7) // 8 * 8
lda, ldb = r8(M if am else K), r8(N if bm else K)
a = torch.randn(lda * (K if am else M), device="cuda").to(torch.bfloat16)
b = torch.randn(ldb * (K if bm else N), device="cuda").to(torch.bfloat16)
c = torch.zeros(M * N, device="cuda")
ts = torch.cuda.Stream()
torch.cuda.synchronize()
lib.weedcu_gemm_set_mode(C.c_int(mode))
fn = lib.weedcu_gemm_bf16
fn.restype = C.c_int
for _ in range(5):
    check(fn(C.c_void_p(a.data_ptr()), C.c_int(am), C.c_uint64(lda), C.c_void_p(b.data_ptr()), C.c_int(bm), C.c_uint64(ldb),
             C.c_void_p(c.data_ptr()), C.c_uint64(M), C.c_uint32(M), C.c_uint32(N), C.c_uint32(K), C.c_int(acc), C.c_void_p(0),
             C.c_void_p(ts.cuda_stream)), "gemm_bf16")
torch.cuda.synchronize()
print("ok", float(c[:16].abs().sum()))
