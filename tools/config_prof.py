"""Per-kernel-class device time of one step of config C2 / C4 (instrumented pass, CUDA events per launch)."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import config_sweep as cs  # noqa: E402
from weed_b200 import weedcu, check  # noqa: E402
from weed_b200.harness import Harness  # noqa: E402

lib = weedcu()
check(lib.weedcu_set_device(C.c_int(0)))
P = Harness.product()
P.config("fused", 1)
names = {1: "gemm_bf16_tcgen05", 2: "gemm_f32_ffma", 3: "pack_bf16", 4: "elementwise", 5: "softmax", 6: "layernorm", 7: "cross_entropy", 8: "optimizer",
         9: "reduce", 10: "embedding", 11: "fill", 12: "nccl", 13: "attention_flash_tcgen05"}


class ProfTimer(cs.Timer):
    def time(self, fn, iters, warmup):
        for i in range(warmup):
            fn(i)
        P.sync()
        lib.weedcu_prof_enable(C.c_int(1))
        for i in range(2):
            fn(i)
        P.sync()
        lib.weedcu_prof_enable(C.c_int(0))
        out = {}
        for cls, nm in names.items():
            t, n, w = C.c_double(), C.c_uint64(), C.c_double()
            lib.weedcu_prof_read(C.c_int(cls), C.byref(t), C.byref(n), C.byref(w))
            if n.value:
                out[nm] = (round(t.value / 2, 3), n.value / 2)
        print(json.dumps(out))
        return cs.Timer.time(self, fn, iters, 0)


t = ProfTimer(lib, P.stream(), check)
for name, job in (("c2", cs.run_c2), ("c4", cs.run_c4)):
    print(name)
    r = job(P, t, {})
    print(name, round(r["ms_per_step"], 3))
