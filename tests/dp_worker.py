"""Worker of tests/test_dp_gloo_cpu.py: one data-parallel rank on the oracle-backed mock device.
The product's host library runs its real DP path (broadcast_parameters, allreduce_gradients, the
1/world gradient scale inside the fused Adam); only the collective itself is swapped: the mock's
weedcu_nccl_* entries call back into this process, which reduces over torch.distributed (gloo)."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
MOCK = os.path.join(ROOT, "tests", "mockdev")


def main():
    out_path, steps = sys.argv[1], int(sys.argv[2])
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from weed_b200.harness import GPU, Harness
    import bench
    P = Harness(os.path.join(MOCK, "libweed_b200_mock_harness.so"), GPU)
    mock = C.CDLL(os.path.join(MOCK, "libweedcu_mock.so"))
    HOOK = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_float), C.c_uint64)
    calls = {"allreduce": 0, "bcast": 0}

    def as_tensor(buf, n):
        return torch.from_numpy(np.ctypeslib.as_array(buf, shape=(n,)))

    def allreduce(buf, n):
        calls["allreduce"] += 1
        dist.all_reduce(as_tensor(buf, n), op=dist.ReduceOp.SUM)
        return 0

    def bcast(buf, n):
        calls["bcast"] += 1
        dist.broadcast(as_tensor(buf, n), 0)
        return 0

    hooks = (HOOK(allreduce), HOOK(bcast))
    mock.weedcu_mock_set_collective_hooks(*hooks)

    P.config("fused", 1)
    P.config("matmul_precision", 0)
    cfg = dict(V=48, d=16, H=2, dff=32, L=2, T=8, B=2)          # per-rank batch
    model, _ = bench.build_model(P, cfg, seed=2000 + 17 * rank)  # ranks start DIFFERENT: broadcast must fix it
    uid = (C.c_uint8 * 128)()
    assert P.lib.wh_dp_load(b"") == 0
    assert P.lib.wh_dp_unique_id(uid) == 0
    assert P.lib.wh_dp_init(uid, C.c_int(rank), C.c_int(world)) == 0
    assert P.lib.wh_dp_broadcast_params(C.c_int64(model)) == 0
    opt = P.adam(model, 1e-2)
    # global batch = world * B sequences, rank r takes sequences [r*B, (r+1)*B)
    gcfg = dict(cfg, B=cfg["B"] * world)
    tok, tgt = bench.make_tokens(gcfg, 5)
    B, T, GB = cfg["B"], cfg["T"], gcfg["B"]
    tok = tok.reshape(T, GB)[:, rank * B:(rank + 1) * B]   # make_tokens lays tokens out [B, T] column-major (b fastest)
    tgt = tgt.reshape(T, GB)[:, rank * B:(rank + 1) * B]
    st = P.symbol(np.ascontiguousarray(tok).ravel(), [B, T])
    sg = P.symbol(np.ascontiguousarray(tgt).ravel(), [B, T])
    losses = [float(P.read(P.train_step_tokens(model, opt, st, sg))[0]) for _ in range(steps)]
    params = [P.read_storage(P.param(model, i)).tolist() for i in range(P.param_count(model))]
    json.dump({"rank": rank, "losses": losses, "params": params, "calls": calls}, open(out_path, "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
