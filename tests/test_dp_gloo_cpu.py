"""Data-parallel path with world_size 2 on CPU: two processes, torch.distributed `gloo`, the
product's host library on the oracle-backed mock device (tests/mockdev). Checks that
  * ranks that start from DIFFERENT weights hold identical weights after broadcast + 3 steps,
  * the result equals ONE process training on the concatenated global batch (sum-all-reduce with the
    1/world scale folded into the fused Adam == gradient of the global mean loss),
  * gradients nothing back-propagated into (W_q/W_k/W_v: the reference's batched matmul has no
    grad_node) are skipped by the all-reduce on every rank alike."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MOCK_DIR = os.path.join(ROOT, "tests", "mockdev")
STEPS = 3


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _single_process_reference(world):
    sys.path.insert(0, ROOT)
    from weed_b200.harness import GPU, Harness
    import bench
    P = Harness(os.path.join(MOCK_DIR, "libweed_b200_mock_harness.so"), GPU)
    P.config("fused", 1)
    P.config("matmul_precision", 0)
    P.config("grad_scale", 1.0)
    cfg = dict(V=48, d=16, H=2, dff=32, L=2, T=8, B=2 * world)
    model, _ = bench.build_model(P, cfg, seed=2000)  # rank 0's weights
    opt = P.adam(model, 1e-2)
    tok, tgt = bench.make_tokens(cfg, 5)
    st, sg = P.symbol(tok, [cfg["B"], cfg["T"]]), P.symbol(tgt, [cfg["B"], cfg["T"]])
    losses = [float(P.read(P.train_step_tokens(model, opt, st, sg))[0]) for _ in range(STEPS)]
    params = [P.read_storage(P.param(model, i)).copy() for i in range(P.param_count(model))]
    P.reset()
    return np.array(losses), params


@pytest.mark.parametrize("overlap,bucket_bytes,chain,split", [(1, 0, 1, 0), (1, 2048, 1, 0), (1, 2048, 0, 0), (1, 2048, 0, 1), (0, 0, 0, 0)],
                         ids=["overlap_default_bucket_adam_chained", "overlap_2KB_buckets_adam_chained", "overlap_2KB_buckets",
                              "overlap_2KB_buckets_adam_split_around_last_bucket", "after_backward"])
def test_two_rank_data_parallel_matches_single_process(tmp_path, overlap, bucket_bytes, chain, split):
    """overlap=1: GradientBuckets hands gradients to the all-reduce as Tensor::backward reports them
    final (tiny buckets force several flushes in the middle of the walk); overlap=0: one grouped
    all-reduce after backward. chain=1: the fused Adam update of each bucket's parameters is issued right behind that
    bucket's all-reduce (GradientBuckets::begin(opt, params)) instead of one adam_step after the last bucket.
    All must equal the single-process global-batch step."""
    subprocess.check_call(["make", "-C", MOCK_DIR], stdout=subprocess.DEVNULL)
    world, port = 2, _free_port()
    procs, outs = [], []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   GLOO_SOCKET_IFNAME="lo", WH_DP_OVERLAP=str(overlap), WH_DP_CHAIN_ADAM=str(chain), WH_DP_SPLIT_ADAM=str(split))
        if bucket_bytes:
            env["WH_DP_BUCKET_BYTES"] = str(bucket_bytes)
        out = str(tmp_path / f"rank{r}.json")
        outs.append(out)
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "dp_worker.py"), out, str(STEPS)], env=env))
    for p in procs:
        assert p.wait(timeout=600) == 0
    res = [json.load(open(o)) for o in outs]
    # identical replicas
    for a, b in zip(res[0]["params"], res[1]["params"]):
        assert np.array_equal(np.array(a, np.float32), np.array(b, np.float32))
    # every rank issued the same collectives; untouched gradients were skipped
    assert res[0]["calls"] == res[1]["calls"]
    n_params = len(res[0]["params"])
    assert res[0]["calls"]["bcast"] == n_params
    assert 0 < res[0]["calls"]["allreduce"] < STEPS * n_params
    # == one process on the global batch
    ref_losses, ref_params = _single_process_reference(world)
    dp_losses = np.mean([r["losses"] for r in res], axis=0)
    assert np.max(np.abs(dp_losses - ref_losses) / np.abs(ref_losses)) <= 1e-4, (dp_losses, ref_losses)
    for i, (a, b) in enumerate(zip(res[0]["params"], ref_params)):
        assert np.max(np.abs(np.array(a, np.float32) - b)) <= 2e-4, f"parameter {i}"
