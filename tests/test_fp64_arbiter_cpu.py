"""Pins the fp64 numpy arbiter (tests/fp64_model.py) to the UNMODIFIED reference CPU build where the reference is
self-consistent (B = 1): encoder layer forward + every parameter gradient, and the token model's logits and
cross-entropy loss value. The arbiter is then the oracle for B > 1 in tests/test_host_gpu.py."""
import os

import numpy as np
import pytest

import cases
import fp64_model as F
import test_host_gpu as G

R = G.R


def test_arbiter_encoder_layer_matches_reference_B1(R):
    B, T, d, Hh = 1, 12, 16, 2
    rng = np.random.default_rng(5)
    x = rng.uniform(-1, 1, size=(B, T, d)).astype(np.float32)
    w = rng.uniform(-1, 1, size=(B, T, d)).astype(np.float32)
    enc = R.module("encoder", d, Hh, 2 * d)
    weights = R.init_params(enc, 9)
    y = R.forward(enc, R.tensor(F.uncol(x), [B, T, d], True))
    R.backward(R.op("sum", [R.op("mul", [y, R.tensor(F.uncol(w), [B, T, d])])]))
    ref_y = F.col(R.read(y), [B, T, d])
    ref_g = []
    for i in range(R.param_count(enc)):
        g = R.grad(R.param(enc, i))
        ref_g.append(R.read_storage(g).astype(np.float64) if g else np.zeros(1))
    R.reset()
    p = F.encoder_params(weights, d, 2 * d)
    ops = F.Ops(False)
    my_y, cache = F.encoder_fwd(ops, x.astype(np.float64), p, Hh)
    _, my_g = F.encoder_bwd(ops, w.astype(np.float64), p, cache)
    assert cases.rel_err(my_y, ref_y) <= 1e-5
    checked = 0
    for i, k in enumerate(F.ENC_PARAMS):
        m = F.uncol(my_g[k]) if my_g[k].ndim == 2 else my_g[k]
        if ref_g[i].size != m.size:
            assert not np.any(ref_g[i]) and not np.any(m), k
            continue
        if not np.any(m):
            assert not np.any(ref_g[i]), k  # W_q / W_k / W_v / norm1: the reference sends no gradient there either (D9)
            continue
        assert cases.rel_err(m, ref_g[i]) <= 2e-5, (k, cases.rel_err(m, ref_g[i]))
        checked += 1
    assert checked == 8


def test_arbiter_token_model_logits_and_loss_match_reference_B1(R):
    """The reference's Embedding gathers token 0 only for indices [1, T] (stride[0] == 0, defect D6), so its model starts
    at the positional encoding on a dense x[1, T, d]; the arbiter gets the same x as an embedding table indexed by
    arange(T)."""
    cfg = dict(V=40, d=16, H=2, dff=32, L=2, T=8)
    T, d = cfg["T"], cfg["d"]
    rng = np.random.default_rng(61)
    x = rng.uniform(-1, 1, size=(1, T, d)).astype(np.float32)
    targets = rng.integers(0, cfg["V"], size=(1, T)).astype(np.int32)
    mods = [R.module("posenc", T, d)] + [R.module("encoder", d, cfg["H"], cfg["dff"]) for _ in range(cfg["L"])]
    mods += [R.module("layernorm", d), R.module("linear", d, cfg["V"], 1)]
    model = R.module("sequential", *mods)
    weights = R.init_params(model, 62)
    logits = R.forward(model, R.tensor(F.uncol(x), [1, T, d]))
    loss = R.cross_entropy(logits, R.symbol(F.uncol(targets), [T]))  # rank-1: a [1, T] symbol has stride[0] == 0 (D6)
    ref_logits = F.col(R.read(logits), [1, T, cfg["V"]])
    ref_loss = float(R.read(loss)[0])
    R.reset()
    table = np.zeros((cfg["V"], d), np.float32)
    table[:T] = x[0]
    my_loss, my_logits, _ = F.token_model([F.uncol(table)] + weights, cfg, np.arange(T, dtype=np.int32)[None], targets)
    assert cases.rel_err(my_logits, ref_logits) <= 1e-5
    assert abs(my_loss - ref_loss) <= 1e-5 * abs(ref_loss)


def test_bf16_round_is_round_to_nearest_even():
    x = np.array([1.0, 1.0078125, 1.00390625, 1.01171875, -3.1415927, 65504.0, 1e-30], np.float32)
    r = F.bf16_round(x)
    assert r[0] == 1.0 and r[1] == 1.0078125  # representable (7 mantissa bits)
    assert r[2] == 1.0                           # tie -> even mantissa
    assert r[3] == 1.015625                      # tie -> even mantissa (up)
    import torch
    want = torch.from_numpy(x).to(torch.bfloat16).to(torch.float64).numpy()
    assert np.array_equal(r, want)
