#!/usr/bin/env python
"""bench.py — headline benchmark of BASELINE.json: train samples/s at GPT-2-small shape.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (config C5): Embedding(50257,768) + LearnedPositionalEncoding(1024,768) +
12 x TransformerEncoderLayer(768, 12 heads, d_ff 3072) + LayerNorm(768) + Linear(768,50257)
(untied, ~163 M parameters), batch 8 sequences of 1024 tokens per GPU, fused cross-entropy, Adam.
Synthetic seeded tokens and random-init weights. One "step" = forward + loss + backward +
(gradient all-reduce) + Adam + zero_grad through Weed's own API (Sequential::forward,
cross_entropy_loss, Tensor::backward, adam_step, zero_grad) on this repo's CUDA host library.

Default arm ("ours"): one process per GPU (torchrun for N > 1), data parallel, NCCL gradient
all-reduce. `value` is measured with inputs resident in HBM; `e2e` repeats the same K steps with the
per-step tokens/targets copied from pinned host memory and the loss read back every step.

--impl reference: the UNMODIFIED reference CPU build (oracle/_ref, compiled from /root/reference by
oracle/Makefile) runs the same model code on the host cores on a BOUNDED sample of the workload
(fewer layers / tokens / vocabulary so a run ends in minutes); its samples/s is extrapolated to the
full configuration by the algorithmic FLOP ratio and labelled as such.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FULL = dict(V=50257, d=768, H=12, dff=3072, L=12, T=1024, B=8)
GEMM_FP32, GEMM_BF16 = 0, 1


# ----------------------------------------------------------------------------------- model
def build_model(H, cfg, seed=2000):
    mods = [H.module("embedding", cfg["V"], cfg["d"]), H.module("posenc", cfg["T"], cfg["d"])]
    mods += [H.module("encoder", cfg["d"], cfg["H"], cfg["dff"]) for _ in range(cfg["L"])]
    mods += [H.module("layernorm", cfg["d"]), H.module("linear", cfg["d"], cfg["V"], 1)]
    model = H.module("sequential", *mods)
    # seeded weights (never std::random_device): uniform(+-lim), LayerNorm gamma=1 beta=0
    rng = np.random.default_rng(seed)
    n_params = 0
    for i in range(H.param_count(model)):
        n = H.param_size(model, i)
        n_params += n
        if n == cfg["d"]:
            continue  # biases / gamma / beta keep their constructor values (0, 1, 0)
        lim = 0.02 if n >= cfg["V"] * cfg["d"] else float(np.sqrt(6.0 / (cfg["d"] + n // cfg["d"])))
        H.param_set(model, i, rng.uniform(-lim, lim, size=n).astype(np.float32))
    return model, n_params


def step_flops(cfg):
    """Algorithmic FLOP of one training step per rank under the reference's autograd semantics:
    every GEMM forward; dA + dB for W_o / ff1 / ff2 / LM head only — no gradient flows through the
    batched attention products or into W_q/W_k/W_v (SURVEY §7 hard part 5(i))."""
    toks = cfg["B"] * cfg["T"]
    d, dff, V, L, T, Hh = cfg["d"], cfg["dff"], cfg["V"], cfg["L"], cfg["T"], cfg["H"]
    proj = 2.0 * toks * d * d
    ffn = 2.0 * toks * d * dff
    attn = 2 * 2.0 * cfg["B"] * Hh * T * T * (d // Hh)
    head = 2.0 * toks * d * V
    fwd = L * (4 * proj + 2 * ffn + attn) + head
    bwd = L * 2 * (proj + 2 * ffn) + 2 * head
    return fwd + bwd


def make_tokens(cfg, seed):
    rng = np.random.default_rng(seed)
    tok = rng.integers(0, cfg["V"], size=cfg["B"] * cfg["T"]).astype(np.int32)
    tgt = rng.integers(0, cfg["V"], size=cfg["B"] * cfg["T"]).astype(np.int32)
    return tok, tgt


# ----------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ----------------------------------------------------------------------------------- reference arm
def reference_run(cfg_small, steps, warmup):
    from weed_b200.harness import Harness
    R = Harness.reference()
    model, _ = build_model(R, cfg_small)
    opt = R.adam(model, 1e-4)
    times = []
    for s in range(warmup + steps):
        tok, tgt = make_tokens(cfg_small, 3000 + s)
        t0 = time.perf_counter()
        st = R.symbol(tok, [cfg_small["B"], cfg_small["T"]])
        sg = R.symbol(tgt, [cfg_small["B"], cfg_small["T"]])
        loss = R.train_step_tokens(model, opt, st, sg)
        lv = float(R.read(loss)[0])
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
        for h in (st, sg, loss):
            R.free(h)
    R.reset()
    return float(np.mean(times)), lv


def pick_sample(budget_s, gflops=0.25):
    """Largest bounded sample of C5 whose estimated CPU time fits the per-step budget."""
    best = None
    for L in (1,):
        # B = 1: the reference's cross_entropy_loss reshapes to {T, V} and throws for B > 1
        # (include/autograd/cross_entropy_loss.hpp:22-27)
        for T, B in ((32, 1), (64, 1), (128, 1), (256, 1)):
            for V in (1024, 2048, 4096):
                if T * V > (1 << 18):
                    # at >= 2*2^18 items the reference's par_for goes multi-threaded and its sparse
                    # one-hot product (unordered_map writes from several threads) segfaults [measured]
                    continue
                c = dict(FULL, L=L, T=T, B=B, V=V)
                est = step_flops(c) / (gflops * 1e9)
                if est <= budget_s and (best is None or step_flops(c) > step_flops(best)):
                    best = c
    return best or dict(FULL, L=1, T=32, B=1, V=1024)


def reference_budget(steps, warmup):
    """Per-step CPU budget of the bounded reference sample: the driver's `--impl reference --steps K --warmup W` run has to
    end within a few minutes, so the sample shrinks with K + W. cpu_baseline (one step inside our arm) uses the SAME
    sample, so the record holds one CPU configuration, not two."""
    return min(30.0, max(3.0, 150.0 / max(1, steps + warmup)))


def cpu_baseline(cfg_full, steps=1, warmup=0, budget_s=20.0):
    small = pick_sample(budget_s)
    t, loss = reference_run(small, steps, warmup)
    ratio = step_flops(cfg_full) / step_flops(small)
    return {"value": cfg_full["B"] / (t * ratio), "unit": "samples/s", "cores": os.cpu_count(), "kind": "reference",
            "sample": (f"unmodified reference CPU build (oracle/_ref), 1 train step on L={small['L']} T={small['T']} B={small['B']} "
                       f"V={small['V']} d=768 dff=3072: {t:.2f} s/step measured ({step_flops(small) / t / 1e9:.2f} GFLOP/s, loss {loss:.3f}); "
                       f"extrapolated x{ratio:.0f} by algorithmic FLOP to the full config (optimistic for the CPU: its step is dominated by "
                       f"per-op overhead that grows with tokens, not FLOP); WEED_BLAS=OFF; par_for runs serial below "
                       f"2*2^18 items so most ops use 1 of the {os.cpu_count()} threads")}, t


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = dict(FULL)
    small = pick_sample(reference_budget(args.steps, args.warmup))
    t, loss = reference_run(small, args.steps, args.warmup)
    ratio = step_flops(cfg) / step_flops(small)
    value = cfg["B"] / (t * ratio)
    sample = (f"reference CPU build on L={small['L']} T={small['T']} B={small['B']} V={small['V']} (d=768, dff=3072): {t:.2f} s/step, "
              f"extrapolated x{ratio:.0f} by algorithmic FLOP to L=12 T=1024 B=8 V=50257")
    line = {"impl": "reference", "metric": "train samples/s (GPT-2-small shape)", "value": value, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * ratio * 1000.0, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C5 GPT-2-small-shape training step (untied, 163M params)", "layers": cfg["L"], "d_model": cfg["d"],
                       "heads": cfg["H"], "d_ff": cfg["dff"], "vocab": cfg["V"], "seq_len": cfg["T"], "batch_per_gpu": cfg["B"],
                       "global_batch": cfg["B"], "parallelism": "cpu", "optimizer": "Adam", "loss": "cross_entropy_loss",
                       "measured_on": "bounded CPU sample, extrapolated by algorithmic FLOP", "sample": sample},
            "cpu_baseline": {"value": value, "unit": "samples/s", "cores": os.cpu_count(), "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------- parity legs
def loss_trajectory(P, cfg, precision, steps, pdl=1, seed=2000, lr=1e-4):
    """`steps` seeded training steps from the seeded initial weights; returns the loss of every step. Everything built
    here is released again."""
    mark = P.mark()
    P.config("matmul_precision", precision)
    P.config("pdl", pdl)
    model, _ = build_model(P, cfg, seed)
    opt = P.adam(model, lr)
    out = []
    for s in range(steps):
        tok, tgt = make_tokens(cfg, 9000 + s)
        st, sg = P.symbol(tok, [cfg["B"], cfg["T"]]), P.symbol(tgt, [cfg["B"], cfg["T"]])
        h = P.train_step_tokens(model, opt, st, sg)
        out.append(float(P.read(h)[0]))
        for x in (h, st, sg):
            P.free(x)
    P.release_since(mark)
    P.config("pdl", 1)
    return out


def parity_legs(P, cfg, steps=5):
    """Full-shape correctness evidence for the configuration being timed (VERDICT r1 #1d/e): the bf16 tensor-core run
    against the fp32 FpMath run of the same seeded steps, and programmatic dependent launch on against off."""
    bf16 = loss_trajectory(P, cfg, GEMM_BF16, steps, pdl=1)
    bf16_nopdl = loss_trajectory(P, cfg, GEMM_BF16, steps, pdl=0)
    fp32 = loss_trajectory(P, cfg, GEMM_FP32, steps, pdl=1)
    rel = lambda a, b: float(max(abs(x - y) / abs(y) for x, y in zip(a, b)))
    return {"steps": steps, "loss_bf16": bf16, "loss_fp32": fp32, "loss_bf16_pdl_off": bf16_nopdl,
            "loss_rel_diff_bf16_vs_fp32": rel(bf16, fp32), "bf16_loss_bound": BF16_LOSS_BOUND,
            "within_bf16_bound": rel(bf16, fp32) <= BF16_LOSS_BOUND,
            "loss_rel_diff_pdl_on_vs_off": rel(bf16, bf16_nopdl), "pdl_bound": 2e-5, "pdl_equal": rel(bf16, bf16_nopdl) <= 2e-5}


# stated bf16 bound on the loss of the benchmarked configuration after 5 seeded steps (north_star: "a stated looser bound
# for bf16"). Measured on B200 in round 2: 3.0e-4 with bf16 logits feeding the loss, 9e-6 with fp32 logits
# (profiles/r02_parity_ab.txt); the fp32 trajectory itself moves by ~1e-5 between builds at step 4-5 (Adam amplifies
# rounding-level differences ~3x per step at this shape), so the bound keeps a factor 3 over the measured figure.
BF16_LOSS_BOUND = 1e-3


# ----------------------------------------------------------------------------------- data-parallel equality
def check_dp(P, args, cfg, rank, world, dist, torch):
    """N ranks on a global batch of GB sequences (GB / N per rank, bucketed gradient all-reduce overlapped with backward,
    1/N folded into the fused Adam) against ONE rank on the same GB sequences, in both GEMM precisions:
      * per-step loss: <= 1e-3 relative with bf16 tensor-core GEMMs, <= 1e-5 with the fp32 path;
      * Adam's first and second moment of every parameter after the steps (linear / quadratic in the exchanged gradients):
        rel-to-max <= 1e-4 at fp32, <= 2e-2 (the bf16 bound) at bf16; parameters with an exactly-zero gradient are listed.
    Parameter checksums are reported, not gated: Adam's step is lr * m / (sqrt(v) + eps) ~ lr * sign(g) wherever a gradient
    is rounding noise around zero (zero-initialised biases), so two correct runs differ by 2 * lr in those elements."""
    steps = int(os.environ.get("CHECK_DP_STEPS", "3"))
    GB = args.batch * world if args.batch else 16
    assert GB % world == 0
    Bl = GB // world
    gcfg, lcfg = dict(cfg, B=GB), dict(cfg, B=Bl)
    T = cfg["T"]

    def run(c, shard, precision):
        mark = P.mark()
        P.config("matmul_precision", precision)
        model, n_params = build_model(P, c)
        if shard is not None and world > 1:
            assert P.lib.wh_dp_broadcast_params(C.c_int64(model)) == 0
        opt = P.adam(model, 1e-4)
        losses = []
        for s in range(steps):
            tok, tgt = make_tokens(gcfg, 11000 + s)
            if shard is not None:
                tok = np.ascontiguousarray(tok.reshape(T, GB)[:, shard * Bl:(shard + 1) * Bl]).ravel()
                tgt = np.ascontiguousarray(tgt.reshape(T, GB)[:, shard * Bl:(shard + 1) * Bl]).ravel()
            st, sg = P.symbol(tok, [c["B"], T]), P.symbol(tgt, [c["B"], T])
            h = P.train_step_tokens(model, opt, st, sg)
            losses.append(float(P.read(h)[0]))
            for x in (h, st, sg):
                P.free(x)
        moments, sums = [], []
        for i in range(P.param_count(model)):
            moments.append((P.read_storage(P.adam_moment(opt, model, i, 0)), P.read_storage(P.adam_moment(opt, model, i, 1))))
            w = P.read_storage(P.param(model, i)).astype(np.float64)
            sums.append((float(w.sum()), float(np.abs(w).sum())))
        P.release_since(mark)
        return losses, moments, np.array(sums)

    result = {"check_dp": True, "n_gpus": world, "global_batch": GB, "batch_per_gpu": Bl, "steps": steps, "layers": cfg["L"], "vocab": cfg["V"],
              "seq_len": T, "adam_chained_per_bucket": os.environ.get("WH_DP_CHAIN_ADAM", "0") != "0", "ok": True}
    # The gated legs run LayerNorm's backward with the analytic gradient. The reference's autograd chain (the default, kept
    # for parity) leaves a term in dx that does not scale with the upstream gradient (core.hpp: layernorm_exact_grad), so
    # with it a gradient of the mean loss over GB sequences and the mean of N gradients over GB / N sequences differ by
    # (1 - 1/N) of that term per LayerNorm -- measured ~0.3 % of the gradient per layer at this shape -- whoever exchanges
    # them. The third leg reports that figure, ungated.
    legs = ((GEMM_FP32, "fp32", 1e-5, 1e-4, 1, True), (GEMM_BF16, "bf16", 1e-3, 2e-2, 1, True),
            (GEMM_BF16, "bf16_reference_layernorm_chain", 1e-3, 2e-2, 0, False))
    for precision, name, loss_bound, moment_bound, exact_ln, gated in legs:
        P.config("layernorm_exact_grad", exact_ln)
        dp_losses, dp_mom, dp_sums = run(lcfg, rank, precision)
        t = torch.tensor(dp_losses, dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        dp_losses = (t / world).cpu().tolist()  # equal shards: the mean of the rank means is the global mean
        if rank == 0:
            P.lib.wh_dp_set_active(C.c_int(0))
            one_losses, one_mom, one_sums = run(gcfg, None, precision)
            P.lib.wh_dp_set_active(C.c_int(1))
            loss_rel = float(max(abs(a - b) / abs(b) for a, b in zip(dp_losses, one_losses)))
            # Parameters whose gradient is exactly zero -- the Q/K/V projections and the LayerNorm in front of them: the
            # reference propagates nothing through its batched matmul (SURVEY defect D9), reproduced here -- or rounding
            # noise only (max|m| below 1e-4 of the median over all parameters) are listed, not compared.
            scale = np.array([float(np.max(np.abs(m0))) for m0, _ in one_mom])
            floor = 1e-4 * float(np.median(scale[scale > 0]))
            worst_m = worst_v = 0.0
            noise, rows = [], []
            for i, ((m1, v1), (m0, v0)) in enumerate(zip(dp_mom, one_mom)):
                if scale[i] <= floor:
                    noise.append({"param": i, "size": int(m0.size), "max_abs_m": scale[i]})
                    continue
                dm = float(np.max(np.abs(m1 - m0)) / np.max(np.abs(m0)))
                dv = float(np.max(np.abs(v1 - v0)) / np.max(np.abs(v0)))
                rows.append((max(dm, dv), i, int(m0.size), scale[i], dm, dv))
                worst_m, worst_v = max(worst_m, dm), max(worst_v, dv)
            if os.environ.get("CHECK_DP_VERBOSE"):
                print(name, [(r[1], round(r[4], 6)) for r in sorted(rows, key=lambda r: r[1])], file=sys.stderr)
            rows.sort(reverse=True)
            chk = float(np.max(np.abs(dp_sums - one_sums) / np.maximum(one_sums[:, 1:2], 1e-30)))
            ok = bool(loss_rel <= loss_bound and worst_m <= moment_bound and worst_v <= moment_bound)
            result[name] = {"loss_dp": dp_losses, "loss_single": one_losses, "loss_rel_diff": loss_rel, "loss_bound": loss_bound,
                            "adam_m_rel_to_max_diff": worst_m, "adam_v_rel_to_max_diff": worst_v, "moment_bound": moment_bound,
                            "params_gated": len(rows), "params_zero_gradient_noise_only": noise, "median_max_abs_m": float(np.median(scale)),
                            "worst_params": [{"param": r[1], "size": r[2], "max_abs_m": r[3], "m_diff": r[4], "v_diff": r[5]} for r in rows[:4]],
                            "param_checksum_rel_diff_not_gated": chk, "layernorm_backward": "analytic" if exact_ln else "reference chain",
                            "gated": gated, "ok": ok}
            if gated:
                result["ok"] = result["ok"] and ok
        dist.barrier()
    P.config("layernorm_exact_grad", 0)
    return result if rank == 0 else None


# ----------------------------------------------------------------------------------- our arm
def main_ours(args):
    import torch
    from weed_b200 import weedcu, check
    from weed_b200.harness import Harness

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    lib = weedcu()
    check(lib.weedcu_set_device(C.c_int(local)))
    dist = None
    if world > 1:
        import torch.distributed as dist
        import datetime
        # a short collective timeout: a rank that falls out of step must fail the run, not hang it
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=180))
    P = Harness.product()
    cfg = dict(FULL)
    if args.layers:
        cfg["L"] = args.layers
    if args.seq:
        cfg["T"] = args.seq
    if args.batch:
        cfg["B"] = args.batch
    if args.vocab:
        cfg["V"] = args.vocab
    precision = GEMM_BF16 if args.precision == "bf16" else GEMM_FP32
    P.config("fused", 1)
    P.config("ref_index_quirks", 0)
    P.config("matmul_precision", precision)
    for kv in filter(None, os.environ.get("WH_CONFIG", "").split(",")):  # backend switches for A/B runs: WH_CONFIG="bf16_act_grad=0,..."
        k, v = kv.split("=")
        P.config(k.strip(), float(v))
    if world > 1:
        import torch as _t
        nccl_path = os.path.join(os.path.dirname(_t.__file__), "..", "nvidia", "nccl", "lib", "libnccl.so.2")
        rc = P.lib.wh_dp_load(os.path.abspath(nccl_path).encode())
        assert rc == 0, "could not load NCCL"
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf = (C.c_uint8 * 128)()
            assert P.lib.wh_dp_unique_id(buf) == 0
            uid.copy_(torch.tensor(list(buf), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        host = (C.c_uint8 * 128)(*uid.cpu().tolist())
        assert P.lib.wh_dp_init(host, C.c_int(rank), C.c_int(world)) == 0, P.lib.wh_last_error()

    if args.check_dp:
        assert world > 1, "--check-dp compares N > 1 ranks against one rank: launch under torchrun"
        res = check_dp(P, args, cfg, rank, world, dist, torch)
        if rank == 0:
            print(json.dumps(res))
        dist.destroy_process_group()
        return 0 if (rank != 0 or res["ok"]) else 1

    model, n_params = build_model(P, cfg)
    if world > 1:
        assert P.lib.wh_dp_broadcast_params(C.c_int64(model)) == 0
    opt = P.adam(model, 1e-4)
    stream = P.stream()
    ntok = cfg["B"] * cfg["T"]

    def ev():
        e = C.c_void_p()
        check(lib.weedcu_event_create(C.byref(e)))
        return e

    def barrier():
        P.sync()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = ev(), ev()
        barrier()
        check(lib.weedcu_event_record(e0, C.c_void_p(stream)))
        for s in range(steps):
            fn(s)
        check(lib.weedcu_event_record(e1, C.c_void_p(stream)))
        check(lib.weedcu_event_sync(e1))
        barrier()
        ms = C.c_float()
        check(lib.weedcu_event_elapsed_ms(e0, e1, C.byref(ms)))
        t = torch.tensor([ms.value], device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # device-resident inputs: one (tokens, targets) pair per step, uploaded before the timed region
    total_steps = args.warmup + args.steps
    pairs = []
    for s in range(total_steps):
        tok, tgt = make_tokens(cfg, 3000 + 1000 * rank + s)
        pairs.append((P.symbol(tok, [cfg["B"], cfg["T"]]), P.symbol(tgt, [cfg["B"], cfg["T"]])))
    losses = []

    def resident_step(s):
        # the loss handle keeps the step's autograd graph (all activations) alive: drop the previous
        # one as soon as the next step is issued; nothing is read back inside the timed loop
        st, sg = pairs[s % len(pairs)]
        h = P.train_step_tokens(model, opt, st, sg)
        for old in losses:
            P.free(old)
        losses.clear()
        losses.append(h)

    first_loss = None
    for s in range(args.warmup):
        resident_step(s)
        if s == 0:
            first_loss = float(P.read(losses[0])[0])
    for h in losses:
        P.free(h)
    losses.clear()

    def host_stats():
        a, b, c, d = C.c_double(), C.c_uint64(), C.c_double(), C.c_uint64()
        lib.weedcu_host_stats(C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        return a.value, b.value, c.value, d.value

    n0, n1 = C.c_uint64(), C.c_uint64()
    sampler = ClockSampler(local)
    if rank == 0 and not args.no_clocks:
        sampler.start()
    lib.weedcu_launch_count(C.byref(n0))
    hs0 = host_stats()
    host_t0 = time.perf_counter()
    host_issue = [0.0]

    def timed_step(s):
        resident_step(args.warmup + s)
        host_issue[0] = time.perf_counter() - host_t0

    ms = timed(timed_step, args.steps)
    hs1 = host_stats()
    lib.weedcu_launch_count(C.byref(n1))
    clocks = sampler.stop() if (rank == 0 and not args.no_clocks) else None
    host = {"issue_ms_per_step": host_issue[0] * 1000.0 / args.steps,
            "pool_malloc_ms_per_step": (hs1[0] - hs0[0]) / args.steps, "pool_mallocs_per_step": (hs1[1] - hs0[1]) / args.steps,
            "pool_free_ms_per_step": (hs1[2] - hs0[2]) / args.steps, "pool_frees_per_step": (hs1[3] - hs0[3]) / args.steps}
    last_loss = float(P.read(losses[-1])[0])
    for h in losses:
        P.free(h)
    losses.clear()
    # host issue time of one step with an EMPTY launch queue (issue_ms_per_step above includes the time
    # the host spends blocked on the full queue once it is ~1000 launches ahead of the device)
    unblocked = []
    for s in range(3):
        P.sync()
        t0 = time.perf_counter()
        resident_step(s)
        unblocked.append((time.perf_counter() - t0) * 1000.0)
    P.sync()
    for h in losses:
        P.free(h)
    losses.clear()
    host["issue_ms_per_step_unblocked"] = float(np.median(unblocked))
    ms_per_step = ms / args.steps
    value = cfg["B"] * world / (ms_per_step / 1000.0)

    # end to end: per-step H2D of tokens+targets from pinned memory, D2H loss read every step
    pin = C.c_void_p()
    check(lib.weedcu_host_alloc(C.byref(pin), C.c_size_t(2 * ntok * 4)))
    pin_arr = np.ctypeslib.as_array(C.cast(pin, C.POINTER(C.c_int32)), shape=(2 * ntok,))
    st, sg = pairs[0]
    host_tokens = [make_tokens(cfg, 7000 + 1000 * rank + s) for s in range(args.steps)]
    e2e_losses = []

    def e2e_step(s):
        tok, tgt = host_tokens[s]
        pin_arr[:ntok] = tok
        pin_arr[ntok:] = tgt
        P.symbol_upload(st, pin.value, ntok)
        P.symbol_upload(sg, pin.value + 4 * ntok, ntok)
        h = P.train_step_tokens(model, opt, st, sg)
        e2e_losses.append(float(P.read(h)[0]))  # blocking 4-byte device->host read
        P.free(h)

    ms_e2e = timed(e2e_step, args.steps)
    e2e_value = cfg["B"] * world / (ms_e2e / args.steps / 1000.0)

    # per-kernel-class device time (instrumented pass over the same step, CUDA events per launch)
    roofline, breakdown = None, {}
    # every rank runs the same steps (a training step contains the gradient all-reduce: a rank that
    # stepped alone would wait for its peers forever); only rank 0 records the per-launch events
    psteps = min(2, args.steps)
    if rank == 0:
        lib.weedcu_prof_enable(C.c_int(1))
    for s in range(psteps):
        resident_step(s)
    P.sync()
    if rank == 0:
        lib.weedcu_prof_enable(C.c_int(0))
    for h in losses:
        P.free(h)
    losses.clear()
    if rank == 0:
        names = {1: "gemm_bf16_tcgen05", 2: "gemm_f32_ffma", 3: "pack_bf16", 4: "elementwise", 5: "softmax", 6: "layernorm", 7: "cross_entropy",
                 8: "optimizer", 9: "reduce", 10: "embedding", 11: "fill", 12: "nccl", 13: "attention_flash_tcgen05"}
        tot = 0.0
        for cls, nm in names.items():
            t, n, w = C.c_double(), C.c_uint64(), C.c_double()
            lib.weedcu_prof_read(C.c_int(cls), C.byref(t), C.byref(n), C.byref(w))
            if n.value:
                breakdown[nm] = {"ms_per_step": t.value / psteps, "launches_per_step": n.value / psteps, "work_per_step": w.value / psteps}
                tot += t.value / psteps
        peaks = {}
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            peaks = json.load(open(pk))
        if breakdown:
            top = max(breakdown, key=lambda k: breakdown[k]["ms_per_step"])
            b = breakdown[top]
            if top.startswith("gemm") or top.startswith("attention"):
                peak = peaks.get("bf16_tflops_sustained", 1400.0)
                ach = b["work_per_step"] / (b["ms_per_step"] / 1000.0) / 1e12
                roofline = {"kernel": top, "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                            "traffic": None, "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1400",
                            "share_of_step": b["ms_per_step"] / ms_per_step, "launches_per_step": b["launches_per_step"]}
            else:
                peak = peaks.get("hbm_gbs", 6650.0)
                ach = b["work_per_step"] / (b["ms_per_step"] / 1000.0) / 1e9
                roofline = {"kernel": top, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                            "share_of_step": b["ms_per_step"] / ms_per_step, "launches_per_step": b["launches_per_step"]}
            roofline["instrumented_ms_per_step"] = tot
            # DRAM bytes per launch of this kernel class from the committed ncu capture of the same command
            # (tools/ncu_summary.py traffic -> profiles/kernel_traffic.json); null when no capture covers the class
            tpath = os.path.join(ROOT, "profiles", "kernel_traffic.json")
            if os.path.exists(tpath):
                tr = json.load(open(tpath)).get(top)
                if tr:
                    roofline["traffic"] = tr["dram_bytes_per_launch"]
                    roofline["traffic_source"] = tr["source"]
            # The instrumented pass brackets every launch with two event records; the classes then sum to more than the
            # un-instrumented step although the same kernels run back to back. That excess, spread evenly over the launches,
            # is taken off each class for the *_net figures (frac / achieved stay the raw, conservative ones).
            n_launch = sum(v["launches_per_step"] for v in breakdown.values())
            over_ms = max(0.0, tot - ms_per_step) / max(1.0, n_launch)
            net_ms = max(1e-9, b["ms_per_step"] - over_ms * b["launches_per_step"])
            unit_div = 1e12 if roofline["bound"] == "tensor" else 1e9
            roofline["instrumentation_overhead_us_per_launch"] = over_ms * 1000.0
            roofline["achieved_net"] = b["work_per_step"] / (net_ms / 1000.0) / unit_div
            roofline["frac_net"] = roofline["achieved_net"] / roofline["peak"]

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu, _ = cpu_baseline(cfg, budget_s=reference_budget(args.steps, args.warmup))
        except Exception as e:  # the oracle build is optional on a box without oracle/_ref
            cpu = {"value": None, "unit": "samples/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"unavailable: {e}"}

    parity = None
    if rank == 0 and world == 1 and not args.no_parity:
        P.free(st)
        P.free(sg)
        for a, b in pairs:
            P.free(a)
            P.free(b)
        parity = parity_legs(P, cfg)
        P.config("matmul_precision", precision)

    extra = None
    if rank == 0 and world == 1 and not args.no_configs:
        # the other named configurations of BASELINE.json (C2, C3 op sweep, C4, the decode half of C5), each with its own
        # roofline fraction and clocks record: tools/config_sweep.py
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import config_sweep
        pk_all = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        extra = {"configs": config_sweep.run_all(P, lib, check, pk_all, (lambda: ClockSampler(local)) if not args.no_clocks else None)}
        P.config("matmul_precision", precision)

    if rank == 0:
        line = {"metric": "train samples/s (GPT-2-small shape)", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16" if precision == GEMM_BF16 else "f32", "data": "synthetic",
                "config": {"workload": "C5 GPT-2-small-shape training step (untied, %.0fM params)" % (n_params / 1e6), "layers": cfg["L"],
                           "d_model": cfg["d"], "heads": cfg["H"], "d_ff": cfg["dff"], "vocab": cfg["V"], "seq_len": cfg["T"],
                           "batch_per_gpu": cfg["B"], "global_batch": cfg["B"] * world, "parallelism": f"dp{world}", "optimizer": "Adam",
                           "loss": "fused cross-entropy", "gemm": "bf16 tcgen05 operands, fp32 accumulate/outputs" if precision == GEMM_BF16 else "fp32 FFMA",
                           "non_gemm": "fp32", "autograd": "reference semantics (batched attention products carry no grad, tensor.cpp:1253-1271)",
                           "l2": "working set >> 126 MB L2 (activations ~10 GB/step); no flush needed", "algorithmic_tflop_per_step": step_flops(cfg) / 1e12,
                           "loss_first": first_loss, "loss_last": last_loss, "parity": parity},
                "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": 2 * ntok * 4, "d2h_bytes_per_step": 4,
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": int(n1.value - n0.value), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
                "kernel_breakdown": breakdown, "host": host, "model_tflops": step_flops(cfg) * world / (ms_per_step / 1000.0) / 1e12, "extra": extra}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--layers", type=int, default=0)
    ap.add_argument("--seq", type=int, default=0)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--vocab", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the full-shape bf16-vs-fp32 / PDL on-vs-off loss legs")
    ap.add_argument("--no-configs", action="store_true", help="skip the C2 / C3 / C4 / decode measurements (extra.configs)")
    ap.add_argument("--check-dp", action="store_true",
                    help="N ranks on a global batch vs one rank on the same batch (loss <= 1e-3, parameter checksum <= 1e-5)")
    ap.add_argument("--no-clocks", action="store_true", help="do not poll nvidia-smi during the timed region")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    # stdout carries exactly ONE JSON line: native libraries that write to fd 1 (NCCL's version banner at
    # communicator creation) are sent to stderr while the run lasts
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    rc = main_reference(args) if args.impl == "reference" else main_ours(args)
    sys.stdout.flush()
    sys.exit(rc)


if __name__ == "__main__":
    main()
