"""CPU-only coverage of the product's HOST logic (weed_b200/host: views, broadcasting, autograd
graph, modules, optimisers) by re-running the host parity tests of test_host_gpu.py against an
oracle-backed host-memory mock of libweedcu.so (tests/mockdev — test infrastructure only).
The device kernels themselves are covered by the -m gpu suites on the real library."""
import os
import subprocess

import pytest

import test_host_gpu as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MOCK_DIR = os.path.join(ROOT, "tests", "mockdev")


@pytest.fixture(scope="module")
def P():
    from weed_b200.harness import GPU, Harness
    subprocess.check_call(["make", "-C", MOCK_DIR], stdout=subprocess.DEVNULL)
    h = Harness(os.path.join(MOCK_DIR, "libweed_b200_mock_harness.so"), GPU)
    assert h.backend() == "weed_b200"
    return h


R = G.R

test_host_op_matches_reference_fixture = G.test_host_op_matches_reference_fixture
test_intended_axis_sum = G.test_intended_axis_sum_differs_from_reference_only_by_permutation
test_config_c1_xor = G.test_config_c1_xor_training_matches_reference
test_config_c2_mlp = G.test_config_c2_tabular_mlp_training_matches_reference
test_mha_forward = G.test_multihead_attention_forward_matches_reference
test_mha_forward_bf16 = G.test_multihead_attention_bf16_fused_matches_reference
test_encoder_fused_B1 = G.test_encoder_layer_fused_matches_reference_B1
test_encoder_faithful_B4 = G.test_encoder_layer_faithful_mode_matches_reference_B4
test_c4_faithful = G.test_config_c4_transformer_training_faithful_mode_matches_reference
test_c4_default = G.test_config_c4_transformer_training_default_mode_runs_and_learns
test_gpt_shape_bf16_vs_fp32 = G.test_gpt_shape_train_step_bf16_vs_fp32
test_operand_cache_and_lazy_zero = G.test_operand_cache_and_lazy_zero_change_nothing
test_deferred_gradients = G.test_deferred_gradients_change_nothing
test_cow_gradients = G.test_cow_gradients_change_nothing
test_mha_kv_cache_decode = G.test_mha_kv_cache_decode_matches_reference
test_mha_kv_cache_prefill = G.test_mha_kv_cache_prefill_then_decode_fused_equals_unfused
test_greedy_decode = G.test_greedy_decode_matches_reference
test_greedy_decode_batched = G.test_greedy_decode_batched_fused_equals_unfused
test_view_keeps_owner = G.test_view_of_graph_tensor_keeps_owner_alive
test_token_model_fp64_arbiter = G.test_token_model_fused_fp32_matches_fp64_arbiter_B3
test_token_model_bf16_arbiter = G.test_token_model_fused_bf16_matches_fp64_bf16_arbiter_B8
test_encoder_fp64_arbiter = G.test_encoder_layer_fused_matches_fp64_arbiter_B4
test_c4_scaled_B1 = G.test_config_c4_scaled_dims_B1_matches_reference
test_gemm_epilogue_fusions = G.test_gemm_epilogue_fusions_match_unfused_passes
test_checkpoint_round_trip = G.test_checkpoint_round_trip_with_reference
test_checkpoint_families = G.test_checkpoint_of_every_module_family_round_trips
test_shared_c_api = G.test_shared_c_api_load_forward_train_step
test_f4_modules = G.test_f4_module_families_match_reference


def test_training_steps_do_not_leak_device_buffers(P):
    """Every buffer allocated during a training step is released once the step's loss handle is
    dropped. (The reference's closures capture their own output tensor strongly, tensor.cpp:418-430,
    a shared_ptr cycle that keeps every graph alive; here the node captures it weakly.)"""
    import ctypes as C

    import numpy as np

    import bench
    mock = C.CDLL(os.path.join(MOCK_DIR, "libweedcu_mock.so"))

    def outstanding():
        m, f = C.c_uint64(), C.c_uint64()
        mock.weedcu_host_stats(None, C.byref(m), None, C.byref(f))
        return m.value - f.value

    for fused in (1, 0):
        G.set_mode(P, fused)
        cfg = dict(V=40, d=16, H=2, dff=32, L=2, T=8, B=2)
        model, _ = bench.build_model(P, cfg)
        opt = P.adam(model, 1e-3)
        tok, tgt = bench.make_tokens(cfg, 1)
        st, sg = P.symbol(tok, [2, 8]), P.symbol(tgt, [2, 8])
        for _ in range(3):
            P.free(P.train_step_tokens(model, opt, st, sg))
        base = outstanding()
        for _ in range(5):
            P.free(P.train_step_tokens(model, opt, st, sg))
        assert outstanding() == base, f"fused={fused}: {outstanding() - base} device buffers leaked over 5 steps"
        P.reset()
    G.set_mode(P, 1)


def test_reference_catch_suite_passes_on_the_mock_device():
    """The reference's own Catch suite (real-dtype TEST_CASEs, filtered at build time by tools/ref_tests/filter_tests.py)
    against the product's host library on the oracle-backed mock device."""
    if not os.path.exists("/root/reference/test/tests.cpp"):
        pytest.skip("reference sources not present")
    subprocess.check_call(["make", "-C", MOCK_DIR, "ref_unittest_mock"], stdout=subprocess.DEVNULL)
    res = subprocess.run([os.path.join(MOCK_DIR, "ref_unittest_mock"), "--device-gpu"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                         timeout=600)
    assert res.returncode == 0, res.stdout[-3000:]
    assert res.stdout.count("All tests passed") == 2, res.stdout[-3000:]
