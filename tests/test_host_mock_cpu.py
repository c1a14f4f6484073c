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
test_encoder_fused_B1 = G.test_encoder_layer_fused_matches_reference_B1
test_encoder_faithful_B4 = G.test_encoder_layer_faithful_mode_matches_reference_B4
test_c4_faithful = G.test_config_c4_transformer_training_faithful_mode_matches_reference
test_c4_default = G.test_config_c4_transformer_training_default_mode_runs_and_learns
test_gpt_shape_bf16_vs_fp32 = G.test_gpt_shape_train_step_bf16_vs_fp32
