"""ctypes binding of include/weedcu.h. Fails loudly when the CUDA extension is not built."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
MAX_RANK = 8


class WeedcuError(RuntimeError):
    pass


class View(C.Structure):
    """weedcu_view: (offset, rank, shape[8], stride[8]) — BaseTensor's view triple
    (reference include/tensors/base_tensor.hpp:25-30)."""
    _fields_ = [("offset", C.c_uint64), ("rank", C.c_int32),
                ("shape", C.c_uint32 * MAX_RANK), ("stride", C.c_uint32 * MAX_RANK)]


class Mat(C.Structure):
    """weedcu_mat: (offset, s0, s1, batch_stride) — MatrixDim of reference src/ops/matmul.cpp:87-122."""
    _fields_ = [("offset", C.c_uint64), ("s0", C.c_uint32), ("s1", C.c_uint32),
                ("batch_stride", C.c_uint64)]


class GemmEpilogue(C.Structure):
    """weedcu_gemm_epilogue (include/weedcu.h): the extended epilogue of weedcu_gemm_bf16_ex"""
    _fields_ = [("col_bias", C.c_void_p), ("residual", C.c_void_p), ("ldr", C.c_uint64), ("activation", C.c_int), ("row_stats", C.c_int),
                ("stats", C.c_void_p), ("stats_capacity_tiles", C.c_uint32), ("stats_tiles", C.POINTER(C.c_uint32)),
                ("stats_tile_cols", C.POINTER(C.c_uint32))]


def make_view(shape, stride, offset=0):
    v = View()
    v.offset = int(offset)
    v.rank = len(shape)
    for d in range(MAX_RANK):
        v.shape[d] = int(shape[d]) if d < len(shape) else 1
        v.stride[d] = int(stride[d]) if d < len(stride) else 0
    return v


def contiguous_stride(shape):
    """BaseTensor::full_contiguous_stride (reference base_tensor.hpp:282-292): column-major,
    extent-1 dims get stride 0."""
    st, acc = [], 1
    for s in shape:
        st.append(0 if s == 1 else acc)
        acc *= s
    return st


def contiguous_view(shape, offset=0):
    return make_view(shape, contiguous_stride(shape), offset)


_lib = None


def weedcu():
    """Load weed_b200/libweedcu.so (built by `make -C weed_b200/csrc` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.path.join(_HERE, "libweedcu.so")
    if not os.path.exists(path):
        raise WeedcuError(
            f"{path} is missing: the CUDA extension is not built. There is no CPU fallback; "
            "run `python -c 'import __graft_entry__ as g; g.build()'` first.")
    lib = C.CDLL(path)  # RTLD_LOCAL: dependants find it through their own DT_NEEDED + $ORIGIN rpath
    lib.weedcu_error_string.restype = C.c_char_p
    lib.weedcu_default_stream.restype = C.c_void_p
    _lib = lib
    return lib


def check(rc, what="weedcu call"):
    if rc != 0:
        msg = weedcu().weedcu_error_string(C.c_int(rc)).decode()
        raise WeedcuError(f"{what} failed with code {rc}: {msg}")


# every symbol include/weedcu.h declares (kept in sync by tests/test_abi_cpu.py)
SYMBOLS = """
weedcu_device_count weedcu_set_device weedcu_get_device weedcu_device_info weedcu_error_string
weedcu_default_stream weedcu_set_default_stream weedcu_stream_create weedcu_stream_create_priority weedcu_stream_destroy
weedcu_stream_sync weedcu_stream_wait_event weedcu_event_create weedcu_event_destroy
weedcu_event_record weedcu_event_sync weedcu_event_elapsed_ms weedcu_malloc weedcu_free weedcu_pool_trim
weedcu_mem_info weedcu_host_alloc weedcu_host_free weedcu_memcpy_h2d weedcu_memcpy_d2h
weedcu_memcpy_d2d weedcu_launch_count weedcu_set_pdl weedcu_host_stats weedcu_fill_real weedcu_fill_int weedcu_binary_real
weedcu_inplace_real weedcu_copy_real weedcu_unary_real weedcu_unary_grad_real weedcu_reduce_real
weedcu_reduce_grad_real weedcu_sum_real weedcu_clamp_real weedcu_clamp_grad_real weedcu_extremum_real weedcu_match_grad_full_real weedcu_extremum_axis_real weedcu_match_grad_real weedcu_softmax_real weedcu_softmax_grad_real
weedcu_attn_softmax_real weedcu_attention_fwd weedcu_attention_fwd_bf16out weedcu_attention_fwd_bf16in weedcu_attention_decode weedcu_matmul_skinny weedcu_matmul_skinny_grouped weedcu_matmul_skinny_residual weedcu_cross_entropy_fwd weedcu_cross_entropy_bwd weedcu_cross_entropy_bwd_pack weedcu_cross_entropy_fwd_stats weedcu_cross_entropy_fwd_bf16in weedcu_cross_entropy_bwd_pack_bf16in weedcu_layernorm_fwd weedcu_layernorm_fwd_bf16 weedcu_layernorm_fwd_stats weedcu_gelu_fwd_bf16 weedcu_gelu_grad_pack weedcu_gelu_grad_pack_bf16dy weedcu_multi_copy
weedcu_layernorm_bwd weedcu_layernorm_bwd_from weedcu_embedding_gather weedcu_embedding_scatter_add weedcu_triu_fill_real
weedcu_argmax_rows weedcu_sgd_step weedcu_adam_step weedcu_adam_step_multi weedcu_adam_step_multi_shadow weedcu_adam_step_multi_zero weedcu_matmul_real weedcu_gemm_bf16 weedcu_gemm_bf16_grouped weedcu_gemm_bf16_residual weedcu_gemm_bf16_ex weedcu_gemm_bf16_grouped_bf16out weedcu_gemm_set_mode weedcu_gemm_set_dynamic
weedcu_pack_bf16 weedcu_pack_bf16_colsum weedcu_gemm_workspace_bytes weedcu_prof_enable weedcu_prof_read weedcu_nccl_load weedcu_nccl_unique_id
weedcu_nccl_init weedcu_nccl_destroy weedcu_nccl_group_start weedcu_nccl_group_end weedcu_nccl_allreduce_sum
weedcu_nccl_broadcast
""".split()
