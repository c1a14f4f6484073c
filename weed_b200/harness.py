"""ctypes driver for the handle-based C-ABI of harness/weed_harness.cpp.

The same class drives both builds of that one client source:
  * weed_b200/libweed_b200_harness.so  — this repo's host library on the CUDA device (the product)
  * oracle/_ref/libweed_ref_harness.so — the unmodified reference CPU build (tests/bench oracle only)
There is no fallback between them: `Harness.product()` raises if the CUDA build is missing.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
CPU, GPU = 2, 3


class HarnessError(RuntimeError):
    pass


class Harness:
    def __init__(self, path, device_tag):
        if not os.path.exists(path):
            raise HarnessError(f"{path} is missing (build it first); no fallback is provided")
        self.path = path
        self.lib = C.CDLL(path)  # RTLD_LOCAL: the two builds export the same Weed:: symbols
        L = self.lib
        L.wh_last_error.restype = C.c_char_p
        L.wh_backend.restype = C.c_char_p
        L.wh_stream.restype = C.c_void_p
        L.wh_wall_seconds.restype = C.c_double
        for f in ("wh_tensor", "wh_scalar", "wh_symbol", "wh_view", "wh_grad", "wh_op", "wh_cross_entropy", "wh_module",
                  "wh_param", "wh_param_size", "wh_forward", "wh_forward_symbol", "wh_adam", "wh_train_step_tokens"):
            getattr(L, f).restype = C.c_int64
        self.device_tag = device_tag
        self._ck(L.wh_init(C.c_int(device_tag)))

    # ---------------------------------------------------------------- construction helpers
    @classmethod
    def product(cls):
        """This repo's host library on the GPU (libweed_b200_harness.so -> libweed_b200.so ->
        libweedcu.so, resolved through $ORIGIN rpaths). weedcu() is loaded first so that a missing
        CUDA extension fails with the explicit no-fallback error."""
        from ._lib import weedcu
        weedcu()
        return cls(os.path.join(_HERE, "libweed_b200_harness.so"), GPU)

    @classmethod
    def reference(cls):
        """The unmodified reference, CPU device (oracle; tests and bench baseline only)."""
        return cls(os.path.join(ROOT, "oracle", "_ref", "libweed_ref_harness.so"), CPU)

    def backend(self):
        return self.lib.wh_backend().decode()

    def _ck(self, rc):
        if rc is None or rc < 0:
            raise HarnessError(f"[{self.backend()}] {self.lib.wh_last_error().decode()}")
        return rc

    # ---------------------------------------------------------------- tensors
    def tensor(self, values, shape, requires_grad=False):
        a = np.ascontiguousarray(values, dtype=np.float32).ravel()
        sh = (C.c_uint32 * len(shape))(*shape)
        return self._ck(self.lib.wh_tensor(a.ctypes.data_as(C.c_void_p), C.c_uint32(a.size), C.c_int(len(shape)), sh,
                                           C.c_int(int(requires_grad))))

    def scalar(self, value, requires_grad=False):
        return self._ck(self.lib.wh_scalar(C.c_float(value), C.c_int(int(requires_grad))))

    def symbol(self, values, shape):
        a = np.ascontiguousarray(values, dtype=np.int32).ravel()
        sh = (C.c_uint32 * len(shape))(*shape)
        return self._ck(self.lib.wh_symbol(a.ctypes.data_as(C.c_void_p), C.c_uint32(a.size), C.c_int(len(shape)), sh))

    def symbol_upload(self, h, src_ptr, n):
        """Refill symbol tensor `h` from host memory at address src_ptr (int32[n]); async H2D."""
        self._ck(self.lib.wh_symbol_upload(C.c_int64(h), C.c_void_p(src_ptr), C.c_uint32(n)))

    def view(self, h, offset, shape, stride):
        sh = (C.c_uint32 * len(shape))(*shape)
        st = (C.c_uint32 * len(stride))(*stride)
        return self._ck(self.lib.wh_view(C.c_int64(h), C.c_uint32(offset), C.c_int(len(shape)), sh, st))

    def info(self, h):
        rank, off, ssz, rg = C.c_int(), C.c_uint32(), C.c_uint32(), C.c_int()
        shape, stride = (C.c_uint32 * 8)(), (C.c_uint32 * 8)()
        self._ck(self.lib.wh_info(C.c_int64(h), C.byref(rank), shape, stride, C.byref(off), C.byref(ssz), C.byref(rg)))
        r = rank.value
        return {"shape": list(shape[:r]), "stride": list(stride[:r]), "offset": off.value, "storage_size": ssz.value,
                "requires_grad": bool(rg.value)}

    def read(self, h):
        """Logical values in flat column-major order through the view."""
        inf = self.info(h)
        n = int(np.prod(inf["shape"])) if inf["shape"] else 0
        out = np.zeros(max(n, 1), np.float32)
        cnt = C.c_uint32()
        self._ck(self.lib.wh_read(C.c_int64(h), out.ctypes.data_as(C.c_void_p), C.c_uint32(out.size), C.byref(cnt)))
        return out[:cnt.value].copy()

    def read_storage(self, h):
        inf = self.info(h)
        out = np.zeros(inf["storage_size"], np.float32)
        cnt = C.c_uint32()
        self._ck(self.lib.wh_read_storage(C.c_int64(h), out.ctypes.data_as(C.c_void_p), C.c_uint32(out.size), C.byref(cnt)))
        return out[:cnt.value].copy()

    def grad(self, h):
        g = self._ck(self.lib.wh_grad(C.c_int64(h)))
        return g if g else None

    def backward(self, h):
        self._ck(self.lib.wh_backward(C.c_int64(h)))

    def op(self, name, ins, floats=(), ints=()):
        hi = (C.c_int64 * max(len(ins), 1))(*ins)
        fl = (C.c_float * max(len(floats), 1))(*floats)
        iv = (C.c_int32 * max(len(ints), 1))(*ints)
        return self._ck(self.lib.wh_op(name.encode(), hi, C.c_int(len(ins)), fl, C.c_int(len(floats)), iv, C.c_int(len(ints))))

    def cross_entropy(self, logits, targets):
        return self._ck(self.lib.wh_cross_entropy(C.c_int64(logits), C.c_int64(targets)))

    def free(self, h):
        self.lib.wh_free(C.c_int64(h))

    def reset(self):
        self.lib.wh_reset()

    def mark(self):
        self.lib.wh_mark.restype = C.c_int64
        return self.lib.wh_mark()

    def release_since(self, mark):
        """Drop every tensor / symbol / module / optimiser handle created since mark()."""
        self.lib.wh_release_since(C.c_int64(mark))

    def config(self, name, value):
        self._ck(self.lib.wh_config(name.encode(), C.c_double(value)))

    def sync(self):
        self._ck(self.lib.wh_sync())

    def stream(self):
        return self.lib.wh_stream()

    # ---------------------------------------------------------------- modules / optimisers
    def module(self, kind, *args):
        a = (C.c_int64 * max(len(args), 1))(*[int(x) for x in args])
        return self._ck(self.lib.wh_module(kind.encode(), a, C.c_int(len(args))))

    def module_save(self, m, path):
        self._ck(self.lib.wh_module_save(C.c_int64(m), path.encode()))

    def module_load(self, path):
        self.lib.wh_module_load.restype = C.c_int64
        return self._ck(self.lib.wh_module_load(path.encode()))

    def module_set(self, m, field, value):
        self._ck(self.lib.wh_module_set(C.c_int64(m), field.encode(), C.c_int64(int(value))))

    def param_count(self, m):
        return self._ck(self.lib.wh_param_count(C.c_int64(m)))

    def param_size(self, m, i):
        return self._ck(self.lib.wh_param_size(C.c_int64(m), C.c_int(i)))

    def param(self, m, i):
        return self._ck(self.lib.wh_param(C.c_int64(m), C.c_int(i)))

    def param_set(self, m, i, values):
        a = np.ascontiguousarray(values, dtype=np.float32).ravel()
        self._ck(self.lib.wh_param_set(C.c_int64(m), C.c_int(i), a.ctypes.data_as(C.c_void_p), C.c_uint32(a.size)))

    def forward(self, m, x):
        return self._ck(self.lib.wh_forward(C.c_int64(m), C.c_int64(x)))

    def forward_symbol(self, m, s):
        return self._ck(self.lib.wh_forward_symbol(C.c_int64(m), C.c_int64(s)))

    def argmax_last(self, logits):
        """Arg-max token of the last position of logits [B, T, V] -> symbol handle [B, 1]."""
        self.lib.wh_argmax_last.restype = C.c_int64
        return self._ck(self.lib.wh_argmax_last(C.c_int64(logits)))

    def read_symbol(self, h, n):
        out = np.zeros(n, np.int32)
        self.lib.wh_read_symbol.restype = C.c_int
        self._ck(self.lib.wh_read_symbol(C.c_int64(h), out.ctypes.data_as(C.POINTER(C.c_int32)), C.c_uint32(n)))
        return out

    def squeeze(self, h, axis):
        self._ck(self.lib.wh_squeeze(C.c_int64(h), C.c_int(axis)))

    def adam(self, module, lr, b1=0.9, b2=0.999, eps=1e-8):
        return self._ck(self.lib.wh_adam(C.c_float(lr), C.c_float(b1), C.c_float(b2), C.c_float(eps), C.c_int64(module)))

    def adam_moment(self, opt, module, i, which=0):
        """Adam's first (which=0) / second (which=1) moment of parameter i as a tensor handle."""
        self.lib.wh_adam_moment.restype = C.c_int64
        return self._ck(self.lib.wh_adam_moment(C.c_int64(opt), C.c_int64(module), C.c_int(i), C.c_int(which)))

    def adam_step(self, opt, module):
        self._ck(self.lib.wh_adam_step(C.c_int64(opt), C.c_int64(module)))

    def sgd_step(self, module, lr):
        self._ck(self.lib.wh_sgd_step(C.c_int64(module), C.c_float(lr)))

    def zero_grad(self, module):
        self._ck(self.lib.wh_zero_grad(C.c_int64(module)))

    def train_step_tokens(self, model, opt, tokens, targets):
        return self._ck(self.lib.wh_train_step_tokens(C.c_int64(model), C.c_int64(opt), C.c_int64(tokens), C.c_int64(targets)))

    # ---------------------------------------------------------------- seeded weight injection
    def init_params(self, m, seed, scheme="uniform"):
        """Deterministic weights (never std::random_device): parameter i gets
        uniform(-lim, lim) with lim = sqrt(6/(fan_in+fan_out)) proxy sqrt(3/size^0.5); 1-D params
        (biases / beta) small uniform, gamma-like params left to the caller."""
        rng = np.random.default_rng(seed)
        out = []
        for i in range(self.param_count(m)):
            n = self.param_size(m, i)
            lim = float(np.sqrt(6.0 / (2.0 * np.sqrt(n)))) if n > 1 else 0.1
            w = rng.uniform(-lim, lim, size=n).astype(np.float32)
            self.param_set(m, i, w)
            out.append(w)
        return out
