"""Two interchangeable executors for the same test cases.

OracleBackend : oracle/liboracle.so (CPU restatement of the reference; TEST INFRASTRUCTURE) on numpy
GpuBackend    : weed_b200/libweedcu.so through the C-ABI of include/weedcu.h on torch CUDA buffers

Both expose  buf(np_array) -> handle(.ptr, .get()),  call(name, *args)  where `name` is the suffix
shared by wo_<name> / weedcu_<name>; the GPU call appends the stream argument.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")


def _ctype_args(args):
    out = []
    for a in args:
        if isinstance(a, (np.floating, float)):
            out.append(C.c_float(float(a)))
        elif isinstance(a, Handle):
            out.append(C.c_void_p(a.ptr))
        elif isinstance(a, (C.Structure,)):
            out.append(C.byref(a))
        elif a is None:
            out.append(C.c_void_p(0))
        else:
            out.append(a)
    return out


class Handle:
    def __init__(self, ptr, getter, keep):
        self.ptr = ptr
        self._get = getter
        self._keep = keep

    def get(self):
        return self._get()


class OracleBackend:
    name = "oracle"

    def __init__(self):
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "oracle"], stdout=subprocess.DEVNULL)
        self.lib = C.CDLL(path)

    def buf(self, arr):
        a = np.ascontiguousarray(arr).copy()
        return Handle(a.ctypes.data, lambda: a.copy(), a)

    def call(self, name, *args):
        fn = getattr(self.lib, "wo_" + name)
        fn.restype = C.c_int
        rc = fn(*_ctype_args(args))
        assert rc == 0, f"oracle wo_{name} returned {rc}"

    def sync(self):
        pass


class GpuBackend:
    name = "gpu"

    def __init__(self):
        import torch
        from weed_b200 import weedcu, check
        assert torch.cuda.is_available(), "GPU tests need a CUDA device"
        self.torch = torch
        self.lib = weedcu()  # raises if the extension is not built: no fallback
        self.check = check
        # A real (non-legacy) stream: handle 0 would mean "library default stream" to weedcu, which
        # is not ordered against torch's copies.
        self.tstream = torch.cuda.Stream()
        self.stream = self.tstream.cuda_stream
        assert self.stream != 0

    def buf(self, arr):
        torch = self.torch
        with torch.cuda.stream(self.tstream):
            t = torch.from_numpy(np.ascontiguousarray(arr).copy()).cuda()

        def get():
            with torch.cuda.stream(self.tstream):
                out = t.cpu()
            self.tstream.synchronize()
            return out.numpy().copy()
        return Handle(t.data_ptr(), get, t)

    def call(self, name, *args):
        fn = getattr(self.lib, "weedcu_" + name)
        fn.restype = C.c_int
        rc = fn(*_ctype_args(args), C.c_void_p(self.stream))
        self.check(rc, "weedcu_" + name)

    def sync(self):
        self.torch.cuda.synchronize()
