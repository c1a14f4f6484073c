"""Generate tests/golden/ref_pins.npz from the UNMODIFIED reference (run in the build container,
where /root/reference exists and `make -C oracle ref` has produced oracle/_ref/).

    python tests/golden/make_golden.py

For every pin in tests/refpins.py it stores the seeded inputs and the outputs of the compiled
reference CPU path, so tests/test_oracle_cpu.py can check the C oracle against real reference
results on machines that have neither /root/reference nor oracle/_ref (e.g. a fresh clone).
"""
import inspect
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import refpins  # noqa: E402
from weed_b200.harness import Harness  # noqa: E402


def seed_of(name):
    return sum(ord(ch) * (i + 1) for i, ch in enumerate(name)) % (2**31)


def main():
    R = Harness.reference()
    store = {}
    for name, fn, _tol in refpins.PINS:
        inp, ref, _orc = fn(np.random.default_rng(seed_of(name)))
        out = ref(R, inp)
        R.reset()
        for k, v in inp.items():
            store[f"{name}/in/{k}"] = np.asarray(v)
        for k, v in out.items():
            store[f"{name}/out/{k}"] = np.asarray(v)
    path = os.path.join(HERE, "ref_pins.npz")
    np.savez_compressed(path, **store)
    print(f"wrote {path}: {len(refpins.PINS)} pins, {len(store)} arrays, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
