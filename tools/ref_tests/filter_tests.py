#!/usr/bin/env python
"""Build-time filter for running the reference's OWN Catch suite on this backend.

Reads /root/reference/test/tests.cpp where it lies and writes a translation unit with only the TEST_CASEs this backend's
scope covers (real dtype, dense storage — SURVEY §8) into a git-ignored build directory (oracle/_ref/): nothing of the
reference is copied into the repository. The exclusion rule is by TEST_CASE name: complex / mixed dtype and sparse
storage are out of scope.

    python tools/ref_tests/filter_tests.py /root/reference/test/tests.cpp oracle/_ref/ref_tests_real.cpp
"""
import re
import sys

OUT_OF_SCOPE = ("complex", "mixed", "sparse")


def split_cases(src):
    """(preamble, [(name, text)]) — TEST_CASE blocks found by brace matching"""
    out, pos, pre_end = [], 0, None
    for m in re.finditer(r'TEST_CASE\("([^"]+)"\)\s*\{', src):
        if m.start() < pos:
            continue
        if pre_end is None:
            pre_end = m.start()
        depth, i = 1, m.end()
        while depth:
            c = src[i]
            depth += (c == "{") - (c == "}")
            i += 1
        out.append((m.group(1), src[m.start():i]))
        pos = i
    return src[:pre_end], out


def main():
    src = open(sys.argv[1]).read()
    pre, cases = split_cases(src)
    # by name, and by body: a few real-dtype cases also exercise a complex or sparse tensor inside
    kept = [(n, t) for n, t in cases if not any(k in n for k in OUT_OF_SCOPE) and not re.search(r"Complex|COMPLEX|Sparse|\bC\(", t)]
    pre = "\n".join(l for l in pre.splitlines() if "complex_scalar.hpp" not in l) + "\n"
    with open(sys.argv[2], "w") as f:
        f.write("// GENERATED at build time by tools/ref_tests/filter_tests.py from the reference's test/tests.cpp — not tracked.\n")
        f.write(pre)
        for _n, t in kept:
            f.write(t + "\n\n")
    print(f"kept {len(kept)} of {len(cases)} TEST_CASEs: " + " ".join(n for n, _ in kept))


if __name__ == "__main__":
    main()
