"""slice_key of the collapse kernels (mirge3.0_b200/csrc/key_format.cuh: a first-level key without its UMI flanks, i.e.
the centre of ``UMIParser`` in mirge/libs/digest.py:305-315 used by the second collapse level, digest.py:164-205)
compiled for the host and held against string slicing: every (f, b) around the key length, keys with N / lower-case
exceptions inside and outside the flanks, lengths around the 16-base word boundaries."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests.util import host_harness_flags
import mirge_b200
from tests.test_annotate_verify_host import pack_key

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hk(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("hk") / "libkey_format_host.so")
    subprocess.check_call(["g++"] + host_harness_flags() + ["-std=c++17", "-shared", "-fPIC", "-I", os.path.join(mirge_b200.PACKAGE_DIR, "csrc"),
                           "-o", so, os.path.join(HERE, "key_format_harness.cpp")])
    lib = C.CDLL(so)
    lib.hk_slice.restype = C.c_uint32
    lib.hk_slice.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    return lib


def test_slice_key_equals_string_slicing(hk):
    rng = np.random.default_rng(12)
    out = np.zeros(1 + 40 + 600, dtype=np.uint32)
    n = 0
    for L in list(range(0, 70)) + [95, 96, 97, 160]:
        for rep in range(6):
            s = list("".join(rng.choice(np.array(list("ACGT")), L)))
            for _ in range(int(rng.integers(0, 4)) if L else 0):
                j = int(rng.integers(L))
                s[j] = "N" if rng.random() < 0.5 else s[j].lower()
            s = "".join(s)
            key = pack_key(s)
            for f, b in [(0, 0), (4, 4), (0, 12), (3, 0), (L, 0), (0, L), (L // 2, L - L // 2), (L, L), (int(rng.integers(0, 20)), int(rng.integers(0, 20)))]:
                cl = max(L - f - b, 0)
                centre = s[f : f + cl] if cl > 0 else ""   # what UMIParser's s[f:-b] / s[f:] leaves
                exp = pack_key(centre)
                nw = hk.hk_slice(key.ctypes.data, f, b, out.ctypes.data)
                assert nw == len(exp) and np.array_equal(out[:nw], exp), (s, f, b, centre)
                n += 1
    assert n > 3000
