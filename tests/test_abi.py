"""CPU tests of the drop-in boundary: libgpann_cuda.so loads without a GPU, exports every symbol include/gpann.h
declares (and the ctypes table covers exactly those), the host-side key helpers order like
std::pair<float,uint32_t>, and gpann_init fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "gpann.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gpann_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(gpann_mod):
    import subprocess
    subprocess.check_call(["python", os.path.join(ROOT, "gp-ann_b200", "build.py")])
    lib = ctypes.CDLL(gpann_mod.lib_path())
    declared = _declared_symbols()
    assert len(declared) >= 24
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/gpann.h but not exported"
    from gpann.lib import SYMBOLS
    assert sorted(SYMBOLS) == declared


def test_key_order_matches_pair_lexicographic(gpann_mod):
    rng = np.random.default_rng(3)
    d = np.concatenate([rng.standard_normal(200).astype(np.float32) * 100, [0.0, -0.0, 1.0, 1.0, np.inf, 8323200.0]])
    d = d.astype(np.float32)
    ids = rng.integers(0, 2**32, len(d), dtype=np.uint64).astype(np.uint32)
    keys = np.array([gpann_mod.pack_key(x, i) for x, i in zip(d, ids)], dtype=np.uint64)
    order_keys = np.argsort(keys, kind="stable")
    dd = np.where(d == 0, 0.0, d)  # -0.0 < +0.0 in key order; as floats they tie -- compare through the key's own float
    back_d, back_i = gpann_mod.unpack_keys(keys)
    assert np.array_equal(back_i, ids)
    assert np.array_equal(back_d.view(np.uint32), d.view(np.uint32))
    pairs = sorted(range(len(d)), key=lambda j: (float(dd[j]), int(ids[j])))
    nz = [j for j in pairs if d[j] != 0]
    assert [j for j in order_keys if d[j] != 0] == nz
    # the C helper and the numpy helper agree
    lib = gpann_mod.load_library()
    fd, fi = ctypes.c_float(), ctypes.c_uint32()
    lib.gpann_unpack_key(int(keys[5]), ctypes.byref(fd), ctypes.byref(fi))
    assert fd.value == d[5] and fi.value == ids[5]
    assert gpann_mod.EMPTY_KEY > int(keys.max())


def test_no_cpu_fallback(gpann_mod):
    """Without a CUDA device the product path must fail, not compute on the host."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present; covered by the gpu tests")
    with pytest.raises(gpann_mod.GpannError):
        gpann_mod.Context(0)
    # and the package holds no reference to the oracle
    pkg = os.path.join(ROOT, "gp-ann_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                # (OracleRecall is a reference function name in recall.h; what must not appear is any use of oracle/)
                for needle in ("from oracle", "import oracle", "oracle/", "libgpann_oracle", "libgpann_ref", "oracle."):
                    assert needle not in text, f"{f} references the test oracle ({needle})"
