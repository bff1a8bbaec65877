"""File formats of the drop-in layer (gp-ann_b200/cxx/standalone/points_io.cpp = src/points_io.cpp:63-179) and
ComputeGroundTruthFromFile (cxx/ground_truth_file.h: .u8bin / .i8bin rows go to the GPU as bytes, chunk by chunk)."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

from conftest import ROOT, golden
from helpers import assert_topk_close, assert_topk_exact

_f = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_u32 = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_sz = C.c_size_t


def _lib():
    lib = C.CDLL(os.path.join(ROOT, "gp-ann_b200", "libgpann_dropin.so"))
    lib.dropin_read_points.argtypes = [C.c_char_p, _f, _sz, C.POINTER(_sz), C.POINTER(_sz)]
    lib.dropin_write_ground_truth.argtypes = [C.c_char_p, _f, _u32, _sz, _sz]
    lib.dropin_ground_truth_file.argtypes = [C.c_char_p, _f, _sz, _sz, C.c_int, _sz, C.c_char_p, _f, _u32]
    return lib


def _write_bin(path, a):
    with open(path, "wb") as f:
        f.write(struct.pack("<II", a.shape[0], a.shape[1]))
        f.write(np.ascontiguousarray(a).tobytes())


def test_point_and_ground_truth_file_formats(tmp_path):
    lib = _lib()
    rng = np.random.default_rng(3)
    for suffix, arr in ((".u8bin", rng.integers(0, 256, (37, 20), dtype=np.uint8)), (".i8bin", rng.integers(-128, 128, (11, 7), dtype=np.int8)),
                        (".fbin", rng.standard_normal((13, 5)).astype(np.float32))):
        p = str(tmp_path / ("pts" + suffix))
        _write_bin(p, arr)
        out = np.zeros(arr.size, np.float32)
        n, d = _sz(0), _sz(0)
        assert lib.dropin_read_points(p.encode(), out, out.size, C.byref(n), C.byref(d)) == 0
        assert (n.value, d.value) == arr.shape
        assert np.array_equal(out.reshape(arr.shape), arr.astype(np.float32))  # bytes widened like src/points_io.cpp:100-104
    dist = rng.standard_normal((6, 4)).astype(np.float32)
    ids = rng.integers(0, 2**32, (6, 4), dtype=np.uint64).astype(np.uint32)
    p = str(tmp_path / "gt.bin")
    assert lib.dropin_write_ground_truth(p.encode(), dist, ids, 6, 4) == 0
    raw = open(p, "rb").read()
    # src/points_io.cpp:157-179: u32 nq, u32 k, all ids, then all distances
    assert raw == struct.pack("<II", 6, 4) + ids.tobytes() + dist.tobytes()


@pytest.mark.gpu
def test_ground_truth_from_file(tmp_path):
    lib = _lib()
    g = golden("gt_u8_d128_k10")
    p = str(tmp_path / "base.u8bin")
    _write_bin(p, g["points"])
    Q = np.ascontiguousarray(g["queries"], np.float32)
    d = np.zeros((len(Q), 10), np.float32)
    i = np.zeros((len(Q), 10), np.uint32)
    gt = str(tmp_path / "gt.bin")
    # 700-row file chunks: five streamed calls, merged lexicographically
    assert lib.dropin_ground_truth_file(p.encode(), Q, len(Q), Q.shape[1], 10, 700, gt.encode(), d, i) == 0
    assert_topk_exact(d, i, g["dist"], g["ids"])
    g = golden("gt_s8_d100_k10")
    p = str(tmp_path / "base.i8bin")
    _write_bin(p, g["points"])
    Q = np.ascontiguousarray(g["queries"], np.float32)
    d = np.zeros((len(Q), 10), np.float32)
    i = np.zeros((len(Q), 10), np.uint32)
    assert lib.dropin_ground_truth_file(p.encode(), Q, len(Q), Q.shape[1], 10, 0, gt.encode(), d, i) == 0
    assert_topk_exact(d, i, g["dist"], g["ids"])
    g = golden("gt_f32_d100_k10")
    p = str(tmp_path / "base.fbin")
    _write_bin(p, g["points"])
    Q = np.ascontiguousarray(g["queries"], np.float32)
    d = np.zeros((len(Q), 10), np.float32)
    i = np.zeros((len(Q), 10), np.uint32)
    assert lib.dropin_ground_truth_file(p.encode(), Q, len(Q), Q.shape[1], 10, 500, gt.encode(), d, i) == 0
    assert_topk_close(d, i, g["dist"], g["ids"])
