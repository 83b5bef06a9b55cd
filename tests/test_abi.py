"""CPU: the C-ABI library loads, exports what include/gmdist.h declares, and the host
logic around it (policy, argument checks, index helpers) behaves.  No compute calls."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from groopm_b200 import _lib, dist, engine

ROOT = Path(__file__).resolve().parents[1]


def header_symbols():
    text = (ROOT / "include" / "gmdist.h").read_text()
    return re.findall(r"GMD_API\s+[\w\s\*]+?\b(gmd_\w+)\s*\(", text)


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    names = header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in gmdist.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), "ctypes signatures out of sync with gmdist.h"


def test_header_cites_reference_call_sites():
    text = (ROOT / "include" / "gmdist.h").read_text()
    for cite in ("torusMesh.py:123-130", "refine.py:653-665", "cluster.py:495", "refine.py:1502-1503",
                 "som.py:348-353", "cluster.py:533-536"):
        assert cite in text


def test_policy_helpers():
    lib = _lib.lib()
    assert lib.gmd_version() >= 100
    assert lib.gmd_pick_form(_lib.L1, 30) == _lib.FORM_DIRECT_L1
    assert lib.gmd_pick_form(_lib.L2, 3) == _lib.FORM_DIRECT_L2
    assert lib.gmd_pick_form(_lib.L2, 30) == _lib.FORM_EXPANDED
    for k in (1, 6, 11, 16, 32, 64):
        kk = lib.gmd_pick_kk(k)
        assert k + 3 <= kk <= k + 8
    assert lib.gmd_tiles_bytes(64, 30) == 31 * 64 * 4
    assert lib.gmd_tiles_bytes(65, 30) == 2 * 31 * 64 * 4
    assert lib.gmd_tiles_bytes(10, 3) == 5 * 64 * 4          # D = 3 is padded to 4
    assert lib.gmd_tiles_bytes(10, 137) == 0                 # beyond GMD_MAX_D
    assert 1 <= lib.gmd_pick_nsplit(100_000, 100_000, 10, 20) <= 16
    assert lib.gmd_pick_nsplit(1_000_000, 4096, 4, 4) <= 4


def test_argument_errors_map_to_reference_exceptions():
    lib = _lib.lib()
    x = np.zeros((4, 3))
    out_i = np.zeros((4, 9), dtype=np.int32)
    rc = lib.gmd_knn_host_f64(x.ctypes.data, 4, x.ctypes.data, 4, 3, 9, 0, out_i.ctypes.data, None)
    assert rc == _lib.ERR_BAD_ARG
    with pytest.raises(ValueError, match="exceeds"):
        _lib.check(rc)
    rc = lib.gmd_knn_host_f64(x.ctypes.data, 4, x.ctypes.data, 4, 200, 1, 0, out_i.ctypes.data, None)
    assert rc == _lib.ERR_UNSUPPORTED
    rc = lib.gmd_knn_host_f64(None, 4, x.ctypes.data, 4, 3, 1, 0, out_i.ctypes.data, None)
    assert rc == _lib.ERR_BAD_ARG
    with pytest.raises(ValueError, match="Unknown Distance Metric"):
        engine._metric("chebyshev")
    with pytest.raises(ValueError, match="2-dimensional"):
        engine._as_rows(np.zeros(3), "XA")


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(_lib.GmdistError, match="no CPU fallback"):
        engine.knn(np.zeros((4, 3)), None, k=1)
    # the pure C-ABI host call reports the CUDA error instead of computing on the CPU
    x = np.zeros((4, 3))
    out_i = np.zeros((4, 1), dtype=np.int32)
    rc = _lib.lib().gmd_knn_host_f64(x.ctypes.data, 4, x.ctypes.data, 4, 3, 1, 0, out_i.ctypes.data, None)
    assert rc == _lib.ERR_CUDA


def test_product_never_imports_oracle():
    for py in (ROOT / "groopm_b200").glob("*.py"):
        text = py.read_text()
        assert "import oracle" not in text and "from oracle" not in text, py


def test_squareform_round_trip():
    from scipy.spatial.distance import squareform as sp_squareform

    v = np.arange(1.0, 11.0)
    sq = engine.squareform(v)
    assert np.array_equal(sq, sp_squareform(v))
    assert np.array_equal(engine.squareform(sq), v)
    with pytest.raises(ValueError):
        engine.squareform(np.arange(4.0))


def test_shard_bounds_cover_rows_exactly():
    for n in (1, 63, 64, 65, 1000, 100_000, 1_000_000, 4_000_001):
        for world in (1, 2, 3, 4, 8):
            edges = [dist.shard_bounds(n, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for (a0, a1), (b0, b1) in zip(edges, edges[1:]):
                assert a1 == b0 and a0 <= a1
            per = dist.shard_rows(n, world)
            assert per % 64 == 0
            # only the last non-empty shard may be short
            sizes = [b - a for a, b in edges if b > a]
            assert all(s == per for s in sizes[:-1])
