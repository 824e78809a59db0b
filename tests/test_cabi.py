"""CPU-side checks of the drop-in boundary: the library builds, loads, and exports
exactly the symbols include/gnode_b200.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import gn_ode_sir_b200 as gn
    gn.build_library()
    return gn


def header_symbols():
    src = open(os.path.join(ROOT, "include", "gnode_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gnode_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    syms = header_symbols()
    for s in ("gnode_graph_create", "gnode_batch_create", "gnode_aggregate", "gnode_odefunc_eval",
              "gnode_rollout_forward", "gnode_rollout_backward", "gnode_last_error"):
        assert s in syms


def test_library_exports_every_declared_symbol(built):
    from gn_ode_sir_b200 import _lib
    handle = ctypes.CDLL(built.LIB_PATH)
    for s in header_symbols():
        assert hasattr(handle, s), "libgnode_b200.so does not export " + s
    assert sorted(_lib.SIGNATURES) == header_symbols()
    assert _lib.lib().gnode_version() >= 100


def test_grad_layout_matches_header(built):
    from gn_ode_sir_b200 import _lib
    src = open(os.path.join(ROOT, "include", "gnode_b200.h")).read()
    assert "#define GNODE_H 64" in src
    assert _lib.GRAD_COUNT == 64 * 64 + 64 + 64 + 64 + 4 * 64 + 4 + 4 + 1


def test_argument_errors_are_reported_without_gpu(built):
    from gn_ode_sir_b200 import _lib
    L = _lib.lib()
    h = ctypes.c_void_p()
    rc = L.gnode_graph_create(0, 0, None, None, ctypes.byref(h))
    assert rc == -1 and b"bad arguments" in L.gnode_last_error()
    with pytest.raises(RuntimeError):
        _lib.check(rc, "gnode_graph_create")


def test_product_path_refuses_cpu_tensors(built):
    import torch
    from gn_ode_sir_b200 import rollout
    with pytest.raises(RuntimeError, match="no CPU path"):
        rollout._check_cuda_f32(torch.zeros(3), "x")


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "gn-ode-sir_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt, os.path.join(dirpath, f)
    for f in ("ode_nn.py", "ode_nn_ngraph_sim.py", "ode_nn_ngraphs.py", "gn_ode_sir_b200.py"):
        p = os.path.join(ROOT, f)
        if os.path.exists(p):
            assert "oracle" not in open(p).read(), f


def test_dropin_state_dict_matches_reference_layout(built):
    import numpy as np
    import scipy.sparse
    import torch
    from oracle import gnode_oracle as orc
    A = scipy.sparse.csr_matrix(np.array([[0, 1], [1, 0]]))
    for seed in (0, 3):
        torch.manual_seed(seed)
        of = built.ode_sim.ODEfunc(A, 0.2, 0.1, 64, "cpu")
        blk = built.ode_sim.ODEBlock(20, 0.5, 2, [0], 64, of, "cpu")
        sd = blk.state_dict()
        assert list(sd.keys()) == list(orc.PARAM_SHAPES.keys())
        want = orc.default_params(64, seed)
        for k, v in sd.items():
            assert tuple(v.shape) == orc.PARAM_SHAPES[k](64)
            assert torch.equal(v, want[k]), k
        torch.manual_seed(seed)
        of2 = built.ode_ngraphs.ODEfunc([A], 64, "cpu")
        blk2 = built.ode_ngraphs.ODEBlock(20, 0.5, 64, of2, "cpu")
        assert all(torch.equal(v, want[k]) for k, v in blk2.state_dict().items())
        assert blk.integration_time.dtype == torch.float64 and len(blk.integration_time) == 40
