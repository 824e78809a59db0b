"""Build-container-only check: the oracle equals the reference's own classes
BIT FOR BIT on fresh seeds (not just on the committed fixtures). Skipped where
/root/reference is absent (the GPU box)."""
import pytest
import torch

from oracle import gnode_oracle as orc
from oracle import ref_harness as rh

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="reference tree not present")


@pytest.mark.parametrize("graph,B,seed", [("karate", 3, 11), ("dolphins", 2, 12), ("fb-food", 1, 13)])
def test_sim_variant_bit_exact(graph, B, seed):
    sim, _ = rh.load_reference("adjoint")
    A = rh.load_reference_graph(graph)
    N, H = A.shape[0], 64
    torch.manual_seed(seed)
    of = sim.ODEfunc(A, 0.2, 0.1, H, "cpu")
    blk = sim.ODEBlock(20, 0.5, N, [0, 1], H, of, "cpu")
    x = torch.stack([orc.synthetic_trial(N, H, 7 * seed + b) for b in range(B)])
    with torch.no_grad():
        S, I, R = blk(x)
    ref = torch.cat((S, I, R), -1)
    params = {k: v.detach().clone() for k, v in blk.state_dict().items()}
    got = orc.forward(x.view(-1, 3 + H), params, orc.batch_coo([A], [0] * B), orc.time_grid(20, 0.5))
    assert torch.equal(got, ref)
    assert all(torch.equal(v, params[k]) for k, v in orc.default_params(H, seed).items())


def test_ngraphs_variant_bit_exact():
    _, ng = rh.load_reference("adjoint")
    names = ["karate", "dolphins"]
    adjs = [rh.load_reference_graph(n) for n in names]
    inst = [1, 0, 1]
    H = 64
    torch.manual_seed(3)
    of = ng.ODEfunc(adjs, H, "cpu")
    blk = ng.ODEBlock(10, 0.5, H, of, "cpu")
    x = torch.cat([orc.synthetic_trial(adjs[g].shape[0], H, 50 + i, graph_marker=g + 1.0)
                   for i, g in enumerate(inst)])
    with torch.no_grad():
        S, I, R = blk(x)
    ref = torch.cat((S, I, R), -1)
    params = {k: v.detach().clone() for k, v in blk.state_dict().items()}
    got = orc.forward(x, params, orc.batch_coo(adjs, inst), orc.time_grid(10, 0.5))
    assert torch.equal(got, ref)


@pytest.mark.skipif(not rh.reference_available(), reason="needs /root/reference (build container only)")
@pytest.mark.parametrize("name", ["karate", "dolphins", "fb-food", "fb-social", "openflights", "wiki-vote", "enron"])
def test_cached_graph_pipeline_equals_reference_create_graph(name, tmp_path, monkeypatch):
    """N2: pickle -> largest component -> CSR without networkx's adjacency_matrix, and its on-disk cache, give exactly the
    matrix the reference's create_graph builds (ode_nn.py:394-414), for every shipped graph."""
    import numpy as np
    import gn_ode_sir_b200  # noqa: F401
    from gn_ode_sir_b200 import harness
    monkeypatch.setenv("GNODE_GRAPH_CACHE", str(tmp_path))
    want = rh.load_reference_graph(name).tocsr()
    want.sort_indices()
    label = rh.REFERENCE_ROOT + "/real_graphs/" + name
    for attempt in ("built", "cached"):
        G, A = harness.load_graph(label)
        assert A.shape == want.shape and np.array_equal(A.indptr, want.indptr) and np.array_equal(A.indices, want.indices), attempt
        assert G.number_of_nodes() == want.shape[0]
    assert len(list(tmp_path.iterdir())) == 1                       # the second load came from the cache file
