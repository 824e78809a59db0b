"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Loads the *unmodified* reference modules from ``/root/reference`` so that the
oracle restatement (``oracle/gnode_oracle.py``) can be pinned against the
reference's own classes and so that golden vectors can be generated
(``tests/golden/make_golden.py``).  ``/root/reference`` exists only in the
build container, never on the GPU box; everything here is therefore used only
by the fixture generator and by ``-m "not gpu"`` tests that skip when the
reference tree is absent.

What is stubbed:
  * matplotlib / ndlib  -- imported by the reference at module scope, unused on
    the rollout path (ode_nn.py:2,16-18; ode_nn_ngraph_sim.py:2,18-20).
  * torchdiffeq==0.2.2 (requirements.txt:59) -- NOT vendored and NOT installed.
    Its fixed-grid Euler solver and adjoint backward are restated below from
    the published algorithm:
      forward   y_{k+1} = y_k + (t_{k+1}-t_k) * f(t_k, y_k); outputs are the grid
                states (grid == the requested times, no step_size option is
                passed at ode_nn_ngraph_sim.py:168 / ode_nn_ngraphs.py:137).
      backward  OdeintAdjointMethod.backward: for i = T-1..1 one Euler step of
                the augmented system from t_i to t_{i-1}, evaluated at the
                stored forward state y_i, then the state slot is reset to
                y_{i-1} and dL/dy_{i-1} is added to the adjoint.
    The reference holds no test that pins torchdiffeq's results on this path
    ("parity unpinned" at that boundary).
"""
import os
import sys
import types

import torch

REFERENCE_ROOT = "/root/reference"


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "ode_nn_ngraph_sim.py"))


# --------------------------------------------------------------------------
# torchdiffeq restatement (forward + adjoint backward) used ONLY to drive the
# reference's own ODEfunc.  The adjoint parameters are func.parameters(), as
# torchdiffeq's odeint_adjoint collects them when adjoint_params is omitted.
# --------------------------------------------------------------------------
class _EulerAdjoint(torch.autograd.Function):
    @staticmethod
    def forward(ctx, func, t, y0, *adjoint_params):
        with torch.no_grad():
            sol = [y0]
            y = y0
            for k in range(len(t) - 1):
                dt = t[k + 1] - t[k]
                y = y + dt * func(t[k], y)
                sol.append(y)
            sol = torch.stack(sol)
        ctx.func = func
        ctx.save_for_backward(t, sol, *adjoint_params)
        return sol

    @staticmethod
    def backward(ctx, grad_sol):
        func = ctx.func
        t, sol, *params = ctx.saved_tensors
        params = tuple(params)
        with torch.no_grad():
            adj_y = grad_sol[-1].clone()
            adj_p = [torch.zeros_like(p) for p in params]
            for i in range(len(t) - 1, 0, -1):
                with torch.enable_grad():
                    y_i = sol[i].detach().requires_grad_(True)
                    f_i = func(t[i], y_i)
                    vjps = torch.autograd.grad(f_i, (y_i,) + params, -adj_y,
                                               allow_unused=True)
                vjp_y = vjps[0] if vjps[0] is not None else torch.zeros_like(y_i)
                dt = t[i - 1] - t[i]                      # negative
                adj_y = adj_y + dt * vjp_y
                for j, v in enumerate(vjps[1:]):
                    if v is not None:
                        adj_p[j] = adj_p[j] + dt * v
                adj_y = adj_y + grad_sol[i - 1]
        return (None, None, adj_y, *adj_p)


def odeint_adjoint_restated(func, y0, t, method="euler", **kw):
    assert method == "euler"
    params = tuple(p for p in func.parameters() if p.requires_grad)
    return _EulerAdjoint.apply(func, t, y0, *params)


def odeint_plain_restated(func, y0, t, method="euler", **kw):
    """Differentiable Euler loop: plain autograd through it is the *discrete* gradient."""
    assert method == "euler"
    sol, y = [y0], y0
    for k in range(len(t) - 1):
        y = y + (t[k + 1] - t[k]) * func(t[k], y)
        sol.append(y)
    return torch.stack(sol)


_loaded = {}


def load_reference(grad_mode="adjoint"):
    """Import the reference's two live model modules, unmodified.

    grad_mode: "adjoint" -> odeint = restated torchdiffeq adjoint;
               "discrete" -> odeint = plain differentiable Euler loop.
    Returns (ode_nn_ngraph_sim, ode_nn_ngraphs).
    """
    if not reference_available():
        raise RuntimeError("reference tree not present at " + REFERENCE_ROOT)

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    if not _loaded:
        for n in ["matplotlib", "matplotlib.pyplot", "ndlib", "ndlib.models",
                  "ndlib.models.ModelConfig", "ndlib.models.CompositeModel",
                  "ndlib.models.compartments"]:
            if n not in sys.modules:
                stub(n)
        if "ndlib.models.epidemics" not in sys.modules:
            stub("ndlib.models.epidemics", SIRModel=object)
        stub("torchdiffeq", odeint_adjoint=odeint_adjoint_restated,
             odeint=odeint_plain_restated)
        # import under private names so the repo-root drop-in scripts of the
        # same file names are never shadowed / confused with the reference
        import importlib.util
        saved_path = list(sys.path)
        sys.path.insert(0, REFERENCE_ROOT)
        try:
            saved_helper = sys.modules.pop("ode_nn", None)
            mods = []
            for fname in ("ode_nn_ngraph_sim", "ode_nn_ngraphs"):
                spec = importlib.util.spec_from_file_location(
                    "_reference_" + fname, os.path.join(REFERENCE_ROOT, fname + ".py"))
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                mods.append(mod)
            _loaded["mods"] = tuple(mods)
            _loaded["ref_helper"] = sys.modules.pop("ode_nn", None)
            if saved_helper is not None:
                sys.modules["ode_nn"] = saved_helper
        finally:
            sys.path[:] = saved_path
        # the reference flips the default dtype to float64 at import time
        # (ode_nn.py:493, ode_nn_ngraph_sim.py:322); main() flips it back before
        # building the model (ode_nn_ngraph_sim.py:433).
        torch.set_default_dtype(torch.float32)
    sim, ng = _loaded["mods"]
    od = odeint_adjoint_restated if grad_mode == "adjoint" else odeint_plain_restated
    sim.odeint = od
    ng.odeint = od
    return sim, ng


def load_reference_graph(name):
    """scipy CSR adjacency exactly as the reference builds it (ode_nn.py:394-414)."""
    import pickle
    import networkx as nx
    with open(os.path.join(REFERENCE_ROOT, "real_graphs", name + ".pkl"), "rb") as fh:
        G = pickle.load(fh)
    G = G.to_undirected()
    G = G.subgraph(max(nx.connected_components(G), key=len))
    return nx.adjacency_matrix(G)
