"""torch.autograd.Function around the C-ABI rollout (forward + reverse sweep).

PyTorch is plumbing here: it owns device memory, the stream and autograd's graph;
all arithmetic of the path runs in libgnode_b200.so.
"""
import ctypes

import numpy as np
import torch

from . import _lib

H = _lib.GNODE_H
PARAM_ORDER = tuple(k for k, _ in _lib.GRAD_LAYOUT)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _params_struct(tensors):
    return _lib.GnodeParams(*[t.data_ptr() for t in tensors])


def _check_cuda_f32(t, name):
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor: the GN-ODE rollout has no CPU path (got %s)" % (name, t.device))
    if t.dtype != torch.float32:
        raise RuntimeError("%s must be float32 (got %s)" % (name, t.dtype))


def dt_array(integration_time):
    """dt_k = float32(t_{k+1} - t_k) computed in float64 like torchdiffeq's fixed grid."""
    t = np.asarray(integration_time.detach().cpu().numpy() if torch.is_tensor(integration_time)
                   else integration_time, dtype=np.float64)
    return np.ascontiguousarray((t[1:] - t[:-1]).astype(np.float32))


class _Rollout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, batch, dt, grad_mode, want_grad, *params):
        L = _lib.lib()
        M, T = batch.M, len(dt) + 1
        _check_cuda_f32(x, "x")
        if x.dim() != 2 or x.size(0) != M or x.size(1) < 5 or x.stride(1) != 1:
            raise RuntimeError("x must be [M=%d, >=5] with unit column stride, got %r" % (M, tuple(x.shape)))
        ps = []
        for k, p in zip(PARAM_ORDER, params):
            _check_cuda_f32(p, k)
            ps.append(p.detach().contiguous())
        # grad mode is always off inside forward(): the caller's grad mode arrives as `want_grad`
        need_grad = bool(want_grad) and any(ctx.needs_input_grad[5:])
        traj = torch.empty((T, 3, M, H), dtype=torch.float32, device=x.device) if need_grad else None
        probs = torch.empty((T, M, 3), dtype=torch.float32, device=x.device)
        ws_bytes = int(L.gnode_rollout_workspace_bytes(batch.handle, 1 if need_grad else 0))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        pstruct = _params_struct(ps)
        _lib.check(L.gnode_rollout_forward(batch.handle, _ptr(x), x.stride(0), ctypes.byref(pstruct), T,
                                           dt.ctypes.data_as(_lib.c_float_p),
                                           _ptr(traj) if need_grad else None, _ptr(probs), _ptr(ws), ws_bytes,
                                           _stream()), "gnode_rollout_forward")
        if need_grad:
            ctx.save_for_backward(x, traj, *ps)
            ctx.batch, ctx.dt, ctx.grad_mode = batch, dt, grad_mode
        return probs

    @staticmethod
    def backward(ctx, grad_probs):
        L = _lib.lib()
        x, traj, *ps = ctx.saved_tensors
        batch, dt = ctx.batch, ctx.dt
        T = len(dt) + 1
        grad_probs = grad_probs.contiguous()
        _check_cuda_f32(grad_probs, "grad_probs")
        grads = torch.empty(_lib.GRAD_COUNT, dtype=torch.float32, device=x.device)
        ws_bytes = int(L.gnode_backward_workspace_bytes(batch.handle))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        pstruct = _params_struct(ps)
        mode = {"adjoint": _lib.GRAD_ADJOINT, "discrete": _lib.GRAD_DISCRETE}[ctx.grad_mode]
        _lib.check(L.gnode_rollout_backward(batch.handle, _ptr(x), x.stride(0), ctypes.byref(pstruct), T,
                                            dt.ctypes.data_as(_lib.c_float_p), _ptr(traj), _ptr(grad_probs), mode,
                                            _ptr(grads), _ptr(ws), ws_bytes, _stream()), "gnode_rollout_backward")
        out, off = [], 0
        for i, (k, shape) in enumerate(_lib.GRAD_LAYOUT):
            n = int(np.prod(shape))
            out.append(grads[off:off + n].view(shape) if ctx.needs_input_grad[5 + i] else None)
            off += n
        return (None, None, None, None, None, *out)


def rollout(x, batch, dt, params, grad_mode="adjoint"):
    """x [M, >=5] -> probabilities [T, M, 3]; params in PARAM_ORDER (state_dict names)."""
    if grad_mode not in ("adjoint", "discrete"):
        raise ValueError("grad_mode must be 'adjoint' or 'discrete'")
    want_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
    return _Rollout.apply(x, batch, dt, grad_mode, want_grad, *params)


def odefunc_eval(y, beta, gamma, batch, params):
    """One f(t, y): y [3, M, H] -> [3, M, H] (dS, dI, dR). Not differentiable."""
    L = _lib.lib()
    for t, n in ((y, "y"), (beta, "beta"), (gamma, "gamma")):
        _check_cuda_f32(t, n)
    y, beta, gamma = y.contiguous(), beta.contiguous(), gamma.contiguous()
    ps = [p.detach().contiguous() for p in params]
    dy = torch.empty_like(y)
    scratch = torch.empty((batch.M, H), dtype=torch.float32, device=y.device)
    pstruct = _params_struct(ps)
    _lib.check(L.gnode_odefunc_eval(batch.handle, _ptr(y), _ptr(beta), _ptr(gamma), ctypes.byref(pstruct),
                                    _ptr(dy), _ptr(scratch), _stream()), "gnode_odefunc_eval")
    return dy


def aggregate(v, batch, transpose=False):
    """out[r] = sum of v over the neighbours of r. v: [M, H]."""
    L = _lib.lib()
    _check_cuda_f32(v, "v")
    v = v.contiguous()
    out = torch.empty_like(v)
    _lib.check(L.gnode_aggregate(batch.handle, _ptr(v), _ptr(out), 1 if transpose else 0, _stream()),
               "gnode_aggregate")
    return out
