"""Drop-in ``ODEfunc`` / ``ODEBlock`` for the multi-graph script.

Mirrors /root/reference/ode_nn_ngraphs.py:37-83 (ODEfunc) and :86-152 (ODEBlock):
the state is the ragged concatenation of instances along the node axis and every
instance names its graph through the marker ``x[first row of instance, 5]`` =
graph index + 1 (ode_nn_ngraphs.py:55,65-67,333).
"""
import numpy as np
import torch
import torch.nn as nn

from . import rollout as _ro
from .graph import BatchCache, DeviceGraph


class ODEfunc(nn.Module):
    def __init__(self, A_list, hidden1, device):
        super().__init__()
        if hidden1 != _ro.H:
            raise NotImplementedError("the B200 kernels are specialised for hidden width %d (got %d)" % (_ro.H, hidden1))
        self.A_list = A_list
        self.ln = nn.LayerNorm(hidden1)
        self.linear = nn.Linear(hidden1, hidden1)
        self.relu = nn.ReLU()
        self.sigmoid = nn.Sigmoid()
        self._graphs = None
        self._batches = BatchCache()

    def device_graphs(self):
        if self._graphs is None:
            self._graphs = [DeviceGraph(A) for A in self.A_list]
        return self._graphs

    def batch_from_markers(self, marker_col):
        """marker_col: x[:, 5] (device or host). Non-zero entries open an instance."""
        mk = marker_col.detach().to("cpu", torch.float32).numpy()
        starts = np.flatnonzero(mk)
        gids = mk[starts].astype(np.int64) - 1
        graphs = self.device_graphs()
        if len(starts) == 0 or starts[0] != 0:
            raise RuntimeError("the first row of the batch must carry a graph marker (ode_nn_ngraphs.py:333)")
        sizes = np.diff(np.append(starts, len(mk)))
        for s, g in zip(sizes, gids):
            if g < 0 or g >= len(graphs) or graphs[g].n != s:
                raise RuntimeError("instance of %d rows does not match graph %d" % (s, g))
        return self._batches.get([graphs[g] for g in gids])

    def forward(self, t, x):
        """f(t, y) on the stacked state [4, sum N, H] (ode_nn_ngraphs.py:54-83); inference-only."""
        batch = self.batch_from_markers(x[3, :, 2])
        dy = _ro.odefunc_eval(x[:3], x[3, :, 0], x[3, :, 1], batch,
                              [self.linear.weight, self.linear.bias] + [self.linear.bias] * 6)
        return torch.cat((dy, torch.zeros_like(x[3:])))


class ODEBlock(nn.Module):
    grad_mode = "adjoint"

    def __init__(self, maxTime, deltaT, hidden1, odefunc, device):
        super().__init__()
        self.maxTime = maxTime
        self.deltaT = deltaT
        self.device = device
        self.integration_time = torch.from_numpy(np.arange(0, self.maxTime, self.deltaT)).to(device)
        self.odefunc = odefunc
        self.hidden1 = hidden1
        self.linearS1 = nn.Linear(1, hidden1)
        self.ln = nn.LayerNorm(hidden1)
        self.linear3 = nn.Linear(hidden1, 4)
        self.relu3 = nn.ReLU()
        self.linearS2 = nn.Linear(4, 1)
        self.softmax = nn.Softmax(dim=2)
        self.relu = nn.ReLU()
        self._dt = _ro.dt_array(self.integration_time)

    def _params(self):
        return [self.odefunc.linear.weight, self.odefunc.linear.bias, self.linearS1.weight, self.linearS1.bias,
                self.linear3.weight, self.linear3.bias, self.linearS2.weight, self.linearS2.bias]

    def forward(self, x):
        batch = self.odefunc.batch_from_markers(x[:, 5])
        probs = _ro.rollout(x, batch, self._dt, self._params(), self.grad_mode)
        if probs.requires_grad:
            probs = probs + 0.0 * (self.odefunc.ln.weight.sum() + self.odefunc.ln.bias.sum())
        S, I, R = probs.chunk(3, dim=-1)
        return S, I, R
