"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the optimizer step of the reference training loop.

Reference: tools/train.py:139-148 builds `torch.optim.SGD(params, lr, momentum, weight_decay, nesterov)` over ALL
parameters in one group; utils/function.py:47-49 calls `optimizer.step()` every iteration and then
`adjust_learning_rate` (utils/utils.py:154-160, poly schedule, power 0.9).  The arithmetic lives in torch
(`torch/optim/sgd.py::_single_tensor_sgd`); this file restates it in numpy fp32 and is pinned against the live
`torch.optim.SGD` in tests/test_sgd_oracle.py (same container torch the reference itself would run on).
Only tests/ may import this module; the product path (pidnet_b200/optim.py -> pidnet_sgd_step) never does."""
import numpy as np


def sgd_step(p, g, buf, lr, momentum=0.9, dampening=0.0, weight_decay=5e-4, nesterov=False, first=False):
    """One SGD step on fp32 arrays; returns (p_new, buf_new).  Follows torch/optim/sgd.py::_single_tensor_sgd:
    d_p = g + wd*p; buf = d_p (first step) | momentum*buf + (1-dampening)*d_p; d_p = d_p + momentum*buf (nesterov) | buf;
    p = p - lr*d_p."""
    f = np.float32
    p = np.asarray(p, f); g = np.asarray(g, f)
    d = g.copy()
    if weight_decay != 0:
        d = (d + f(weight_decay) * p).astype(f)
    if momentum != 0:
        if first or buf is None:
            buf = d.copy()
        else:
            buf = (f(momentum) * np.asarray(buf, f) + f(1 - dampening) * d).astype(f)
        d = (d + f(momentum) * buf).astype(f) if nesterov else buf
    return (p - f(lr) * d).astype(f), buf


def adjust_learning_rate(base_lr, max_iters, cur_iters, power=0.9):
    """utils/utils.py:154-160: lr = base_lr * (1 - cur/max) ** power (python float arithmetic)."""
    return base_lr * ((1 - float(cur_iters) / max_iters) ** power)
