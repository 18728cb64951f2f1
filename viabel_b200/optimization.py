"""Stochastic optimisers -- mirror of viabel/optimization.py's SGD family (:52-476).

`optimize()` keeps the variational parameter, the optimiser state and all histories on the
device; RMSProp and Adam use the fused CUDA step (state update + parameter update in one
launch).  `descent_direction(grad)` is kept for drop-in use with numpy arrays or tensors and
reproduces the reference's first-step behaviour.
"""
from abc import ABC, abstractmethod
from collections import defaultdict

import numpy as np
import torch
import tqdm

from . import _lib
from ._tensor import F64, device, is_host, to_dev

__all__ = ['Optimizer', 'StochasticGradientOptimizer', 'RMSProp', 'Adam', 'Adagrad', 'WindowedAdagrad',
           'AveragedRMSProp', 'AveragedAdam']


def _sqrt(x):
    return torch.sqrt(x) if isinstance(x, torch.Tensor) else np.sqrt(x)


def _to_numpy_results(results):
    out = {}
    for key, hist in results.items():
        if isinstance(hist, list) and len(hist) and isinstance(hist[0], torch.Tensor):
            out[key] = torch.stack([h.reshape(-1) if h.dim() else h for h in hist]).cpu().numpy()
        elif isinstance(hist, torch.Tensor):
            out[key] = hist.cpu().numpy()
        else:
            out[key] = np.array(hist)
    return out


class Optimizer(ABC):
    @abstractmethod
    def optimize(self, n_iters, objective, init_param, **kwargs):
        """Return a dict with at least `opt_param`."""


class StochasticGradientOptimizer(Optimizer):
    """Plain SGD and the base of the adaptive methods (optimization.py:52-144)."""

    def __init__(self, learning_rate, *, weight_decay=0, iterate_avg_prop=0.2, diagnostics=False):
        self._learning_rate = learning_rate
        self._weight_decay = weight_decay
        if iterate_avg_prop is not None and (iterate_avg_prop > 1.0 or iterate_avg_prop <= 0.0):
            raise ValueError('"iterate_avg_prop" must be None or between 0 and 1')
        self._iterate_avg_prop = iterate_avg_prop
        self._diagnostics = diagnostics
        self.progress = True
        self.reset_state()

    def reset_state(self):
        pass

    def descent_direction(self, grad):
        return grad

    def _fused_step(self, var_param, grad, want_dir):
        """In-place `var_param -= lr * direction` on device tensors; returns direction or None.
        Subclasses with a CUDA step override this."""
        direction = self.descent_direction(grad)
        var_param -= self._learning_rate * direction
        return direction if want_dir else None

    def optimize(self, n_iters, objective, init_param, init_hamflow_model_param=None,
                 init_hamflow_rho_param=None):
        """The reference loop (optimization.py:83-127) with device-resident state."""
        from .objectives import VariationalObjective
        var_param = to_dev(init_param).clone()
        iap = self._iterate_avg_prop
        results = defaultdict(list)
        plain_update = type(objective).update is VariationalObjective.update \
            if isinstance(objective, VariationalObjective) else False
        k = -1
        bar = tqdm.trange(n_iters, disable=not self.progress)
        try:
            for k in bar:
                value, grad = objective(var_param)
                if not isinstance(grad, torch.Tensor):      # host objective (user code)
                    grad = to_dev(grad)
                    value = torch.as_tensor(float(value), dtype=F64, device=grad.device)
                if plain_update:
                    direction = self._fused_step(var_param, grad, self._diagnostics)
                else:
                    direction = self.descent_direction(grad)
                    var_param = to_dev(objective.update(var_param, self._learning_rate * direction))
                if var_param.dim() == 2:
                    var_param *= (1 - self._weight_decay)
                results['value_history'].append(value.detach().reshape(()))
                if self._diagnostics or iap is not None:
                    results['variational_param_history'].append(var_param.clone())
                    if iap is not None and len(results['variational_param_history']) > iap * k:
                        results['variational_param_history'].pop(0)
                if self._diagnostics:
                    results['descent_dir_history'].append(direction.clone())
                if self.progress and k % 10 == 0:
                    recent = torch.stack(results['value_history'][max(0, k - 1000):k + 1])
                    bar.set_description('average loss = {:,.5g}'.format(float(recent.mean())))
        except (KeyboardInterrupt, StopIteration):  # pragma: no cover
            pass
        finally:
            bar.close()
        if iap is not None and len(results['variational_param_history']):
            window = max(1, int(k * iap))
            tail = torch.stack(results['variational_param_history'][-window:])
            results['opt_param'] = tail.mean(dim=0)
        else:
            # (for n_iters <= 5 the reference's window is empty and it returns NaN; the last
            # iterate is returned instead)
            results['opt_param'] = var_param.clone()
        return _to_numpy_results(results)


class RMSProp(StochasticGradientOptimizer):
    """RMSProp (optimization.py:147-197): nu starts at grad**2."""

    def __init__(self, learning_rate, *, weight_decay=0, iterate_avg_prop=0.2, beta=0.9, jitter=1e-8,
                 diagnostics=False):
        self._beta = beta
        self._jitter = jitter
        super().__init__(learning_rate, weight_decay=weight_decay, iterate_avg_prop=iterate_avg_prop,
                         diagnostics=diagnostics)

    def reset_state(self):
        self._avg_grad_sq = None

    def descent_direction(self, grad):
        g2 = grad ** 2
        nu = g2 if self._avg_grad_sq is None else self._avg_grad_sq
        nu = nu * self._beta
        nu = nu + (1. - self._beta) * g2
        self._avg_grad_sq = nu
        return grad / _sqrt(self._jitter + nu)

    def _fused_step(self, var_param, grad, want_dir):
        first = self._avg_grad_sq is None or not isinstance(self._avg_grad_sq, torch.Tensor)
        if first:
            self._avg_grad_sq = torch.empty_like(grad)
        direction = torch.empty_like(grad) if want_dir else None
        _lib.check(_lib.lib.vb_rmsprop_step_f64(
            _lib.ptr(var_param), _lib.ptr(grad), _lib.ptr(self._avg_grad_sq), _lib.ptr(direction),
            grad.numel(), float(self._learning_rate), float(self._beta), float(self._jitter), int(first),
            _lib.stream()))
        return direction


class Adam(StochasticGradientOptimizer):
    """Adam (optimization.py:260-326), including the first-step aliasing of `momentum = grad`."""

    def __init__(self, learning_rate, *, beta1=0.9, beta2=0.999, jitter=1e-8, iterate_avg_prop=0.2,
                 diagnostics=False):
        self._beta1 = beta1
        self._beta2 = beta2
        self._jitter = jitter
        super().__init__(learning_rate, iterate_avg_prop=iterate_avg_prop, diagnostics=diagnostics)

    def reset_state(self):
        self._momentum = None
        self._avg_grad_sq = None

    def descent_direction(self, grad):
        b1, b2 = self._beta1, self._beta2
        if self._momentum is None:
            scaled = grad * b1                       # the aliased array after `momentum *= beta1`
            m = scaled + (1. - b1) * scaled
            nu = (grad ** 2) * b2
            nu = nu + (1. - b2) * m ** 2
        else:
            m = self._momentum * b1
            m = m + (1. - b1) * grad
            nu = self._avg_grad_sq * b2
            nu = nu + (1. - b2) * grad ** 2
        self._momentum, self._avg_grad_sq = m, nu
        return m / _sqrt(self._jitter + nu)

    def _fused_step(self, var_param, grad, want_dir):
        first = self._momentum is None or not isinstance(self._momentum, torch.Tensor)
        if first:
            self._momentum = torch.empty_like(grad)
            self._avg_grad_sq = torch.empty_like(grad)
        direction = torch.empty_like(grad) if want_dir else None
        _lib.check(_lib.lib.vb_adam_step_f64(
            _lib.ptr(var_param), _lib.ptr(grad), _lib.ptr(self._momentum), _lib.ptr(self._avg_grad_sq),
            _lib.ptr(direction), grad.numel(), float(self._learning_rate), float(self._beta1),
            float(self._beta2), float(self._jitter), int(first), _lib.stream()))
        return direction


class AveragedRMSProp(StochasticGradientOptimizer):
    """optimization.py:200-258 (beta_k = 1 - 1/k)."""

    def __init__(self, learning_rate, *, jitter=1e-8, diagnostics=False, component_wise=True):
        self._jitter = jitter
        self._component_wise = component_wise
        super().__init__(learning_rate, diagnostics=diagnostics)

    def reset_state(self):
        self._avg_grad_sq = None
        self._t = None

    def descent_direction(self, grad):
        g2 = grad ** 2
        if self._avg_grad_sq is None:
            nu, t = g2, 1
        else:
            nu, t = self._avg_grad_sq, self._t + 1
        beta = 1 - 1 / t
        nu = nu * beta + (1. - beta) * g2
        self._avg_grad_sq, self._t = nu, t
        denom = nu if self._component_wise else nu.sum()
        return grad / _sqrt(self._jitter + denom)


class AveragedAdam(StochasticGradientOptimizer):
    """optimization.py:328-396."""

    def __init__(self, learning_rate, *, beta1=0.9, jitter=1e-8, diagnostics=False, component_wise=True):
        self._beta1 = beta1
        self._jitter = jitter
        self._component_wise = component_wise
        super().__init__(learning_rate, diagnostics=diagnostics)

    def reset_state(self):
        self._momentum = None
        self._avg_grad_sq = None
        self._t = None

    def descent_direction(self, grad):
        b1 = self._beta1
        if self._momentum is None:
            # same aliasing as Adam: grad is scaled in place before it is reused
            scaled = grad * b1
            m = scaled + (1. - b1) * scaled
            g_eff2 = m ** 2
            nu, t = grad ** 2, 1
        else:
            m = self._momentum * b1 + (1. - b1) * grad
            g_eff2 = grad ** 2
            nu, t = self._avg_grad_sq, self._t + 1
        beta2 = 1 - 1 / t
        nu = nu * beta2 + (1. - beta2) * g_eff2
        self._momentum, self._avg_grad_sq, self._t = m, nu, t
        denom = nu if self._component_wise else nu.sum()
        return m / _sqrt(self._jitter + denom)


class Adagrad(StochasticGradientOptimizer):
    """optimization.py:398-433."""

    def __init__(self, learning_rate, *, weight_decay=0, jitter=1e-8, iterate_avg_prop=0.2, diagnostics=False):
        self._jitter = jitter
        super().__init__(learning_rate, weight_decay=weight_decay, iterate_avg_prop=iterate_avg_prop,
                         diagnostics=diagnostics)

    def reset_state(self):
        self._sum_grad_sq = 0

    def descent_direction(self, grad):
        self._sum_grad_sq = self._sum_grad_sq + grad ** 2
        return grad / _sqrt(self._jitter + self._sum_grad_sq)


class WindowedAdagrad(StochasticGradientOptimizer):
    """optimization.py:435-476."""

    def __init__(self, learning_rate, *, weight_decay=0, window_size=10, jitter=1e-8, diagnostics=False):
        self._window_size = window_size
        self._jitter = jitter
        super().__init__(learning_rate, weight_decay=weight_decay, diagnostics=diagnostics)

    def reset_state(self):
        self._history = []

    def descent_direction(self, grad):
        self._history.append(grad ** 2)
        if len(self._history) > self._window_size:
            self._history.pop(0)
        mean_sq = sum(self._history) / len(self._history)
        return grad / _sqrt(self._jitter + mean_sq)
