"""Minimal TensorFlow-2 operator shim on top of torch float64 (TEST INFRASTRUCTURE ONLY).

Purpose: TensorFlow cannot be installed in this image (no wheel, no network), but the
reference (/root/reference/equation.py, solver.py) is pure Python over ``tf.*`` calls.  With
this package first on ``sys.path`` the reference's own, unmodified source files import and
run, so golden vectors can be generated from the reference's control flow and formulas
(tests/golden/make_golden.py).  Only the ~35 TF entry points those two files touch exist here,
each following the documented TF2 semantics (sign(0)=0, BatchNormalization inference formula,
Keras Adam with epsilon outside the sqrt, PiecewiseConstantDecay with ``step <= boundary``).

Never imported by the product package.
"""
from __future__ import annotations

import math as _math
import types as _types

import numpy as _np
import torch as _torch

_DT = _torch.float64
_FLOATX = "float64"


def _raw(x):
    if isinstance(x, Tensor):
        return x.t
    if isinstance(x, _torch.Tensor):
        return x
    if isinstance(x, _np.ndarray):
        return _torch.as_tensor(x, dtype=_DT) if x.dtype.kind in "fiub" else _torch.as_tensor(x)
    return x  # python scalar


def _tt(x):
    r = _raw(x)
    if not isinstance(r, _torch.Tensor):
        r = _torch.as_tensor(r, dtype=_DT)
    return r


class Tensor:
    """Eager tensor / variable: wraps a torch tensor, interoperates with numpy operands."""
    __array_ufunc__ = None       # make ndarray.__op__(Tensor) defer to our reflected ops
    __array_priority__ = 10000

    def __init__(self, t):
        self.t = t

    # numpy-ish surface
    @property
    def shape(self):
        return tuple(self.t.shape)

    @property
    def dtype(self):
        return self.t.dtype

    def numpy(self):
        return self.t.detach().cpu().numpy()

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a.astype(dtype) if dtype is not None else a

    def __len__(self):
        return self.t.shape[0]

    def __getitem__(self, idx):
        return Tensor(self.t[idx])

    def __float__(self):
        return float(self.t)

    def __repr__(self):
        return f"tf_shim.Tensor({self.t!r})"

    # arithmetic
    def __add__(self, o): return Tensor(self.t + _raw(o))
    def __radd__(self, o): return Tensor(_raw(o) + self.t)
    def __sub__(self, o): return Tensor(self.t - _raw(o))
    def __rsub__(self, o): return Tensor(_raw(o) - self.t)
    def __mul__(self, o): return Tensor(self.t * _raw(o))
    def __rmul__(self, o): return Tensor(_raw(o) * self.t)
    def __truediv__(self, o): return Tensor(self.t / _raw(o))
    def __rtruediv__(self, o): return Tensor(_raw(o) / self.t)
    def __pow__(self, o): return Tensor(self.t ** _raw(o))
    def __neg__(self): return Tensor(-self.t)
    def __lt__(self, o): return Tensor(self.t < _raw(o))
    def __gt__(self, o): return Tensor(self.t > _raw(o))


class Variable(Tensor):
    def __init__(self, value, name=None):
        super().__init__(_tt(value).clone().detach().requires_grad_(True))
        self.name = name

    def assign(self, value):
        with _torch.no_grad():
            self.t.copy_(_tt(value).reshape(self.t.shape))
        return self


def _W(t):
    return Tensor(t)


# ------------------------------------------------------------------ plain ops used by the reference
def reshape(x, shape): return _W(_tt(x).reshape(*[int(s) for s in shape]))
def concat(values, axis): return _W(_torch.cat([_tt(v) for v in values], dim=axis))
def sqrt(x): return _W(_torch.sqrt(_tt(x)))
def square(x): return _W(_tt(x) ** 2)
def sign(x): return _W(_torch.sign(_tt(x)))          # tf.sign(0) == 0, like torch
def abs(x): return _W(_torch.abs(_tt(x)))            # noqa: A001
def maximum(x, y): return _W(_torch.maximum(_tt(x), _tt(y)))
def where(c, x, y): return _W(_torch.where(_raw(c), _tt(x), _tt(y)))


def _reduce(fn, x, axis, keepdims):
    t = _tt(x)
    if axis is None:
        return _W(fn(t))
    return _W(fn(t, dim=axis, keepdim=keepdims))


def reduce_sum(x, axis=None, keepdims=False): return _reduce(_torch.sum, x, axis, keepdims)
def reduce_mean(x, axis=None, keepdims=False): return _reduce(_torch.mean, x, axis, keepdims)


def reduce_max(x, axis=None, keepdims=False):
    t = _tt(x)
    return _W(_torch.max(t)) if axis is None else _W(_torch.max(t, dim=axis, keepdim=keepdims).values)


math = _types.SimpleNamespace(
    ceil=lambda x: _W(_torch.ceil(_tt(x))),
    floor=lambda x: _W(_torch.floor(_tt(x))),
    sign=sign,
    exp=lambda x: _W(_torch.exp(_tt(x))),
)

nn = _types.SimpleNamespace(relu=lambda x: _W(_torch.relu(_tt(x))))

linalg = _types.SimpleNamespace(
    matvec=lambda a, b: _W(_torch.einsum("bij,bj->bi", _tt(a), _tt(b))),
    diag=lambda x: _W(_torch.diag_embed(_tt(x))),
)


def function(fn=None, **_kw):
    """@tf.function: graph tracing is an execution detail; eager semantics are identical."""
    if fn is None:
        return lambda f: f
    return fn


class GradientTape:
    def __init__(self, persistent=False):
        self.persistent = persistent

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def watch(self, x):
        pass

    def gradient(self, target, sources):
        single = not isinstance(sources, (list, tuple))
        src = [sources] if single else list(sources)
        g = _torch.autograd.grad(_tt(target), [s.t for s in src], allow_unused=True, retain_graph=True)
        out = [None if gi is None else _W(gi) for gi in g]
        return out[0] if single else out


# ------------------------------------------------------------------ initialisers
class random_normal_initializer:
    def __init__(self, mean=0.0, stddev=0.05):
        self.mean, self.stddev = mean, stddev

    def __call__(self, shape):
        return _np.random.normal(self.mean, self.stddev, size=shape)


class random_uniform_initializer:
    def __init__(self, minval=-0.05, maxval=0.05):
        self.minval, self.maxval = minval, maxval

    def __call__(self, shape):
        return _np.random.uniform(self.minval, self.maxval, size=shape)


# ------------------------------------------------------------------ keras subset
class _Layer:
    def __init__(self):
        self.built = False

    def __call__(self, *a, **k):
        if not self.built:
            self.build(_tt(a[0]).shape)
            self.built = True
        return self.call(*a, **k)


class _BatchNormalization(_Layer):
    """Inference formula (training=False): gamma*(x-moving_mean)/sqrt(moving_var+eps)+beta, with
    moving_mean=0, moving_var=1 (Keras initial values; never updated when training=False)."""

    def __init__(self, momentum=0.99, epsilon=1e-3, beta_initializer=None, gamma_initializer=None):
        super().__init__()
        self.epsilon = epsilon
        self.beta_initializer, self.gamma_initializer = beta_initializer, gamma_initializer

    def build(self, shape):
        n = int(shape[-1])
        self.gamma = Variable(self.gamma_initializer([n]) if self.gamma_initializer else _np.ones(n))
        self.beta = Variable(self.beta_initializer([n]) if self.beta_initializer else _np.zeros(n))
        self.moving_mean = _torch.zeros(n, dtype=_DT)
        self.moving_variance = _torch.ones(n, dtype=_DT)

    def call(self, x, training=False):
        if training:
            raise NotImplementedError("the reference always passes training=False")
        t = _tt(x)
        return _W((t - self.moving_mean) * (self.gamma.t / _torch.sqrt(self.moving_variance + self.epsilon)) + self.beta.t)

    @property
    def trainable_variables(self):
        return [self.gamma, self.beta] if self.built else []


class _Dense(_Layer):
    def __init__(self, units, use_bias=True, activation=None):
        super().__init__()
        assert activation is None
        self.units, self.use_bias = int(units), use_bias

    def build(self, shape):
        fan_in = int(shape[-1])
        lim = _math.sqrt(6.0 / (fan_in + self.units))            # Glorot uniform (Keras default)
        self.kernel = Variable(_np.random.uniform(-lim, lim, size=[fan_in, self.units]))
        self.bias = Variable(_np.zeros(self.units)) if self.use_bias else None

    def call(self, x):
        y = _tt(x) @ self.kernel.t
        if self.bias is not None:
            y = y + self.bias.t
        return _W(y)

    @property
    def trainable_variables(self):
        if not self.built:
            return []
        return [self.kernel] + ([self.bias] if self.bias is not None else [])


class _Model:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return self.call(*a, **k)

    @property
    def trainable_variables(self):
        out = []

        def visit(o):
            if isinstance(o, (_Layer, _Model)):
                out.extend(o.trainable_variables)
            elif isinstance(o, (list, tuple)):
                for e in o:
                    visit(e)

        for _, v in self.__dict__.items():      # attribute-creation order, as Keras tracks
            visit(v)
        return out


class _PiecewiseConstantDecay:
    def __init__(self, boundaries, values):
        self.boundaries, self.values = list(boundaries), list(values)

    def __call__(self, step):
        for b, v in zip(self.boundaries, self.values):
            if step <= b:
                return v
        return self.values[-1]


class _Adam:
    """tf.keras.optimizers.Adam (non-amsgrad) dense update."""

    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.learning_rate, self.b1, self.b2, self.eps = learning_rate, beta_1, beta_2, epsilon
        self.iterations = 0
        self._slots = {}

    def apply_gradients(self, grads_and_vars):
        lr = self.learning_rate(self.iterations) if callable(self.learning_rate) else self.learning_rate
        t = self.iterations + 1
        lr_t = lr * _math.sqrt(1 - self.b2 ** t) / (1 - self.b1 ** t)
        with _torch.no_grad():
            for g, v in grads_and_vars:
                if g is None:
                    continue
                g = _tt(g)
                m, vv = self._slots.setdefault(id(v), (_torch.zeros_like(v.t), _torch.zeros_like(v.t)))
                m += (g - m) * (1 - self.b1)
                vv += (g * g - vv) * (1 - self.b2)
                v.t -= lr_t * m / (_torch.sqrt(vv) + self.eps)
        self.iterations = t


def _set_floatx(name):
    global _FLOATX
    assert name == "float64", "shim computes in float64 (every shipped config uses float64)"
    _FLOATX = name


keras = _types.SimpleNamespace(
    Model=_Model,
    layers=_types.SimpleNamespace(BatchNormalization=_BatchNormalization, Dense=_Dense),
    optimizers=_types.SimpleNamespace(
        Adam=_Adam,
        schedules=_types.SimpleNamespace(PiecewiseConstantDecay=_PiecewiseConstantDecay),
    ),
    backend=_types.SimpleNamespace(set_floatx=_set_floatx, floatx=lambda: _FLOATX),
)
