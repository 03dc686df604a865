"""Equation classes with the reference's names and methods (reference equation.py), backed by
libdeeppde_b200: LQR, VDP, ekn (alias EKN), LQR_var.

* ``sample_normal / sample_bounded / sample0`` are the reference's host samplers (NumPy global
  RNG, same stream consumption as equation.py:13-44) -- used for parity; the solver's default
  training path samples on the device instead (Philox, ``dpb_sample_x`` + in-kernel increments).
* closed forms (``V_true, u_true, V_grad_true, Z_tf, w_tf``) and ``propagate_naive / propagate_adaptive``
  run on the GPU through the C ABI.  Arguments may be NumPy arrays or CUDA tensors; results are
  CUDA tensors of the engine's dtype.
"""
from __future__ import annotations

import numpy as np

from . import _cabi
from .engine import Engine, _get


class Equation(object):
    """Base class (equation.py:5-142)."""

    def __init__(self, eqn_config):
        self.eqn_config = eqn_config
        self.dim = _get(eqn_config, "dim")                      # equation.py:8
        self.gamma = _get(eqn_config, "discount")               # equation.py:9
        self.R = _get(eqn_config, "R")                          # equation.py:10
        self.control_dim = _get(eqn_config, "control_dim")      # equation.py:11
        self.sigma_Up = np.sqrt(2.0)
        self._engine = None
        self._engine_kw = {}

    # ---- engine plumbing ---------------------------------------------------------------------
    def bind(self, engine):
        """Use the solver's engine (same dtype/device) for closed forms and propagate."""
        self._engine = engine

    def engine(self):
        if self._engine is None:
            net = {"num_hiddens_actor": [8], "num_hiddens_critic": [8]}
            self._engine = Engine(self.eqn_config, net, {"scheme": "adaptive", "TD_type": "TD1"}, **self._engine_kw)
        return self._engine

    # ---- samplers: host, NumPy global RNG, as the reference (equation.py:13-44) ----------------
    def _ball(self, num_sample):
        r = np.random.uniform(low=0, high=self.R, size=[num_sample, 1])
        r = r ** (1 / self.dim) * (self.R ** ((self.dim - 1) / self.dim))
        g = np.random.standard_normal(size=[num_sample, self.dim])
        return r * g / np.sqrt(np.sum(g ** 2, 1, keepdims=True))

    def _sphere(self, num_sample):
        g = np.random.standard_normal(size=[num_sample, self.dim])
        return self.R * g / np.sqrt(np.sum(np.square(g), 1, keepdims=True))

    def sample_normal(self, num_sample, N):
        x0 = self._ball(num_sample)
        dw_sample = np.random.standard_normal(size=[num_sample, self.dim, N])
        return x0, dw_sample, self._sphere(num_sample)

    def sample_bounded(self, num_sample, N):
        x0 = self._ball(num_sample)
        dw_sample = np.floor((np.random.randint(6, size=[num_sample, self.dim, N]) - 1) / 4) * np.sqrt(3.0)
        return x0, dw_sample, self._sphere(num_sample)

    def sample0(self, num_sample, N):
        x0 = np.zeros(shape=[num_sample, self.dim]) + 0.01
        dw_sample = np.random.standard_normal(size=[num_sample, self.dim, N])
        return x0, dw_sample, self._sphere(num_sample)

    # ---- schemes (equation.py:46-106) -----------------------------------------------------------
    def _propagate(self, scheme, num_sample, x0, dw_sample, NN_control, training, T, N, cheat):
        eng = NN_control.engine if (not cheat and hasattr(NN_control, "engine")) else self.engine()
        if not cheat and not hasattr(NN_control, "theta"):
            raise TypeError("NN_control must be a DeepNN (its weights are evaluated inside the CUDA rollout)")
        if eng.cfg.scheme != _cabi.SCHEME_IDS[scheme]:
            eng = _rebuild(eng, scheme)
        x0d, dwd = eng.tensor(x0), eng.tensor(dw_sample)
        assert x0d.shape[0] == num_sample
        r = eng.critic_step(None if cheat else NN_control.theta, None, None, x0d, dwd, None, N, T,
                            cheat_control=bool(cheat), propagate_only=True, want=("x_smp", "dt", "coef"))
        return r["x_smp"], r["dt"], r["coef"]

    def propagate_naive(self, num_sample, x0, dw_sample, NN_control, training, T, N, cheat):
        return self._propagate("naive", num_sample, x0, dw_sample, NN_control, training, T, N, cheat)

    def propagate_adaptive(self, num_sample, x0, dw_sample, NN_control, training, T, N, cheat):
        return self._propagate("adaptive", num_sample, x0, dw_sample, NN_control, training, T, N, cheat)

    # ---- closed forms -----------------------------------------------------------------------------
    def _cf(self, which, x, u=None):
        eng = self.engine()
        return eng.closed_form(which, eng.tensor(x), None if u is None else eng.tensor(u))

    def w_tf(self, x, u):
        return self._cf(_cabi.CF_W, x, u)

    def Z_tf(self, x):
        return self._cf(_cabi.CF_Z, x)

    def V_true(self, x):
        return self._cf(_cabi.CF_V_TRUE, x)

    def u_true(self, x):
        return self._cf(_cabi.CF_U_TRUE, x)

    def V_grad_true(self, x):
        return self._cf(_cabi.CF_V_GRAD_TRUE, x)

    # ---- SDE coefficients (equation.py:132-142 and the subclasses' :169-176, :229-238, :267-276, :304-311)
    def sigma(self, x, u, num_sample):
        """diffusion coefficient, num_sample x dim x dim_w (diagonal for all four equations)"""
        eng = self.engine()
        xd = eng.tensor(x)
        assert xd.shape[0] == num_sample
        return eng.closed_form(_cabi.CF_SIGMA, xd, None if u is None else eng.tensor(u))

    def drift(self, x, u):
        """drift in the SDE, num_sample x dim"""
        eng = self.engine()
        return eng.closed_form(_cabi.CF_DRIFT, eng.tensor(x), eng.tensor(u))

    def diffusion(self, x, u, dw, num_sample):
        """sigma(x, u) . dw, num_sample x dim"""
        eng = self.engine()
        xd = eng.tensor(x)
        assert xd.shape[0] == num_sample
        return eng.diffusion(xd, None if u is None else eng.tensor(u), eng.tensor(dw))

    def b_np(self, x):  # equation.py:116-118
        return np.sum(x ** 2, axis=1, keepdims=True) - (self.R ** 2)

    def b_tf(self, x):  # equation.py:120-122
        xd = self.engine().tensor(x)
        return (xd * xd).sum(1, keepdim=True) - (self.R ** 2)


def _rebuild(eng, scheme):
    """An engine identical to ``eng`` but for the scheme (cached on the engine)."""
    cache = eng.__dict__.setdefault("_scheme_variants", {})
    if scheme not in cache:
        import copy
        e2 = object.__new__(Engine)
        e2.__dict__.update({k: v for k, v in eng.__dict__.items() if k not in ("handle", "_scheme_variants", "cfg")})
        e2.cfg = copy.copy(eng.cfg)
        e2.cfg.scheme = _cabi.SCHEME_IDS[scheme]
        import ctypes as C
        e2.handle = C.c_void_p()
        _cabi.check(e2.lib, None, e2.lib.dpb_create(C.byref(e2.handle), C.byref(e2.cfg)))
        e2._ws = None
        cache[scheme] = e2
    return cache[scheme]


class LQR(Equation):
    """equation.py:144-176."""

    def __init__(self, eqn_config):
        super(LQR, self).__init__(eqn_config)
        self.p, self.q, self.beta = _get(eqn_config, "p"), _get(eqn_config, "q"), _get(eqn_config, "beta")
        self.k = (((self.gamma ** 2) * (self.q ** 2) + 4 * self.p * self.q * (self.beta ** 2)) ** 0.5 - self.q * self.gamma) / (self.beta ** 2) / 2


class VDP(Equation):
    """equation.py:179-238."""

    def __init__(self, eqn_config):
        super(VDP, self).__init__(eqn_config)
        self.a, self.epsl, self.q = _get(eqn_config, "a"), _get(eqn_config, "epsilon"), _get(eqn_config, "q")


class ekn(Equation):
    """equation.py:240-276.  ``sigma_fix=True`` uses sigma = sqrt(2*epsl) (the dynamics V_true actually
    solves) instead of the reference's sqrt(2)."""

    def __init__(self, eqn_config, sigma_fix=False):
        super(ekn, self).__init__(eqn_config)
        self.a2, self.a3 = _get(eqn_config, "a2"), _get(eqn_config, "a3")
        self.epsl = 1 / 2 / self.a2 / self.dim
        self.sigma_fix = bool(sigma_fix or _get(eqn_config, "sigma_fix", False))
        self._engine_kw = {"ekn_sigma_fix": self.sigma_fix}


EKN = ekn   # the shipped ekn_*.json say "EKN" (configs/ekn_d20.json:4) while the reference class is `ekn`


class LQR_var(Equation):
    """equation.py:278-311."""

    def __init__(self, eqn_config):
        super(LQR_var, self).__init__(eqn_config)
        self.k = (np.sqrt(5) - 1) / 2
        self.q, self.beta, self.epsilon = _get(eqn_config, "q"), _get(eqn_config, "beta"), _get(eqn_config, "epsilon")
