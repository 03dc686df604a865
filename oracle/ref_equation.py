"""Float64 torch restatement of /root/reference/equation.py (TEST ORACLE, not product code).

Every function cites the reference lines it follows.  Tensors are torch (float64 by
default) so that ``torch.autograd`` plays the part of ``tf.GradientTape``; ``sign``,
``floor`` and ``ceil`` have zero gradient in both frameworks.
"""
from __future__ import annotations

import math

import numpy as np
import torch


def _cfg(cfg, key, default=None):
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


class RefEquation:
    """equation.py:5-142 (base class)."""

    name = "base"

    def __init__(self, eqn_config):
        self.dim = int(_cfg(eqn_config, "dim"))                    # equation.py:8
        self.gamma = float(_cfg(eqn_config, "discount"))           # equation.py:9
        self.R = float(_cfg(eqn_config, "R"))                      # equation.py:10
        self.control_dim = int(_cfg(eqn_config, "control_dim"))    # equation.py:11
        self.sigma_Up = math.sqrt(2.0)

    # ---- samplers: equation.py:13-44 (NumPy global RNG, as the reference) -------------
    def sample_normal(self, num_sample, N, rng=np.random):
        r_sample = rng.uniform(low=0, high=self.R, size=[num_sample, 1])          # :14
        r = r_sample ** (1 / self.dim) * (self.R ** ((self.dim - 1) / self.dim))  # :15
        angle = rng.standard_normal(size=[num_sample, self.dim])                  # :16
        norm = np.sqrt(np.sum(angle ** 2, 1, keepdims=True))                      # :17
        x0 = r * angle / norm                                                     # :18
        dw = rng.standard_normal(size=[num_sample, self.dim, N])                  # :19
        xb = rng.standard_normal(size=[num_sample, self.dim])                     # :20
        xb = self.R * xb / np.sqrt(np.sum(np.square(xb), 1, keepdims=True))       # :21-22
        return x0, dw, xb

    def sample_bounded(self, num_sample, N, rng=np.random):
        r_sample = rng.uniform(low=0, high=self.R, size=[num_sample, 1])          # :26
        r = r_sample ** (1 / self.dim) * (self.R ** ((self.dim - 1) / self.dim))  # :27
        angle = rng.standard_normal(size=[num_sample, self.dim])                  # :28
        norm = np.sqrt(np.sum(angle ** 2, 1, keepdims=True))
        x0 = r * angle / norm                                                     # :30
        dw = rng.randint(6, size=[num_sample, self.dim, N])                       # :31
        dw = np.floor((dw - 1) / 4) * np.sqrt(3.0)                                # :32
        xb = rng.standard_normal(size=[num_sample, self.dim])                     # :33
        xb = self.R * xb / np.sqrt(np.sum(np.square(xb), 1, keepdims=True))
        return x0, dw, xb

    def sample0(self, num_sample, N, rng=np.random):
        x0 = np.zeros(shape=[num_sample, self.dim]) + 0.01                        # :39
        dw = rng.standard_normal(size=[num_sample, self.dim, N])                  # :40
        xb = rng.standard_normal(size=[num_sample, self.dim])                     # :41
        xb = self.R * xb / np.sqrt(np.sum(np.square(xb), 1, keepdims=True))
        return x0, dw, xb

    # ---- schemes -----------------------------------------------------------------------
    def propagate_naive(self, x0, dw, control, T, N):
        """equation.py:46-71.  ``control`` maps x[B,d] -> u[B,m]."""
        B = x0.shape[0]
        delta_t = T / N
        sqrt_delta_t = math.sqrt(delta_t)
        xs = [x0]
        coefs = []
        x_i = x0
        flag = torch.ones(B, dtype=x0.dtype)
        for i in range(N):
            u_i = control(x_i)                                                     # :54-57
            delta_x = self.drift(x_i, u_i) * delta_t + self.diffusion(x_i, u_i, dw[:, :, i]) * sqrt_delta_t  # :58
            x_tmp = x_i + delta_x                                                  # :59
            Exit = self.b(x_tmp).reshape(B)                                        # :60
            Exit = torch.ceil((torch.sign(Exit) + 1) / 2)                          # :61  (>=0 -> 1)
            coef_i = flag * (1 - Exit)                                             # :62
            coefs.append(coef_i)
            x_i = x_i + delta_x * coef_i.reshape(B, 1)                             # :67
            xs.append(x_i)
            flag = flag * (1 - Exit)                                               # :69
        dt = torch.ones(B, N, dtype=x0.dtype) * delta_t                            # :70
        return torch.stack(xs, dim=2), dt, torch.stack(coefs, dim=1)

    def _flag(self, xnorm, delta_t):
        # equation.py:80,82 / :94-95
        temp = torch.sign(self.R - xnorm - self.sigma_Up * math.sqrt(3 * self.dim * delta_t)) + torch.sign(self.R - xnorm)
        return 1.0 + torch.floor(temp / 2)

    def propagate_adaptive(self, x0, dw, control, T, N):
        """equation.py:73-106."""
        B = x0.shape[0]
        delta_t = T / N
        xs = [x0]
        coefs, dts = [], []
        x_i = x0
        flag = self._flag(torch.sqrt(torch.sum(x0 ** 2, 1)), delta_t)             # :78-82
        for i in range(N):
            xi_norm = torch.sqrt(torch.sum(x_i ** 2, 1))                           # :84
            dt_i = (2 * flag - flag ** 2) * ((self.R - xi_norm) ** 2) / (3 * self.dim * self.sigma_Up ** 2) \
                + (flag ** 2 - 2 * flag + 1) * delta_t                             # :85
            dt_i = torch.maximum(dt_i, torch.full_like(dt_i, delta_t * 1e-4))      # :86
            u_i = control(x_i)                                                     # :87-90
            delta_x = self.drift(x_i, u_i) * dt_i.reshape(B, 1) \
                + self.diffusion(x_i, u_i, dw[:, :, i]) * torch.sqrt(dt_i).reshape(B, 1)  # :91
            x_tmp = x_i + delta_x                                                  # :92
            tmp_norm = torch.sqrt(torch.sum(x_tmp ** 2, 1))                        # :93
            new_flag = self._flag(tmp_norm, delta_t) * torch.sign(flag)            # :94-95
            coef_i = torch.sign(flag) * torch.sign(new_flag)                       # :96
            coefs.append(coef_i)
            dts.append(dt_i)
            x_i = x_i + delta_x * coef_i.reshape(B, 1)                             # :103
            xs.append(x_i)
            flag = new_flag                                                        # :105
        return torch.stack(xs, dim=2), torch.stack(dts, dim=1), torch.stack(coefs, dim=1)

    def b(self, x):
        return torch.sum(x ** 2, 1, keepdim=True) - self.R ** 2                    # :120-122

    # sigma is diagonal in all four equations (equation.py:170,230,268,305); the dense
    # [B,d,d] matvec of :176 reduces to an elementwise product with the diagonal.
    def sigma_diag(self, x, u):
        raise NotImplementedError

    def diffusion(self, x, u, dw):
        return self.sigma_diag(x, u) * dw


class RefLQR(RefEquation):
    """equation.py:144-176."""
    name = "LQR"

    def __init__(self, c):
        super().__init__(c)
        self.p, self.q, self.beta = float(_cfg(c, "p")), float(_cfg(c, "q")), float(_cfg(c, "beta"))
        self.k = (((self.gamma ** 2) * (self.q ** 2) + 4 * self.p * self.q * (self.beta ** 2)) ** 0.5
                  - self.q * self.gamma) / (self.beta ** 2) / 2                    # :151

    def w(self, x, u):
        return torch.sum(self.p * x ** 2, 1, keepdim=True) + torch.sum(self.q * u ** 2, 1, keepdim=True) \
            - 2 * self.k * self.dim                                                # :155

    def Z(self, x):
        return 0 * torch.sum(x, 1, keepdim=True) + self.k * self.R ** 2            # :158

    def V_true(self, x):
        return torch.sum(x ** 2, 1, keepdim=True) * self.k                         # :161

    def u_true(self, x):
        return -self.beta * self.k / self.q * x                                    # :164

    def V_grad_true(self, x):
        return 2 * self.k * x                                                      # :167

    def sigma_diag(self, x, u):
        return torch.full_like(x, math.sqrt(2.0))                                  # :170

    def drift(self, x, u):
        return self.beta * u                                                       # :173


class RefVDP(RefEquation):
    """equation.py:179-238."""
    name = "VDP"

    def __init__(self, c):
        super().__init__(c)
        self.a, self.epsl, self.q = float(_cfg(c, "a")), float(_cfg(c, "epsilon")), float(_cfg(c, "q"))

    def _split(self, x):
        d = self.control_dim
        x1, x2 = x[:, 0:d], x[:, d:self.dim]
        px1 = torch.cat([x1[:, 1:d], x1[:, 0:1]], 1)                               # :192
        px2 = torch.cat([x2[:, 1:d], x2[:, 0:1]], 1)
        nx1 = torch.cat([x1[:, d - 1:d], x1[:, 0:d - 1]], 1)                       # :194
        nx2 = torch.cat([x2[:, d - 1:d], x2[:, 0:d - 1]], 1)
        return x1, x2, px1, px2, nx1, nx2

    def w(self, x, u):
        x1, x2, px1, px2, nx1, nx2 = self._split(x)
        dv1 = 2 * self.a * x1 - self.epsl * (px1 + nx1)                            # :196
        dv2 = 2 * self.a * x2 - self.epsl * (px2 + nx2)                            # :197
        temp = -self.gamma * self.epsl * (x1 * px1 + x2 * px2) + (dv2 ** 2) / 4 / self.q \
            - x2 * dv1 - ((1 - x1 ** 2) * x2 - x1) * dv2                           # :198
        return torch.sum(temp + self.q * (u ** 2), 1, keepdim=True) \
            + self.gamma * self.a * torch.sum(x ** 2, 1, keepdim=True) - 2 * self.a * self.dim  # :199

    def Z(self, x):
        return self.V_true(x)                                                      # :202

    def V_true(self, x):
        x1, x2, px1, px2, _, _ = self._split(x)
        return self.a * torch.sum(x ** 2, 1, keepdim=True) \
            - self.epsl * torch.sum(x1 * px1 + x2 * px2, 1, keepdim=True)          # :210

    def u_true(self, x):
        _, x2, _, px2, _, nx2 = self._split(x)
        return -(2 * self.a * x2 - self.epsl * (px2 + nx2)) / 2 / self.q           # :217

    def V_grad_true(self, x):
        x1, x2, px1, px2, nx1, nx2 = self._split(x)
        return torch.cat([2 * self.a * x1 - self.epsl * (px1 + nx1),
                          2 * self.a * x2 - self.epsl * (px2 + nx2)], 1)           # :227

    def sigma_diag(self, x, u):
        return torch.full_like(x, math.sqrt(2.0))                                  # :230

    def drift(self, x, u):
        x1 = x[:, 0:self.control_dim]
        x2 = x[:, self.control_dim:self.dim]
        return torch.cat([x2, (1 - x1 ** 2) * x2 - x1 + u], 1)                     # :235


class RefEKN(RefEquation):
    """equation.py:240-276.  ``sigma_fix`` switches sqrt(2) -> sqrt(2*epsl);
    the reference itself is ``sigma_fix=False``."""
    name = "ekn"

    def __init__(self, c, sigma_fix=False):
        super().__init__(c)
        self.a2, self.a3 = float(_cfg(c, "a2")), float(_cfg(c, "a3"))
        self.epsl = 1 / 2 / self.a2 / self.dim                                     # :246
        self.sig = math.sqrt(2.0 * self.epsl) if sigma_fix else math.sqrt(2.0)

    def w(self, x, u):
        return 0 * torch.sum(x, 1, keepdim=True) + 1                               # :250

    def Z(self, x):
        return self.V_true(x)

    def V_true(self, x):
        r = torch.sum(x ** 2, 1, keepdim=True) ** 0.5
        return self.a3 * r ** 3 - self.a2 * r ** 2                                 # :257

    def u_true(self, x):
        r = torch.sum(x ** 2, 1, keepdim=True) ** 0.5
        return x / r                                                               # :261

    def V_grad_true(self, x):
        r = torch.sum(x ** 2, 1, keepdim=True) ** 0.5
        return (3 * self.a3 * r - 2 * self.a2) * x                                 # :265

    def sigma_diag(self, x, u):
        return torch.full_like(x, self.sig)                                        # :268

    def drift(self, x, u):
        r = torch.sum(x ** 2, 1, keepdim=True) ** 0.5
        c = 3 * (self.dim + 1) * self.a3 / 2 / self.a2 / self.dim / (2 * self.a2 - 3 * self.a3 * r)  # :272
        return c * u


class RefLQRVar(RefEquation):
    """equation.py:278-311."""
    name = "LQR_var"

    def __init__(self, c):
        super().__init__(c)
        self.k = (math.sqrt(5) - 1) / 2                                            # :282
        self.q, self.beta, self.epsilon = float(_cfg(c, "q")), float(_cfg(c, "beta")), float(_cfg(c, "epsilon"))

    def w(self, x, u):
        temp = torch.sum(self.k ** 2 * (self.beta + 2 * self.epsilon) ** 2 * x ** 2
                         / (self.q + 2 * self.k * self.epsilon ** 2 * x ** 2), 1, keepdim=True)  # :289
        return temp + torch.sum(self.gamma * self.k * x ** 2 + self.q * u ** 2, 1, keepdim=True) \
            - 2 * self.k * self.dim                                                # :290

    def Z(self, x):
        return 0 * torch.sum(x, 1, keepdim=True) + self.k * self.R ** 2            # :293

    def V_true(self, x):
        return torch.sum(x ** 2, 1, keepdim=True) * self.k                         # :296

    def u_true(self, x):
        return -(self.beta + 2 * self.epsilon) * x / (self.q / self.k + 2 * self.epsilon ** 2 * x ** 2)  # :299

    def V_grad_true(self, x):
        return 2 * self.k * x                                                      # :302

    def sigma_diag(self, x, u):
        return math.sqrt(2.0) * (1 + self.epsilon * x * u)                         # :305 (diagonal)

    def drift(self, x, u):
        return self.beta * u                                                       # :308


def make_ref_equation(eqn_config, ekn_sigma_fix=False):
    name = _cfg(eqn_config, "eqn_name")
    if name == "LQR":
        return RefLQR(eqn_config)
    if name == "VDP":
        return RefVDP(eqn_config)
    if name in ("ekn", "EKN"):      # shipped configs say "EKN", class is "ekn"
        return RefEKN(eqn_config, sigma_fix=ekn_sigma_fix)
    if name == "LQR_var":
        return RefLQRVar(eqn_config)
    raise ValueError(f"unknown eqn_name {name!r}")
