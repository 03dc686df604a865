"""FLOAT32 restatement of the reference's two Euler-Maruyama schemes under the true control (TEST ORACLE, not product).

Follows /root/reference/equation.py line by line -- propagate_naive :46-71, propagate_adaptive :73-106, the four
equations' u_true / drift / sigma (:163-176, :212-238, :259-276, :298-311) -- in NumPy float32, vectorised over the paths
and SEQUENTIAL over the state components (sums over the d components are accumulated left to right in float32, every
product and sum is a separate IEEE operation, no fused multiply-add).  That is the evaluation order any scalar FP32
implementation of the reference's formulas has, and it is what the CUDA kernels must reproduce BIT FOR BIT on
`coef`, `dt`, the exit index and every state (BASELINE.json north_star: "exit-step indices and the adaptive step
schedule must be bit-exact").  Constants that depend only on the configuration are evaluated in float64 (as Python /
NumPy do in the reference's constructors, equation.py:146-156,181-187,242-247,280-287) and then rounded to float32 once.

Written independently of deeppde_actorcritic_b200/csrc/dpb_eqn.h: tests/test_host_math.py checks the g++ build of that
header against this file on the CPU, tests/test_gpu_parity.py checks both CUDA implementations against it on the GPU.
Pinned to the reference by the float64 oracle it shadows: for the same inputs its exit pattern equals ref_equation's
(float64, itself pinned by tests/golden) except where a proposal lies within float32 rounding of the boundary.
"""
from __future__ import annotations

import math

import numpy as np

F = np.float32


def _c(eqn_config, key, default=0.0):
    v = eqn_config.get(key, default)
    return float(v if v is not None else default)


class ScheduleF32:
    def __init__(self, eqn_config, scheme, T, N, ekn_sigma_fix=False):
        e = eqn_config
        self.name = {"EKN": "ekn"}.get(e["eqn_name"], e["eqn_name"])
        self.d, self.m = int(e["dim"]), int(e["control_dim"])
        self.scheme, self.N = scheme, int(N)
        d = float(self.d)
        R, gamma = _c(e, "R"), _c(e, "discount")
        self.R, self.R2 = F(R), F(R * R)
        sig = math.sqrt(2.0)                                           # equation.py:170,230,268,305
        beta, q, p = _c(e, "beta"), _c(e, "q"), _c(e, "p")
        if self.name == "LQR":                                         # equation.py:151,164
            k = (math.sqrt(gamma * gamma * q * q + 4.0 * p * q * (beta * beta)) - q * gamma) / (beta * beta) / 2.0
            self.cu, self.beta = F(-beta * k / q), F(beta)
        elif self.name == "VDP":                                       # equation.py:183-187
            self.a, self.eps, self.q = F(_c(e, "a")), F(_c(e, "epsilon")), F(q)
        elif self.name == "ekn":                                       # equation.py:244-246,272
            a2, a3 = _c(e, "a2"), _c(e, "a3")
            self.a2, self.a3 = F(a2), F(a3)
            self.C0 = F(3.0 * (d + 1.0) * a3 / 2.0 / a2 / d)
            if ekn_sigma_fix:
                sig = math.sqrt(2.0 * (1.0 / 2.0 / a2 / d))
        else:                                                          # LQR_var, equation.py:282-285,299,305
            k = (math.sqrt(5.0) - 1.0) / 2.0
            eps = _c(e, "epsilon")
            self.beta, self.eps = F(beta), F(eps)
            self.lv_un, self.lv_ud, self.lv_ue = F(beta + 2.0 * eps), F(q / k), F(2.0 * (eps * eps))
        self.sig = F(sig)
        delta_t = float(T) / float(N)                                  # equation.py:48,75
        sigU = math.sqrt(2.0)                                          # equation.py:8 (sigma_Up)
        self.delta_t, self.sqrt_delta_t = F(delta_t), F(math.sqrt(delta_t))
        self.hb = F(sigU * math.sqrt(3.0 * d * delta_t))               # equation.py:80
        self.c3 = F((3.0 * d) * (sigU * sigU))                         # equation.py:85
        self.hmin = F(delta_t * 1e-4)                                  # equation.py:86

    # ---- pieces ------------------------------------------------------------------------------------------------
    def norm2(self, x):
        s = np.zeros(x.shape[0], F)
        for k in range(self.d):
            s = s + x[:, k] * x[:, k]
        return s

    def u_true(self, x):
        d, m = self.d, self.m
        u = np.zeros((x.shape[0], m), F)
        if self.name == "LQR":                                         # equation.py:163-164
            for k in range(d):
                u[:, k] = self.cu * x[:, k]
        elif self.name == "VDP":                                       # equation.py:212-217
            for j in range(m):
                x2, px2, nx2 = x[:, m + j], x[:, m + (j + 1) % m], x[:, m + (j - 1) % m]
                u[:, j] = -((F(2) * self.a) * x2 - self.eps * (px2 + nx2)) / F(2) / self.q
        elif self.name == "ekn":                                       # equation.py:259-261
            r = np.sqrt(self.norm2(x))
            for k in range(d):
                u[:, k] = x[:, k] / r
        else:                                                          # equation.py:298-299
            for k in range(d):
                xk = x[:, k]
                u[:, k] = (-self.lv_un) * xk / (self.lv_ud + (self.lv_ue * xk) * xk)
        return u

    def flag(self, nrm):                                               # equation.py:80-82,94-95
        t2 = self.R - nrm
        t1 = self.R - nrm - self.hb
        return np.where(t2 > 0, np.where(t1 > 0, 2, 1), 0).astype(np.int32)

    def _drift(self, cc, x, u, k):
        if self.name in ("LQR", "LQR_var"):                            # equation.py:172,307
            return self.beta * u[:, k]
        if self.name == "VDP":                                         # equation.py:232-235
            m = self.m
            if k < m:
                return x[:, m + k]
            j = k - m
            x1, x2 = x[:, j], x[:, m + j]
            return (F(1) - x1 * x1) * x2 - x1 + u[:, j]
        return cc * u[:, k]                                            # equation.py:270-273

    def _sigma(self, x, u, k):
        if self.name == "LQR_var":                                     # equation.py:305
            return self.sig * (F(1) + self.eps * x[:, k] * u[:, k])
        return np.full(x.shape[0], self.sig, F)

    # ---- the schemes -------------------------------------------------------------------------------------------
    def propagate(self, x0, dw):
        """x0 [B,d], dw [B,d,N] float32 -> x_smp [B,d,N+1], dt [B,N], coef [B,N] (float32), exit index [B] (int)"""
        with np.errstate(all="ignore"):
            return self._propagate(np.asarray(x0, F), np.asarray(dw, F))

    def _propagate(self, x0, dw):
        B, d, N = x0.shape[0], self.d, self.N
        adaptive = self.scheme == "adaptive"
        x = x0.copy()
        xs = np.zeros((B, d, N + 1), F)
        dts, coefs = np.zeros((B, N), F), np.zeros((B, N), F)
        xs[:, :, 0] = x
        if adaptive:
            flag = self.flag(np.sqrt(self.norm2(x)))                   # equation.py:80-82
        else:
            flag = np.ones(B, np.int32)                                # equation.py:51
        for t in range(N):
            xnorm = np.sqrt(self.norm2(x)) if adaptive else np.zeros(B, F)
            if adaptive:                                               # equation.py:84-86
                g = self.R - xnorm
                dt = np.where(flag == 1, g * g / self.c3, self.delta_t).astype(F)
                dt = np.where(dt > self.hmin, dt, self.hmin).astype(F)
            else:
                dt = np.full(B, self.delta_t, F)                       # equation.py:48
            sqdt = np.sqrt(dt) if adaptive else np.full(B, self.sqrt_delta_t, F)
            u = self.u_true(x)                                         # equation.py:54,87 (cheat)
            cc = None
            if self.name == "ekn":
                r = xnorm if adaptive else np.sqrt(self.norm2(x))
                cc = self.C0 / (F(2) * self.a2 - (F(3) * self.a3) * r)
            n2 = np.zeros(B, F)
            dk = np.zeros((B, d), F)
            for k in range(d):                                         # equation.py:58-60,91-93
                sd = self._sigma(x, u, k) * dw[:, k, t]
                dk[:, k] = self._drift(cc, x, u, k) * dt + sd * sqdt
                pk = x[:, k] + dk[:, k]
                n2 = n2 + pk * pk
            if adaptive:                                               # equation.py:94-98
                nf = self.flag(np.sqrt(n2))
                newflag = np.where(flag > 0, nf, 0).astype(np.int32)
                coef = ((flag > 0) & (newflag > 0)).astype(np.int32)
            else:                                                      # equation.py:61-65
                ex = (n2 - self.R2 >= 0).astype(np.int32)
                coef = flag * (1 - ex)
                newflag = coef
            x = np.where(coef[:, None] > 0, x + dk, x).astype(F)       # equation.py:66,99
            flag = newflag
            xs[:, :, t + 1] = x
            dts[:, t] = dt
            coefs[:, t] = coef
        return xs, dts, coefs, coefs.sum(1).astype(np.int32)
