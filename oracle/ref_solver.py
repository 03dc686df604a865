"""Float64 torch restatement of /root/reference/solver.py:73-278 (TEST ORACLE, not product).

Weights are injectable as one flat vector per network.  Flat layout (shared *by
specification* with the product, see DESIGN.md "parameter layout"):

    bn0.gamma[in] bn0.beta[in]
    for each hidden layer i: W_i[in_i, H_i] (row-major, Keras kernel orientation)
                             bn_i.gamma[H_i] bn_i.beta[H_i]
    W_last[H_L, out] (row-major)  b_last[out]  bn_last.gamma[out] bn_last.beta[out]

All BatchNorm layers run with ``training=False`` and never-updated moving statistics
(mean 0, var 1) at every call site of the reference (solver.py:42,46,47,101,106),
so each is the affine map  z -> z * gamma / sqrt(1 + 1e-6) + beta  (solver.py:239-246).
"""
from __future__ import annotations

import math

import numpy as np
import torch

DELTA_CLIP = 50.0          # solver.py:5
BN_EPS = 1e-6              # solver.py:242


# ------------------------------------------------------------------------------- networks
def net_dims(config, kind):
    """(in_dim, hiddens, out_dim, ekn_head) for kind in {"actor","critic","critic_grad"}; solver.py:228-258."""
    eqn, net = config["eqn_config"], config["net_config"]
    dim, m = int(eqn["dim"]), int(eqn["control_dim"])
    if kind == "actor":
        hid = list(net["num_hiddens_actor"])                       # :236
    else:
        hid = list(net["num_hiddens_critic"])                      # :238
    ekn_head = False
    if kind == "critic":
        out = 1                                                    # :252
    elif kind == "critic_grad":
        out = dim                                                  # :254
    elif eqn["eqn_name"] in ("ekn", "EKN"):
        out, ekn_head = m + 1, True                                # :256  (Q1: both spellings)
    else:
        out = m                                                    # :258
    return dim, hid, out, ekn_head


def param_count(in_dim, hid, out):
    n = 2 * in_dim
    prev = in_dim
    for h in hid:
        n += prev * h + 2 * h
        prev = h
    n += prev * out + 3 * out
    return n


def unflatten(theta, in_dim, hid, out):
    """Views into the flat vector following the layout in the module docstring."""
    o = 0

    def take(*shape):
        nonlocal o
        n = int(np.prod(shape))
        v = theta[o:o + n].reshape(*shape)
        o += n
        return v

    p = {"g0": take(in_dim), "b0": take(in_dim), "hidden": []}
    prev = in_dim
    for h in hid:
        p["hidden"].append((take(prev, h), take(h), take(h)))
        prev = h
    p["W"] = take(prev, out)
    p["b"] = take(out)
    p["g"] = take(out)
    p["be"] = take(out)
    assert o == theta.numel()
    return p


def init_params(in_dim, hid, out, rng):
    """Reference initialisers (solver.py:239-258): BN gamma~U(0.1,0.5), beta~N(0,0.1^2); Dense
    kernels Glorot-uniform, last bias zeros.  ``rng`` is a numpy RandomState."""
    parts = [rng.uniform(0.1, 0.5, in_dim), rng.normal(0.0, 0.1, in_dim)]
    prev = in_dim
    for h in hid:
        lim = math.sqrt(6.0 / (prev + h))
        parts += [rng.uniform(-lim, lim, prev * h), rng.uniform(0.1, 0.5, h), rng.normal(0.0, 0.1, h)]
        prev = h
    lim = math.sqrt(6.0 / (prev + out))
    parts += [rng.uniform(-lim, lim, prev * out), np.zeros(out), rng.uniform(0.1, 0.5, out), rng.normal(0.0, 0.1, out)]
    return np.concatenate(parts)


def mlp(theta, x, in_dim, hid, out, ekn_head=False, d_ctrl=None):
    """DeepNN.call, solver.py:260-278, with BN == fixed affine (Q4)."""
    p = unflatten(theta, in_dim, hid, out)
    c = 1.0 / math.sqrt(1.0 + BN_EPS)
    y = x * (p["g0"] * c) + p["b0"]                                # :265
    for W, g, b in p["hidden"]:
        y = y @ W                                                  # :267
        y = y * (g * c) + b                                        # :268
        y = y + torch.relu(y)                                      # :269
    y = y @ p["W"] + p["b"]                                        # :270
    y = y * (p["g"] * c) + p["be"]                                 # :271
    if ekn_head:                                                   # :272-274
        d = d_ctrl
        norm_y = torch.sum(y[:, 0:d] ** 2, 1, keepdim=True) ** 0.5
        y = y[:, 0:d] / (0.000000000000001 + torch.relu(y[:, d:d + 1]) + norm_y)
    return y


class RefNets:
    """The three networks of the reference (solver.py:145-146,200) with flat weights."""

    def __init__(self, config, theta_actor, theta_V, theta_G):
        self.config = config
        self.dims = {k: net_dims(config, k) for k in ("actor", "critic", "critic_grad")}
        self.theta = {"actor": theta_actor, "critic": theta_V, "critic_grad": theta_G}
        self.m = int(config["eqn_config"]["control_dim"])

    def __call__(self, kind, x):
        i, h, o, ek = self.dims[kind]
        return mlp(self.theta[kind], x, i, h, o, ek, self.m)


# ----------------------------------------------------------------------------- loss graphs
def _propagate(eqn, config, x0, dw, control, T, N):
    if config["train_config"]["scheme"] == "naive":                # solver.py:148-151
        return eqn.propagate_naive(x0, dw, control, T, N)
    return eqn.propagate_adaptive(x0, dw, control, T, N)


def critic_delta(eqn, config, nets, inputs, cheat_control):
    """CriticModel.call, solver.py:159-191.  Returns (delta, delta_bdry, aux)."""
    x0, dw, xb = inputs
    ec = config["eqn_config"]
    T, N = float(ec["total_time_critic"]), int(ec["num_time_interval_critic"])
    control = (lambda x: eqn.u_true(x)) if cheat_control else (lambda x: nets("actor", x))  # :153-157
    x, dt, coef = _propagate(eqn, config, x0, dw, control, T, N)   # :165
    y = 0
    discount = torch.ones(x0.shape[0], 1, dtype=x0.dtype)          # :164
    td1 = config["train_config"]["TD_type"] == "TD1"
    for t in range(N):
        xt = x[:, :, t]
        u = control(xt)                                            # :167
        w = eqn.w(xt, u)                                           # :168
        y = y + w * discount * coef[:, t:t + 1] * dt[:, t:t + 1]   # :170-174
        if td1:                                                    # :177-184
            dif = eqn.sigma_diag(xt, u) * dw[:, :, t]
            dif = torch.sum(dif * nets("critic_grad", xt), 1, keepdim=True) * discount
            y = y - dif * coef[:, t:t + 1] * torch.sqrt(dt[:, t:t + 1])
        discount = discount * torch.exp(-eqn.gamma * dt[:, t:t + 1] * coef[:, t:t + 1])    # :187
    delta = nets("critic", x[:, :, 0]) - y - nets("critic", x[:, :, -1]) * discount         # :189
    delta_bdry = nets("critic", xb) - eqn.Z(xb)                    # :190
    return delta, delta_bdry, {"x": x, "dt": dt, "coef": coef, "y": y, "discount": discount}


def actor_cost(eqn, config, nets, inputs, cheat_value, cheat_control):
    """ActorModel.call, solver.py:207-224.  Returns (y, aux)."""
    x0, dw, xb = inputs
    ec = config["eqn_config"]
    T, N = float(ec["total_time_actor"]), int(ec["num_time_interval_actor"])
    control = (lambda x: eqn.u_true(x)) if cheat_control else (lambda x: nets("actor", x))
    x, dt, coef = _propagate(eqn, config, x0, dw, control, T, N)   # :211
    y = 0
    discount = torch.ones(x0.shape[0], 1, dtype=x0.dtype)
    for t in range(N):
        xt = x[:, :, t]
        w = eqn.w(xt, control(xt))                                 # :214-217
        y = y + coef[:, t:t + 1] * w * dt[:, t:t + 1] * discount   # :218
        discount = discount * torch.exp(-eqn.gamma * dt[:, t:t + 1] * coef[:, t:t + 1])   # :219
    if cheat_value:
        y = y + eqn.V_true(x[:, :, -1]) * discount                 # :223
    else:
        y = y + nets("critic", x[:, :, -1]) * discount             # :221
    return y, {"x": x, "dt": dt, "coef": coef, "discount": discount}


def huber_clip(delta):
    """solver.py:76-77."""
    return torch.where(torch.abs(delta) < DELTA_CLIP, delta ** 2, 2 * DELTA_CLIP * torch.abs(delta) - DELTA_CLIP ** 2)


def loss_critic(eqn, config, nets, inputs, cheat_control):
    delta, delta_bdry, aux = critic_delta(eqn, config, nets, inputs, cheat_control)
    loss = (torch.mean(huber_clip(delta)) + torch.mean(huber_clip(delta_bdry))) * 100      # :76-78
    return loss, delta, delta_bdry, aux


def loss_actor(eqn, config, nets, inputs, cheat_value, cheat_control):
    y, aux = actor_cost(eqn, config, nets, inputs, cheat_value, cheat_control)
    return torch.mean(y), y, aux                                   # :82


def grad_critic(eqn, config, thetas, inputs, cheat_control):
    """solver.py:85-90: d loss_critic / d (theta_V, theta_G).  Unused nets (TD2) -> zeros."""
    th = {k: v.detach().clone().requires_grad_(k != "actor") for k, v in thetas.items()}
    nets = RefNets(config, th["actor"], th["critic"], th["critic_grad"])
    loss, delta, delta_bdry, aux = loss_critic(eqn, config, nets, inputs, cheat_control)
    gV, gG = torch.autograd.grad(loss, [th["critic"], th["critic_grad"]], allow_unused=True)
    gV = torch.zeros_like(th["critic"]) if gV is None else gV
    gG = torch.zeros_like(th["critic_grad"]) if gG is None else gG
    return loss.detach(), gV, gG, delta.detach(), delta_bdry.detach(), aux


def grad_actor(eqn, config, thetas, inputs, cheat_value, cheat_control=False):
    """solver.py:92-97: d loss_actor / d theta_actor (BPTT through the rollout)."""
    th = {k: v.detach().clone().requires_grad_(k == "actor") for k, v in thetas.items()}
    nets = RefNets(config, th["actor"], th["critic"], th["critic_grad"])
    loss, y, aux = loss_actor(eqn, config, nets, inputs, cheat_value, cheat_control)
    (gA,) = torch.autograd.grad(loss, [th["actor"]], allow_unused=True)
    gA = torch.zeros_like(th["actor"]) if gA is None else gA
    return loss.detach(), gA, y.detach(), aux


# ------------------------------------------------------------------------------- optimizer
class RefKerasAdam:
    """tf.keras.optimizers.Adam(lr=PiecewiseConstantDecay(boundaries, values), epsilon=1e-8)
    as used at solver.py:16-21:  t += 1;  lr_t = lr(t-1) * sqrt(1-b2^t)/(1-b1^t);
    m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  theta -= lr_t * m / (sqrt(v) + eps).
    PiecewiseConstantDecay: values[0] while step <= boundaries[0], ... (step = iterations
    counter *before* the update)."""

    def __init__(self, n, boundaries, values, beta1=0.9, beta2=0.999, eps=1e-8, dtype=torch.float64):
        self.m = torch.zeros(n, dtype=dtype)
        self.v = torch.zeros(n, dtype=dtype)
        self.t = 0
        self.boundaries, self.values = list(boundaries), list(values)
        self.b1, self.b2, self.eps = beta1, beta2, eps

    def lr(self, step):
        for b, v in zip(self.boundaries, self.values):
            if step <= b:
                return v
        return self.values[-1]

    def step(self, theta, grad):
        lr = self.lr(self.t)
        self.t += 1
        self.m = self.b1 * self.m + (1 - self.b1) * grad
        self.v = self.b2 * self.v + (1 - self.b2) * grad * grad
        lr_t = lr * math.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
        return theta - lr_t * self.m / (torch.sqrt(self.v) + self.eps)


# ------------------------------------------------------------------- whole train iteration
class RefSolver:
    """Minimal restatement of ActorCriticSolver (solver.py:7-136) around the functions above;
    used as the CPU baseline (bench.py) and for short convergence checks."""

    def __init__(self, config, eqn, seed=0, dtype=torch.float64):
        self.config, self.eqn, self.dtype = config, eqn, dtype
        rng = np.random.RandomState(seed)
        self.thetas = {}
        for k in ("actor", "critic", "critic_grad"):
            i, h, o, _ = net_dims(config, k)
            self.thetas[k] = torch.tensor(init_params(i, h, o, rng), dtype=dtype)
        net = config["net_config"]
        nV, nG = self.thetas["critic"].numel(), self.thetas["critic_grad"].numel()
        self.opt_critic = RefKerasAdam(nV + nG, net["lr_boundaries_critic"], net["lr_values_critic"], dtype=dtype)
        self.opt_actor = RefKerasAdam(self.thetas["actor"].numel(), net["lr_boundaries_actor"], net["lr_values_actor"], dtype=dtype)
        self.sample = eqn.sample_normal if config["train_config"]["sample_type"] == "normal" else eqn.sample_bounded
        tr = config["train_config"]["train"]
        self.cheat_control_in_critic = tr == "critic"              # solver.py:28-34
        self.cheat_value_in_actor = tr == "actor"

    def _t(self, arrs):
        return tuple(torch.as_tensor(a, dtype=self.dtype) for a in arrs)

    def train_step_critic(self, data):
        _, gV, gG, _, _, _ = grad_critic(self.eqn, self.config, self.thetas, self._t(data), self.cheat_control_in_critic)
        nV = gV.numel()
        new = self.opt_critic.step(torch.cat([self.thetas["critic"], self.thetas["critic_grad"]]), torch.cat([gV, gG]))
        self.thetas["critic"], self.thetas["critic_grad"] = new[:nV].clone(), new[nV:].clone()

    def train_step_actor(self, data):
        _, gA, _, _ = grad_actor(self.eqn, self.config, self.thetas, self._t(data), self.cheat_value_in_actor, False)
        self.thetas["actor"] = self.opt_actor.step(self.thetas["actor"], gA)

    def train_iteration(self):
        """solver.py:67-70 (one loop body, host sampling included)."""
        ec, nc, tr = self.config["eqn_config"], self.config["net_config"], self.config["train_config"]["train"]
        if tr in ("actor-critic", "critic"):
            self.train_step_critic(self.sample(nc["batch_size"], ec["num_time_interval_critic"]))
        if tr in ("actor-critic", "actor"):
            self.train_step_actor(self.sample(nc["batch_size"], ec["num_time_interval_actor"]))

    def errors(self, x0):
        """solver.py:109-130 on a validation x0."""
        x0 = torch.as_tensor(x0, dtype=self.dtype)
        nets = RefNets(self.config, self.thetas["actor"], self.thetas["critic"], self.thetas["critic_grad"])
        with torch.no_grad():
            Vt, ut, Gt = self.eqn.V_true(x0), self.eqn.u_true(x0), self.eqn.V_grad_true(x0)
            V, u, G = nets("critic", x0), nets("actor", x0), nets("critic_grad", x0)
            return {
                "err_value": float(torch.sqrt(torch.sum((Vt - V) ** 2) / torch.sum(Vt ** 2))),
                "err_control": float(torch.sqrt(torch.sum((ut - u) ** 2) / torch.sum(ut ** 2))),
                "err_value_grad": float(torch.sqrt(torch.sum((Gt - G) ** 2) / torch.sum(Gt ** 2))),
                "err_value_infty": float(torch.max(torch.abs(Vt - V))),
            }
