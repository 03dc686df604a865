#!/usr/bin/env python
"""Generate golden vectors from the REFERENCE'S OWN SOURCE (/root/reference/equation.py and
solver.py, imported unmodified) executed under the TensorFlow->torch shim in oracle/tf_shim.

Run here (build container, where /root/reference exists):

    python tests/golden/make_golden.py

Writes tests/golden/*.npz.  Tests never read /root/reference; they read these fixtures.
Weights use the flat layout of oracle/ref_solver.py; inputs come from the reference samplers
under ``np.random.seed``.
"""
import importlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "tf_shim"))
sys.path.insert(0, "/root/reference")

import tensorflow as tf  # noqa: E402  (the shim)
assert "tf_shim" in tf.__file__

ref_eqn = importlib.import_module("equation")
ref_solver = importlib.import_module("solver")
assert ref_eqn.__file__.startswith("/root/reference"), ref_eqn.__file__
assert ref_solver.__file__.startswith("/root/reference"), ref_solver.__file__

from oracle.ref_solver import init_params, net_dims  # noqa: E402  (layout + initialisers only)


class AttrDict(dict):
    """munch stand-in: attribute access over a dict (main.py:33 uses munch.munchify)."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError:
            raise AttributeError(k)
        return AttrDict(v) if isinstance(v, dict) else v


EQN_BASE = {
    "LQR": {"eqn_name": "LQR", "discount": 1.0, "p": 1.0, "q": 1.0, "beta": 1.0, "R": 1.0},
    "VDP": {"eqn_name": "VDP", "discount": 1.0, "a": 1.0, "epsilon": 0.1, "q": 1.0, "R": 1.0},
    "ekn": {"eqn_name": "ekn", "discount": 0, "a2": 1.2, "a3": 0.2, "R": 1.0},
    "LQR_var": {"eqn_name": "LQR_var", "discount": 1.0, "q": 1.0, "beta": 1.0, "epsilon": 0.01, "R": 1.0},
}


def make_config(eqn, dim, cdim, N, T, hid_c, hid_a, scheme, td, sample, train="actor-critic", B=24):
    e = dict(EQN_BASE[eqn])
    e.update({"dim": dim, "control_dim": cdim, "total_time_critic": T, "total_time_actor": T,
              "num_time_interval_critic": N, "num_time_interval_actor": N})
    return {
        "eqn_config": e,
        "net_config": {"num_hiddens_critic": hid_c, "num_hiddens_actor": hid_a,
                       "lr_values_critic": [1e-3, 1e-4], "lr_boundaries_critic": [2],
                       "lr_values_actor": [2e-3, 1e-4], "lr_boundaries_actor": [1],
                       "num_iterations": 3, "batch_size": B, "valid_size": B,
                       "logging_frequency": 1, "dtype": "float64", "verbose": False},
        "train_config": {"sample_type": sample, "scheme": scheme, "TD_type": td, "train": train},
    }


# name -> config.  T is enlarged / x0 pushed outward so that exits and boundary-layer steps occur.
CASES = {
    "lqr_d5_naive_normal_td1": make_config("LQR", 5, 5, 10, 0.1, [24, 16], [24, 16], "naive", "TD1", "normal"),
    "lqr_d5_adaptive_normal_td1": make_config("LQR", 5, 5, 10, 1.5, [24, 16], [24, 16], "adaptive", "TD1", "normal"),
    "vdp_d10_adaptive_bounded_td2": make_config("VDP", 10, 5, 12, 1.2, [16, 16, 16], [16, 16, 16], "adaptive", "TD2", "bounded"),
    "vdp_d4_naive_bounded_td1": make_config("VDP", 4, 2, 8, 0.3, [16, 8], [8, 16], "naive", "TD1", "bounded"),
    "ekn_d6_adaptive_normal_td1": make_config("ekn", 6, 6, 10, 1.0, [16, 16], [16, 16], "adaptive", "TD1", "normal"),
    "ekn_d6_naive_normal_td2": make_config("ekn", 6, 6, 10, 0.05, [16, 16], [16, 16], "naive", "TD2", "normal"),
    "lqr_var_d8_adaptive_normal_td1": make_config("LQR_var", 8, 8, 10, 2.0, [16, 24, 16], [16, 24, 16], "adaptive", "TD1", "normal"),
    "lqr_var_d8_naive_bounded_td2": make_config("LQR_var", 8, 8, 10, 0.05, [16, 24, 16], [16, 24, 16], "naive", "TD2", "bounded"),
}


def flat_from_net(net):
    """Reference DeepNN variables -> flat layout of oracle/ref_solver.py."""
    parts = [net.bn_layers[0].gamma.numpy(), net.bn_layers[0].beta.numpy()]
    L = len(net.dense_layers) - 1
    for i in range(L):
        parts += [net.dense_layers[i].kernel.numpy().ravel(), net.bn_layers[i + 1].gamma.numpy(), net.bn_layers[i + 1].beta.numpy()]
    parts += [net.dense_layers[-1].kernel.numpy().ravel(), net.dense_layers[-1].bias.numpy(),
              net.bn_layers[-1].gamma.numpy(), net.bn_layers[-1].beta.numpy()]
    return np.concatenate(parts)


def load_flat_into_net(net, theta, in_dim, hid, out):
    o = 0

    def take(var, shape):
        nonlocal o
        n = int(np.prod(shape))
        var.assign(theta[o:o + n].reshape(shape))
        o += n

    take(net.bn_layers[0].gamma, [in_dim]); take(net.bn_layers[0].beta, [in_dim])
    prev = in_dim
    for i, h in enumerate(hid):
        take(net.dense_layers[i].kernel, [prev, h]); take(net.bn_layers[i + 1].gamma, [h]); take(net.bn_layers[i + 1].beta, [h])
        prev = h
    take(net.dense_layers[-1].kernel, [prev, out]); take(net.dense_layers[-1].bias, [out])
    take(net.bn_layers[-1].gamma, [out]); take(net.bn_layers[-1].beta, [out])
    assert o == theta.size


def grads_flat(grads, net_list):
    """tape.gradient output (list over model.trainable_variables) -> one flat vector per DeepNN in
    OUR layout.  trainable_variables order = bn_layers (gamma,beta)*, dense_layers (kernel[,bias])*
    per DeepNN, nets in attribute order (solver.py:145-146,239,247)."""
    out = []
    i = 0
    for net in net_list:
        nb, nd = len(net.bn_layers), len(net.dense_layers)
        bn = []
        for _ in range(nb):
            bn.append((grads[i], grads[i + 1])); i += 2
        dn = []
        for j in range(nd):
            if j == nd - 1:
                dn.append((grads[i], grads[i + 1])); i += 2
            else:
                dn.append((grads[i],)); i += 1

        def z(g, like):
            return np.zeros(like.shape) if g is None else g.numpy()

        parts = [z(bn[0][0], net.bn_layers[0].gamma), z(bn[0][1], net.bn_layers[0].beta)]
        for j in range(nd - 1):
            parts += [z(dn[j][0], net.dense_layers[j].kernel).ravel(), z(bn[j + 1][0], net.bn_layers[j + 1].gamma), z(bn[j + 1][1], net.bn_layers[j + 1].beta)]
        parts += [z(dn[-1][0], net.dense_layers[-1].kernel).ravel(), z(dn[-1][1], net.dense_layers[-1].bias),
                  z(bn[-1][0], net.bn_layers[-1].gamma), z(bn[-1][1], net.bn_layers[-1].beta)]
        out.append(np.concatenate(parts))
    assert i == len(grads)
    return out


def run_case(name, cfg, seed):
    config = AttrDict(cfg)
    # the reference resolves the class by name (main.py:34); "ekn" is the class that exists (Q1)
    bsde = getattr(ref_eqn, config.eqn_config.eqn_name)(config.eqn_config)
    solver = ref_solver.ActorCriticSolver(config, bsde)
    B = cfg["net_config"]["batch_size"]
    N = cfg["eqn_config"]["num_time_interval_critic"]
    T = cfg["eqn_config"]["total_time_critic"]
    dim, m = cfg["eqn_config"]["dim"], cfg["eqn_config"]["control_dim"]

    np.random.seed(seed)
    x0, dw, xb = solver.sample(B, N)                      # reference sampler, numpy global RNG
    # build all variables (Keras builds lazily) then inject seeded weights
    nets = {"actor": solver.model_actor.NN_control, "critic": solver.model_critic.NN_value,
            "critic_grad": solver.model_critic.NN_value_grad}
    rng = np.random.RandomState(1000 + seed)
    thetas = {}
    for k, net in nets.items():
        net(x0, False, need_grad=False)
        i, h, o, _ = net_dims(cfg, k)
        thetas[k] = init_params(i, h, o, rng)
        # the reference initialises the last bias to zeros; use a non-zero one so that every
        # term of the layout is exercised
        thetas[k][-3 * o:-2 * o] = rng.normal(0.0, 0.1, o)
        load_flat_into_net(net, thetas[k], i, h, o)
        assert np.allclose(flat_from_net(net), thetas[k])

    out = {"x0": x0, "dw": dw, "xb": xb, "config_json": np.array(json.dumps(cfg))}
    for k in thetas:
        out["theta_" + k] = thetas[k]
    inputs = (x0, dw, xb)

    # raw network outputs on x0
    for k, net in nets.items():
        out["net_" + k] = net(x0, False, need_grad=False).numpy()
    # closed forms
    u_true = bsde.u_true(x0)
    out["u_true"] = np.asarray(u_true.numpy() if hasattr(u_true, "numpy") else u_true)
    for fn in ("V_true", "V_grad_true", "Z_tf"):
        v = getattr(bsde, fn)(xb if fn == "Z_tf" else x0)
        out[fn] = v.numpy() if hasattr(v, "numpy") else np.asarray(v)
    u_nn = nets["actor"](x0, False, need_grad=False)
    w = bsde.w_tf(x0, u_nn)
    out["w_tf"] = w.numpy()
    # SDE coefficients (equation.py:169-176,229-238,267-276,304-311) at (x0, NN control, first increment)
    npy = lambda v: v.numpy() if hasattr(v, "numpy") else np.asarray(v)
    out["sigma"] = npy(bsde.sigma(x0, u_nn, B))
    out["drift"] = npy(bsde.drift(x0, u_nn))
    out["diffusion"] = npy(bsde.diffusion(x0, u_nn, dw[:, :, 0], B))
    # CriticModel.control (solver.py:153-157), both settings
    out["control_nn"] = npy(solver.model_critic.control(x0, False, solver.model_actor))
    out["control_cheat"] = npy(solver.model_critic.control(x0, True, solver.model_actor))

    # propagate (both cheat settings) through the scheme the config selects
    prop = solver.model_critic.propagate
    for cheat in (False, True):
        xs, dt, coef = prop(B, x0, dw, nets["actor"], False, T, N, cheat)
        tag = "cheat" if cheat else "nn"
        out[f"prop_{tag}_x"] = xs.numpy()
        out[f"prop_{tag}_dt"] = np.asarray(dt.numpy() if hasattr(dt, "numpy") else dt)
        out[f"prop_{tag}_coef"] = coef.numpy()

    # critic: residuals, loss, gradient
    for cheat in (False, True):
        tag = "cheat" if cheat else "nn"
        delta, delta_b = solver.model_critic(inputs, solver.model_actor, False, cheat)
        out[f"critic_{tag}_delta"] = delta.numpy()
        out[f"critic_{tag}_delta_bdry"] = delta_b.numpy()
        out[f"critic_{tag}_loss"] = solver.loss_critic(inputs, False, cheat).numpy()
        g = solver.grad_critic(inputs, False, cheat)
        gV, gG = grads_flat(g, [nets["critic"], nets["critic_grad"]])
        out[f"critic_{tag}_grad_V"] = gV
        out[f"critic_{tag}_grad_G"] = gG

    # actor: cost, loss, gradient (cheat_value both ways; cheat_control False as in train_step_actor)
    for cheat_v in (False, True):
        tag = "cheatV" if cheat_v else "nn"
        y = solver.model_actor(inputs, solver.model_critic, False, cheat_v, False)
        out[f"actor_{tag}_y"] = y.numpy()
        out[f"actor_{tag}_loss"] = solver.loss_actor(inputs, False, cheat_v, False).numpy()
        g = solver.grad_actor(inputs, False, cheat_v, False)
        (gA,) = grads_flat(g, [nets["actor"]])
        out[f"actor_{tag}_grad"] = gA
    out["actor_true_loss"] = solver.loss_actor(inputs, False, True, True).numpy()

    # error metrics (solver.py:109-136)
    out["err_value"] = solver.err_value(inputs).numpy()
    out["err_control"] = solver.err_control(inputs).numpy()
    out["err_value_grad"] = solver.err_value_grad(inputs).numpy()
    out["err_value_infty"] = solver.err_value_infty(inputs).numpy()
    out["err_cost"] = solver.err_cost(inputs).numpy()

    # three training iterations on fixed data: exercises Keras Adam + PiecewiseConstantDecay
    solver.cheat_value_in_actor = False
    solver.cheat_control_in_critic = False
    for _ in range(3):
        solver.train_step_critic(inputs)
        solver.train_step_actor(inputs)
    for k, net in nets.items():
        out["theta_after3_" + k] = flat_from_net(net)

    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: live-frac nn={out['prop_nn_coef'].mean():.3f} cheat={out['prop_cheat_coef'].mean():.3f} "
          f"loss_c={float(out['critic_nn_loss']):.4e} loss_a={float(out['actor_nn_loss']):.4e}")


def sampler_stream_fixture():
    """Pin the reference samplers (equation.py:13-44) under np.random.seed for the product's
    host-side samplers: scipy's multivariate_normal.rvs must consume the same stream."""
    cfg = AttrDict(make_config("LQR", 5, 5, 4, 0.2, [8], [8], "naive", "TD1", "normal"))
    bsde = ref_eqn.LQR(cfg.eqn_config)
    out = {}
    for fn in ("sample_normal", "sample_bounded", "sample0"):
        np.random.seed(7)
        x0, dw, xb = getattr(bsde, fn)(6, 4)
        out[fn + "_x0"], out[fn + "_dw"], out[fn + "_xb"] = x0, dw, xb
    np.savez_compressed(os.path.join(HERE, "samplers_seed7.npz"), **out)


if __name__ == "__main__":
    for n, (name, cfg) in enumerate(CASES.items()):
        run_case(name, cfg, seed=11 + n)
    sampler_stream_fixture()
