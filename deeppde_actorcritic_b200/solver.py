"""ActorCriticSolver / CriticModel / ActorModel / DeepNN with the reference's names, arguments and
return values (reference solver.py), every arithmetic step running in libdeeppde_b200.

Differences from the reference that a user can see:
  * weights live in one flat device vector per network (layout: include/deeppde_b200.h);
  * ``net_config.dtype`` ("float64" in every shipped config) is honoured only with
    ``compute_dtype="config"``; the default compute type is float32 (BASELINE.json north_star);
  * ``train_config.sampler`` (optional, default "device"): "device" draws x0/x_bdry and the Brownian
    increments on the GPU (Philox4x32-10 keyed by the GLOBAL path index, so 1/2/4/8-GPU runs see the
    same paths); "host" uses the reference's NumPy samplers (equation.py:13-44);
  * with ``torch.distributed`` initialised the batch is sharded over ranks and each phase ends
    with ONE sum-all-reduce of [flat gradient | loss scalars]; Adam is replicated.
"""
from __future__ import annotations

import logging
import math
import os
import time

import numpy as np
import torch

from . import _cabi
from .engine import Engine, _get

DELTA_CLIP = 50.0   # solver.py:5 (applied inside the critic kernel)


def _rng_to_plain(st):
    """np.random.get_state() as plain Python / torch objects (loadable with weights_only=True)"""
    return {"name": st[0], "keys": torch.from_numpy(np.asarray(st[1], dtype=np.int64)), "pos": int(st[2]), "has_gauss": int(st[3]), "gauss": float(st[4])}


def _rng_from_plain(d):
    return (d["name"], np.asarray(d["keys"].numpy(), dtype=np.uint32), d["pos"], d["has_gauss"], d["gauss"])


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist
    return None


class DeepNN(object):
    """solver.py:227-278.  ``theta`` is the flat parameter vector (device)."""

    def __init__(self, config, AC, engine, rng=None):
        self.AC = AC
        self.engine = engine
        self.eqn = _get(config.eqn_config, "eqn_name")
        self.n = engine.n_params[AC]
        rng = rng if rng is not None else np.random
        self.theta = engine.tensor(self._init(config, AC, rng))

    @staticmethod
    def dims(config, AC):
        e, n = config.eqn_config, config.net_config
        dim, m = int(_get(e, "dim")), int(_get(e, "control_dim"))
        hid = list(_get(n, "num_hiddens_actor") if AC == "actor" else _get(n, "num_hiddens_critic"))   # solver.py:235-238
        if AC == "critic":
            out = 1
        elif AC == "critic_grad":
            out = dim
        elif _get(e, "eqn_name") in ("ekn", "EKN"):
            out = m + 1                                                                              # solver.py:255-256
        else:
            out = m
        return dim, hid, out

    @classmethod
    def _init(cls, config, AC, rng):
        """Reference initialisers (solver.py:239-258): BN gamma~U(0.1,0.5), beta~N(0,0.1^2); Dense kernels
        Glorot-uniform (Keras default), last bias zeros."""
        in_dim, hid, out = cls.dims(config, AC)
        parts = [rng.uniform(0.1, 0.5, in_dim), rng.normal(0.0, 0.1, in_dim)]
        prev = in_dim
        for h in hid:
            lim = math.sqrt(6.0 / (prev + h))
            parts += [rng.uniform(-lim, lim, prev * h), rng.uniform(0.1, 0.5, h), rng.normal(0.0, 0.1, h)]
            prev = h
        lim = math.sqrt(6.0 / (prev + out))
        parts += [rng.uniform(-lim, lim, prev * out), np.zeros(out), rng.uniform(0.1, 0.5, out), rng.normal(0.0, 0.1, out)]
        return np.concatenate(parts)

    def __call__(self, x, training=False, need_grad=False):
        if need_grad:
            raise NotImplementedError("need_grad=True is dead code in the reference (SURVEY Q5)")
        return self.engine.mlp_forward(self.AC, self.theta, self.engine.tensor(x))

    call = __call__

    @property
    def trainable_variables(self):
        return [self.theta]


def _unpack(engine, inputs):
    x0, dw, xb = inputs
    return engine.tensor(x0), (None if dw is None else engine.tensor(dw)), (None if xb is None else engine.tensor(xb))


class CriticModel(object):
    """solver.py:138-191."""

    def __init__(self, config, bsde, engine=None, rng=None):
        self.eqn_config, self.net_config, self.train_config = config.eqn_config, config.net_config, config.train_config
        self.bsde = bsde
        self.engine = engine
        self.NN_value = DeepNN(config, "critic", engine, rng)
        self.NN_value_grad = DeepNN(config, "critic_grad", engine, rng)
        self.gamma = _get(self.eqn_config, "discount")
        self.propagate = bsde.propagate_naive if _get(self.train_config, "scheme") == "naive" else bsde.propagate_adaptive

    def control(self, x, cheat_control, model_actor):
        """solver.py:153-157"""
        if cheat_control == False:  # noqa: E712  (the reference's own comparison)
            return model_actor.NN_control(x, training=False, need_grad=False)
        return self.bsde.u_true(x)

    def _step(self, inputs, model_actor, cheat_control, need_grad=False, want=("delta", "delta_bdry"), **kw):
        x0, dw, xb = _unpack(self.engine, inputs)
        return self.engine.critic_step(model_actor.NN_control.theta, self.NN_value.theta, self.NN_value_grad.theta, x0, dw, xb,
                                       int(_get(self.eqn_config, "num_time_interval_critic")),
                                       float(_get(self.eqn_config, "total_time_critic")),
                                       cheat_control=bool(cheat_control), need_grad=need_grad, want=want, **kw)

    def __call__(self, inputs, model_actor, training=False, cheat_control=False):
        r = self._step(inputs, model_actor, cheat_control)
        return r["delta"], r["delta_bdry"]

    call = __call__

    @property
    def trainable_variables(self):
        return [self.NN_value.theta, self.NN_value_grad.theta]


class ActorModel(object):
    """solver.py:193-224."""

    def __init__(self, config, bsde, engine=None, rng=None):
        self.eqn_config, self.net_config, self.train_config = config.eqn_config, config.net_config, config.train_config
        self.bsde = bsde
        self.engine = engine
        self.NN_control = DeepNN(config, "actor", engine, rng)
        self.gamma = _get(self.eqn_config, "discount")
        self.propagate = bsde.propagate_naive if _get(self.train_config, "scheme") == "naive" else bsde.propagate_adaptive

    def _step(self, inputs, model_critic, cheat_value, cheat_control, need_grad=False, want=("delta",), **kw):
        x0, dw, _ = _unpack(self.engine, inputs)
        return self.engine.actor_step(self.NN_control.theta, model_critic.NN_value.theta, x0, dw,
                                      int(_get(self.eqn_config, "num_time_interval_actor")),
                                      float(_get(self.eqn_config, "total_time_actor")),
                                      cheat_control=bool(cheat_control), cheat_value=bool(cheat_value), need_grad=need_grad,
                                      want=want, **kw)

    def __call__(self, inputs, model_critic, training=False, cheat_value=False, cheat_control=False):
        return self._step(inputs, model_critic, cheat_value, cheat_control)["delta"]

    call = __call__

    @property
    def trainable_variables(self):
        return [self.NN_control.theta]


class KerasAdam(object):
    """tf.keras.optimizers.Adam(learning_rate=PiecewiseConstantDecay(boundaries, values), epsilon=1e-8)
    as used at solver.py:14-19, on flat device vectors: lr_t = lr(step) * sqrt(1-b2^t)/(1-b1^t);
    theta -= lr_t * m / (sqrt(v) + eps).  PiecewiseConstantDecay returns values[i] while
    step <= boundaries[i] (step = number of updates applied so far)."""

    def __init__(self, engine, params, boundaries, values, beta_1=0.9, beta_2=0.999, epsilon=1e-8):
        self.engine = engine
        self.params = params
        self.m = [torch.zeros_like(p) for p in params]
        self.v = [torch.zeros_like(p) for p in params]
        self.iterations = 0
        self.boundaries, self.values = list(boundaries), list(values)
        self.b1, self.b2, self.eps = beta_1, beta_2, epsilon

    def learning_rate(self, step):
        for b, v in zip(self.boundaries, self.values):
            if step <= b:
                return v
        return self.values[-1]

    def next_lr_t(self):
        """advance the step counter and return this step's bias-corrected rate"""
        lr = self.learning_rate(self.iterations)
        self.iterations += 1
        t = self.iterations
        return lr * math.sqrt(1.0 - self.b2 ** t) / (1.0 - self.b1 ** t)

    def apply_gradients(self, grads, lr_dev=None):
        """lr_dev: one-element float64 device tensor the Adam kernels read the rate from (a captured iteration is replayed
        with the host writing next_lr_t() there); None: the rate is passed by value"""
        lr_t = self.next_lr_t() if lr_dev is None else 0.0
        for p, g, m, v in zip(self.params, grads, self.m, self.v):
            if g is None:
                continue
            self.engine.adam_step(p, g, m, v, lr_t, self.b1, self.b2, self.eps, lr_dev=lr_dev)


class ActorCriticSolver(object):
    """solver.py:7-136."""

    def __init__(self, config, bsde, compute_dtype="float32", device=None, seed=None, impl=None):
        self.eqn_config, self.net_config, self.train_config = config.eqn_config, config.net_config, config.train_config
        self.bsde = bsde
        dtype = _get(self.net_config, "dtype", "float64") if compute_dtype == "config" else compute_dtype
        if impl is None:                       # tensor cores compute in float32 (bf16x3); float64 is the exact path's
            impl = "tensor" if dtype == "float32" else "exact"
        dist = _dist()
        self.rank = dist.get_rank() if dist else 0
        self.world = dist.get_world_size() if dist else 1
        self.engine = Engine(self.eqn_config, self.net_config, self.train_config, dtype=dtype, device=device,
                             ekn_sigma_fix=getattr(bsde, "sigma_fix", False), impl=impl)
        bsde.bind(self.engine)
        rng = np.random.RandomState(seed) if seed is not None else None      # the reference seeds nothing (Q7)
        self.model_critic = CriticModel(config, bsde, self.engine, rng)
        self.model_actor = ActorModel(config, bsde, self.engine, rng)
        if dist and self.world > 1:                                          # identical replicas
            for t in self.model_critic.trainable_variables + self.model_actor.trainable_variables:
                dist.broadcast(t, src=0)
        n = self.net_config
        self.optimizer_critic = KerasAdam(self.engine, self.model_critic.trainable_variables,
                                          _get(n, "lr_boundaries_critic"), _get(n, "lr_values_critic"))
        self.optimizer_actor = KerasAdam(self.engine, self.model_actor.trainable_variables,
                                         _get(n, "lr_boundaries_actor"), _get(n, "lr_values_actor"))
        self.x = None
        self.gamma = _get(self.eqn_config, "discount")
        st = _get(self.train_config, "sample_type")
        if st == "normal":
            self.sample = self.bsde.sample_normal
        if st == "bounded":
            self.sample = self.bsde.sample_bounded
        tr = _get(self.train_config, "train")
        self.cheat_value_in_actor = False
        self.cheat_control_in_critic = False
        if tr == "critic":
            self.cheat_control_in_critic = True
        elif tr == "actor":
            self.cheat_value_in_actor = True
        self.sampler = _get(self.train_config, "sampler", "device")
        self.seed = int(seed if seed is not None else np.random.randint(1 << 31))
        if dist and self.world > 1:                                          # one Philox key for all ranks: the device sampler is keyed by
            sd = torch.tensor([self.seed], dtype=torch.int64, device=self.engine.device)   # (seed, iteration, GLOBAL path index)
            dist.broadcast(sd, src=0)
            self.seed = int(sd.item())
        self._dw_mode = _cabi.DW_PHILOX_NORMAL if st == "normal" else _cabi.DW_PHILOX_BOUNDED
        self._iter = 0
        self.N_c = int(_get(self.eqn_config, "num_time_interval_critic"))
        self.N_a = int(_get(self.eqn_config, "num_time_interval_actor"))
        self.T_c = float(_get(self.eqn_config, "total_time_critic"))
        self.T_a = float(_get(self.eqn_config, "total_time_actor"))

    # ------------------------------------------------------------------ sharding / reductions
    def _shard(self, B):
        """contiguous global path indices [offset, offset+B_local) of this rank"""
        per = (B + self.world - 1) // self.world
        lo = min(self.rank * per, B)
        hi = min(lo + per, B)
        return lo, hi - lo

    def _allreduce(self, tensors):
        dist = _dist()
        if not dist or self.world == 1:
            return tensors
        flat = torch.cat([t.reshape(-1) for t in tensors])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        out, o = [], 0
        for t in tensors:
            out.append(flat[o:o + t.numel()].view_as(t))
            o += t.numel()
        return out

    def _shard_inputs(self, inputs):
        if self.world == 1:
            return inputs, 0, inputs[0].shape[0]
        B = inputs[0].shape[0]
        lo, n = self._shard(B)
        return tuple(None if a is None else a[lo:lo + n] for a in inputs), lo, B

    # ------------------------------------------------------------------ losses / gradients
    def loss_critic(self, inputs, training=False, cheat_control=False):
        sh, lo, B = self._shard_inputs(inputs)
        r = self.model_critic._step(sh, self.model_actor, cheat_control, want=(), B_global=B, path_offset=lo)
        (loss,) = self._allreduce([r["loss"]])
        return loss[0] + loss[1]                                           # solver.py:78

    def loss_actor(self, inputs, training=False, cheat_value=False, cheat_control=False):
        sh, lo, B = self._shard_inputs(inputs)
        r = self.model_actor._step(sh, self.model_critic, cheat_value, cheat_control, want=(), B_global=B, path_offset=lo)
        (loss,) = self._allreduce([r["loss"]])
        return loss[0]                                                     # solver.py:82

    def grad_critic(self, inputs, training=False, cheat_control=False, **kw):
        sh, lo, B = self._shard_inputs(inputs)
        r = self.model_critic._step(sh, self.model_actor, cheat_control, need_grad=True, want=(), B_global=B, path_offset=lo, **kw)
        gV, gG, _ = self._allreduce([r["grad_V"], r["grad_G"], r["loss"]])
        return [gV, gG]

    def grad_actor(self, inputs, training=False, cheat_value=False, cheat_control=False, **kw):
        sh, lo, B = self._shard_inputs(inputs)
        r = self.model_actor._step(sh, self.model_critic, cheat_value, cheat_control, need_grad=True, want=(), B_global=B, path_offset=lo, **kw)
        gA, _ = self._allreduce([r["grad_actor"], r["loss"]])
        return [gA]

    def train_step_critic(self, train_data):
        grad = self.grad_critic(train_data, training=False, cheat_control=self.cheat_control_in_critic)
        self.optimizer_critic.apply_gradients(grad)

    def train_step_actor(self, train_data):
        grad = self.grad_actor(train_data, training=False, cheat_value=self.cheat_value_in_actor, cheat_control=False)
        self.optimizer_actor.apply_gradients(grad)

    # device-sampled training steps: x0/x_bdry from dpb_sample_x, increments generated in-kernel
    def _device_batch(self, B, phase, graph=None):
        lo, n = self._shard(B)
        # replayed graph: the iteration lives in device memory (graph["stream"] = iteration << 1), the argument is the phase
        stream_id = phase if graph else (self._iter << 1) | phase
        x0, xb = self.engine.sample_x(self.seed, stream_id, lo, n, want_xb=(phase == 0), stream_base=graph["stream"] if graph else None)
        return x0, xb, lo, stream_id

    def train_step_critic_device(self, B, graph=None):
        x0, xb, lo, sid = self._device_batch(B, 0, graph)
        r = self.engine.critic_step(self.model_actor.NN_control.theta, self.model_critic.NN_value.theta,
                                    self.model_critic.NN_value_grad.theta, x0, None, xb, self.N_c, self.T_c, B_global=B,
                                    path_offset=lo, cheat_control=self.cheat_control_in_critic, need_grad=True,
                                    dw_mode=self._dw_mode, seed=self.seed, stream_id=sid, stream_base=graph["stream"] if graph else None)
        gV, gG, loss = self._allreduce([r["grad_V"], r["grad_G"], r["loss"]])
        self.optimizer_critic.apply_gradients([gV, gG], lr_dev=graph["lr"][0:1] if graph else None)
        return loss

    def train_step_actor_device(self, B, graph=None):
        x0, _, lo, sid = self._device_batch(B, 1, graph)
        r = self.engine.actor_step(self.model_actor.NN_control.theta, self.model_critic.NN_value.theta, x0, None, self.N_a,
                                   self.T_a, B_global=B, path_offset=lo, cheat_value=self.cheat_value_in_actor, need_grad=True,
                                   dw_mode=self._dw_mode, seed=self.seed, stream_id=sid, stream_base=graph["stream"] if graph else None)
        gA, loss = self._allreduce([r["grad_actor"], r["loss"]])
        self.optimizer_actor.apply_gradients([gA], lr_dev=graph["lr"][1:2] if graph else None)
        return loss

    # ------------------------------------------------------------------ one iteration as ONE CUDA graph
    def enable_cuda_graph(self):
        """Capture a whole device-sampled training iteration -- sample_x, weight packing, critic rollout + TD gradient, slab
        reduction, Adam, the same for the actor -- into one CUDA graph and replay it from then on: one launch per iteration
        instead of ~40.  What changes from iteration to iteration (the Philox stream id = iteration << 1, the two
        bias-corrected Adam rates) lives in device memory the kernels read (dpb_inputs.stream_base, dpb_adam_step's lr_t_dev);
        the host writes it before each replay.  Single process only: with torch.distributed the NCCL all-reduce stays
        outside a graph and the iteration is launched kernel by kernel as before."""
        if self.world > 1 or self.sampler != "device":
            return False
        dev = self.engine.device
        g = {"stream": torch.zeros(1, dtype=torch.int64, device=dev), "lr": torch.zeros(2, dtype=torch.float64, device=dev),
             "graph": None, "launches": 0, "replays": 0}
        self._graph = g
        return True

    def _graph_params(self, g):
        tr = _get(self.train_config, "train")
        lr = [0.0, 0.0]
        if tr in ("actor-critic", "critic"):
            lr[0] = self.optimizer_critic.next_lr_t()
        if tr in ("actor-critic", "actor"):
            lr[1] = self.optimizer_actor.next_lr_t()
        # pageable sources on purpose: the driver stages them before cudaMemcpyAsync returns, so the next iteration's values
        # can be written while the GPU is still several iterations behind (a pinned buffer would be read when the copy RUNS)
        g["stream"].copy_(torch.tensor([self._iter << 1], dtype=torch.int64))
        g["lr"].copy_(torch.tensor(lr, dtype=torch.float64))

    def _graph_body(self, g):
        tr = _get(self.train_config, "train")
        B = int(_get(self.net_config, "batch_size"))
        if tr in ("actor-critic", "critic"):
            self.train_step_critic_device(B, g)
        if tr in ("actor-critic", "actor"):
            self.train_step_actor_device(B, g)

    def _graph_iteration(self):
        g = self._graph
        self._graph_params(g)
        if g["graph"] is None:
            # the first iteration runs eagerly on a side stream (allocator warm-up, cudaFuncSetAttribute calls), then the
            # very same call sequence is captured; the captured iteration is NOT executed by the capture itself
            self.engine.set_timing(False)
            side = torch.cuda.Stream(device=self.engine.device)
            side.wait_stream(torch.cuda.current_stream(self.engine.device))
            with torch.cuda.stream(side):
                self._graph_body(g)
            torch.cuda.current_stream(self.engine.device).wait_stream(side)
            torch.cuda.synchronize(self.engine.device)
            l0 = self.engine.launch_count()
            cg = torch.cuda.CUDAGraph()
            with torch.cuda.graph(cg):
                self._graph_body(g)
            g["launches"] = self.engine.launch_count() - l0
            g["graph"] = cg
            self._iter += 1
            return
        g["graph"].replay()
        g["replays"] += 1
        self._iter += 1

    def train_iteration(self):
        """One loop body of train() (solver.py:67-70)."""
        if getattr(self, "_graph", None) is not None:
            return self._graph_iteration()
        tr = _get(self.train_config, "train")
        B = int(_get(self.net_config, "batch_size"))
        if tr in ("actor-critic", "critic"):
            if self.sampler == "device":
                self.train_step_critic_device(B)
            else:
                self.train_step_critic(self.sample(B, self.N_c))
        if tr in ("actor-critic", "actor"):
            if self.sampler == "device":
                self.train_step_actor_device(B)
            else:
                self.train_step_actor(self.sample(B, self.N_a))
        self._iter += 1

    # ------------------------------------------------------------------ checkpoint / resume (absent in the reference)
    def state_dict(self):
        opt = lambda o: {"m": [t.cpu() for t in o.m], "v": [t.cpu() for t in o.v], "iterations": o.iterations}
        return {"theta_actor": self.model_actor.NN_control.theta.cpu(), "theta_critic": self.model_critic.NN_value.theta.cpu(),
                "theta_critic_grad": self.model_critic.NN_value_grad.theta.cpu(), "opt_critic": opt(self.optimizer_critic),
                "opt_actor": opt(self.optimizer_actor), "iter": self._iter, "seed": self.seed, "dtype": self.engine.dtype_name,
                # the run record so far (main.py writes the whole history), the wall clock already spent, and the host RNG
                # (the reference's NumPy samplers, used by sampler="host" and by the validation sets)
                "history": [list(map(float, r)) for r in getattr(self, "_history", [])], "elapsed": float(getattr(self, "_elapsed", 0.0)),
                "np_rng": _rng_to_plain(np.random.get_state())}

    def load_state_dict(self, sd):
        self.model_actor.NN_control.theta.copy_(sd["theta_actor"])
        self.model_critic.NN_value.theta.copy_(sd["theta_critic"])
        self.model_critic.NN_value_grad.theta.copy_(sd["theta_critic_grad"])
        for o, k in ((self.optimizer_critic, "opt_critic"), (self.optimizer_actor, "opt_actor")):
            for dst, src in zip(o.m, sd[k]["m"]):
                dst.copy_(src)
            for dst, src in zip(o.v, sd[k]["v"]):
                dst.copy_(src)
            o.iterations = int(sd[k]["iterations"])
        self._iter, self.seed = int(sd["iter"]), int(sd["seed"])      # device sampling is keyed by (seed, iteration): the exact path continues
        self._history = [list(r) for r in sd.get("history", [])]      # bit for bit (the tensor path up to its FP32 slab-reduction order)
        self._elapsed = float(sd.get("elapsed", 0.0))
        if sd.get("np_rng") is not None:
            np.random.set_state(_rng_from_plain(sd["np_rng"]))

    def save_checkpoint(self, path):
        if self.rank == 0:
            tmp = path + ".tmp"                      # write-then-rename: a kill mid-write never destroys the last good checkpoint
            torch.save(self.state_dict(), tmp)
            os.replace(tmp, path)

    def load_checkpoint(self, path):
        self.load_state_dict(torch.load(path, map_location="cpu", weights_only=True))

    # ------------------------------------------------------------------ train loop (solver.py:36-71)
    def train(self):
        training_history = [list(r) for r in getattr(self, "_history", [])]      # rows logged before a resume
        start_time = time.time() - float(getattr(self, "_elapsed", 0.0))         # elapsed time continues across a resume
        self._history = training_history
        n = self.net_config
        valid_size = int(_get(n, "valid_size"))
        dev = lambda data: tuple(self.engine.tensor(a) for a in data)
        valid_data_critic = dev(self.sample(valid_size, self.N_c))
        valid_data_actor = dev(self.sample(valid_size, self.N_a))
        valid_data_cost = dev(self.bsde.sample0(valid_size, self.N_a))
        true_loss_actor = float(self.loss_actor(valid_data_actor, training=False, cheat_value=True, cheat_control=True))
        num_iterations = int(_get(n, "num_iterations"))
        elapsed_time = 0.0
        ckpt_path, ckpt_every = getattr(self, "checkpoint_path", None), int(getattr(self, "checkpoint_every", 0) or 0)
        start = self._iter
        for step in range(start, num_iterations + 1):
            if ckpt_path and ckpt_every and step > start and step % ckpt_every == 0:
                self._elapsed = time.time() - start_time
                self.save_checkpoint(ckpt_path)
            if step % int(_get(n, "logging_frequency")) == 0:
                loss_critic = float(self.loss_critic(valid_data_critic, training=False, cheat_control=False))
                loss_actor = float(self.loss_actor(valid_data_actor, training=False, cheat_value=False, cheat_control=False))
                err_value = float(self.err_value(valid_data_critic))
                err_control = float(self.err_control(valid_data_actor))
                err_value_grad = float(self.err_value_grad(valid_data_critic))
                err_value_infty = float(self.err_value_infty(valid_data_critic))
                err_cost = float(self.err_cost(valid_data_cost))
                elapsed_time = time.time() - start_time
                training_history.append([step, loss_critic, loss_actor, err_value, err_value_infty, err_control, err_value_grad, err_cost, elapsed_time])
                if _get(n, "verbose") and self.rank == 0:
                    logging.info("step: %5u, loss_critic: %.4e, loss_actor: %.4e, err_value: %.4e, err_value_infty: %.4e, err_control: %.4e, err_value_grad: %.4e, err_cost: %.4e, elapsed time: %3u" % (
                        step, loss_critic, loss_actor, err_value, err_value_infty, err_control, err_value_grad, err_cost, elapsed_time))
            if step == num_iterations:
                x0, dw_sample, x_bdry = valid_data_critic
                y = self.model_critic.NN_value(x0, training=False, need_grad=False)
                true_y = self.bsde.V_true(x0)
                grad_y = self.model_critic.NN_value_grad(x0, training=False, need_grad=False)
                z = self.model_actor.NN_control(x0, training=False, need_grad=False)
                true_z = self.bsde.u_true(x0)
                if self.rank == 0:
                    print("true loss actor: ", true_loss_actor)
                training_history.append([0, 0.0, true_loss_actor, 0.0, 0.0, 0.0, 0.0, 0.0, elapsed_time])
            self.train_iteration()
        cpu = lambda t: t.detach().cpu().numpy()
        return np.array(training_history), cpu(x0), cpu(y), cpu(true_y), cpu(z), cpu(true_z), cpu(grad_y)

    # ------------------------------------------------------------------ error metrics (solver.py:109-136)
    def _rel_l2(self, true, approx):
        m = self.engine.err_metrics(true, approx)
        return torch.sqrt(m[0] / m[1])

    def err_value(self, inputs):
        x0 = self.engine.tensor(inputs[0])
        return self._rel_l2(self.bsde.V_true(x0), self.model_critic.NN_value(x0, training=False, need_grad=False))

    def err_control(self, inputs):
        x0 = self.engine.tensor(inputs[0])
        return self._rel_l2(self.bsde.u_true(x0), self.model_actor.NN_control(x0, training=False, need_grad=False))

    def err_value_grad(self, inputs):
        x0 = self.engine.tensor(inputs[0])
        return self._rel_l2(self.bsde.V_grad_true(x0), self.model_critic.NN_value_grad(x0, training=False, need_grad=False))

    def err_value_infty(self, inputs):
        x0 = self.engine.tensor(inputs[0])
        return self.engine.err_metrics(self.bsde.V_true(x0), self.model_critic.NN_value(x0, training=False, need_grad=False))[2]

    def err_cost(self, inputs):
        x0 = self.engine.tensor(inputs[0])
        y = self.model_actor(inputs, self.model_critic, training=False, cheat_value=False, cheat_control=False)
        y0 = self.model_critic.NN_value(x0, training=False, need_grad=False)
        return torch.mean(y - y0)
