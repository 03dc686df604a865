"""Thin Python owner of one libdeeppde_b200 handle: turns the reference's JSON config into a
``dpb_config``, owns the device workspace (a torch uint8 tensor -- torch is only the allocator and
the stream provider) and forwards every arithmetic call to the C ABI.

No arithmetic of the hot path happens in Python or in torch ops.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _cabi


def _get(cfg, key, default=None):
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


TORCH_DTYPES = {"float32": torch.float32, "float64": torch.float64}


def make_dpb_config(eqn_config, net_config, train_config, dtype="float32", ekn_sigma_fix=False, impl="exact", lifetime_sort=True):
    c = _cabi.dpb_config()
    name = _get(eqn_config, "eqn_name")
    if name not in _cabi.EQN_IDS:
        raise ValueError(f"unknown eqn_name {name!r}")
    c.dtype = _cabi.DTYPE_IDS[dtype]
    c.eqn = _cabi.EQN_IDS[name]
    c.dim = int(_get(eqn_config, "dim"))
    c.control_dim = int(_get(eqn_config, "control_dim"))
    c.scheme = _cabi.SCHEME_IDS[_get(train_config, "scheme", "adaptive")]
    c.td_type = _cabi.TD_IDS[_get(train_config, "TD_type", "TD1")]
    c.ekn_sigma_fix = 1 if ekn_sigma_fix else 0
    ha = list(_get(net_config, "num_hiddens_actor"))
    hc = list(_get(net_config, "num_hiddens_critic"))
    if len(ha) > _cabi.DPB_MAX_HIDDEN or len(hc) > _cabi.DPB_MAX_HIDDEN:
        raise ValueError("at most 6 hidden layers per network")
    c.n_hidden_actor, c.n_hidden_critic = len(ha), len(hc)
    for i, h in enumerate(ha):
        c.hidden_actor[i] = int(h)
    for i, h in enumerate(hc):
        c.hidden_critic[i] = int(h)
    c.impl = {"exact": _cabi.IMPL_EXACT, "tensor": _cabi.IMPL_TENSOR}[impl]
    c.reserved[0] = 0 if lifetime_sort else 1
    c.R = float(_get(eqn_config, "R"))
    c.discount = float(_get(eqn_config, "discount"))
    for k in ("p", "q", "beta", "a", "epsilon", "a2", "a3"):
        setattr(c, k, float(_get(eqn_config, k, 0.0) or 0.0))
    return c


class Engine:
    """One handle + one device.  All tensors passed in must be CUDA, contiguous, of ``self.dtype``."""

    def __init__(self, eqn_config, net_config, train_config, dtype="float32", device=None, ekn_sigma_fix=False, impl="exact", lifetime_sort=True):
        self.lib = _cabi.load()
        if not torch.cuda.is_available():
            raise RuntimeError("deeppde_actorcritic_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        torch.cuda.set_device(self.device)
        self.dtype_name = dtype
        self.dtype = TORCH_DTYPES[dtype]
        self.cfg = make_dpb_config(eqn_config, net_config, train_config, dtype, ekn_sigma_fix, impl, lifetime_sort)
        self.dim, self.control_dim = self.cfg.dim, self.cfg.control_dim
        self.handle = C.c_void_p()
        rc = self.lib.dpb_create(C.byref(self.handle), C.byref(self.cfg))
        _cabi.check(self.lib, None, rc)
        self._ws = None
        self.n_params = {k: int(self.lib.dpb_param_count(self.handle, i))
                         for k, i in (("actor", 0), ("critic", 1), ("critic_grad", 2))}

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.dpb_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _chk(self, rc):
        _cabi.check(self.lib, self.handle, rc)

    def workspace(self, nbytes):
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return self._ws

    def tensor(self, a):
        """host array / tensor -> contiguous device tensor of the compute dtype."""
        return torch.as_tensor(a, dtype=self.dtype).to(self.device).contiguous()

    def _p(self, t):
        if t is None:
            return None
        assert t.is_cuda and t.is_contiguous() and t.dtype == self.dtype, "need a contiguous CUDA tensor of the compute dtype"
        return C.c_void_p(t.data_ptr())

    def launch_count(self):
        return int(self.lib.dpb_launch_count(self.handle))

    def set_timing(self, enable):
        """CUDA-event timing of the rollout kernels (switched off while an iteration is captured into a CUDA graph)"""
        self._chk(self.lib.dpb_set_timing(self.handle, 1 if enable else 0))

    def _inputs(self, x0, dw, xb, dw_mode, seed, stream_id, stream_base=None):
        i = _cabi.dpb_inputs()
        i.x0 = x0.data_ptr()
        i.dw = dw.data_ptr() if dw is not None else None
        i.x_bdry = xb.data_ptr() if xb is not None else None
        i.dw_mode, i.seed, i.stream = dw_mode, seed, stream_id
        i.stream_base = stream_base.data_ptr() if stream_base is not None else None     # device int64 added to stream_id (graph replays)
        return i

    def _outs(self, B, N, want):
        o = _cabi.dpb_path_outputs()
        res = {}
        d = self.dim
        shapes = {"x_smp": (B, d, N + 1), "dt": (B, N), "coef": (B, N), "delta": (B, 1), "delta_bdry": (B, 1)}
        for k in want:
            if k == "exit_index":
                t = torch.empty(B, dtype=torch.int32, device=self.device)
            else:
                t = torch.empty(shapes[k], dtype=self.dtype, device=self.device)
            res[k] = t
            setattr(o, k, t.data_ptr())
        return o, res

    # ------------------------------------------------------------------ steps
    def critic_step(self, theta_actor, theta_V, theta_G, x0, dw, xb, N, T, *, B_global=None, path_offset=0,
                    cheat_control=False, need_grad=False, propagate_only=False, want=(), dw_mode=_cabi.DW_EXTERNAL,
                    seed=0, stream_id=0, stream_base=None):
        """CriticModel.call / loss_critic / grad_critic (solver.py:73-78,85-90,159-191).
        Returns dict(loss[2] device tensor, grad_V, grad_G, + requested per-path outputs)."""
        B = x0.shape[0]
        Bg = B if B_global is None else B_global
        flags = (1 if cheat_control else 0) | (4 if need_grad else 0) | (8 if propagate_only else 0)
        nb = self.lib.dpb_workspace_bytes(self.handle, B, N)
        ws = self.workspace(nb)
        inp = self._inputs(x0, dw, xb, dw_mode, seed, stream_id, stream_base)
        o, res = self._outs(B, N, want)
        loss = torch.zeros(2, dtype=self.dtype, device=self.device)
        gV = torch.empty(self.n_params["critic"], dtype=self.dtype, device=self.device) if need_grad else None
        gG = torch.empty(self.n_params["critic_grad"], dtype=self.dtype, device=self.device) if need_grad else None
        rc = self.lib.dpb_critic_step(self.handle, self._p(theta_actor), self._p(theta_V), self._p(theta_G), C.byref(inp),
                                      B, path_offset, Bg, N, float(T), flags, self._p(loss), self._p(gV), self._p(gG),
                                      C.byref(o), C.c_void_p(ws.data_ptr()), ws.numel(), self._stream())
        self._chk(rc)
        res.update(loss=loss, grad_V=gV, grad_G=gG)
        return res

    def actor_step(self, theta_actor, theta_V, x0, dw, N, T, *, B_global=None, path_offset=0, cheat_control=False,
                   cheat_value=False, need_grad=False, want=(), dw_mode=_cabi.DW_EXTERNAL, seed=0, stream_id=0, stream_base=None):
        """ActorModel.call / loss_actor / grad_actor (solver.py:80-83,92-97,207-224)."""
        B = x0.shape[0]
        Bg = B if B_global is None else B_global
        flags = (1 if cheat_control else 0) | (2 if cheat_value else 0) | (4 if need_grad else 0)
        nb = self.lib.dpb_workspace_bytes(self.handle, B, N)
        ws = self.workspace(nb)
        inp = self._inputs(x0, dw, None, dw_mode, seed, stream_id, stream_base)
        o, res = self._outs(B, N, want)
        loss = torch.zeros(2, dtype=self.dtype, device=self.device)
        gA = torch.empty(self.n_params["actor"], dtype=self.dtype, device=self.device) if need_grad else None
        rc = self.lib.dpb_actor_step(self.handle, self._p(theta_actor), self._p(theta_V), C.byref(inp), B, path_offset, Bg,
                                     N, float(T), flags, self._p(loss), self._p(gA), C.byref(o),
                                     C.c_void_p(ws.data_ptr()), ws.numel(), self._stream())
        self._chk(rc)
        res.update(loss=loss, grad_actor=gA)
        return res

    # host-buffer entry points (end-to-end timing): x0 / x_bdry are pinned HOST tensors, the losses come
    # back to the host; dw is generated in-kernel (Philox) unless a host dw tensor is given
    def _host_ws(self, B, N, dw_mode):
        nb = self.lib.dpb_workspace_bytes(self.handle, B, N) + self.lib.dpb_staging_bytes(self.handle, B, N, dw_mode)
        return self.workspace(nb)

    def critic_step_host(self, theta_actor, theta_V, theta_G, x0_h, dw_h, xb_h, N, T, *, B_global=None, path_offset=0,
                         cheat_control=False, need_grad=True, dw_mode=_cabi.DW_PHILOX_NORMAL, seed=0, stream_id=0):
        B = x0_h.shape[0]
        Bg = B if B_global is None else B_global
        if dw_h is not None:
            dw_mode = _cabi.DW_EXTERNAL
        flags = (1 if cheat_control else 0) | (4 if need_grad else 0)
        ws = self._host_ws(B, N, dw_mode)
        inp = self._inputs(x0_h, dw_h, xb_h, dw_mode, seed, stream_id)
        loss_h = torch.zeros(2, dtype=self.dtype)
        gV = torch.empty(self.n_params["critic"], dtype=self.dtype, device=self.device) if need_grad else None
        gG = torch.empty(self.n_params["critic_grad"], dtype=self.dtype, device=self.device) if need_grad else None
        rc = self.lib.dpb_critic_step_host(self.handle, self._p(theta_actor), self._p(theta_V), self._p(theta_G), C.byref(inp),
                                           B, path_offset, Bg, N, float(T), flags, C.c_void_p(loss_h.data_ptr()), self._p(gV),
                                           self._p(gG), C.c_void_p(ws.data_ptr()), ws.numel(), self._stream())
        self._chk(rc)
        return dict(loss=loss_h, grad_V=gV, grad_G=gG)

    def actor_step_host(self, theta_actor, theta_V, x0_h, dw_h, N, T, *, B_global=None, path_offset=0, cheat_value=False,
                        need_grad=True, dw_mode=_cabi.DW_PHILOX_NORMAL, seed=0, stream_id=0):
        B = x0_h.shape[0]
        Bg = B if B_global is None else B_global
        if dw_h is not None:
            dw_mode = _cabi.DW_EXTERNAL
        flags = (2 if cheat_value else 0) | (4 if need_grad else 0)
        ws = self._host_ws(B, N, dw_mode)
        inp = self._inputs(x0_h, dw_h, None, dw_mode, seed, stream_id)
        loss_h = torch.zeros(2, dtype=self.dtype)
        gA = torch.empty(self.n_params["actor"], dtype=self.dtype, device=self.device) if need_grad else None
        rc = self.lib.dpb_actor_step_host(self.handle, self._p(theta_actor), self._p(theta_V), C.byref(inp), B, path_offset, Bg,
                                          N, float(T), flags, C.c_void_p(loss_h.data_ptr()), self._p(gA),
                                          C.c_void_p(ws.data_ptr()), ws.numel(), self._stream())
        self._chk(rc)
        return dict(loss=loss_h, grad_actor=gA)

    def tc_stats(self, B, N):
        """cycle counters of CTA 0 of the last tensor-path launch (diagnostics)"""
        import numpy as np
        out = np.zeros(16, dtype=np.int64)
        self._chk(self.lib.dpb_tc_stats(self.handle, C.c_void_p(self._ws.data_ptr()), B, N, out.ctypes.data_as(C.c_void_p)))
        return out

    def tc_trace(self, B, N):
        """event trace of CTA 0 of the last tensor-path launch: uint64[3][4096] (stats builds only; diagnostics)"""
        import numpy as np
        out = np.zeros((3, 4096), dtype=np.uint64)
        self._chk(self.lib.dpb_tc_trace(self.handle, C.c_void_p(self._ws.data_ptr()), B, N, out.ctypes.data_as(C.c_void_p)))
        return out

    def last_kernel_ms(self):
        """device time of the most recent critic/actor kernel launch (CUDA events recorded by the library)"""
        return float(self.lib.dpb_last_kernel_ms(self.handle))

    def mlp_forward(self, kind, theta, x):
        """DeepNN.call (solver.py:260-278)."""
        which = {"actor": 0, "critic": 1, "critic_grad": 2}[kind]
        n = x.shape[0]
        od = {"actor": self.control_dim, "critic": 1, "critic_grad": self.dim}[kind]
        out = torch.empty((n, od), dtype=self.dtype, device=self.device)
        ws = self.workspace(max(self.lib.dpb_workspace_bytes(self.handle, 1, 1), 1 << 20))
        rc = self.lib.dpb_mlp_forward(self.handle, which, self._p(theta), self._p(x), n, self._p(out),
                                      C.c_void_p(ws.data_ptr()), ws.numel(), self._stream())
        self._chk(rc)
        return out

    def closed_form(self, which, x, u=None):
        n = x.shape[0]
        od = {_cabi.CF_V_TRUE: (1,), _cabi.CF_Z: (1,), _cabi.CF_W: (1,), _cabi.CF_U_TRUE: (self.control_dim,), _cabi.CF_V_GRAD_TRUE: (self.dim,),
              _cabi.CF_DRIFT: (self.dim,), _cabi.CF_SIGMA: (self.dim, self.dim)}[which]
        out = torch.empty((n,) + od, dtype=self.dtype, device=self.device)
        rc = self.lib.dpb_closed_form(self.handle, which, self._p(x), self._p(u), n, self._p(out), self._stream())
        self._chk(rc)
        return out

    def diffusion(self, x, u, dw):
        """Equation.diffusion: sigma(x,u) . dw (equation.py:175-176,237-238,275-276,310-311)"""
        out = torch.empty_like(x)
        rc = self.lib.dpb_diffusion(self.handle, self._p(x), self._p(u), self._p(dw), x.shape[0], self._p(out), self._stream())
        self._chk(rc)
        return out

    def err_metrics(self, truth, approx):
        """{sum (t-a)^2, sum t^2, max |t-a|} (solver.py:109-130) as a 3-vector on the device"""
        out = torch.empty(3, dtype=self.dtype, device=self.device)
        t, a = truth.contiguous(), approx.contiguous()
        self._chk(self.lib.dpb_err_metrics(self.handle, self._p(t), self._p(a), t.numel(), self._p(out), self._stream()))
        return out

    def adam_step(self, theta, grad, m, v, lr_t, beta1=0.9, beta2=0.999, eps=1e-8, lr_dev=None):
        """lr_dev: optional one-element float64 device tensor that overrides lr_t when the kernel runs (graph replays)"""
        rc = self.lib.dpb_adam_step(self.handle, self._p(theta), self._p(grad), self._p(m), self._p(v), theta.numel(),
                                    float(lr_t), C.c_void_p(lr_dev.data_ptr()) if lr_dev is not None else None,
                                    beta1, beta2, eps, self._stream())
        self._chk(rc)

    def philox_dw(self, dw_mode, seed, stream_id, path_offset, B, N):
        dw = torch.empty((B, self.dim, N), dtype=self.dtype, device=self.device)
        rc = self.lib.dpb_philox_dw(self.handle, dw_mode, seed, stream_id, path_offset, B, N, self._p(dw), self._stream())
        self._chk(rc)
        return dw

    def sample_x(self, seed, stream_id, path_offset, B, want_xb=True, stream_base=None):
        x0 = torch.empty((B, self.dim), dtype=self.dtype, device=self.device)
        xb = torch.empty((B, self.dim), dtype=self.dtype, device=self.device) if want_xb else None
        rc = self.lib.dpb_sample_x(self.handle, seed, stream_id, C.c_void_p(stream_base.data_ptr()) if stream_base is not None else None,
                                   path_offset, B, self._p(x0), self._p(xb), self._stream())
        self._chk(rc)
        return x0, xb
