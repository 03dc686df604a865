#!/usr/bin/env python
"""bench.py -- path-steps/sec of the fused rollout + TD training iteration (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload lqr_d20_2p20]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one training iteration = critic phase (rollout + VR-LSTD residual + grad + all-reduce +
Adam) then actor phase (rollout + BPTT + all-reduce + Adam) on a fresh batch, as reference
solver.py:67-70.  Workload (default): `lqr_d20` scaled to 2^20 paths per iteration (BASELINE.json
configs[4]; adaptive / normal / TD1, d=20, N=100, nets 3x200), the global batch sharded over the N
ranks (strong scaling).  value = 2*B*N*K nominal path-steps / max-over-ranks device time.

value : inputs resident in HBM (x0/x_bdry sampled on the device, increments generated in-kernel).
e2e   : the same iteration through the C-ABI host entry points with x0 / x_bdry coming from pinned
        HOST memory every step and the losses read back to the host inside the timed region.
roofline : both rollout kernels are measured (roofline_kernels.critic / .actor: algorithmic FLOPs of SURVEY 8(d)
        for the live path-steps of that launch / its CUDA-event duration, peak from MEASURED_PEAKS.json); `roofline`
        is the one with the longer launch (the dominant kernel).  `traffic` is reported only when profiles/traffic.json
        holds an ncu capture taken from exactly the kernel sources of this build (source hash), else null.
cpu_baseline / --impl reference : the torch-CPU restatement of the reference (oracle/), all host threads, on a
        bounded sample (the config's own batch of 2048 paths) -- float64 as the reference's configs ask, with the
        float32 figure of the same port beside it; the reference line's `config` states the batch it really ran.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (config file, global batch override)
    "lqr_d20_2p20": ("configs/lqr_d20.json", 1 << 20),
    "lqr_d20_2p17": ("configs/lqr_d20.json", 1 << 17),
    "lqr_d20": ("configs/lqr_d20.json", None),
    "lqr_d5": ("configs/bench_lqr_d5_naive_normal_td1.json", None),
    "vdp_d10": ("configs/bench_vdp_d10_adaptive_bounded_td2.json", None),
    "ekn_d20": ("configs/bench_ekn_d20_adaptive_normal_td1.json", None),
    "lqr_var_d20": ("configs/bench_lqr_var_d20_adaptive_normal_td1.json", None),
}


def load_cfg(workload):
    path, B = WORKLOADS[workload]
    with open(os.path.join(ROOT, path)) as f:
        cfg = json.load(f)
    if B is not None:
        cfg["net_config"]["batch_size"] = B
    return cfg


# ----------------------------------------------------------------------------- algorithmic FLOPs
def dense_macs(in_dim, hid, out):
    m, prev = 0, in_dim
    for h in hid:
        m += prev * h
        prev = h
    return m + prev * out


def flop_model(cfg):
    """FLOPs per live path-step and per path, 2 x MAC, one actor evaluation per step."""
    e, n, t = cfg["eqn_config"], cfg["net_config"], cfg["train_config"]
    d, m = e["dim"], e["control_dim"]
    ekn = e["eqn_name"] in ("ekn", "EKN")
    MA = dense_macs(d, n["num_hiddens_actor"], m + 1 if ekn else m)
    MG = dense_macs(d, n["num_hiddens_critic"], d)
    MV = dense_macs(d, n["num_hiddens_critic"], 1)
    f = d * n["num_hiddens_critic"][0]
    critic_step = 2 * (MA + 3 * MG - f) if t["TD_type"] == "TD1" else 2 * MA
    return {"critic_step": critic_step, "actor_step": 2 * 3 * MA, "critic_path": 2 * 3 * (3 * MV - f), "actor_path": 2 * 2 * MV}


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler(threading.Thread):
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU reference arm
def cpu_reference(cfg, steps, warmup, sample_B=None, with_f32=True):
    """Times oracle.RefSolver.train_iteration (torch-CPU restatement of solver.py:67-70, host sampling
    included) on a bounded sample of the workload: float64 (the reference's dtype) and, beside it, float32."""
    import torch
    from oracle import ref_equation as RE
    from oracle import ref_solver as RS
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    c = json.loads(json.dumps(cfg))
    B = sample_B or min(c["net_config"]["batch_size"], 2048)
    c["net_config"]["batch_size"] = B
    eqn = RE.make_ref_equation(c["eqn_config"])
    N = c["eqn_config"]["num_time_interval_critic"]

    def run(dtype, k, w):
        s = RS.RefSolver(c, eqn, seed=0, dtype=dtype)
        for _ in range(w):
            s.train_iteration()
        t0 = time.perf_counter()
        for _ in range(k):
            s.train_iteration()
        return (time.perf_counter() - t0) / k

    dt = run(torch.float64, steps, warmup)
    out = {"value": 2.0 * B * N / dt, "unit": "path-steps/s", "cores": cores, "kind": "port", "sample_paths": B,
           "sample": f"{steps} train iterations (critic+actor step, host sampling included) at B={B} of the workload's paths, "
                     f"N={N}, float64, torch {torch.__version__} CPU, {warmup} warm-up", "sec_per_iter": dt}
    if with_f32:
        try:
            dt32 = run(torch.float32, max(1, min(steps, 2)), 1)
            out["value_f32"] = 2.0 * B * N / dt32
            out["sec_per_iter_f32"] = dt32
        except Exception as ex:                       # the float32 figure is a courtesy; the float64 one is the baseline
            out["value_f32"] = None
            out["f32_error"] = repr(ex)[:200]
    return out


def kernel_source_hash():
    """sha256 over the CUDA sources of the library: ties an ncu traffic capture to the build it was taken from"""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "deeppde_actorcritic_b200", "csrc")
    for f in sorted(os.listdir(d)):
        h.update(f.encode())
        h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def cpu_model_name():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


# ----------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "exact", "tensor"])
    ap.add_argument("--workload", default="lqr_d20_2p20", choices=list(WORKLOADS))
    ap.add_argument("--dtype", default="float32")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cuda-graph", default="auto", choices=["auto", "on", "off"],
                    help="replay each training iteration as ONE captured CUDA graph (solver.enable_cuda_graph); auto: on for "
                         "config-size batches (<= 8192 paths per rank, where launch latency shows), single process only")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: keep a private handle to it and point file descriptor 1 at stderr, so that
    # nothing a library prints (NCCL's version banner goes to fd 1) can end up in front of the line
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = load_cfg(args.workload)
    e, n, t = cfg["eqn_config"], cfg["net_config"], cfg["train_config"]
    B, N = n["batch_size"], e["num_time_interval_critic"]
    config_desc = {"workload": f"{args.workload}: {e['eqn_name']} d={e['dim']} m={e['control_dim']} N={N} T={e['total_time_critic']} "
                               f"nets {len(n['num_hiddens_actor'])}x{n['num_hiddens_actor'][0]} {t['scheme']}/{t['sample_type']}/{t['TD_type']}/{t['train']}",
                   "global_paths_per_iteration": B, "path_steps_per_step": 2 * B * N, "parallelism": f"paths sharded over {world} rank(s)",
                   "cache": "per-iteration working set (x0, trajectory scratch, gradient slabs) exceeds the 126 MB L2; fresh paths every step"}

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference(cfg, max(1, args.steps), max(1, min(args.warmup, 1)))
        Bs = r["sample_paths"]
        config_desc = dict(config_desc, global_paths_per_iteration=Bs, path_steps_per_step=2 * Bs * N,
                           parallelism=f"{r['cores']} host threads (torch intra-op), no GPU",
                           sample_of=f"bounded sample: {Bs} of the workload's {B} paths per iteration (the config's own batch); "
                                     f"path-steps/s is per-path work, so the rate carries over to the full batch",
                           cache="host memory")
        line = {"impl": "reference", "metric": "path-steps/sec (train iteration: fused rollout + TD grad, critic+actor)", "value": r["value"],
                "unit": "path-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["sec_per_iter"] * 1e3,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config_desc, "cpu_baseline": {k: r.get(k) for k in ("value", "unit", "cores", "kind", "sample", "value_f32")},
                "cpu_model": cpu_model_name(), "iters_per_sec": 1.0 / r["sec_per_iter"],
                "e2e": {"value": r["value"], "unit": "path-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), file=json_out, flush=True)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from deeppde_actorcritic_b200 import equation, munchify
    from deeppde_actorcritic_b200.solver import ActorCriticSolver
    torch.cuda.set_device(local_rank)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # NCCL's banner / debug lines go to stderr: stdout is the one JSON line
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    impl = "exact" if args.impl == "exact" else "tensor"          # "ours" = the tensor path (tcgen05, bf16x3 products, FP32 accumulation)
    config = munchify(cfg)
    config.train_config["sampler"] = "device"
    bsde = getattr(equation, config.eqn_config.eqn_name)(config.eqn_config)
    solver = ActorCriticSolver(config, bsde, compute_dtype=args.dtype, seed=2024, impl=impl)
    eng = solver.engine
    dev = eng.device

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------------------------------------------------------- value: inputs resident in HBM
    use_graph = world == 1 and (args.cuda_graph == "on" or (args.cuda_graph == "auto" and B <= 8192))
    if use_graph:
        use_graph = solver.enable_cuda_graph()
    for _ in range(max(args.warmup, 2 if use_graph else 0)):      # (the first graph iteration runs eagerly, the second captures)
        solver.train_iteration()
    sync_all()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.25)
    l0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    ev0.record()
    for _ in range(args.steps):
        solver.train_iteration()
    ev1.record()
    sync_all()
    ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count() - l0
    if use_graph:                                                # replayed graphs: the library's launch counter does not see them
        launches += args.steps * solver._graph["launches"]
        solver._graph = None                                      # the roofline / e2e sections below launch kernel by kernel
        eng.set_timing(True)
    clocks = sampler.stop() if sampler else None
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms)
    value = 2.0 * B * N * args.steps / (ms * 1e-3)

    # ---------------------------------------------------------------- roofline of the two rollout kernels
    fm = flop_model(cfg)
    lo, nloc = solver._shard(B)
    thA_, thV_, thG_ = solver.model_actor.NN_control.theta, solver.model_critic.NN_value.theta, solver.model_critic.NN_value_grad.theta
    meas = {"critic": [], "actor": []}
    for i in range(max(2, min(args.steps, 3)) + 1):                        # first launch of each: warm-up, dropped
        x0, xb, _, sid = solver._device_batch(B, 0)
        r = eng.critic_step(thA_, thV_, thG_, x0, None, xb, solver.N_c, solver.T_c, B_global=B, path_offset=lo, need_grad=True,
                            want=("exit_index",), dw_mode=solver._dw_mode, seed=solver.seed, stream_id=sid + 1000 + 2 * i)
        k = eng.last_kernel_ms()
        live = torch.clamp(r["exit_index"].to(torch.int64) + 1, max=solver.N_c).sum().item()     # steps with a proposal computed
        meas["critic"].append((k, live * fm["critic_step"] + nloc * fm["critic_path"], live / float(nloc * solver.N_c)))
        a = eng.actor_step(thA_, thV_, x0, None, solver.N_a, solver.T_a, B_global=B, path_offset=lo, need_grad=True,
                           want=("exit_index",), dw_mode=solver._dw_mode, seed=solver.seed, stream_id=sid + 1001 + 2 * i)
        k = eng.last_kernel_ms()
        live = torch.clamp(a["exit_index"].to(torch.int64) + 1, max=solver.N_a).sum().item()
        meas["actor"].append((k, live * fm["actor_step"] + nloc * fm["actor_path"], live / float(nloc * solver.N_a)))
    peaks = {}
    pk_src = "fallback (B200_PROFILING.md)"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        pk_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
    except Exception:
        pass
    peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
    traffic_db, src_hash = {}, kernel_source_hash()
    try:
        traffic_db = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        pass
    kname = {"critic": "critic_tc_kernel" if impl == "tensor" else "critic_kernel<%s>" % ("float" if args.dtype == "float32" else "double"),
             "actor": "actor_tc_kernel" if impl == "tensor" else "actor_kernel<%s>" % ("float" if args.dtype == "float32" else "double")}
    rk = {}
    for which in ("critic", "actor"):
        rows = meas[which][1:]
        kms_avg = sum(r_[0] for r_ in rows) / len(rows)
        kfl_avg = sum(r_[1] for r_ in rows) / len(rows)
        achieved = kfl_avg / (kms_avg * 1e-3) / 1e12
        tr = (traffic_db.get(args.workload) or {}).get(which) if (impl == "tensor" and world == 1) else None
        fresh = bool(tr) and tr.get("source_hash") == src_hash
        rk[which] = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "traffic": (tr["dram_bytes_read"] + tr["dram_bytes_write"]) if fresh else None,
                     "traffic_source": (tr.get("capture") if fresh else "no ncu capture of this build's kernel sources (profiles/traffic.json is keyed by source hash)"),
                     "kernel": kname[which], "kernel_ms": kms_avg, "algorithmic_flop_per_launch": kfl_avg, "live_fraction": rows[-1][2],
                     "frac_of_bf16x3_ceiling": achieved / (peak / 3.0), "frac_of_fp32_peak": achieved / fp32_peak}
    dominant = "critic" if rk["critic"]["kernel_ms"] >= rk["actor"]["kernel_ms"] else "actor"
    tot_ms = rk["critic"]["kernel_ms"] + rk["actor"]["kernel_ms"]
    tot_fl = rk["critic"]["algorithmic_flop_per_launch"] + rk["actor"]["algorithmic_flop_per_launch"]
    roofline = dict(rk[dominant], peak_source=pk_src, dominant=dominant,
                    iteration={"achieved": tot_fl / (tot_ms * 1e-3) / 1e12, "frac": tot_fl / (tot_ms * 1e-3) / 1e12 / peak,
                               "kernel_ms": tot_ms, "note": "critic + actor rollout kernels of one iteration"},
                    note=("impl=tensor: every product is formed as 3 bf16 MMAs (hi*hi + hi*lo + lo*hi, FP32 accumulation) to hold FP32 tolerance, so the "
                          "ceiling of these kernels is peak/3 = %.0f TFLOP/s algorithmic; executed/algorithmic MMA work is ~1.3x (recompute in the reverse sweeps)" % (peak / 3.0)
                          if impl == "tensor" else
                          "impl=exact: FP32 CUDA-core FMA path; its own ceiling is the FP32 FMA peak %.1f TFLOP/s at the sampled %.0f MHz" % (fp32_peak, sm_mhz)))

    # ---------------------------------------------------------------- e2e: host buffers through the C-ABI host entry points
    pool = 2
    pin = lambda *s: torch.empty(*s, dtype=eng.dtype).pin_memory()
    hx0c, hxbc, hx0a = [pin(nloc, eng.dim) for _ in range(pool)], [pin(nloc, eng.dim) for _ in range(pool)], [pin(nloc, eng.dim) for _ in range(pool)]
    for i in range(pool):                                                # synthetic host inputs (uniform in ball / on sphere)
        a, b = eng.sample_x(7, 100 + i, lo, nloc)
        c, _ = eng.sample_x(7, 200 + i, lo, nloc, want_xb=False)
        hx0c[i].copy_(a); hxbc[i].copy_(b); hx0a[i].copy_(c)
    torch.cuda.synchronize()
    thA, thV, thG = solver.model_actor.NN_control.theta, solver.model_critic.NN_value.theta, solver.model_critic.NN_value_grad.theta

    def e2e_iteration(i):
        j = i % pool
        r = eng.critic_step_host(thA, thV, thG, hx0c[j], None, hxbc[j], solver.N_c, solver.T_c, B_global=B, path_offset=lo,
                                 dw_mode=solver._dw_mode, seed=solver.seed, stream_id=(10_000 + i) << 1)
        gV, gG, lc = solver._allreduce([r["grad_V"], r["grad_G"], r["loss"].to(dev)])
        solver.optimizer_critic.apply_gradients([gV, gG])
        a = eng.actor_step_host(thA, thV, hx0a[j], None, solver.N_a, solver.T_a, B_global=B, path_offset=lo,
                                dw_mode=solver._dw_mode, seed=solver.seed, stream_id=((10_000 + i) << 1) | 1)
        gA, la = solver._allreduce([a["grad_actor"], a["loss"].to(dev)])
        solver.optimizer_actor.apply_gradients([gA])
        if world > 1:                                                    # global losses (the per-rank ones are already on the host)
            return float(lc.sum()), float(la[0])
        return float(r["loss"].sum()), float(a["loss"][0])

    for i in range(min(args.warmup, 2)):
        e2e_iteration(i)
    sync_all()
    ev0.record()
    for i in range(args.steps):
        losses = e2e_iteration(100 + i)
    ev1.record()
    sync_all()
    tms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    e2e_value = 2.0 * B * N * args.steps / (float(tms) * 1e-3)
    esz = 4 if args.dtype == "float32" else 8
    e2e = {"value": e2e_value, "unit": "path-steps/s", "h2d_bytes_per_step": 3 * nloc * eng.dim * esz * world,
           "d2h_bytes_per_step": 4 * esz * world, "ms_per_step": float(tms) / args.steps,
           "note": "x0/x_bdry (critic) and x0 (actor) copied from pinned host memory each step by dpb_*_step_host; Brownian increments "
                   "generated in-kernel (Philox); both phases' losses copied back to the host each step"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference(cfg, 2, 1)
        cpu = {k: r.get(k) for k in ("value", "unit", "cores", "kind", "sample", "value_f32")}
        cpu["cpu_model"] = cpu_model_name()

    if rank == 0:
        line = {"metric": "path-steps/sec (train iteration: fused rollout + TD grad, critic+actor)", "value": value, "unit": "path-steps/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": ("f32 (bf16x3 tensor-core products, FP32 accumulate)" if impl == "tensor" else "f32" if args.dtype == "float32" else "f64"), "data": "synthetic",
                "config": config_desc, "impl": impl, "iters_per_sec": args.steps / (ms * 1e-3), "clocks": clocks, "e2e": e2e,
                "gpu_launches": launches, "cuda_graph": bool(use_graph), "roofline": roofline, "roofline_kernels": rk, "cpu_baseline": cpu, "last_losses": losses,
                "kernel_source_hash": src_hash}
        print(json.dumps(line), file=json_out, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
