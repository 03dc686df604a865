"""lqr_d5 shape (d=5, N=50, nets 2x200) at B paths on the tensor path: per-LIVE-path-step rate of the naive scheme with and
without the lifetime sort, and of the adaptive scheme (naive within 1.3x of adaptive per live step)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deeppde_actorcritic_b200.engine import Engine
from oracle import ref_solver as RS
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cfg = json.load(open(os.path.join(ROOT, "configs", "bench_lqr_d5_naive_normal_td1.json")))
e, net = cfg["eqn_config"], cfg["net_config"]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 17
N, T = int(e["num_time_interval_critic"]), float(e["total_time_critic"])
rng = np.random.RandomState(12)
th_np = {}
for k in ("actor", "critic", "critic_grad"):
    i, h, o, _ = RS.net_dims(cfg, k)
    th_np[k] = RS.init_params(i, h, o, rng)
rows = []
for scheme, sort in (("naive", False), ("naive", True), ("adaptive", True)):
    tr = dict(cfg["train_config"], scheme=scheme)
    eng = Engine(e, net, tr, dtype="float32", impl="tensor", lifetime_sort=sort)
    th = {k: eng.tensor(v) for k, v in th_np.items()}
    x0, xb = eng.sample_x(5, 1, 0, B)
    kw = dict(dw_mode=1, seed=5, stream_id=3, need_grad=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = []
    for it in range(5):
        torch.cuda.synchronize()
        ev0.record()
        r = eng.critic_step(th["actor"], th["critic"], th["critic_grad"], x0, None, xb, N, T, want=("exit_index",), **kw)
        a = eng.actor_step(th["actor"], th["critic"], x0, None, N, T, want=("exit_index",), **kw)
        ev1.record()
        torch.cuda.synchronize()
        ms.append(ev0.elapsed_time(ev1))
    live = (torch.clamp(r["exit_index"].long() + 1, max=N).sum() + torch.clamp(a["exit_index"].long() + 1, max=N).sum()).item()
    t = float(np.mean(ms[2:]))
    rows.append((scheme, sort, t, live, live / (2.0 * B * N), live / t / 1e3, 2.0 * B * N / t / 1e3))
    print(f"{scheme:8s} lifetime_sort={sort!s:5s}: critic+actor step {t:8.3f} ms (all launches incl. the sort pre-pass), live path-steps {live} "
          f"({live / (2.0 * B * N):.3f}), {live / t / 1e3:8.1f} M live-steps/s, {2.0 * B * N / t / 1e3:8.1f} M nominal path-steps/s")
print(f"naive speed-up from the lifetime sort: {rows[0][2] / rows[1][2]:.2f}x; per-live-step rate naive(sorted) / adaptive = {rows[1][5] / rows[2][5]:.2f}")
