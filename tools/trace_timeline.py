"""Event timeline of CTA 0 of the tensor kernels (library built with DPB_TC_STATS=1): where a tile-step's cycles go.

usage: python tools/trace_timeline.py [paths] [critic|actor] [first_event] [n_events]
Prints, per role, the mean duration of every (event -> next event) transition and a merged excerpt of the three roles'
events in time order.  Event ids: deeppde_actorcritic_b200/csrc/dpb_tc_nets.cuh (TC_TRACE).
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import collections
import numpy as np
from deeppde_actorcritic_b200.engine import Engine
from oracle import ref_solver as RS

NAMES = {40: "starved: weights", 41: "starved: chunk", 6: "slot landed", 7: "chunk seen", 8: "slot issued", 1: "TS ready", 2: "TS issued", 3: "dW ready", 4: "dW issued", 5: "op enter", 10: "epi enter", 11: "acc seen", 12: "epi done",
         13: "drain enter", 14: "dW seen", 15: "drain done", 20: "own publish", 21: "own wait", 22: "own has result"}
ROLE = ("own", "help", "ctrl")

e = {"eqn_name": "LQR", "discount": 1.0, "p": 1.0, "q": 1.0, "beta": 1.0, "R": 1.0, "dim": 20, "control_dim": 20,
     "total_time_critic": 0.2, "total_time_actor": 0.2, "num_time_interval_critic": 100, "num_time_interval_actor": 100}
net = {"num_hiddens_actor": [200, 200, 200], "num_hiddens_critic": [200, 200, 200]}
tr = {"scheme": "adaptive", "TD_type": "TD1"}
cfg = {"eqn_config": e, "net_config": net, "train_config": tr}
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 17
which = sys.argv[2] if len(sys.argv) > 2 else "critic"
first = int(sys.argv[3]) if len(sys.argv) > 3 else 0
count = int(sys.argv[4]) if len(sys.argv) > 4 else 120
N, T = 100, 0.2
eng = Engine(e, net, tr, dtype="float32", impl="tensor")
rng = np.random.RandomState(12)
th = {}
for k in ("actor", "critic", "critic_grad"):
    i, h, o, _ = RS.net_dims(cfg, k)
    th[k] = eng.tensor(RS.init_params(i, h, o, rng))
x0, xb = eng.sample_x(5, 1, 0, B)
kw = dict(dw_mode=1, seed=5, stream_id=3)
for it in range(2):
    if which == "critic":
        eng.critic_step(th["actor"], th["critic"], th["critic_grad"], x0, None, xb, N, T, need_grad=True, **kw)
    else:
        eng.actor_step(th["actor"], th["critic"], x0, None, N, T, need_grad=True, **kw)
ms = eng.last_kernel_ms()
tr_ = eng.tc_trace(B, N)
print(f"{which} kernel {ms:.3f} ms at {B} paths")
ev = []
for r in range(3):
    w = tr_[r]
    w = w[w != 0]
    t = (w >> np.uint64(8)).astype(np.int64)
    i = (w & np.uint64(255)).astype(np.int64)
    # (entries an earlier launch left behind: keep the monotone prefix)
    n = len(t)
    for j in range(1, len(t)):
        if t[j] < t[j - 1]:
            n = j
            break
    t, i = t[:n], i[:n]
    print(f"\n[{ROLE[r]}] {n} events")
    d = collections.defaultdict(list)
    for j in range(n - 1):
        d[(int(i[j]), int(i[j + 1]))].append(int(t[j + 1] - t[j]))
    tot = sum(sum(v) for v in d.values())
    for (a, b), v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        print(f"  {NAMES.get(a, a):>14s} -> {NAMES.get(b, b):<14s} n={len(v):5d} mean {np.mean(v):8.0f} median {np.median(v):8.0f}  share {sum(v) / max(tot, 1):6.1%}")
    ev += [(int(t[j]), r, int(i[j])) for j in range(n)]
ev.sort()
t0 = ev[0][0] if ev else 0
print("\nmerged excerpt (cycles since the first event):")
for (t, r, i) in ev[first:first + count]:
    print(f"  {t - t0:10d}  {'        ' * r}{ROLE[r]:5s} {NAMES.get(i, i)}")
