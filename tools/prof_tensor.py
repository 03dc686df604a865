"""Small driver for ncu: a few tensor-path training launches (critic + actor) at B paths."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deeppde_actorcritic_b200.engine import Engine
from oracle import ref_solver as RS
e = {"eqn_name": "LQR", "discount": 1.0, "p": 1.0, "q": 1.0, "beta": 1.0, "R": 1.0, "dim": 20, "control_dim": 20,
     "total_time_critic": 0.2, "total_time_actor": 0.2, "num_time_interval_critic": 100, "num_time_interval_actor": 100}
net = {"num_hiddens_actor": [200, 200, 200], "num_hiddens_critic": [200, 200, 200]}
tr = {"scheme": "adaptive", "TD_type": "TD1"}
cfg = {"eqn_config": e, "net_config": net, "train_config": tr}
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128
impl = sys.argv[2] if len(sys.argv) > 2 else "tensor"
eng = Engine(e, net, tr, dtype="float32", impl=impl)
rng = np.random.RandomState(12)
th = {}
for k in ("actor", "critic", "critic_grad"):
    i, h, o, _ = RS.net_dims(cfg, k)
    th[k] = eng.tensor(RS.init_params(i, h, o, rng))
x0, xb = eng.sample_x(5, 1, 0, B)
kw = dict(dw_mode=1, seed=5, stream_id=3)
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
kcs, kas = [], []
for it in range(reps):
    r = eng.critic_step(th["actor"], th["critic"], th["critic_grad"], x0, None, xb, 100, 0.2, need_grad=True, **kw)
    kcs.append(eng.last_kernel_ms())
    a = eng.actor_step(th["actor"], th["critic"], x0, None, 100, 0.2, need_grad=True, **kw)
    kas.append(eng.last_kernel_ms())
torch.cuda.synchronize()
kcs, kas = kcs[2:], kas[2:]          # two warm-up launches
print(f"{impl} B={B}: critic {np.mean(kcs):.3f} ms (min {min(kcs):.3f}) actor {np.mean(kas):.3f} ms (min {min(kas):.3f}) over {len(kcs)} launches; "
      f"loss {float(r['loss'].sum()):.5f} {float(a['loss'][0]):.5f}")
