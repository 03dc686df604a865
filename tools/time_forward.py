"""Times the forward-only critic/actor kernels (loss evaluation) of both implementations."""
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
B, N, T = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 17, 100, 0.2
need_grad = len(sys.argv) > 2 and sys.argv[2] == "grad"
for impl in ("tensor", "exact"):
    eng = Engine(e, net, tr, dtype="float32", impl=impl)
    rng = np.random.RandomState(12)
    th = {}
    for k in ("actor", "critic", "critic_grad"):
        i, h, o, _ = RS.net_dims(cfg, k)
        th[k] = eng.tensor(RS.init_params(i, h, o, rng))
    x0, xb = eng.sample_x(5, 1, 0, B)
    kw = dict(dw_mode=1, seed=5, stream_id=3)
    for it in range(3):
        r = eng.critic_step(th["actor"], th["critic"], th["critic_grad"], x0, None, xb, N, T, need_grad=need_grad, want=("exit_index",), **kw)
        kc = eng.last_kernel_ms()
        stc = eng.tc_stats(B, N) if impl == "tensor" else None
        a = eng.actor_step(th["actor"], th["critic"], x0, None, N, T, need_grad=need_grad, **kw)
        ka = eng.last_kernel_ms()
        sta = eng.tc_stats(B, N) if impl == "tensor" else None
    if impl == "tensor":
        for nm, st in (("critic", stc), ("actor", sta)):
            T = max(int(st[0]), 1)
            print(f"  {nm} CTA0 cycles: total {st[0]:,} ops {st[3]:,} ({st[0]/max(st[3],1):.0f} cyc/op) | control warp: waits owners/helpers {st[1]/T:.1%}, "
                  f"dW waits operands {st[2]/T:.1%}, in issue loops {st[7]/T:.1%}, waits MMAs reading ACT {st[8]/T:.1%} | owner thread 0: waits network "
                  f"results {st[4]/T:.1%}, writes inputs {st[5]/T:.1%} | helper thread 128: waits tensor pipe {st[9]/T:.1%}, hidden epilogues (incl. wait) "
                  f"{st[6]/T:.1%}, dW drains {st[10]/T:.1%}")
    live = torch.clamp(r["exit_index"].long() + 1, max=N).sum().item()
    print(f"{impl:7s} B={B} grad={need_grad}: critic kernel {kc:9.3f} ms, actor kernel {ka:9.3f} ms, live path-steps {live} "
          f"({live / (B * N):.3f}); critic loss {float(r['loss'].sum()):.6f} actor loss {float(a['loss'][0]):.6f}; "
          f"critic {live / kc / 1e3:.1f} M live-steps/s, actor {live / ka / 1e3:.1f} M live-steps/s")
