"""CPU oracle for the rollout + TD training step of MoZhou1995/DeepPDE_ActorCritic.

TEST INFRASTRUCTURE ONLY.  Nothing under ``deeppde_actorcritic_b200/`` imports this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may use it, and there only as the checker / timed baseline.

Pinning status: the reference ships no tests and no golden vectors, and its arithmetic
lives in TensorFlow 2 (un-vendored, not installable here).  The restatement in
``ref_equation.py`` / ``ref_solver.py`` is therefore pinned against the reference's OWN
source files (``/root/reference/equation.py`` and ``solver.py``, imported unmodified)
executed under the small TensorFlow->torch operator shim in ``oracle/tf_shim`` --
see ``tests/golden/make_golden.py`` and ``tests/test_oracle_vs_golden.py``.
"""
