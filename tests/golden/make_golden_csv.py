#!/usr/bin/env python
"""Golden CSV files written by the REFERENCE's own main.py (/root/reference/main.py, unmodified): its ActorCriticSolver is
replaced by a fake whose train() returns the fixed arrays of csv_case.py, so that only the reference's writer code
(main.py:43-68) runs.  munch / matplotlib are not installed: they are stubbed (main.py uses munch.munchify only and never
touches matplotlib).  Run in the build container:  python tests/golden/make_golden_csv.py  ->  tests/golden/csv/*"""
import json
import os
import shutil
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "tf_shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, HERE)



class Munch(dict):
    """stand-in for munch.Munch: attribute access, and dir() lists the keys (main.py:47-48 relies on that)"""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __dir__(self):
        return list(self.keys())


def munchify(obj):
    if isinstance(obj, dict):
        return Munch((k, munchify(v)) for k, v in obj.items())
    if isinstance(obj, list):
        return [munchify(v) for v in obj]
    return obj


munch = types.ModuleType("munch")
munch.munchify = munchify
sys.modules["munch"] = munch
mpl = types.ModuleType("matplotlib")
mpl.pyplot = types.ModuleType("matplotlib.pyplot")
sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, mpl.pyplot

import csv_case  # noqa: E402
import main as ref_main  # noqa: E402
assert ref_main.__file__.startswith("/root/reference"), ref_main.__file__


class FakeSolver:
    def __init__(self, config, bsde):
        pass

    def train(self):
        return csv_case.fake_train_result()


ref_main.ActorCriticSolver = FakeSolver
import tensorflow as tf  # noqa: E402  (the shim)
if not hasattr(tf.keras.backend, "set_floatx"):
    tf.keras.backend.set_floatx = lambda dtype: None

with tempfile.TemporaryDirectory() as tmp:
    cfg_path = os.path.join(tmp, "csvcase.json")
    json.dump(csv_case.CONFIG, open(cfg_path, "w"))
    os.chdir(tmp)
    ref_main.FLAGS(["main.py", "--config_path=" + cfg_path])
    ref_main.main([])
    out = os.path.join(HERE, "csv")
    os.makedirs(out, exist_ok=True)
    for f in sorted(os.listdir(os.path.join(tmp, "logs"))):
        shutil.copy(os.path.join(tmp, "logs", f), os.path.join(out, f))
        print("wrote", f)
