"""The drop-in main.py writes the reference's files byte for byte: the two CSVs (training history, figure data) are
compared with the golden files that the REFERENCE's own main.py wrote from the same arrays
(tests/golden/make_golden_csv.py); the config dump is compared as JSON.  CPU only: the solver is replaced by a fake whose
train() returns the fixed arrays of tests/golden/csv_case.py, so only the writer code of main.py (reference main.py:43-68) runs."""
import json
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "csv")


def test_main_writes_the_reference_files(tmp_path):
    script = textwrap.dedent(f"""
        import json, os, sys
        sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'tests', 'golden')!r})
        import csv_case
        from deeppde_actorcritic_b200 import main as M
        class FakeSolver:
            def __init__(self, *a, **k): pass
            def train(self): return csv_case.fake_train_result()
        M.ActorCriticSolver = FakeSolver
        json.dump(csv_case.CONFIG, open('csvcase.json', 'w'))
        M.FLAGS(['main.py', '--config_path=csvcase.json'])
        M.main([])
    """)
    r = subprocess.run([sys.executable, "-c", script], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    logs = tmp_path / "logs"
    names = sorted(os.listdir(GOLD))
    assert sorted(os.listdir(logs)) == names
    for f in names:
        got, ref = (logs / f).read_bytes(), open(os.path.join(GOLD, f), "rb").read()
        if f.endswith(".json"):
            assert json.loads(got) == json.loads(ref), f
        else:
            assert got == ref, f"{f} differs from the file the reference's main.py wrote"
