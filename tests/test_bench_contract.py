"""bench.py's contract, as far as it can be checked without a GPU: the reference arm (the oracle restatement on the
host) prints exactly one JSON line with the keys the driver reads, does not load the product library, and the
workload table names BASELINE.json's configurations."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "c2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "bin-states/s" and d["unit"] == "bin-states/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1
    assert "configs[1]" in d["config"]["workload"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "bin-states/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_does_not_load_the_product_library():
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--config', 'c2', '--steps', '1', '--warmup', '0'];"
            "runpy.run_path(%r, run_name='__main__');"
            "import os; maps = open('/proc/self/maps').read();"
            "assert 'libehmm_b200' not in maps, 'the product library was mapped';"
            "assert 'epigrahmm_b200' not in sys.modules") % os.path.join(ROOT, "bench.py")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]


def test_workloads_name_the_baseline_configurations():
    sys.path.insert(0, ROOT)
    import importlib
    bench = importlib.import_module("bench")
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert set(bench.CONFIGS) == {"c2", "c3", "c4", "c5"}
    for name, idx, words in (("c2", 1, ("15", "4 replicates", "consensus")), ("c3", 2, ("3 conditions", "2 replicates", "differential")),
                             ("c4", 3, ("5 epigenomic marks", "2 replicates", "500 bp")), ("c5", 4, ("60", "4 conditions", "3 rep"))):
        w = bench.CONFIGS[name]["workload"]
        assert "configs[%d]" % idx in w
        for word in words:
            assert word.lower() in w.lower() and word.split()[0].lower() in base["configs"][idx].lower(), (name, word)
    # the default configuration is the one the metric's target is quoted on (north_star: 3 conditions x 2 replicates)
    assert "--config" in open(os.path.join(ROOT, "bench.py")).read() and 'default="c3"' in open(os.path.join(ROOT, "bench.py")).read()
