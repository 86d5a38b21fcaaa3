"""The measurement contract of bench.py, checked on the CPU through its reference arm (the oracle port timed on
a bounded sample of the bench workload): ONE JSON line on stdout with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, PFX_BENCH_SAMPLE_SIZES="12,16,20", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "sen_shear", "--steps", "1",
                        "--warmup", "1"],
                       capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "staggered_load_steps_per_second" and d["unit"] == "steps/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["config"]["workload"].startswith("sen_shear_2d_spectral_n") and d["value"] > 0
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and d["cpu_baseline"]["kind"] == "port"
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] == 1
    assert len(d["cpu_baseline"]["points"]) == 3 and d["cpu_baseline"]["extrapolated"] is True
    assert 0.3 < d["cpu_baseline"]["fit_exponent"] < 3.0
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_parity_block_compares_against_the_golden_records():
    """`parity` of the bench line: reactions / energies / stagger count of a load step against the committed golden
    run (tests/golden/bench_golden.json); run-independent logic checked on a synthetic golden."""
    sys.path.insert(0, ROOT)
    import bench
    gold = {"w": {"how": "synthetic", "steps": {"4": {"stagger": 3, "reactions": [[1.0, -2.0, 0.0], [0.0, 0.0, 0.0]],
                                                      "energies": {"E": 2.0, "W": 0.0, "gamma": 1e-3}}}}}
    ok = bench.parity_block("w", 4, [[1.0 + 1e-7, -2.0, 0.0], [0.0, 0.0, 0.0]], {"E": 2.0 * (1 + 5e-9), "W": 0.0, "gamma": 1e-3}, 3, golden=gold)
    assert ok["available"] and ok["ok"] and abs(ok["reaction_rel_diff"] - 5e-8) < 1e-12 and abs(ok["energy_rel_diff"] - 5e-9) < 1e-12
    bad_r = bench.parity_block("w", 4, [[1.0, -2.0 - 1e-5, 0.0], [0.0, 0.0, 0.0]], gold["w"]["steps"]["4"]["energies"], 3, golden=gold)
    assert not bad_r["ok"] and bad_r["reaction_rel_diff"] > 1e-6
    bad_e = bench.parity_block("w", 4, gold["w"]["steps"]["4"]["reactions"], {"E": 2.0, "W": 0.0, "gamma": 1e-3 * (1 + 1e-6)}, 3, golden=gold)
    assert not bad_e["ok"] and bad_e["worst_energy"] == "gamma"
    bad_s = bench.parity_block("w", 4, gold["w"]["steps"]["4"]["reactions"], gold["w"]["steps"]["4"]["energies"], 4, golden=gold)
    assert not bad_s["ok"]
    assert bench.parity_block("w", 5, [], {}, 0, golden=gold)["available"] is False
    assert bench.parity_block("other", 4, [], {}, 0, golden=gold)["available"] is False
