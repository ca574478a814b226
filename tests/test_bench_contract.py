"""The committed bench line (profiles/bench_r01_*.json, produced by `python bench.py` on a B200) carries every key the
measurement contract asks for; the CPU arm of bench.py runs here."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _latest():
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "bench_r0*_v*.json")))
    assert files, "no committed bench line"
    return json.load(open(files[-1]))


def test_committed_bench_line_has_the_contract_keys():
    d = _latest()
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "clocks"):
        assert k in d, k
    assert d["unit"] == "tok/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "f16" and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    assert abs(d["value"] - d["n_gpus"] * (d["config"]["n_cached"] + d["config"]["q_len"]) / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    e = d["e2e"]
    assert e["unit"] == "tok/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or 0.9 < r["traffic"] / r["algorithmic_bytes_per_launch"] < 1.1  # no wasted re-reads
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"])


def _run_reference_arm(extra_env, *flags):
    env = dict(os.environ, **extra_env)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", *flags],
                         capture_output=True, text=True, env=env, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "tok/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"] and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    return d


def test_reference_arm_runs_the_unmodified_reference_on_cpu():
    """`bench.py --impl reference` = the reference's OWN CacheEngine.process -> GenerationEngine.generate from the staged
    copy oracle/_ref (here with one decoder layer to stay within the CPU suite's budget; the driver runs all 32)."""
    from oracle import ref_shim
    if not ref_shim.reference_available():
        import pytest
        pytest.skip("oracle/_ref not staged")
    d = _run_reference_arm({}, "--config", "C1", "--layers", "1")
    assert d["cpu_baseline"]["kind"] == "reference" and "UNMODIFIED reference" in d["cpu_baseline"]["sample"]
    assert d["config"]["n_cached"] == 600 and d["config"]["q_len"] == 16


def test_reference_arm_falls_back_to_the_oracle_port():
    d = _run_reference_arm({"PC_BENCH_CPU": "port", "PC_BENCH_CPU_LAYERS": "1"}, "--config", "C1", "--layers", "2")
    assert d["cpu_baseline"]["kind"] == "port"


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0 and not [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
