"""The eval_sys-style harness (tools/eval_sys.py; reference eval_sys.py:88-215,300-375) runs end to end on the GPU."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode,lens", [("critical_point", ["16", "300", "1024"]), ("latency", ["256", "1024"])])
def test_eval_sys_harness(tmp_path, mode, lens):
    out = tmp_path / f"{mode}.json"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "eval_sys.py"), mode, "--layers", "2", "--repeat", "1", "--out", str(out),
                        "--seq-lens", *lens], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    d = json.load(open(out))
    assert len(d["results"]) == len(lens)
    for rec in d["results"]:
        if mode == "critical_point":
            assert rec["upload_ms"] > 0 and rec["splice_copy_ms"] > 0 and rec["recompute_ms"] > 0 and rec["virtual_splice_ms"] > 0
        else:
            assert rec["cache_response_time_ms"] > 0 and rec["no_cache_response_time_ms"] > 0 and rec["speedup"] > 0
            assert rec["no_cache_cache_time_ms"] < 5.0  # no_cache: nothing is spliced
