"""Run bench.py --no-extras --no-cpu-baseline in-process N times is overkill: this prints the few headline numbers of one run.
   python tools/bench_short.py [label] [extra bench args...]"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
label = sys.argv[1] if len(sys.argv) > 1 else ""
out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-extras", "--no-cpu-baseline", *sys.argv[2:]],
                     capture_output=True, text=True)
lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
if not lines:
    print(label, "FAILED", out.stderr[-1500:])
    sys.exit(1)
d = json.loads(lines[0])
print(f"{label}: ttft {d['ms_per_step']:.4f} ms  e2e {d['e2e']['ttft_ms']:.4f} ms  decode {d['decode_ms_per_token']:.4f} ms/tok  "
      f"copied-splice step {d.get('ms_per_step_with_copied_splice', float('nan')):.4f} ms")
