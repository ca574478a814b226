"""Time pc_prefill_attn alone at given (H, Hkv, D, q, N) points (same method as bench.py's suffix sweep).
   python tools/attn_point.py 32,32,128,1024,8192 71,1,64,256,2048 ..."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    from bench import attention_point, load_peaks
    from promptcache_b200 import _native as nat, build
    build.build()
    nat.load()
    peak, _, tf = load_peaks()
    dev = torch.device("cuda", 0)
    for spec in sys.argv[1:]:
        H, Hkv, D, q, N = (int(x) for x in spec.split(","))
        p = attention_point(nat, dev, H, Hkv, D, q, N, peak, tf, rounding=(0 if D == 64 else None), launches=8 if q >= 256 else 16)
        print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in p.items() if not k.startswith("algorithmic")}))


if __name__ == "__main__":
    main()
