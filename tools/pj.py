"""print the key fields of a bench.py JSON line (stdin)"""
import json, sys
tag = sys.argv[1] if len(sys.argv) > 1 else ""
try:
    d = json.loads(sys.stdin.read().strip().splitlines()[-1])
    r = d["roofline"]
    print(f"{tag:28s} ms/it {d['ms_per_step']:8.2f}  Mupd/s {d['value']/1e6:7.2f}  dot {r['avg_launch_ms']*1e3:7.1f} us  frac {r['frac']:.3f}  dot_share {r['dot_share_of_step']:.2f}  "
          f"us/step {r.get('per_step_us')}  e2e {d['e2e']['value']/1e6:6.2f}  pub/it {d['chain']['published_per_iter']:.0f}  sigE {d['chain']['sigmaE']:.3f}  clk {d['clocks']['sm_mhz']}")
except Exception as ex:
    print(tag, "FAILED", ex)
