"""Development aid: the numbers of a bench.py JSON line that matter while iterating."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d.get("roofline", {})
e = d.get("e2e", {})
print(f"  value {d['value']:.1f} {d['unit']} ({d['ms_per_step']:.4f} ms/step, n_gpus {d['n_gpus']}, steps {d['steps']}); e2e {e.get('value', 0):.1f} "
      f"(init_ms {e.get('init_ms')}); root {r.get('ms_per_launch')} ms frac {r.get('frac')}; child {r.get('child_hist_ms_per_step')} ms "
      f"frac {r.get('child_hist_frac')}; lambda {r.get('lambda_ms_per_step')} ms; with events {r.get('ms_per_step_with_event_nodes')}; "
      f"hash {d.get('tree_hash')}; ndcg {d['config'].get('ndcg_at_10_T')}")
if "parity" in d:
    print("  parity", d["parity"])
if "cpu_baseline" in d:
    print("  cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
