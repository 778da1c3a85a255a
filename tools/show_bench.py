"""Print the numbers of one bench.py JSON line that matter at a glance:  python tools/show_bench.py <bench.json>"""
import json
import sys

d = json.load(open(sys.argv[1]))
k = d["roofline"]["kernels"]
print(f"N={d['n_gpus']}: {d['ms_per_step']:.2f} ms per price (device loop), e2e {d['e2e']['ms_per_step']:.2f} ms, {d['value']:.3e} {d['unit']}; "
      f"generator {k['rbergomi_paths_kernel']['ms']:.2f} ms, LSM {1e3 * d['lsm_price_time_s']:.2f} ms "
      f"({1e3 * k['lsm_sweep_kernel']['avg_ms_per_sweep_step']:.1f} us per step, {k['lsm_sweep_kernel']['frac_hbm']:.3f} of HBM); price {d['price']:.6f}")
print("clocks", d["clocks"])
if d.get("parity_mode"):
    print(f"parity mode: {d['parity_mode']['us_per_sweep_step']:.1f} us per step, {d['parity_mode']['frac_hbm']:.3f} of HBM on 20 B")
if d.get("policy_value") and "policy_value" in d["policy_value"]:
    p = d["policy_value"]
    print(f"policy value {p['policy_value']:.5f} +- {p['policy_value_std_error']:.5f} (in-sample {p['in_sample_price']:.5f}), pass {p['pass_ms']:.1f} ms")
if d.get("shard_parity"):
    print("shard parity", d["shard_parity"]["shard_parity_rel"], "first-exercise mismatches", d["shard_parity"]["first_exercise_mismatches"])
for c, v in (d.get("configs") or {}).items():
    print(c, {a: b for a, b in v.items() if a not in ("workload", "ncu")})
if d.get("cpu_baseline"):
    print(f"cpu baseline {d['cpu_baseline']['value']:.3e} on {d['cpu_baseline']['cores']} cores")
