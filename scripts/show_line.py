"""Prints the headline fields of bench.py JSON lines: python scripts/show_line.py file..."""
import json
import sys

for path in sys.argv[1:]:
    for x in open(path).read().strip().splitlines():
        if not x.startswith("{"):
            print(path, "|", x[:200])
            continue
        d = json.loads(x)
        r = d.get("roofline") or {}
        e = d.get("e2e") or {}
        print(f"{path}: N={d.get('n_gpus')} {d.get('scaling')} value={d.get('value'):.4g} step={d.get('ms_per_step'):.3f}ms "
              f"match={d.get('match_ms', 0):.3f} xchg={d.get('exchange_gather_ms', 0) or 0:.3f} ransac={d.get('ransac_ms', 0) or 0:.3f} "
              f"hyps/s={d.get('ransac_hyps_per_s') or 0:.3g} kern={r.get('kernel_ms', 0):.3f}ms frac={r.get('frac', 0):.3f} "
              f"launches={d.get('gpu_launches')} e2e={e.get('value', 0):.4g} e2e_match_ms={e.get('match_ms', 0) or 0:.3f} "
              f"pair_ms={(e.get('pair') or {}).get('ms', 0) or 0:.3f} pairs/s={d.get('pairs_per_s') or 0:.1f}")
        if d.get("roofline_scoring"):
            s = d["roofline_scoring"]
            print("    scoring: %.3f ms, frac %.3f, evaluated %.3f" % (s["kernel_ms"], s["frac"], s["evaluated_fraction"]), s["pruning"])
        if d.get("cpu_baseline"):
            print("    cpu:", d["cpu_baseline"]["value"], d["cpu_baseline"].get("ransac_hyps_per_s"), d["cpu_baseline"]["cores"], "ctx:", d.get("cpu_context"))
