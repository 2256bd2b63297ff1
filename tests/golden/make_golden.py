"""Regenerates tests/golden/scenario_plans.json: the oracle's results for every plan of scenarios.plans(), minmax_plans(),
in_plans(), multi_group_plans() (minus the f64-SUM one) and limit_plans() on dirty_region(1, n_keys=600).

The oracle is pinned on the reference's own vectors; these files freeze its answers so that (a) a change to
the oracle that moves any of them is visible in review, and (b) the CUDA path is also compared with committed bytes, not
only with whatever the oracle computes on the day.  Run from the repo root:  python tests/golden/make_golden.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import orc  # noqa: E402
import scenarios as sc  # noqa: E402


def cases():
    out = []
    for group in (sc.plans(), sc.minmax_plans(), sc.in_plans(), [p for p in sc.multi_group_plans() if p[0] != "mg_same_expr_twice"], sc.limit_plans()):
        out += [(n, p) for n, p in group]
    return out


def region():
    return sc.dirty_region(1, n_keys=600).build(read_ts=sc.READ_TS, n_write_blocks=2)


def unordered(name):
    return sc.is_agg(name) or "group" in name or name.startswith(("mg_", "minmax_"))


def canon(rows, name):
    """JSON-able rows: floats as hex strings (exact), aggregation results sorted (group order is unspecified)."""
    enc = [[("f:" + float(v).hex()) if isinstance(v, float) else v for v in r] for r in rows]
    if unordered(name):
        enc.sort(key=lambda r: [(0, "") if v is None else (1, str(v)) for v in r])
    return enc


def main():
    reg = region()
    doc = {}
    for name, plan in cases():
        res = orc.dag_handle(plan, sc.split_ranges(), reg)
        doc[name] = {"status": res.status, "rows": canon(res.rows(), name)}
    with open(os.path.join(HERE, "scenario_plans.json"), "w") as f:
        json.dump(doc, f, separators=(",", ":"), sort_keys=True)
    print(len(doc), "plans,", sum(len(v["rows"]) for v in doc.values()), "rows")


if __name__ == "__main__":
    main()
