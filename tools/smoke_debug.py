"""GPU debugging aid: the steps of smoke() one by one, each reported (run with env switches to bisect a failure)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import orc
import scenarios as sc
from compare import assert_same_rows
from tikv_b200 import smoke
from tikv_b200.executor import DagHandler, DeviceRegion

seed, keys = int(os.environ.get("SEED", 3)), int(os.environ.get("KEYS", 2000))
host = sc.dirty_region(seed, n_keys=keys).build(read_ts=sc.READ_TS, n_write_blocks=int(os.environ.get("BLOCKS", 2)))
dev = DeviceRegion(host)
scan_filter, hash_agg = smoke.plans()
order = os.environ.get("ORDER", "sf:host,sf:dev,agg:host,agg:dev").split(",")
for step in order:
    pn, rn = step.split(":")
    plan, region = (scan_filter if pn == "sf" else hash_agg), (host if rn == "host" else dev)
    got = DagHandler(plan, sc.WHOLE, region).handle_request()
    exp = orc.dag_handle(plan, sc.WHOLE, host)
    try:
        assert_same_rows(got, exp, ordered=pn == "sf", ctx=step)
        print(step, "ok", got.n_rows, flush=True)
    except AssertionError as e:
        print(step, "FAILED", str(e)[:200], flush=True)
        break
