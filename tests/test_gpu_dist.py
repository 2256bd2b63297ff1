"""`-m gpu`: the multi-GPU merge path fed by real device partials (b2_exec_agg_partials), world size 1..N.
With one GPU the collectives degenerate to the identity; the multi-rank behaviour of the same functions is covered on
CPU by tests/test_dist_gloo.py and on 2+ GPUs by `bench.py --gpus N`."""
import pytest
import torch

import kvfmt
import orc
import scenarios as sc
from tikv_b200 import dist as bd
from tikv_b200 import ffi
from tikv_b200.executor import BatchExecutor, DeviceRegion
from tikv_b200.plan import Plan, col, const_int

pytestmark = pytest.mark.gpu


def test_agg_partials_merge_two_shards():
    """Two region shards scanned separately on the GPU; their device-resident partial tables are merged like two ranks'."""
    region = sc.dirty_region(6, n_keys=1500).build(read_ts=sc.READ_TS)
    dev = DeviceRegion(region)
    plan = (Plan().table_scan(sc.TABLE, sc.COLUMNS)
            .aggregation([("count", const_int(1)), ("sum", col(sc.C1)), ("avg", col(sc.C3, unsigned=True))], group_by=[col(sc.C2)]).build())
    parts = []
    for lo, hi in ((-1000, 2000), (2000, 10000)):
        with BatchExecutor(plan, [kvfmt.table_range(sc.TABLE, lo, hi)], dev) as ex:
            r = ex.next_batch(1 << 30)
            assert r.error is None and r.is_drained
            parts.append(tuple(t.clone() for t in bd.agg_partials_as_tensors(ex, 0)))
    keys = torch.cat([p[0] for p in parts]); nul = torch.cat([p[1] for p in parts]); acc = torch.cat([p[2] for p in parts])
    k, n, a = bd.merge_agg_partials(keys, nul, acc)
    assert a.shape[1] == 1 + 3 + 3  # COUNT | SUM(int) | AVG(int)
    got = [((None if n[i] else int(k[i])), int(a[i, 0]), bd.limbs_to_int(int(a[i, 2]), int(a[i, 3])), int(a[i, 4]),
            bd.limbs_to_int(int(a[i, 5]), int(a[i, 6]), unsigned=True)) for i in range(k.shape[0])]
    exp = orc.dag_handle(plan, sc.WHOLE, region)
    # oracle columns: count, sum, avg_count, avg_sum, key
    want = [(r[4], r[0], r[1] if r[1] is not None else 0, r[2], r[3] if r[3] is not None else 0) for r in exp.rows()]
    key = lambda t: (t[0] is not None, t[0] or 0)
    assert sorted(got, key=key) == sorted(want, key=key) and len(want) > 50


def test_checksum_partition_merge():
    region = sc.dirty_region(7, n_keys=800).build(read_ts=sc.READ_TS, n_write_blocks=2)
    from tikv_b200.executor import checksum
    dev = DeviceRegion(region)
    parts = [checksum([kvfmt.table_range(sc.TABLE, lo, hi)], dev)[1] for lo, hi in ((-1000, 300), (300, 10000))]
    x = parts[0][0] ^ parts[1][0]
    assert (x, parts[0][1] + parts[1][1], parts[0][2] + parts[1][2]) == orc.checksum(sc.WHOLE, region)[1]
    assert bd.merge_checksum(*parts[0]) == parts[0]


def test_min_max_partials_merge_two_shards():
    """MAX / MIN partial states of two shards: counts add, extremum keys merge by unsigned maximum (max_word_mask)."""
    import ctypes as C
    region = sc.dirty_region(8, n_keys=1200).build(read_ts=sc.READ_TS)
    dev = DeviceRegion(region)
    plan = (Plan().table_scan(sc.TABLE, sc.COLUMNS)
            .aggregation([("max", col(sc.C1)), ("min", col(sc.C3, unsigned=True)), ("count", const_int(1))], group_by=[col(sc.C6, tp=ffi.TP_LONG)]).build())
    parts, mask = [], 0
    for lo, hi in ((-1000, 1500), (1500, 10000)):
        with BatchExecutor(plan, [kvfmt.table_range(sc.TABLE, lo, hi)], dev) as ex:
            r = ex.next_batch(1 << 30)
            assert r.error is None and r.is_drained
            parts.append(tuple(t.clone() for t in bd.agg_partials_as_tensors(ex, 0)))
            p = ffi.AggPartials()
            assert ffi.lib().b2_exec_agg_partials(ex._h, C.byref(p)) == 0
            mask = p.max_word_mask
    assert mask == 0b1010  # words: [cnt, key] [cnt, key] [cnt]
    keys = torch.cat([p[0] for p in parts]); nul = torch.cat([p[1] for p in parts]); acc = torch.cat([p[2] for p in parts])
    k, n, a = bd.merge_agg_partials(keys, nul, acc, max_words=[w for w in range(acc.shape[1]) if mask >> w & 1])
    def dec(key, unsigned, is_min):
        key &= (1 << 64) - 1
        if is_min:
            key ^= (1 << 64) - 1
        if not unsigned:
            key ^= 1 << 63
            return key - (1 << 64) if key >= 1 << 63 else key
        return key
    got = sorted(((None if n[i] else int(k[i])), dec(int(a[i, 1]), False, False) if a[i, 0] else None, dec(int(a[i, 3]), True, True) if a[i, 2] else None, int(a[i, 4]))
                 for i in range(k.shape[0]))
    exp = orc.dag_handle(plan, sc.WHOLE, region)
    want = sorted((r[3], r[0], (r[1] & ((1 << 64) - 1)) if r[1] is not None else None, r[2]) for r in exp.rows())
    assert got == want and len(want) >= 10


def test_multi_column_partials_merge_two_shards():
    """BatchSlowHashAggregation partial states of two shards (b2_agg_partials.key_words = 2): composite keys + NULL masks
    merge to the whole-table result."""
    region = sc.dirty_region(9, n_keys=1200).build(read_ts=sc.READ_TS)
    dev = DeviceRegion(region)
    plan = (Plan().table_scan(sc.TABLE, sc.COLUMNS)
            .aggregation([("count", const_int(1)), ("count", col(sc.C1))], group_by=[col(sc.C6, tp=ffi.TP_LONG), col(sc.C2)]).build())
    parts = []
    for lo, hi in ((-1000, 1500), (1500, 10000)):
        with BatchExecutor(plan, [kvfmt.table_range(sc.TABLE, lo, hi)], dev) as ex:
            r = ex.next_batch(1 << 30)
            assert r.error is None and r.is_drained
            parts.append(tuple(t.clone() for t in bd.agg_partials_as_tensors(ex, 0)))
    keys = torch.cat([p[0] for p in parts]); nul = torch.cat([p[1] for p in parts]); acc = torch.cat([p[2] for p in parts])
    assert keys.dim() == 2 and keys.shape[1] == 2
    k, n, a = bd.merge_agg_partials(keys, nul, acc)
    nk = lambda t: tuple((0, 0) if x is None else (1, x) for x in t)
    got = sorted(((int(a[i, 0]), int(a[i, 1]), None if int(n[i]) & 1 else int(k[i, 0]), None if int(n[i]) & 2 else int(k[i, 1])) for i in range(k.shape[0])), key=nk)
    want = sorted(orc.dag_handle(plan, sc.WHOLE, region).rows(), key=nk)
    assert got == want and len(want) >= 20
