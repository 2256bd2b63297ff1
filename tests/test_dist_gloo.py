"""`-m "not gpu"`: the N>1 path on CPU — two gloo ranks, each owning half of the regions, partial results merged with
tikv_b200.dist exactly as the NCCL ranks do on GPUs.  Partial results here come from the oracle (no GPU); the merged
answer must equal the oracle's answer over the whole key space."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _i64(v):
    v &= (1 << 64) - 1
    return v - (1 << 64) if v >= (1 << 63) else v


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import kvfmt
    import orc
    import scenarios as sc
    from tikv_b200 import dist as bd
    from tikv_b200.plan import Plan, col, const_int

    region = sc.dirty_region(5, n_keys=400).build(read_ts=sc.READ_TS)
    # this rank's share of the key space (regions are disjoint handle ranges)
    cuts = [-1000, 500, 5000]
    mine = [kvfmt.table_range(sc.TABLE, cuts[rank], cuts[rank + 1])]

    # hash aggregation: partial (count, sum limbs) per group -> all_gather + re-aggregate
    plan = Plan().table_scan(sc.TABLE, sc.COLUMNS).aggregation([("count", const_int(1)), ("sum", col(sc.C1))], group_by=[col(sc.C2)]).build()
    part = orc.dag_handle(plan, mine, region)
    keys = torch.tensor([0 if r[2] is None else r[2] for r in part.rows()], dtype=torch.int64)
    nul = torch.tensor([r[2] is None for r in part.rows()], dtype=torch.bool)
    acc = torch.tensor([[r[0], 1, _i64(r[1] & 0xFFFFFFFF), _i64(r[1] >> 32)] for r in part.rows()], dtype=torch.int64).reshape(-1, 4)
    k, n, a = bd.merge_agg_partials(keys, nul, acc)
    def keyf(t):
        return (t[0] is not None, t[0] or 0)
    merged = sorted((((None if n[i] else int(k[i])), int(a[i, 0]), bd.limbs_to_int(int(a[i, 2]), int(a[i, 3]))) for i in range(k.shape[0])), key=keyf)
    whole = orc.dag_handle(plan, sc.WHOLE, region)
    expect = sorted(((r[2], r[0], r[1]) for r in whole.rows()), key=keyf)
    ok_agg = merged == expect and len(expect) > 10

    # composite group keys (BatchSlowHashAggregation partials: [n, 2] key words + NULL masks), MAX merged by unsigned maximum
    mplan = Plan().table_scan(sc.TABLE, sc.COLUMNS).aggregation([("count", const_int(1)), ("max", col(sc.C6))], group_by=[col(sc.C2), col(sc.C6)]).build()
    part = orc.dag_handle(mplan, mine, region)
    rows = part.rows()  # (count, max, c2, c6)
    mkeys = torch.tensor([[0 if r[2] is None else r[2], 0 if r[3] is None else r[3]] for r in rows], dtype=torch.int64).reshape(-1, 2)
    mnul = torch.tensor([(1 if r[2] is None else 0) | (2 if r[3] is None else 0) for r in rows], dtype=torch.int64)
    # MAX state as the device keeps it: [count of non-NULL inputs, order-preserving key = value ^ sign bit]
    macc = torch.tensor([[r[0], 0 if r[1] is None else 1, 0 if r[1] is None else _i64(r[1] ^ (1 << 63))] for r in rows], dtype=torch.int64).reshape(-1, 3)
    k2, n2, a2 = bd.merge_agg_partials(mkeys, mnul, macc, max_words=(2,))
    nk = lambda t: tuple((0, 0) if x is None else (1, x) for x in t)
    mm = sorted(((int(a2[i, 0]), (_i64(int(a2[i, 2]) ^ (1 << 63)) if int(a2[i, 1]) else None), None if int(n2[i]) & 1 else int(k2[i, 0]), None if int(n2[i]) & 2 else int(k2[i, 1]))
                 for i in range(k2.shape[0])), key=nk)
    ok_agg = ok_agg and mm == sorted(orc.dag_handle(mplan, sc.WHOLE, region).rows(), key=nk) and len(mm) > 30

    # TopN: ORDER BY c2 DESC, c1 ASC LIMIT 40
    tplan = Plan().table_scan(sc.TABLE, sc.COLUMNS).topn([(col(sc.C2), True), (col(sc.C1), False)], 40).build(output_offsets=[sc.C_H, sc.C1, sc.C2])
    tp = orc.dag_handle(tplan, mine, region)
    cols = [torch.tensor([0 if v is None else v for v in c], dtype=torch.int64) for c in tp.columns]
    nls = [torch.tensor([v is None for v in c], dtype=torch.bool) for c in tp.columns]
    mc, mn = bd.merge_topn(cols, nls, [(2, True, "i64"), (1, False, "i64")], 40)
    got_rows = [tuple(None if mn[j][i] else int(mc[j][i]) for j in range(3)) for i in range(mc[0].shape[0])]
    ok_topn = got_rows == orc.dag_handle(tplan, sc.WHOLE, region).rows()

    # checksum: XOR of partial folds
    st, (c, kv, by), _ = orc.checksum(mine, region)
    mcsum = bd.merge_checksum(c, kv, by)
    ok_ck = mcsum == orc.checksum(sc.WHOLE, region)[1]

    q.put((rank, bool(ok_agg), bool(ok_topn), bool(ok_ck), bd.shard_blocks(7, world, rank)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_merge_matches_oracle():
    import subprocess
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1:4] for r in res] == [(True, True, True), (True, True, True)], res
    assert res[0][4] + res[1][4] == list(range(7))  # every block assigned exactly once, in order


def test_merge_single_process_semantics():
    sys.path.insert(0, ROOT)
    from tikv_b200 import dist as bd
    keys = torch.tensor([5, 7, 5, 0, 0], dtype=torch.int64)
    nul = torch.tensor([False, False, False, True, True])
    acc = torch.tensor([[1, 10, 0], [2, 20, 0], [3, (1 << 32) - 1, -1], [4, 1, 0], [5, 2, 0]], dtype=torch.int64)
    k, n, a = bd.merge_agg_partials(keys, nul, acc)
    rows = sorted((bool(n[i]), int(k[i]), [int(x) for x in a[i]]) for i in range(3))
    assert rows == [(False, 5, [4, 10 + (1 << 32) - 1, -1]), (False, 7, [2, 20, 0]), (True, 0, [9, 3, 0])]
    assert bd.limbs_to_int(10 + (1 << 32) - 1, -1) == 9 and bd.limbs_to_int(5, 3, unsigned=True) == 3 * (1 << 32) + 5
    # f64 word merge
    f = torch.tensor([[1, 0], [1, 0]], dtype=torch.int64)
    f[:, 1] = torch.tensor([1.5, 2.25], dtype=torch.float64).view(torch.int64)
    k, n, a = bd.merge_agg_partials(torch.tensor([1, 1]), torch.tensor([False, False]), f, real_words=(1,))
    assert a[0, 0] == 2 and a[0, 1:2].view(torch.float64).item() == 3.75
    # exact Real SUM words (66 carry-save digits, b2_device.h f64_acc_add): partial sums merge by integer addition and
    # round once: 1e308 + 1.0 - 1e308 + 2^-1074 is exactly 1 + 2^-1074 -> 1.0 (a sequential f64 sum would give 0.0 or 5e-324)
    vals = [1e308, 1.0, -1e308, 5e-324]
    parts = torch.tensor([[1] + bd.f64_acc_digits(v) for v in vals], dtype=torch.int64)
    k, n, a = bd.merge_agg_partials(torch.tensor([3, 3, 3, 3]), torch.tensor([False] * 4), parts)
    assert int(a[0, 0]) == 4 and bd.f64_acc_value(a[0, 1:].tolist()) == 1.0
    import math
    import random
    rng = random.Random(7)
    xs = [rng.uniform(-1, 1) * 10.0 ** rng.randrange(-300, 300) for _ in range(300)]
    tot = [sum(d) for d in zip(*[bd.f64_acc_digits(x) for x in xs])]
    assert bd.f64_acc_value(tot) == math.fsum(xs)
    # MAX / MIN extremum keys merge by unsigned maximum (word 1), their counts by addition (word 0)
    m = torch.tensor([[2, 5], [1, -3], [4, 9]], dtype=torch.int64)  # -3 is a huge unsigned key
    k, n, a = bd.merge_agg_partials(torch.tensor([8, 8, 9]), torch.tensor([False, False, False]), m, max_words=(1,))
    assert sorted((int(k[i]), int(a[i, 0]), int(a[i, 1])) for i in range(2)) == [(8, 3, -3), (9, 4, 9)]
    # composite keys (b2_agg_partials.key_words = 2): [n, 2] key words + NULL mask per group
    mk = torch.tensor([[1, 2], [1, 2], [1, 0], [1, 0], [0, 0]], dtype=torch.int64)
    mn = torch.tensor([0, 0, 2, 0, 3], dtype=torch.int64)  # (1,2) (1,2) (1,NULL) (1,0) (NULL,NULL)
    k, n, a = bd.merge_agg_partials(mk, mn, torch.tensor([[1], [2], [4], [8], [16]], dtype=torch.int64))
    assert sorted((int(n[i]), [int(x) for x in k[i]], int(a[i, 0])) for i in range(k.shape[0])) == [(0, [1, 0], 8), (0, [1, 2], 3), (2, [1, 0], 4), (3, [0, 0], 16)]
    # TopN with NULLs: asc puts NULL first, desc puts NULL last
    c, nl = bd.merge_topn([torch.tensor([3, 1, 0, 2])], [torch.tensor([False, False, True, False])], [(0, False, "i64")], 3)
    assert [None if nl[0][i] else int(c[0][i]) for i in range(3)] == [None, 1, 2]
    c, nl = bd.merge_topn([torch.tensor([3, 1, 0, 2])], [torch.tensor([False, False, True, False])], [(0, True, "i64")], 4)
    assert [None if nl[0][i] else int(c[0][i]) for i in range(4)] == [3, 2, 1, None]
    c, nl = bd.merge_topn([torch.tensor([-1, 1, 5])], [torch.tensor([False, False, False])], [(0, False, "u64")], 3)
    assert [int(x) for x in c[0]] == [1, 5, -1]  # -1 is u64::MAX
    assert bd.merge_checksum(7, 1, 2) == (7, 1, 2)
