"""`-m gpu`: the device block reader (b2_sst_decode) and its tooling twin (b2_sst_encode) against the oracle's RocksDB
data-block builder / iterator, bit-exact; requests over decoded blocks against the oracle's answers; a bench-sized
encode -> decode round trip."""
import ctypes as C

import numpy as np
import pytest

import kvfmt
import orc
import scenarios as sc
import sstfmt
from compare import assert_same_rows
from tikv_b200 import ffi
from tikv_b200.executor import B2Error, DagHandler, DeviceRegion, SstDecoder, SstRegion, checksum

pytestmark = pytest.mark.gpu


def device_block_kvs(blk):
    """device-resident b2_cf_block -> [(key, value)] (+ checks the padding contract: 16 readable bytes past the heaps)"""
    L = ffi.lib()
    n = blk.n
    ko, vo = np.zeros(n + 1, dtype=np.uint32), np.zeros(n + 1, dtype=np.uint32)
    assert L.b2_copy_to_host(0, ko.ctypes.data, blk.key_offs, 4 * (n + 1)) == 0
    assert L.b2_copy_to_host(0, vo.ctypes.data, blk.val_offs, 4 * (n + 1)) == 0
    assert ko[0] == 0 and vo[0] == 0
    keys, vals = np.zeros(int(ko[n]) + 16, dtype=np.uint8), np.zeros(int(vo[n]) + 16, dtype=np.uint8)
    assert L.b2_copy_to_host(0, keys.ctypes.data, blk.keys, len(keys)) == 0
    assert L.b2_copy_to_host(0, vals.ctypes.data, blk.vals, len(vals)) == 0
    kb, vb = keys.tobytes(), vals.tobytes()
    return [(kb[ko[i]:ko[i + 1]], vb[vo[i]:vo[i + 1]]) for i in range(n)]


OPTS = [dict(), dict(restart_interval=1), dict(restart_interval=3, block_size=700), dict(block_size=0, entries_per_block=37),
        dict(key_prefix_len=0, key_suffix_len=0, trailer_len=0, block_size=4096), dict(key_prefix_len=3, key_prefix_byte=0x7a, trailer_len=0)]


@pytest.mark.parametrize("opts", OPTS)
def test_decode_matches_the_block_iterator(opts):
    host = sc.dirty_region(7, n_keys=900).build(read_ts=sc.READ_TS)
    fmt = {k: opts[k] for k in ("trailer_len", "key_prefix_len", "key_suffix_len") if k in opts}
    with SstDecoder() as dec:
        for blk in host.wblocks + [host.dblock]:
            data, offs = sstfmt.build(blk, **opts)
            rc, exp = sstfmt.decode(data, offs, **fmt)
            assert rc == 0 and exp == blk.kvs
            got, st = dec.decode(data, offs, **fmt)  # (the handle is reused: its buffers only grow)
            assert device_block_kvs(got) == exp
            assert st.n_entries == blk.n and st.key_bytes == sum(len(k) for k, _ in exp) and st.val_bytes == sum(len(v) for _, v in exp)
            assert st.h2d_bytes == len(data) + 8 * len(offs)
        # blocks with gaps between them (BlockHandle offsets into a file image) and a device-resident image
        blk = host.wblocks[0]
        data, offs = sstfmt.build(blk, **dict(opts, block_size=900) if not opts.get("entries_per_block") else opts)
        img, at, noffs = b"", [], []
        for a, b in zip(offs, offs[1:]):
            img += b"\xee" * 3
            noffs.append((len(img), len(img) + b - a))
            img += data[a:b]
        img += b"\xee" * 16
        # gaps only work when slices are given one by one; the ABI takes ascending offsets, so decode block by block here
        for (a, b), (lo, hi) in list(zip(noffs, zip(offs, offs[1:])))[:5]:
            got, _ = dec.decode(img, [a, b], **fmt)
            assert device_block_kvs(got) == sstfmt.decode(data[lo:hi], [0, hi - lo], **fmt)[1]
        import torch
        t = torch.from_numpy(np.frombuffer(data + b"\0" * 16, dtype=np.uint8).copy()).cuda()
        got, st = dec.decode(t.data_ptr(), offs, location=ffi.LOC_DEVICE, **fmt)
        assert device_block_kvs(got) == blk.kvs and st.h2d_bytes == 8 * len(offs)
        got, st = dec.decode(b"", [0], **fmt)
        assert got.n == 0 and device_block_kvs(got) == []


def test_long_keys_and_concurrent_handles():
    """keys beyond the 128 bytes a warp holds in registers take the sequential path from that entry on (mixed with short keys
    inside one restart interval); two threads decode at once (two staging buffers per device)."""
    import random
    from concurrent.futures import ThreadPoolExecutor
    rng = random.Random(5)
    kvs, k = [], b""
    for i in range(3000):
        stem = b"k%06d" % i
        k = stem + (bytes(rng.randrange(97, 123) for _ in range(rng.choice((0, 3, 60, 118, 119, 120, 121, 130, 200, 400)))) if rng.random() < 0.3 else b"")
        kvs.append((k, bytes(rng.randrange(256) for _ in range(rng.choice((0, 1, 31, 32, 33, 100, 300))))))
    blk = kvfmt.HostBlock(kvs)
    runs = [sstfmt.build(blk, restart_interval=r, block_size=bs) for r, bs in ((16, 32768), (5, 3000), (16, 100000), (1, 4096))]
    decs = [SstDecoder() for _ in runs]
    try:
        for d, (data, offs) in zip(decs, runs):
            assert device_block_kvs(d.decode(data, offs)[0]) == blk.kvs
        with ThreadPoolExecutor(max_workers=4) as pool:
            for _ in range(3):
                outs = list(pool.map(lambda a: device_block_kvs(a[0].decode(a[1][0], a[1][1])[0]), zip(decs, runs)))
                assert all(o == blk.kvs for o in outs)
    finally:
        for d in decs:
            d.close()


def test_rejected_blocks():
    blk = kvfmt.HostBlock([(b"apple", b"v0"), (b"apply", b""), (b"banana", b"abc")])
    good, offs = sstfmt.build(blk, restart_interval=2)
    foot = b"\x01" + b"\0" * 7
    pos = good.index(foot)
    with SstDecoder() as dec:
        assert device_block_kvs(dec.decode(good, offs)[0]) == blk.kvs
        for bad, status in ((good[:pos] + b"\0" + good[pos + 1:], ffi.B2_ERR_UNSUPPORTED),   # Delete tombstone
                            (good[:-6] + b"\x80" + good[-5:], ffi.B2_ERR_UNSUPPORTED),       # hash index flag
                            (good[:7], ffi.B2_ERR_STORAGE), (b"\x7f" * 40, ffi.B2_ERR_STORAGE)):
            with pytest.raises(B2Error) as e:
                dec.decode(bad, [0, len(bad)])
            assert e.value.status == status
        # single bit flips: same verdict class as the oracle's iterator, never a crash; accepted blocks decode identically
        import random
        rng = random.Random(11)
        for _ in range(120):
            bad = bytearray(good)
            bad[rng.randrange(len(bad) - 5)] ^= 1 << rng.randrange(8)
            rc, exp = sstfmt.decode(bytes(bad), [0, len(bad)])
            try:
                got = device_block_kvs(dec.decode(bytes(bad), [0, len(bad)])[0])
                assert rc == 0 and got == exp
            except B2Error as e:
                assert rc != 0 and e.status == (ffi.B2_ERR_UNSUPPORTED if rc == 2 else ffi.B2_ERR_STORAGE)


def test_encoder_twin_is_byte_identical():
    host = sc.dirty_region(9, n_keys=600).build(read_ts=sc.READ_TS)
    dev = DeviceRegion(host)
    L = ffi.lib()
    for per_block, restart, pl, sl, tl in ((37, 16, 1, 8, 5), (5, 1, 0, 0, 0), (1000, 3, 2, 8, 0), (16, 16, 1, 0, 5)):
        h, enc = C.c_void_p(), ffi.SstEncoded()
        assert L.b2_sst_encode(0, C.byref(dev._w[0]), per_block, restart, pl, ord("z"), sl, tl, C.byref(h), C.byref(enc)) == 0
        data = np.zeros(enc.data_len, dtype=np.uint8)
        offs = np.zeros(enc.n_blocks + 1, dtype=np.uint64)
        assert L.b2_copy_to_host(0, data.ctypes.data, enc.data, enc.data_len) == 0
        assert L.b2_copy_to_host(0, offs.ctypes.data, enc.block_offs, 8 * len(offs)) == 0
        exp, exp_offs = sstfmt.build(host.wblocks[0], restart_interval=restart, block_size=0, entries_per_block=per_block, key_prefix_len=pl, key_suffix_len=sl, trailer_len=tl)
        assert data.tobytes() == exp and [int(x) for x in offs] == exp_offs
        L.b2_sst_free(h)


PLANS = sc.plans()
SOME = [p for p in PLANS if p[0] in ("scan_all", "sel_lt_const", "count_col_sum_avg", "group_by_small")] + [(t[0], t[1]) for t in sc.topn_plans()[:1]]


@pytest.mark.parametrize("name,plan", SOME, ids=[n for n, _ in SOME])
def test_requests_over_decoded_blocks_match_oracle(name, plan):
    """the whole path: data blocks in host memory -> b2_sst_decode -> request over the decoded HBM-resident blocks"""
    host = sc.dirty_region(1, n_keys=900).build(read_ts=sc.READ_TS, n_write_blocks=2)
    runs = [sstfmt.build(b, block_size=2048) for b in host.wblocks]
    reg = SstRegion(host, runs)
    try:
        for ranges in (sc.WHOLE, sc.split_ranges()):
            exp = orc.dag_handle(plan, ranges, host)
            got = DagHandler(plan, ranges, reg).handle_request()
            assert_same_rows(got, exp, ordered=not sc.is_agg(name), ctx=name)
            assert got.stats.write_processed_keys == exp.stats["processed_keys"] and got.stats.processed_size == exp.stats["processed_size"]
        assert checksum(sc.WHOLE, reg) == checksum(sc.WHOLE, host)
        assert checksum(sc.WHOLE, reg)[:2] == orc.checksum(sc.WHOLE, host)[:2]
    finally:
        reg.close()


def test_bench_sized_round_trip():
    """3 M generated rows: encode on the device (32 KiB-ish blocks, restart interval 16, 'z' prefix, internal-key footer,
    trailer), decode from HOST memory, every byte and offset equal to the generator's block; checksum request equal."""
    L = ffi.lib()
    n = 3_000_000
    spec = ffi.GenSpec()
    spec.table_id, spec.first_handle, spec.n_rows, spec.n_cols, spec.row_format, spec.seed = 1000, 0, n, 2, 2, 0x525C682A2F7CE3DB
    spec.commit_ts, spec.newer_ts = 20, 1 << 62
    spec.extra_versions_per_million, spec.delete_per_million, spec.lock_rec_per_million = 30000, 20000, 10000
    g, blk = C.c_void_p(), ffi.GenBlock()
    assert L.b2_gen_create(0, C.byref(spec), C.byref(g), C.byref(blk)) == 0
    h, enc = C.c_void_p(), ffi.SstEncoded()
    assert L.b2_sst_encode(0, C.byref(blk.block), 600, 16, 1, ord("z"), 8, 5, C.byref(h), C.byref(enc)) == 0
    flat_bytes = blk.key_bytes + blk.val_bytes + 8 * (blk.block.n + 1)
    assert enc.data_len < 0.8 * flat_bytes  # prefix compression pays on table keys
    data = np.zeros(enc.data_len + 16, dtype=np.uint8)
    offs = np.zeros(enc.n_blocks + 1, dtype=np.uint64)
    assert L.b2_copy_to_host(0, data.ctypes.data, enc.data, enc.data_len) == 0
    assert L.b2_copy_to_host(0, offs.ctypes.data, enc.block_offs, 8 * len(offs)) == 0
    L.b2_sst_free(h)
    with SstDecoder() as dec:
        got, st = dec.decode(data, offs)
        assert got.n == blk.block.n and st.key_bytes == blk.key_bytes and st.val_bytes == blk.val_bytes

        def same(a, b, nbytes):
            ha, hb = np.zeros(nbytes, dtype=np.uint8), np.zeros(nbytes, dtype=np.uint8)
            assert L.b2_copy_to_host(0, ha.ctypes.data, a, nbytes) == 0 and L.b2_copy_to_host(0, hb.ctypes.data, b, nbytes) == 0
            return np.array_equal(ha, hb)
        assert same(got.keys, blk.block.keys, blk.key_bytes) and same(got.vals, blk.block.vals, blk.val_bytes)
        assert same(got.key_offs, blk.block.key_offs, 4 * (got.n + 1)) and same(got.val_offs, blk.block.val_offs, 4 * (got.n + 1))

        class Src:
            pass
        outs = []
        for b in (got, blk.block):
            s = Src()
            s.arr = (ffi.CfBlock * 1)(b)
            s.c = ffi.RegionSource()
            s.c.location, s.c.device, s.c.write, s.c.n_write = ffi.LOC_DEVICE, 0, s.arr, 1
            s.c.read_ts, s.c.isolation_level = (1 << 63) - 1, ffi.ISO_SI
            outs.append(checksum([kvfmt.table_range(1000)], s))
        assert outs[0] == outs[1] and outs[0][0] == 0 and outs[0][1][1] > 0.9 * n
    L.b2_gen_destroy(g)
