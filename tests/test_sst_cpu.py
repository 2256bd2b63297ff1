"""`-m "not gpu"`: the oracle's RocksDB data-block builder / iterator (oracle/orc_sst.h) against a hand-assembled block
that follows the published layout byte by byte, and against each other (round trips, rejected inputs).  RocksDB is not in
the reference tree, so this file is what pins the block format for the device decoder's tests (tests/test_gpu_sst.py)."""
import random
import struct

import pytest

import kvfmt
import scenarios as sc
import sstfmt

FOOT = b"\x01" + b"\0" * 7  # fixed64 LE of (sequence 0 << 8 | kTypeValue)


def hand_block():
    """Three TiKV data keys, restart interval 2, stored form (5-byte trailer)."""
    e0 = bytes([0, 6 + 8, 2]) + b"zapple" + FOOT + b"v0"            # restart: shared 0
    e1 = bytes([5, 1 + 8, 0]) + b"y" + FOOT                          # "zappl" shared with the previous key, empty value
    e2 = bytes([0, 7 + 8, 3]) + b"zbanana" + FOOT + b"abc"           # second restart point
    ents = e0 + e1 + e2
    restarts = struct.pack("<II", 0, len(e0) + len(e1))
    return ents + restarts + struct.pack("<I", 2) + b"\0" * 5


def test_builder_and_iterator_match_the_published_layout():
    blk = kvfmt.HostBlock([(b"apple", b"v0"), (b"apply", b""), (b"banana", b"abc")])
    data, offs = sstfmt.build(blk, restart_interval=2)
    assert data == hand_block() and offs == [0, len(data)]
    rc, kvs = sstfmt.decode(hand_block(), [0, len(hand_block())])
    assert rc == 0 and kvs == [(b"apple", b"v0"), (b"apply", b""), (b"banana", b"abc")]
    # contents-only form without prefix / footer: the same entries with plain user keys
    data2, offs2 = sstfmt.build(blk, restart_interval=2, key_prefix_len=0, key_suffix_len=0, trailer_len=0)
    exp = bytes([0, 5, 2]) + b"applev0" + bytes([4, 1, 0]) + b"y" + bytes([0, 6, 3]) + b"bananaabc" + struct.pack("<III", 0, 14, 2)
    assert data2 == exp
    assert sstfmt.decode(data2, offs2, trailer_len=0, key_prefix_len=0, key_suffix_len=0) == (0, blk.kvs)


def test_a_key_that_extends_its_predecessor_shares_into_the_footer():
    """"ab" then "ab\\x01": the full keys are z ab 01 00.. and z ab 01 01 00..: four bytes are shared, one of them the
    first footer byte of the predecessor."""
    blk = kvfmt.HostBlock([(b"ab", b"1"), (b"ab\x01", b"2")])
    data, offs = sstfmt.build(blk)
    assert data[:3 + 11 + 1] == bytes([0, 11, 1]) + b"zab" + FOOT + b"1"
    assert data[15:18] == bytes([4, 8, 1])
    assert sstfmt.decode(data, offs) == (0, blk.kvs)


@pytest.mark.parametrize("opts", [dict(), dict(restart_interval=1), dict(restart_interval=3, block_size=700), dict(block_size=0, entries_per_block=37),
                                  dict(key_prefix_len=0, key_suffix_len=0, trailer_len=0, block_size=4096), dict(key_prefix_len=3, key_prefix_byte=0x7a, trailer_len=0)])
def test_round_trip_of_generated_regions(opts):
    host = sc.dirty_region(5, n_keys=700).build(read_ts=sc.READ_TS)
    for blk in host.wblocks + ([host.dblock] if host.dblock else []):
        data, offs = sstfmt.build(blk, **opts)
        fmt = {k: opts[k] for k in ("trailer_len", "key_prefix_len", "key_suffix_len") if k in opts}
        rc, kvs = sstfmt.decode(data, offs, **fmt)
        assert rc == 0 and kvs == blk.kvs
        assert offs[0] == 0 and offs[-1] == len(data) and offs == sorted(offs)
        if opts.get("entries_per_block"):
            assert len(offs) - 1 == -(-blk.n // 37)
        elif opts.get("block_size", 32768) and blk.n:
            assert max(b - a for a, b in zip(offs, offs[1:])) < opts.get("block_size", 32768) + 1400  # one entry past the threshold at most


def test_empty_input_and_rejected_blocks():
    assert sstfmt.build(kvfmt.HostBlock([])) == (b"", [0])
    assert sstfmt.decode(b"", [0]) == (0, [])
    good = hand_block()
    rng = random.Random(3)
    # a Delete tombstone (value type 0) needs RocksDB's merging iterator: unsupported, not corrupted
    pos = good.index(FOOT)
    assert sstfmt.decode(good[:pos] + b"\0" + good[pos + 1:], [0, len(good)])[0] == 2
    # data-block hash index flag in the footer
    assert sstfmt.decode(good[:-6] + b"\x80" + good[-5:], [0, len(good)])[0] == 2
    # truncated / damaged blocks never read outside the slice
    assert sstfmt.decode(good[:7], [0, 7])[0] == 1
    assert sstfmt.decode(good, [0, 3])[0] == 1
    for _ in range(300):
        bad = bytearray(good)
        bad[rng.randrange(len(bad) - 5)] ^= 1 << rng.randrange(8)
        rc, kvs = sstfmt.decode(bytes(bad), [0, len(bad)])
        assert rc in (0, 1, 2)


def py_decode_block(block, trailer_len=5, key_prefix_len=1, key_suffix_len=8):
    """A third reading of the layout, in plain Python (no restart array needed for a forward walk)."""
    def var32(p):
        v = s = 0
        while True:
            b = block[p]; p += 1
            v |= (b & 0x7F) << s; s += 7
            if not b & 0x80:
                return v, p
    end = len(block) - trailer_len
    nr = struct.unpack_from("<I", block, end - 4)[0]
    lim = end - 4 - 4 * nr
    restarts = struct.unpack_from("<%dI" % nr, block, lim)
    p, key, out, starts = 0, b"", [], []
    while p < lim:
        starts.append(p)
        sh, p = var32(p); ns, p = var32(p); vl, p = var32(p)
        key = key[:sh] + block[p:p + ns]; p += ns
        assert key[len(key) - key_suffix_len:] == (b"\x01" + b"\0" * 7)[:key_suffix_len] or not key_suffix_len
        out.append((key[key_prefix_len:len(key) - key_suffix_len], block[p:p + vl])); p += vl
    assert p == lim and all(r in starts for r in restarts) and restarts[0] == 0
    return out


def test_builder_output_read_by_an_independent_python_walker():
    host = sc.dirty_region(11, n_keys=500).build(read_ts=sc.READ_TS)
    for opts in (dict(), dict(restart_interval=4, block_size=1500), dict(block_size=0, entries_per_block=50)):
        blk = host.wblocks[0]
        data, offs = sstfmt.build(blk, **opts)
        kvs = []
        for a, b in zip(offs, offs[1:]):
            kvs += py_decode_block(data[a:b])
        assert kvs == blk.kvs
        ri = opts.get("restart_interval", 16)
        for a, b in zip(offs, offs[1:]):  # a restart point every `restart_interval` entries
            nr = struct.unpack_from("<I", data, b - 5 - 4)[0]
            assert nr == -(-len(py_decode_block(data[a:b])) // ri)
