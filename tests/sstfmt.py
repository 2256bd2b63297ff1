"""ctypes access to the oracle's RocksDB data-block builder / iterator (oracle/orc_sst.h).  TEST INFRASTRUCTURE."""
import ctypes as C

import numpy as np

import orc
from tikv_b200 import ffi

DEFAULT = dict(restart_interval=16, block_size=32 * 1024, entries_per_block=0, key_prefix_len=1, key_prefix_byte=ord("z"), key_suffix_len=8, trailer_len=5)


def _lib():
    L = orc.lib()
    if not getattr(L, "_sst_ready", False):
        L.orc_sst_build.argtypes = [C.POINTER(ffi.CfBlock), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint8, C.c_uint32, C.c_uint32]
        L.orc_sst_build.restype = C.c_void_p
        L.orc_sst_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p)]
        L.orc_sst_data.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.POINTER(C.c_uint64)), C.POINTER(C.c_uint32)]
        L.orc_sst_data.restype = C.c_void_p
        L.orc_sst_flat.argtypes = [C.c_void_p, C.POINTER(ffi.CfBlock)]
        L.orc_sst_free.argtypes = [C.c_void_p]
        L._sst_ready = True
    return L


def build(host_block, **opts):
    """kvfmt.HostBlock -> (bytes of the data blocks, [n_blocks + 1 offsets])."""
    o = dict(DEFAULT, **opts)
    L = _lib()
    h = L.orc_sst_build(C.byref(host_block.c), o["restart_interval"], o["block_size"], o["entries_per_block"], o["key_prefix_len"], o["key_prefix_byte"],
                        o["key_suffix_len"], o["trailer_len"])
    ln, offs, nb = C.c_uint64(), C.POINTER(C.c_uint64)(), C.c_uint32()
    p = L.orc_sst_data(h, C.byref(ln), C.byref(offs), C.byref(nb))
    data = C.string_at(p, ln.value) if ln.value else b""
    out = [offs[i] for i in range(nb.value + 1)]
    L.orc_sst_free(h)
    return data, out


def decode(data, offs, trailer_len=5, key_prefix_len=1, key_suffix_len=8):
    """(status, [(key, value)]) by the oracle's block iterator: 0 ok, 1 corrupted, 2 unsupported."""
    L = _lib()
    buf = np.frombuffer(bytes(data) + b"\0" * 8, dtype=np.uint8)
    o = np.asarray(offs, dtype=np.uint64)
    h = C.c_void_p()
    rc = L.orc_sst_decode(buf.ctypes.data, o.ctypes.data, len(offs) - 1, trailer_len, key_prefix_len, key_suffix_len, C.byref(h))
    blk = ffi.CfBlock()
    L.orc_sst_flat(h, C.byref(blk))
    kvs = flat_kvs(blk) if rc == 0 else []
    L.orc_sst_free(h)
    return rc, kvs


def flat_kvs(blk):
    """host-resident b2_cf_block -> [(key, value)]"""
    n = blk.n
    if n == 0:
        return []
    ko = np.ctypeslib.as_array(C.cast(blk.key_offs, C.POINTER(C.c_uint32)), shape=(n + 1,))
    vo = np.ctypeslib.as_array(C.cast(blk.val_offs, C.POINTER(C.c_uint32)), shape=(n + 1,))
    keys = C.string_at(blk.keys, int(ko[n]))
    vals = C.string_at(blk.vals, int(vo[n])) if vo[n] else b""
    return [(keys[ko[i]:ko[i + 1]], vals[vo[i]:vo[i + 1]]) for i in range(n)]
