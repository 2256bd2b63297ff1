"""GPU probe: b2_sst_decode of one generated C3 region from pinned host memory (timings per call; run under ncu for per-kernel times)."""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench
from tikv_b200 import ffi
from tikv_b200.executor import SstDecoder

L = ffi.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
gens, blks = bench.gen_blocks(ffi, 0, "c3", n, 1)
b = blks[0]
h, enc = C.c_void_p(), ffi.SstEncoded()
assert L.b2_sst_encode(0, C.byref(b.block), 630, 16, 1, ord("z"), 8, 5, C.byref(h), C.byref(enc)) == 0
p = L.b2_host_alloc_pinned_near(0, enc.data_len + 64)
offs = np.zeros(enc.n_blocks + 1, dtype=np.uint64)
assert L.b2_copy_to_host(0, p, enc.data, enc.data_len) == 0 and L.b2_copy_to_host(0, offs.ctypes.data, enc.block_offs, 8 * len(offs)) == 0
flat = b.key_bytes + b.val_bytes + 8 * b.block.n
print(f"entries {b.block.n} flat {flat/1e6:.1f} MB encoded {enc.data_len/1e6:.1f} MB ratio {enc.data_len/flat:.3f} blocks {enc.n_blocks}")
with SstDecoder(0) as d:
    for i in range(4):
        t0 = time.perf_counter()
        blk, st = d.decode(p, offs)
        dt = (time.perf_counter() - t0) * 1e3
        print(f"decode call {dt:.2f} ms wall, device passes {st.decode_ms:.2f} ms, h2d {st.h2d_bytes/1e6:.1f} MB -> {st.h2d_bytes/dt/1e6:.1f} GB/s incl. expansion; "
              f"expansion {flat/st.decode_ms/1e6:.0f} GB/s of flat bytes")
    # device-resident image: the passes alone
    blk, st = d.decode(enc.data, offs, location=ffi.LOC_DEVICE)
    print(f"device-resident image: passes {st.decode_ms:.2f} ms")
