"""A few small inputs through the device BGZF coder and one small generator run, for compute-sanitizer:
    compute-sanitizer --tool racecheck python tools/bgzf_sanity.py"""
import ctypes as C
import gzip
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jackalope_b200 as J

ctx = J.Context(0)
rng = np.random.default_rng(1)
cases = [
    bytes(rng.choice(np.frombuffer(b"TCAG\nGJC@-0123", np.uint8), 2 * 0xff00 + 777)),
    rng.integers(0, 256, 70000, dtype=np.uint8).tobytes(),
    b"".join(bytes([65 + k]) * f for k, f in enumerate([1, 2, 3, 5, 8, 13, 21, 34, 55, 89, 144, 233, 377, 610, 987, 1597, 2584, 4181, 6765, 10946, 17711])),
    b"G" * 1000,
]
for data in cases:
    n = C.c_uint64()
    cap = len(data) + 1024
    buf = C.create_string_buffer(cap)
    assert ctx.lib.jlp_bgzf_device(ctx.h, 6, data, len(data), buf, cap, C.byref(n)) == 0
    assert gzip.decompress(buf.raw[:n.value]) == data
g = J.random_genome(2, 20000, seed=3)
r1, r2, _ = J.illumina(g, "", 4000, 150, True, seed=5, ctx=ctx, sink="memory")
z1, z2, _ = J.illumina(g, "", 4000, 150, True, seed=5, ctx=ctx, sink="memory", compress=True, comp_engine="device", batch_pairs=700)
assert gzip.decompress(z1) == r1 and gzip.decompress(z2) == r2
print("ok")
