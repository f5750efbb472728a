"""Construction time of the automaton (host): the dense-table image (acb_build) and the sieve image (acb_sieve_build),
next to the CPU oracle's builder, at 4 244 (names.txt), 50 000 and 100 000 patterns.  SURVEY.md 8(f4)."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from ahocorasick_rs_b200 import _capi, workloads as W
from oracle import Oracle


def timed(pats, kind=0):
    L = _capi.lib()
    offs = np.zeros(len(pats) + 1, dtype=np.uint64)
    np.cumsum([len(p) for p in pats], out=offs[1:])
    blob = np.frombuffer(b"".join(pats), dtype=np.uint8)
    h = C.c_void_p()
    t0 = time.perf_counter()
    assert L.acb_build(blob.ctypes.data, offs.ctypes.data, len(pats), kind, -1, C.byref(h)) == 0
    t1 = time.perf_counter()
    n = L.acb_sieve_build(h, 180 * 1024, 0)
    t2 = time.perf_counter()
    states, image = L.acb_num_states(h), L.acb_image_bytes(h)
    L.acb_free(h)
    t3 = time.perf_counter()
    Oracle(pats, kind)
    t4 = time.perf_counter()
    return t1 - t0, t2 - t1, t4 - t3, states, image, n


for name, pats in [("names.txt (4 244)", [p.encode() for p in W.patterns_long()]),
                   ("config 3 (10 000 tokens)", W.config3(10_000, 10)[0]),
                   ("config 5 (50 000 x 5-12 letters)", W.random_lowercase_patterns(50_000, 5, 12, 5)),
                   ("config 4 (100 000 x 5-8 letters)", W.random_lowercase_patterns(100_000, 5, 8, 4))]:
    d, s, o, states, image, sieve = timed(pats)
    print(f"{name}: dense table {d * 1e3:.0f} ms ({states} states, {image / 1e6:.1f} MB image) | sieve image {s * 1e3:.0f} ms "
          f"({sieve / 1e6:.1f} MB) | oracle builder {o * 1e3:.0f} ms | host threads {os.cpu_count()}", flush=True)
