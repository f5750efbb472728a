"""CPU tests of the PRODUCT's host side: the automaton builder / device image
(csrc/automaton.cpp) and the control flow of both kernels, executed by the
Python image interpreter (tests/image_interp.py) and compared with the oracle.
No GPU needed; the GPU parity tests proper are in test_gpu_parity.py."""
import json
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import Oracle
from tests import image_interp as ii
from tests.spec_bruteforce import spec_find

HERE = os.path.dirname(os.path.abspath(__file__))
KINDS = ["Standard", "LeftmostFirst", "LeftmostLongest"]
KID = {"Standard": 0, "LeftmostFirst": 1, "LeftmostLongest": 2}

with open(os.path.join(HERE, "golden", "reference_vectors.json"), encoding="utf-8") as f:
    VECTORS = [v for v in json.load(f)["vectors"] if not v.get("error")]


@pytest.mark.parametrize("vec", VECTORS, ids=[f"{i}:{v['src']}" for i, v in enumerate(VECTORS)])
def test_image_reproduces_reference_vectors(vec):
    pats = [p.encode("utf-8") for p in vec["patterns"]]
    im = ii.Image(pats, KID[vec["kind"]])
    hay = vec["haystack"]
    raw = hay.encode("utf-8")
    cp = vec["cls"] == "str"
    got = ii.find(im, raw, vec["overlapping"], cp)
    mode = 2 if vec["overlapping"] else (0 if vec["kind"] == "Standard" else 1)
    for H in (2, 3, 7, 40):
        for base in (0, 5, 64 - 3):
            assert ii.staged_lane(im, raw, mode, H, base_addr=base, cp=cp) == got
    if "expect_strings" in vec:
        if cp:
            assert [hay[s:e] for (_, s, e) in got] == vec["expect_strings"]
        else:
            assert [raw[s:e].decode() for (_, s, e) in got] == vec["expect_strings"]
    if "expect_indexes" in vec:
        assert [list(m) for m in got] == vec["expect_indexes"]


def alpha(k, max_size):
    return st.binary(min_size=1, max_size=max_size).map(lambda b: bytes(97 + (x % k) for x in b))


@settings(max_examples=300, deadline=None)
@given(st.lists(alpha(2, 5), min_size=1, max_size=8),
       st.binary(max_size=120).map(lambda b: bytes(97 + (x % 2) for x in b)),
       st.sampled_from(KINDS), st.integers(2, 30), st.integers(0, 63))
def test_kernel_control_flow_matches_oracle_ab(patterns, haystack, kind, H, base):
    orc = Oracle(patterns, kind)
    im = ii.Image(patterns, KID[kind])
    exp = orc.find(haystack)
    assert ii.find(im, haystack) == exp
    mode = 0 if kind == "Standard" else 1
    assert ii.staged_lane(im, haystack, mode, H, base_addr=base) == exp
    if kind == "Standard":
        expo = orc.find(haystack, overlapping=True)
        assert ii.find(im, haystack, overlapping=True) == expo
        assert ii.staged_lane(im, haystack, 2, H, base_addr=base) == expo


@settings(max_examples=200, deadline=None)
@given(st.lists(alpha(3, 6), min_size=1, max_size=12),
       st.binary(max_size=200).map(lambda b: bytes(97 + (x % 3) for x in b)),
       st.sampled_from(KINDS), st.integers(2, 60), st.integers(0, 63))
def test_kernel_control_flow_matches_oracle_abc(patterns, haystack, kind, H, base):
    orc = Oracle(patterns, kind)
    im = ii.Image(patterns, KID[kind])
    exp = orc.find(haystack)
    mode = 0 if kind == "Standard" else 1
    assert ii.staged_lane(im, haystack, mode, H, base_addr=base) == exp
    assert exp == spec_find(patterns, haystack, kind)


@settings(max_examples=150, deadline=None)
@given(st.lists(st.binary(min_size=1, max_size=4), min_size=1, max_size=10), st.binary(max_size=100),
       st.sampled_from(KINDS), st.integers(2, 40))
def test_image_binary_patterns(patterns, haystack, kind, H):
    orc = Oracle(patterns, kind)
    im = ii.Image(patterns, KID[kind])
    exp = orc.find(haystack)
    assert ii.find(im, haystack) == exp
    assert ii.staged_lane(im, haystack, 0 if kind == "Standard" else 1, H) == exp


@settings(max_examples=150, deadline=None)
@given(st.lists(st.text(alphabet="abé☃\U0001F926", min_size=1, max_size=3), min_size=1, max_size=6),
       st.text(alphabet="abé☃\U0001F926 ", max_size=60), st.sampled_from(KINDS), st.integers(2, 30),
       st.integers(0, 15))
def test_code_points(patterns, haystack, kind, H, base):
    pats = [p.encode() for p in patterns]
    orc = Oracle(pats, kind)
    im = ii.Image(pats, KID[kind])
    exp = orc.find_str(haystack)
    raw = haystack.encode()
    assert ii.find(im, raw, cp=True) == exp
    assert ii.staged_lane(im, raw, 0 if kind == "Standard" else 1, H, base_addr=base, cp=True) == exp
    if kind == "Standard":
        expo = orc.find_str(haystack, overlapping=True)
        assert ii.staged_lane(im, raw, 2, H, base_addr=base, cp=True) == expo
        assert ii.find_chunked(im, raw, 64, cp=True) == expo
        assert ii.find_chunked(im, raw, 64, H=H, cp=True, base_addr=base) == expo


@settings(max_examples=100, deadline=None)
@given(st.lists(alpha(2, 6), min_size=1, max_size=8),
       st.binary(min_size=100, max_size=400).map(lambda b: bytes(97 + (x % 2) for x in b)),
       st.sampled_from([64, 80, 128]), st.integers(2, 30), st.integers(0, 63))
def test_chunked_overlapping_equals_serial(patterns, haystack, chunk, H, base):
    orc = Oracle(patterns, "Standard")
    im = ii.Image(patterns, 0)
    exp = orc.find(haystack, overlapping=True)
    assert ii.find_chunked(im, haystack, chunk) == exp
    assert ii.find_chunked(im, haystack, chunk, H=H, base_addr=base) == exp


def test_names_automaton_shape_and_parity():
    pats = [l.strip().encode() for l in open(os.path.join(HERE, "golden", "patterns_long.txt"))]
    im = ii.Image(pats, 0, 2)
    assert im.n_states == 11163 + 1          # 11 163 trie states (+ the dead state)
    assert im.col_mode == 0 and im.n_cols == 27 and im.col_lo == ord("a")
    line = ("no one who had ever seen charlotte in her infancy would have supposed her born to be an heroine. "
            "her name was whatevs—and isabella had never been handsome 12345.").encode()
    orc = Oracle(pats, "Standard")
    exp = orc.find(line)
    assert exp and ii.find(im, line) == exp
    for H in (100, 1500, 4000):
        assert ii.staged_lane(im, line, 0, H, base_addr=7) == exp
    for kind in ("LeftmostFirst", "LeftmostLongest"):
        im2 = ii.Image(pats, KID[kind])
        assert ii.staged_lane(im2, line, 1, 1500) == Oracle(pats, kind).find(line)


def test_profiled_hot_set_keeps_results_and_cuts_traps():
    import struct
    from ahocorasick_rs_b200 import workloads as W
    pats, data, offs = W.config2(3)
    bp = [p.encode() for p in pats]
    im = ii.Image(bp, 0, 2)
    hay = bytes(data[offs[1]:offs[2]])
    # visit counts as acb_profile would produce them from haystack 0
    visits = np.zeros(im.n_states, dtype=np.uint32)
    s = 1
    for b in bytes(data[offs[0]:offs[1]]):
        e = int(im.trans[s, im.col(b)])
        s = e & ii.MASK
        if e & ii.FLAG:
            s = 1
        visits[s] += 1
    exp = Oracle(bp, "Standard").find_str(hay.decode())
    st_bfs, st_prof = {}, {}
    assert ii.staged_lane(im, hay, 0, 256, cp=True, stats=st_bfs) == exp
    assert ii.staged_lane(im, hay, 0, 256, cp=True, stats=st_prof, visits=visits) == exp
    assert st_prof.get("traps", 0) * 20 < st_bfs["traps"]      # 256 profiled rows beat 256 shallowest rows by far
    table, h2f, f2h, rows = im.hot_image(visits, 300)
    assert rows == 300 and h2f[0] == 1 and f2h[1] == 0 and f2h[0] == 0xFFFF
    assert len(set(h2f[:rows].tolist())) == rows and all(f2h[h2f[i]] == i for i in range(rows))


def test_builder_errors():
    with pytest.raises(ValueError):
        ii.Image([b"a", b""])
    im = ii.Image([b"ab", b"ab", b"b"], 0)
    assert ii.find(im, b"xab") == [(0, 1, 3)]
    assert ii.find(im, b"xab", overlapping=True) == [(0, 1, 3), (1, 1, 3), (2, 2, 3)]
