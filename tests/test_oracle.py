"""CPU tests: the oracle (oracle/ac_oracle.c) against the reference's golden
vectors and against the brute-force statement of the semantics."""
import json
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import Oracle
from tests.spec_bruteforce import spec_find

HERE = os.path.dirname(os.path.abspath(__file__))
KINDS = ["Standard", "LeftmostFirst", "LeftmostLongest"]

with open(os.path.join(HERE, "golden", "reference_vectors.json"), encoding="utf-8") as f:
    VECTORS = json.load(f)["vectors"]


@pytest.mark.parametrize("use_dfa", [True, False])
@pytest.mark.parametrize("vec", VECTORS, ids=[f"{i}:{v['src']}" for i, v in enumerate(VECTORS)])
def test_oracle_reproduces_reference_vectors(vec, use_dfa):
    pats = [p.encode("utf-8") for p in vec["patterns"]]
    orc = Oracle(pats, vec["kind"])
    hay = vec["haystack"]
    if vec.get("error"):
        with pytest.raises(ValueError):
            orc.find(hay.encode(), overlapping=vec["overlapping"], use_dfa=use_dfa)
        return
    if vec["cls"] == "str":
        got = orc.find_str(hay, vec["overlapping"], use_dfa)
        if "expect_strings" in vec:
            assert [hay[s:e] for (_, s, e) in got] == vec["expect_strings"]
            assert [vec["patterns"][p] for (p, _, _) in got] == vec["expect_strings"]
    else:
        raw = hay.encode()
        got = orc.find(raw, vec["overlapping"], use_dfa)
        if "expect_strings" in vec:
            assert [raw[s:e].decode() for (_, s, e) in got] == vec["expect_strings"]
    if "expect_indexes" in vec:
        assert [list(m) for m in got] == vec["expect_indexes"]


small_alpha = st.binary(min_size=1, max_size=5).map(lambda b: bytes(97 + (x % 2) for x in b))
abc_alpha = st.binary(min_size=1, max_size=6).map(lambda b: bytes(97 + (x % 3) for x in b))


@settings(max_examples=400, deadline=None)
@given(st.lists(small_alpha, min_size=1, max_size=8),
       st.binary(max_size=40).map(lambda b: bytes(97 + (x % 2) for x in b)),
       st.sampled_from(KINDS), st.booleans())
def test_oracle_matches_bruteforce_ab(patterns, haystack, kind, use_dfa):
    orc = Oracle(patterns, kind)
    assert orc.find(haystack, False, use_dfa) == spec_find(patterns, haystack, kind, False)
    if kind == "Standard":
        assert orc.find(haystack, True, use_dfa) == spec_find(patterns, haystack, kind, True)


@settings(max_examples=300, deadline=None)
@given(st.lists(abc_alpha, min_size=1, max_size=12),
       st.binary(max_size=60).map(lambda b: bytes(97 + (x % 3) for x in b)),
       st.sampled_from(KINDS), st.booleans())
def test_oracle_matches_bruteforce_abc(patterns, haystack, kind, use_dfa):
    orc = Oracle(patterns, kind)
    assert orc.find(haystack, False, use_dfa) == spec_find(patterns, haystack, kind, False)
    if kind == "Standard":
        assert orc.find(haystack, True, use_dfa) == spec_find(patterns, haystack, kind, True)


@settings(max_examples=150, deadline=None)
@given(st.lists(st.binary(min_size=1, max_size=4), min_size=1, max_size=10), st.binary(max_size=64),
       st.sampled_from(KINDS))
def test_oracle_matches_bruteforce_binary(patterns, haystack, kind):
    orc = Oracle(patterns, kind)
    assert orc.find(haystack, False) == spec_find(patterns, haystack, kind, False)
    assert orc.find(haystack, False, use_dfa=False) == spec_find(patterns, haystack, kind, False)


def test_duplicates_and_restart_rules():
    # duplicate patterns: lower id wins when not overlapping, both reported (ascending) when overlapping
    orc = Oracle([b"ab", b"ab", b"b"], "Standard")
    assert orc.find(b"xab") == [(0, 1, 3)]
    assert orc.find(b"xab", overlapping=True) == [(0, 1, 3), (1, 1, 3), (2, 2, 3)]
    # restart from the start state at the end of each match
    assert Oracle([b"aa"], "Standard").find(b"aaaaa") == [(0, 0, 2), (0, 2, 4)]
    assert Oracle([b"aa"], "Standard").find(b"aaaaa", overlapping=True) == [(0, i, i + 2) for i in range(4)]
    assert Oracle([b"aba"], "LeftmostLongest").find(b"ababa") == [(0, 0, 3)]
    for kind in ("LeftmostFirst", "LeftmostLongest"):
        with pytest.raises(ValueError):
            Oracle([b"a"], kind).find(b"", overlapping=True)   # refused before any byte is read
    with pytest.raises(ValueError):
        Oracle([b"a", b""])


def test_code_point_mapping():
    orc = Oracle(["há".encode(), "l\U0001F926l".encode()], "Standard")
    hay = "☃☃ há l\U0001F926l"
    assert [hay[s:e] for (_, s, e) in orc.find_str(hay)] == ["há", "l\U0001F926l"]


def test_batch_driver_equals_per_haystack_calls():
    rng = np.random.default_rng(0)
    pats = [bytes(rng.integers(97, 100, size=rng.integers(1, 5)).astype(np.uint8)) for _ in range(20)]
    hays = [bytes(rng.integers(97, 100, size=rng.integers(0, 50)).astype(np.uint8)) for _ in range(64)]
    data = np.frombuffer(b"".join(hays), dtype=np.uint8)
    offs = np.zeros(len(hays) + 1, dtype=np.int64)
    np.cumsum([len(h) for h in hays], out=offs[1:])
    for kind in KINDS:
        orc = Oracle(pats, kind)
        total, counts, rec = orc.scan_batch(data, offs)
        exp = [(h, p, s, e) for h, hay in enumerate(hays) for (p, s, e) in orc.find(hay)]
        assert total == len(exp)
        assert [tuple(int(x) for x in r) for r in rec] == exp
        assert counts.sum() == total
        t2, c2, _ = orc.scan_batch(data, offs, nthreads=4, want_records=False)
        assert t2 == total and (c2 == counts).all()
