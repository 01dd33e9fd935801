"""Randomised pins of the oracle restatements (hypothesis, derandomised and bounded so the CPU suite
stays fast and repeatable):
the three statements of DBSCAN agree on arbitrary small clouds, the AEC restatement follows the
reference's own classes on arbitrary parameters, the EVT 3.0 codec round-trips arbitrary ordered
streams and its decoder accepts arbitrary words."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import dbscan_cases as D
from oracle import aec, dbscan

SET = dict(deadline=None, derandomize=True, suppress_health_check=[HealthCheck.too_slow,
                                                 HealthCheck.function_scoped_fixture])


@settings(max_examples=120, **SET)
@given(seed=st.integers(0, 2**31), n=st.integers(1, 260), span=st.integers(4, 60),
       eps=st.sampled_from([1.0, 1.5, 2.0, 3.0, 5.0]), min_pts=st.integers(1, 7),
       mn=st.integers(1, 6), dims=st.sampled_from([2, 3]))
def test_dbscan_three_statements_agree(orc, seed, n, span, eps, min_pts, mn, dims):
    r = np.random.default_rng(seed)
    pts = r.integers(0, span, size=(n, dims)).astype(np.float32)   # small lattice: many ties
    lo, co, so = dbscan.oracle(pts, eps, min_pts, mn, 10**6)
    lc, cc, sc = dbscan.contract(pts, eps, min_pts, mn, 10**6)
    assert D.canon(co) == D.canon(cc) and (lo == lc).all() and (so == sc).all()
    if dbscan.ref_available():
        assert D.canon(dbscan.reference(pts, eps, min_pts, mn, 10**6)) == D.canon(co)


@pytest.mark.skipif(not aec.ref_available(), reason="oracle/_ref/libref_aec.so not built")
@settings(max_examples=80, **SET)
@given(seed=st.integers(0, 2**31), n=st.integers(1, 700), sz=st.integers(1, 120),
       radius=st.sampled_from([2.0, 5.0, 12.5, 40.0]), kappa=st.integers(0, 12),
       alpha=st.sampled_from([0.1, 0.3, 0.5, 0.9, 1.0]), min_n=st.integers(0, 8),
       tie=st.integers(1, 40), rand_seed=st.integers(0, 2**31))
def test_aec_restatement_follows_the_reference(orc, seed, n, sz, radius, kappa, alpha, min_n, tie,
                                               rand_seed):
    r = np.random.default_rng(seed)
    e = np.zeros((n, 4))
    e[:, 0] = 3.0 + np.floor(np.arange(n) / tie) * 1e-3
    e[:, 1:3] = r.integers(0, 90, size=(n, 2))
    e[:, 3] = r.integers(0, 2, n)
    init = dict(sz_buffer=sz, radius=radius, kappa=kappa, alpha=alpha, min_n=min_n)
    o, ref = aec.Oracle(init, rand_seed=rand_seed), aec.Reference(init, rand_seed=rand_seed)
    half = n // 2
    for part in (e[:half], e[half:]):
        o.update(part)
        ref.update(part)
        a, b = o.state(), ref.state()
        assert (a["ids"] == b["ids"]).all() and (a["n"] == b["n"]).all()
        assert (a["mu"] == b["mu"]).all() and a["last"] == b["last"]
        for p, q in zip(a["pts"], b["pts"]):
            for x, y in zip(p, q):
                assert (x == y).all()


@settings(max_examples=120, **SET)
@given(seed=st.integers(0, 2**31), n=st.integers(0, 3000), rows=st.integers(1, 40),
       cols=st.integers(1, 300), dt_max=st.integers(0, 5000))
def test_evt3_round_trip_on_arbitrary_ordered_streams(orc, seed, n, rows, cols, dt_max):
    r = np.random.default_rng(seed)
    ev = np.zeros(n, orc.EVENT_DTYPE)
    ev["t"] = np.cumsum(r.integers(0, dt_max + 1, n)) if n else 0
    ev["y"], ev["x"], ev["p"] = r.integers(0, rows, n), r.integers(0, cols, n), r.integers(0, 2, n)
    o = np.lexsort((ev["x"], ev["p"], ev["y"], ev["t"]))             # ordered inside a timestamp
    ev = ev[o]
    w = orc.evt3_encode(ev)
    assert orc.evt3_decode(w).tobytes() == ev.tobytes()


@settings(max_examples=80, **SET)
@given(seed=st.integers(0, 2**31), n=st.integers(0, 4000))
def test_evt3_decoder_takes_any_words(orc, seed, n):
    """every 16-bit word sequence decodes (no crash, counts consistent with the vector masks)"""
    r = np.random.default_rng(seed)
    w = r.integers(0, 65536, n).astype(np.uint16)
    d = orc.evt3_decode(w)
    ty, v = w >> 12, w & 0xFFF
    want = int((ty == 2).sum()) + int(sum(bin(int(x)).count("1") for x in v[ty == 4])) \
        + int(sum(bin(int(x) & 0xFF).count("1") for x in v[ty == 5]))
    assert len(d) == want
    assert (np.diff(d["t"]) >= 0).all() or (ty == 8).sum() > 0 or (ty == 6).sum() > 0
