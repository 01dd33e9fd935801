"""GPU parity of the RAW EVT 3.0 ingest (evk_load_evt3 / evk_load_raw, csrc/evk_evt3.cu): the decode
on the device is byte-identical to the oracle's sequential decoder (oracle/evk_oracle.c
orc_evt3_decode) on encoded streams, on streams whose state crosses block boundaries, and on
arbitrary random words."""
import os

import numpy as np
import pytest

import evk_loader

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def evk():
    m = evk_loader.load()
    m.lib()
    return m


def check(evk, h, orc, words):
    want = orc.evt3_decode(words)
    assert h.load_evt3(words) == len(want)
    got = h.get_events()
    assert got.tobytes() == want.tobytes()
    return want


def test_evt3_decode_matches_oracle(evk, orc):
    with evk.Evk(4_000_000) as h:
        # sparse Gen4 stream (about two words per event), sizes around the 4096-word block
        ev = orc.synth(orc.synth_params(0xE7CA0003, 1_000_000, 1280, 720, 100_000_000, 64))
        ev["p"] = ev["p"] > 0
        w = orc.evt3_encode(ev)
        for n in (0, 1, 15, 16, 17, 4095, 4096, 4097, 8192, 100_003, len(w)):
            got = check(evk, h, orc, w[:n])
        assert got.tobytes() == ev.tobytes()
        # dense rows: vector words, continued bases
        r = np.random.default_rng(2)
        n = 600_000
        t = np.sort(r.integers(0, 5000, n)).astype(np.int64)
        y, x, p = r.integers(0, 16, n), r.integers(0, 400, n), r.integers(0, 2, n)
        o = np.lexsort((x, p, y, t))
        dense = np.zeros(n, orc.EVENT_DTYPE)
        dense["t"], dense["y"], dense["x"], dense["p"] = t[o], y[o], x[o], p[o]
        assert check(evk, h, orc, orc.evt3_encode(dense)).tobytes() == dense.tobytes()
        # > 2^24 us: wraps counted across blocks
        long_ev = np.zeros(300_000, orc.EVENT_DTYPE)
        long_ev["t"] = np.arange(300_000, dtype=np.int64) * 400          # 120 s
        long_ev["x"], long_ev["y"] = np.arange(300_000) % 1280, np.arange(300_000) % 720
        assert check(evk, h, orc, orc.evt3_encode(long_ev)).tobytes() == long_ev.tobytes()


def test_evt3_arbitrary_words(evk, orc):
    """any word sequence decodes like the sequential decoder: random words with every type, long
    stretches without time / row / base words (state carried over many blocks), vector bases
    advancing past 2047"""
    r = np.random.default_rng(3)
    with evk.Evk(3_000_000) as h:
        for trial in range(4):
            n = 200_000 + 977 * trial
            kinds = np.array([0x0, 0x2, 0x3, 0x4, 0x5, 0x6, 0x8, 0xA, 0xE, 0x7, 0xF, 0x1, 0x9])
            prob = np.array([4, 20, 4, 6, 6, 4, 2, 1, 1, 1, 1, 1, 1], float)
            if trial == 1:       # almost no state words: the entry state of a block comes from far back
                prob = np.array([.01, 30, .01, 8, 8, .01, .01, 1, 1, 1, 1, 1, 1], float)
            if trial == 2:       # time-high words only: many wraps
                prob = np.array([1, 5, 1, 1, 1, 1, 30, 0, 0, 0, 0, 0, 0], float)
            ty = r.choice(kinds, size=n, p=prob / prob.sum())
            w = ((ty << 12) | r.integers(0, 4096, n)).astype(np.uint16)
            check(evk, h, orc, w)


def test_evt3_raw_file_and_capacity(evk, orc, tmp_path):
    ev = orc.synth(orc.synth_params(5, 50_000, 1280, 720, 1_000_000, 8))
    ev["p"] = ev["p"] > 0
    w = orc.evt3_encode(ev)
    path = os.path.join(tmp_path, "rec.raw")
    with open(path, "wb") as f:
        f.write(b"% date 2026-10-18\n% evt 3.0\n% format EVT3;height=720;width=1280\n")
        f.write(w.tobytes())
    with evk.Evk(len(ev)) as h:
        assert h.load_raw(path) == len(ev)
        assert h.get_events().tobytes() == ev.tobytes()
        # the fused step runs on the decoded stream like on any other
        ds, km = evk.ds_params(1280, 720, 4, 4, 1000, 0, 1), evk.km_params(8, 2, iters=1)
        a = h.downsample_kmeans(ds, km, True)
        h.load_events(ev)
        assert h.downsample_kmeans(ds, km, True) == a
    with evk.Evk(1000) as h:
        with pytest.raises(evk.EvkError) as e:
            h.load_evt3(w)
        assert e.value.status == -5
