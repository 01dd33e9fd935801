"""GPU parity of the RAW EVT 2.0 ingest (evk_load_evt2 / evk_load_raw) against the oracle's decoder:
the decoded stream must be byte-identical, whatever mix of word types, and the pipeline behind it
must give the same voxels and centroids as the 16-byte loader."""
import os
import tempfile

import numpy as np
import pytest

import evk_loader

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def evk():
    m = evk_loader.load()
    m.lib()
    return m


def test_evt2_decode_matches_oracle(evk, orc):
    rng = np.random.default_rng(11)
    with evk.Evk(3_000_000) as h:
        for n, rate in ((0, 1), (1, 1_000_000), (4095, 10_000_000), (4097, 100_000_000),
                        (1_000_003, 100_000_000), (2_500_000, 3_000_000)):
            ev = orc.synth(orc.synth_params(0xE7CA0003, n, 1280, 720, rate, 64)) if n else \
                np.zeros(0, dtype=evk.EVENT_DTYPE)
            w = orc.evt2_encode(ev)
            assert h.load_evt2(w) == n
            assert h.get_events().tobytes() == ev.tobytes()
        # other word types sprinkled in (triggers, OTHERS, CONTINUED), CD words before the first
        # EVT_TIME_HIGH, long gaps without events: positions and times must not shift
        ev = orc.synth(orc.synth_params(5, 300_000, 1280, 720, 50_000_000, 16))
        w = orc.evt2_encode(ev)
        junk = rng.choice(np.array([0xA0000101, 0xE0000000, 0xF1234567, 0xEFFFFFFF], np.uint32),
                          size=50_000)
        pos = np.sort(rng.integers(0, len(w) + 1, size=len(junk)))
        mixed = np.insert(w, pos, junk)
        mixed = np.concatenate([np.array([(1 << 28) | (7 << 22) | (3 << 11) | 4], np.uint32),
                                mixed, np.full(10_000, 0x80001234, np.uint32)])
        want = orc.evt2_decode(mixed)
        assert h.load_evt2(mixed) == len(want) == len(ev) + 1
        assert h.get_events().tobytes() == want.tobytes()
        # capacity is reported, never fatal
        big = orc.evt2_encode(orc.synth(orc.synth_params(1, 3_000_001, 1280, 720, 10**8, 4)))
        with pytest.raises(evk.EvkError) as e:
            h.load_evt2(big)
        assert e.value.status == -5


def test_raw_file_and_pipeline(evk, orc):
    n, W, H, K = 1_500_000, 1280, 720, 32
    ev = orc.synth(orc.synth_params(0xE7CA0004, n, W, H, 100_000_000, 32))
    w = orc.evt2_encode(ev)
    ds, km = evk.ds_params(W, H, 2, 2, 500, 0, 1), evk.km_params(K, 2, iters=1)
    with tempfile.TemporaryDirectory() as d, evk.Evk(n) as h:
        path = os.path.join(d, "rec.raw")
        with open(path, "wb") as f:
            f.write(b"% Date 2026-01-01 00:00:00\n% evt 2.0\n% geometry 1280x720\n% end\n")
            f.write(w.tobytes())
        assert h.load_raw(path) == n
        a = h.downsample_kmeans(ds, km, True)
        ka, _, fa = h.get_voxels(reps=False)
        ca, na = h.get_centroids(K, 2)
        h.load_events(ev)
        b = h.downsample_kmeans(ds, km, True)
        kb, _, fb = h.get_voxels(reps=False)
        cb, nb = h.get_centroids(K, 2)
        assert a == b and (ka == kb).all() and (fa == fb).all()
        assert (ca == cb).all() and (na == nb).all()
        with open(path, "wb") as f:
            f.write(b"% evt 2.1\n% end\n" + w.tobytes())   # (EVT 3.0: tests/test_gpu_evt3.py)
        with pytest.raises(evk.EvkError) as e:
            h.load_raw(path)
        assert e.value.status == -6
