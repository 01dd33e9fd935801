"""Time surface + corner test (reference: the event callback of
event-cam-tracking/event-cam-fast-corner-tracker/metavision_time_surface_periodic_group_track.cpp
:883-1063).  The reference code sits in a lambda of main() that needs the Metavision SDK, so it
cannot run here: PARITY UNPINNED by the reference.  The oracle's literal restatement
(oracle/evk_oracle.c orc_ts_corners) is pinned by hand-built surfaces with known answers; the CUDA
path (evk_ts_*, csrc/evk_corner.cu) must equal the oracle index for index and surface for surface."""
import numpy as np
import pytest

import evk_loader

W, H = 1280, 720
C3 = [(0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3), (0, -3), (-1, -3), (-2, -2),
      (-3, -1), (-3, 0), (-3, 1), (-2, 2), (-1, 3)]                        # FCT:44 (dy, dx)
C4 = [(0, 4), (1, 4), (2, 3), (3, 2), (4, 1), (4, 0), (4, -1), (3, -2), (2, -3), (1, -4), (0, -4),
      (-1, -4), (-2, -3), (-3, -2), (-4, -1), (-4, 0), (-4, 1), (-3, 2), (-2, 3), (-1, 4)]  # FCT:45


@pytest.fixture(scope="module")
def evk():
    m = evk_loader.load()
    m.lib()
    return m


def events(orc, rows):
    ev = np.zeros(len(rows), orc.EVENT_DTYPE)
    if len(rows):
        a = np.array(rows, np.int64)
        ev["x"], ev["y"], ev["t"] = a[:, 0], a[:, 1], a[:, 2]
    ev["p"] = 1
    return ev


def moving_corner(orc, steps=60, arm=14, cx=200, cy=150, t0=1000, dt=120, seed=0):
    """the tip of an L-shaped edge moving diagonally: events on both arms, shuffled per step"""
    r = np.random.default_rng(seed)
    rows = []
    for k in range(steps):
        pts = [(cx + k, cy + k - a) for a in range(arm)] + [(cx + k - a, cy + k) for a in range(1, arm)]
        r.shuffle(pts)
        rows += [(x, y, t0 + k * dt + j) for j, (x, y) in enumerate(pts)]
    return events(orc, rows)


def arc_surface(streak3, streak4, x=100, y=80, ramp=0):
    """a surface whose circles around (x, y) hold a newest arc of the given lengths starting at
    circle index 2 (arc pixels = 1000 + ramp * k, everything else 10).  With ramp = 0 the arc pixels
    tie, so an arc longer than the largest streak size cannot qualify (a pixel outside any streak
    is as new as the streak's oldest member)."""
    s = np.full((H, W), 10, np.int64)
    for k in range(streak3):
        dy, dx = C3[(2 + k) % 16]
        s[y + dy, x + dx] = 1000 + ramp * k
    for k in range(streak4):
        dy, dx = C4[(2 + k) % 20]
        s[y + dy, x + dx] = 1000 + ramp * k
    return s


def test_oracle_known_answers(orc):
    """streak lengths inside / outside 3..6 (radius 3) and 4..8 (radius 4), ties, border handling"""
    probe = events(orc, [(100, 80, 5000)])     # stamps only its own pixel (not on a circle)
    for s3, s4, want in ((3, 4, 1), (6, 8, 1), (4, 6, 1), (2, 6, 0), (7, 6, 0), (4, 3, 0), (4, 9, 0),
                         (0, 0, 0), (16, 20, 0)):
        s = arc_surface(s3, s4)
        assert len(orc.ts_corners(probe, W, H, s, True)) == want, (s3, s4)
        assert s[80, 100] == 5000
    # strictly increasing times along a 7-arc: its 6 newest pixels are a valid streak (the 7th is
    # older than all of them), so it IS a corner; likewise 9 on the outer circle
    assert len(orc.ts_corners(probe, W, H, arc_surface(7, 9, ramp=1), True)) == 1
    # a pixel outside the arc as new as the arc's oldest member kills the streak (tj >= min_t)
    s = arc_surface(4, 6)
    dy, dx = C3[10]
    s[80 + dy, 100 + dx] = 1000
    assert len(orc.ts_corners(probe, W, H, s, True)) == 0
    # border: as written the first border event ends the range; skip mode tests the others
    s = arc_surface(4, 6)
    rng_ev = events(orc, [(100, 80, 5000), (2, 80, 5001), (100, 80, 5002)])
    assert orc.ts_corners(rng_ev, W, H, s.copy(), True).tolist() == [0]
    assert orc.ts_corners(rng_ev, W, H, s.copy(), False).tolist() == [0, 2]
    assert orc.ts_corners(rng_ev[1:], W, H, s.copy(), True).tolist() == []
    # the stamps of the WHOLE range are in place before the first test (FCT:888-923 precede :931)
    s = np.full((H, W), 10, np.int64)
    arc = [(100 + dx, 80 + dy, 2000 + k) for k, (dy, dx) in enumerate(C3[2:6])] + \
          [(100 + dx, 80 + dy, 2000 + k) for k, (dy, dx) in enumerate(C4[2:8])]
    first_then_arc = events(orc, [(100, 80, 1999)] + arc)
    assert 0 in orc.ts_corners(first_then_arc, W, H, s, True).tolist()
    ev = moving_corner(orc)
    s = np.zeros((H, W), np.int64)
    tot = sum(len(orc.ts_corners(ev[i:i + 200], W, H, s, True)) for i in range(0, len(ev), 200))
    assert tot == 844            # regression pin of the restatement on the moving-corner stream


@pytest.mark.gpu
def test_cuda_matches_oracle(evk, orc):
    cases = [("moving corner", moving_corner(orc), 200),
             ("moving corner, one range", moving_corner(orc, seed=3), 10_000),
             ("synthetic blobs", orc.synth(orc.synth_params(0xE7CA0003, 300_000, W, H, 10_000_000, 16)),
              5000),
             ("synthetic, big ranges", orc.synth(orc.synth_params(7, 400_000, W, H, 50_000_000, 32)),
              100_000)]
    for literal in (True, False):
        for name, ev, chunk in cases:
            s = np.zeros((H, W), np.int64)
            total = 0
            with evk.Evk(max(chunk, 1024)) as h:
                h.ts_create(W, H)
                for i in range(0, len(ev), chunk):
                    want = orc.ts_corners(ev[i:i + chunk], W, H, s, literal)
                    h.load_events(ev[i:i + chunk])
                    got = h.ts_corners(literal)
                    assert (got == want).all() and len(got) == len(want), (name, literal, i)
                    total += len(got)
                assert (h.ts_surface() == s).all(), name
            if not literal:
                assert total > 100, (name, total)


@pytest.mark.gpu
def test_cuda_known_surfaces_and_errors(evk, orc):
    with evk.Evk(1024) as h:
        with pytest.raises(evk.EvkError) as e:
            h.ts_corners()
        assert e.value.status == -4
        h.ts_create(W, H)
        h.load_events(events(orc, []))
        assert len(h.ts_corners()) == 0
        # build the arc with events (the surface starts at zero), then probe
        arc = [(100 + dx, 80 + dy, 2000 + k) for k, (dy, dx) in enumerate(C3[2:6])] + \
              [(100 + dx, 80 + dy, 2000 + k) for k, (dy, dx) in enumerate(C4[2:8])]
        ev = events(orc, arc + [(100, 80, 3000), (2, 80, 3001), (100, 80, 3002), (1279, 719, 3003)])
        s = np.zeros((H, W), np.int64)
        for literal in (True, False):
            h.ts_create(W, H)
            s[:] = 0
            h.load_events(ev)
            want = orc.ts_corners(ev, W, H, s, literal)
            assert h.ts_corners(literal).tolist() == want.tolist()
            assert (h.ts_surface() == s).all()
        with pytest.raises(evk.EvkError):
            h.ts_create(4, 4)


def test_arc_growth_equals_the_literal_streak_loops():
    """The CUDA kernel does not run the reference's nested start x size loops: it grows ONE arc from
    the newest pixel (always taking the newer neighbour) and checks min(arc) > max(rest) per size.
    This pins that reformulation against the literal loops (FCT:958-1003) on random circles with
    heavy ties -- the same statement csrc/evk_corner.cu `streak` implements."""
    def literal(t, smin, smax):
        n = len(t)
        for i in range(n):
            for sz in range(smin, smax + 1):
                if t[i] < t[(i - 1) % n] or t[(i + sz - 1) % n] < t[(i + sz) % n]:
                    continue
                mn = min(t[(i + j) % n] for j in range(sz))
                if all(t[(i + j) % n] < mn for j in range(sz, n)):
                    return True
        return False

    def grown(t, smin, smax):
        n = len(t)
        m = max(range(n), key=lambda k: (t[k], -k))
        lo = hi = m
        cur, mins, added = t[m], {1: t[m]}, {}
        for size in range(2, smax + 1):
            L, R = t[(lo - 1) % n], t[(hi + 1) % n]
            if L >= R:
                lo, a = (lo - 1) % n, L
            else:
                hi, a = (hi + 1) % n, R
            cur = min(cur, a)
            mins[size], added[size] = cur, a
        rest, k = None, (hi + 1) % n
        while k != lo:
            rest = t[k] if rest is None else max(rest, t[k])
            k = (k + 1) % n
        for size in range(smax, smin - 1, -1):
            if mins[size] > rest:
                return True
            rest = max(rest, added[size])
        return False

    r = np.random.default_rng(0)
    pos = 0
    for trial in range(30_000):
        n, smin, smax = (16, 3, 6) if trial % 2 == 0 else (20, 4, 8)
        mode = trial % 4
        if mode == 0:
            t = r.integers(0, 4, n)
        elif mode == 1:
            t = r.integers(0, 1000, n)
        else:
            t = r.integers(0, 10, n) if mode == 2 else np.full(n, 10)
            s, L = r.integers(0, n), r.integers(1, 10)
            for j in range(L):
                t[(s + j) % n] = 100 + (r.integers(0, 5) if mode == 2 else 0)
        t = [int(v) for v in t]
        a = literal(t, smin, smax)
        assert a == grown(t, smin, smax), t
        pos += a
    assert pos > 3000
