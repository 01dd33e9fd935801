"""The asynchronous event clustering consumer (reference: ACCEL/
AEClustering.{h,cpp}, ACCEL/MyCluster.{h,cpp}, hand-off and per-slice report in
ACCEL/metavision_sdk_get_started5_opencl_store.cpp:435-445,461-521).

Three implementations of the SAME sequential algorithm must agree state for state -- cluster order,
ids, event counts, moving averages bit for bit (IEEE doubles), every stored event:
  Reference  the reference's own sources compiled where they lie (oracle/_ref/libref_aec.so)
  Oracle     oracle/aec_oracle.c (the restatement)
  CUDA       evk_aec_* through the C-ABI (csrc/evk_aec.cu)
CPU tests pin Oracle == Reference (live where the library exists, and through the committed golden
digests everywhere); GPU tests pin CUDA == Oracle and CUDA == the reference's golden digests."""
import ctypes
import json
import os

import numpy as np
import pytest

import aec_streams as S
import evk_loader
from oracle import aec

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "aec_golden.json")


@pytest.fixture(scope="module")
def gold():
    with open(GOLDEN) as f:
        return json.load(f)["cases"]


@pytest.fixture(scope="module")
def evk():
    m = evk_loader.load()
    m.lib()
    return m


# ---------------------------------------------------------------- CPU: oracle vs the reference
def test_glibc_rand_restatement_matches_libc(orc):
    """std::rand() of the reference build is glibc's TYPE_3 generator: the restatement the oracle
    (and the CUDA kernel) carry equals this libc's rand() for several seeds"""
    libc = ctypes.CDLL("libc.so.6")
    for seed in (1, 42, 0, 123456789, 0xFFFFFFFF):
        libc.srand(seed)
        want = np.array([libc.rand() for _ in range(2000)], np.int32)
        assert (aec.glibc_rand_seq(seed, 2000) == want).all()


@pytest.mark.parametrize("case", S.CASES, ids=[c[0] for c in S.CASES])
def test_oracle_matches_reference_golden(orc, gold, case):
    name, init, kind, seed, n, chunk = case
    e = S.make(kind, seed, n)
    o = aec.Oracle(S.INITS[init], rand_seed=1)
    for k, i in enumerate(range(0, n, chunk)):
        o.update(e[i:i + chunk])
        assert S.digest(o.state()) == gold[name]["digests"][k], (name, k)
    st = o.state()
    g = gold[name]
    assert st["ids"].tolist() == g["ids"] and st["n"].tolist() == g["n"] and st["last"] == g["last"]
    assert [[float(v).hex() for v in row] for row in st["mu"]] == g["mu_hex"]


def test_cases_cover_the_hard_paths(orc):
    """the streams above do reach merges of large clusters, random sampling and the exhaustive
    minimum (otherwise the parity tests would prove little)"""
    cov = {}
    for name, init, kind, seed, n, chunk in S.CASES:
        o = aec.Oracle(S.INITS[init])
        o.update(S.make(kind, seed, n))
        cov[name] = o.coverage()
    assert cov["default_distinct_t"][0] > 50 and cov["default_distinct_t"][1] > 100
    assert cov["paper_sampling"][2] > 100_000 and cov["paper_sampling"][0] > 100
    assert cov["exhaustive_min"][2] == 0 and cov["exhaustive_min"][0] > 50
    assert cov["unsorted_times"][0] > 50 and cov["dense_merges"][0] > 200


@pytest.mark.skipif(not aec.ref_available(), reason="oracle/_ref/libref_aec.so not built")
@pytest.mark.parametrize("init", list(S.INITS))
def test_oracle_matches_reference_live(orc, init):
    """fresh random streams (not the golden ones), compared after every chunk"""
    for seed, tie in ((101, 1), (102, 40), (103, 5)):
        e = S.stream(seed, 3000, tie=tie)
        o, r = aec.Oracle(S.INITS[init], rand_seed=9), aec.Reference(S.INITS[init], rand_seed=9)
        for i in range(0, len(e), 250):
            o.update(e[i:i + 250])
            r.update(e[i:i + 250])
            S.same_state(o.state(), r.state())
    e = S.unsorted_stream(104, 3000)
    o, r = aec.Oracle(S.INITS[init]), aec.Reference(S.INITS[init])
    o.update(e)
    r.update(e)
    S.same_state(o.state(), r.state())


def test_handoff_and_report_as_written(orc):
    """hand-off loop (store.cpp:435-445): flat array stepped by 4, bounded by the pair count, one
    pseudo-time per slice; report (:461-521): clusters with n >= minN, previous centroid, arrow
    only once both previous coordinates are positive"""
    uc = np.arange(40, dtype=np.int32) + 100          # pairs (100,101) (102,103) ...
    ev = aec.handoff(uc, 10, 4321)                    # i = 0, 4, 8
    assert ev.tolist() == [[4.321, 100, 101, 0], [4.321, 104, 105, 0], [4.321, 108, 109, 0]]
    assert len(aec.handoff(uc, 0, 1)) == 0
    o = aec.Oracle(dict(sz_buffer=100, radius=5.0, kappa=0, alpha=0.5, min_n=3))
    pts = [[0.0, 50, 60, 0], [0.0, 51, 60, 0], [0.0, 50, 61, 0], [0.0, 300, 300, 0]]
    o.update(np.array(pts, float))
    r1 = o.report()
    assert len(r1) == 1 and r1[0][0] == 0 and r1[0][1] == 3
    assert np.allclose(r1[0][2:4], [151 / 3, 181 / 3]) and r1[0][6] == 0 and (r1[0][4:6] == 0).all()
    o.update(np.array([[0.001, 52, 62, 1]], float))
    r2 = o.report()
    assert r2[0][6] == 1 and (r2[0][4:6] == r1[0][2:4]).all()
    assert np.allclose(r2[0][7:9], r2[0][2:4], rtol=0, atol=1e-12)


# ---------------------------------------------------------------- GPU: CUDA vs oracle / golden
@pytest.mark.gpu
@pytest.mark.parametrize("case", S.CASES, ids=[c[0] for c in S.CASES])
def test_cuda_matches_oracle_and_reference_golden(evk, orc, gold, case):
    name, init, kind, seed, n, chunk = case
    e = S.make(kind, seed, n)
    o = aec.Oracle(S.INITS[init], rand_seed=1)
    with evk.Evk(1024) as h:
        h.aec_create(S.INITS[init], rand_seed=1)
        for k, i in enumerate(range(0, n, chunk)):
            o.update(e[i:i + chunk])
            h.aec_update(e[i:i + chunk])
            st = h.aec_state()
            S.same_state(st, o.state())
            assert S.digest(st) == gold[name]["digests"][k], (name, k)
            rep, orep = h.aec_report(), o.report()
            assert len(rep) == len(orep)
            if len(rep):
                assert (rep["id"] == orep[:, 0]).all() and (rep["n"] == orep[:, 1]).all()
                assert (rep["centroid"] == orep[:, 2:4]).all() and (rep["prev"] == orep[:, 4:6]).all()
                assert (rep["has_arrow"] == orep[:, 6]).all()
                assert (rep["arrow_end"] == orep[:, 7:9]).all()


@pytest.mark.gpu
def test_cuda_event_at_a_time_equals_one_batch(evk, orc):
    """n calls with one event == one call with n events (state lives on the device between calls)"""
    e = S.stream(11, 600, tie=3)
    init = S.INITS["paper"]
    with evk.Evk(1024) as a, evk.Evk(1024) as b:
        a.aec_create(init)
        b.aec_create(init)
        a.aec_update(e)
        for row in e:
            b.aec_update(row[None, :])
        S.same_state(a.aec_state(), b.aec_state())
        b.aec_update(e[:0])
        S.same_state(a.aec_state(), b.aec_state())


@pytest.mark.gpu
def test_cuda_slice_pipeline_stays_on_device(evk, orc):
    """downsample a slice, hand its voxel representatives to the consumer without leaving the
    device (evk_aec_update_voxels), report -- equals the oracle fed with evk_get_voxels' output,
    both with every voxel and with the reference's as-written stride (store.cpp:435-445)"""
    W, H, n = 1280, 720, 60_000
    ev = orc.synth(orc.synth_params(0xE7CA0005, n, W, H, 1_000_000, 12))
    ds = evk.ds_params(W, H, 8, 8, 10_000, 0, 0)
    for literal in (False, True):
        o = aec.Oracle(None)
        with evk.Evk(n) as h:
            h.aec_create(None, max_points=8192)
            total = 0
            for s in range(0, n, 10_000):            # 10 ms slices at 1 Mev/s
                h.load_events(ev[s:s + 10_000])
                U, _ = h.downsample(ds)
                total += U
                t = total / 1000.0                   # uniqueCount / 1000.0, store.cpp:440
                _, reps, _ = h.get_voxels(keys=False, first=False)
                if literal:
                    flat = np.zeros(2 * U + 2, np.int32)
                    flat[0:2 * U:2], flat[1:2 * U:2] = reps["x"], reps["y"]
                    e = aec.handoff(flat, U, total)
                    assert h.aec_update_voxels(t, 0, 2, (U + 3) // 4) == len(e)
                else:
                    e = np.zeros((U, 4))
                    e[:, 0], e[:, 1], e[:, 2] = t, reps["x"], reps["y"]
                    assert h.aec_update_voxels(t) == U
                o.update(e)
                S.same_state(h.aec_state(), o.state())
                rep, orep = h.aec_report(), o.report()
                assert len(rep) == len(orep) and (rep["centroid"] == orep[:, 2:4]).all()
                assert (rep["has_arrow"] == orep[:, 6]).all()
            assert len(o.state()["ids"]) > 3 and o.coverage()[0] > 0


@pytest.mark.gpu
def test_cuda_capacity_and_state_errors(evk, orc):
    e = S.stream(21, 3000)
    with evk.Evk(1024) as h:
        with pytest.raises(evk.EvkError) as x:
            h.aec_update(e[:1])
        assert x.value.status == -4                       # no consumer yet
        h.aec_create(None, max_clusters=8)
        with pytest.raises(evk.EvkError) as x:
            h.aec_update(e)
        assert x.value.status == -5 and "clusters" in str(x.value)
        h.aec_create(dict(sz_buffer=5000, radius=4000.0, kappa=0, alpha=0.5, min_n=5), max_points=64)
        with pytest.raises(evk.EvkError) as x:
            h.aec_update(e)                               # one giant cluster outgrows its ring
        assert x.value.status == -5 and "events" in str(x.value)
        with pytest.raises(evk.EvkError):
            h.aec_create(dict(sz_buffer=0, radius=1.0, kappa=0, alpha=0.5, min_n=5))
        h.aec_create(None)                                # usable again after re-creation
        h.aec_update(e[:100])
        with pytest.raises(evk.EvkError):
            h.aec_update_voxels(0.0)                      # no voxel shard on this handle


@pytest.mark.gpu
def test_cpp_host_replay_with_consumer(evk, orc):
    """store_replay ... aec: the C++ host layer runs the reference's whole slice callback --
    downsample, as-written hand-off (every 2nd pair, <= 2048 per slice), asynchronous event
    clustering, flow arrows -- and its arrows equal the oracle's fed from the oracle's downsample"""
    import re
    import subprocess
    exe = os.path.join(evk_loader.PKG_DIR, "store_replay")
    assert os.path.exists(exe), "host demo not built"
    r = subprocess.run([exe, "synth:1500000", "8", "aec"], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stderr
    got, cur = [], None
    for line in r.stdout.splitlines():
        if line.startswith("slice "):
            cur = []
            got.append(cur)
        elif line.startswith("  cluster "):
            cur.append(tuple(float(v) for v in re.findall(r"-?\d+\.?\d*", line)))
    ev = orc.synth(orc.synth_params(0xE7CA0005, 1_500_000, 1280, 720, 10_000_000, 8))
    o = aec.Oracle(None)
    total, want = 0, []
    for s in range(3):
        e0 = ev[(ev["t"] >= s * 50_000) & (ev["t"] < (s + 1) * 50_000)]
        ok, of, _ = orc.downsample(e0, orc.ds_params(1280, 720, 4, 4, 1000, s * 50_000, 1))
        reps = e0[of]
        U = len(ok)
        total += U
        cnt = min((U + 3) // 4, 2048)
        e = np.zeros((cnt, 4))
        e[:, 0], e[:, 1], e[:, 2] = total / 1000.0, reps["x"][::2][:cnt], reps["y"][::2][:cnt]
        o.update(e)
        rep = o.report()
        want.append([(r_[0], r_[1], round(r_[4], 1), round(r_[5], 1), round(r_[7], 1),
                      round(r_[8], 1)) for r_ in rep if r_[6]])
    assert len(got) == 3
    assert sum(len(w) for w in want) > 0
    for g, w in zip(got, want):
        assert len(g) == len(w)
        for a, b in zip(g, w):
            assert a[:2] == b[:2] and np.allclose(a[2:], b[2:], atol=0.051)
