"""CPU tests of the drop-in boundary: libevk.so builds for sm_100a, loads, exports every symbol
include/evk.h declares, and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import evk_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def evk():
    evk_loader.build()
    return evk_loader.load()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "evk.h")).read()
    return sorted(set(re.findall(r"EVK_API\s+[\w\s\*]+?\b(evk_\w+)\s*\(", src)))


def test_header_and_binding_agree(evk):
    syms = header_symbols()
    assert len(syms) >= 30
    assert sorted(evk.SYMBOLS) == syms


def test_library_exports_every_declared_symbol(evk):
    L = evk.lib()
    for s in header_symbols():
        assert hasattr(L, s), s
    out = subprocess.check_output(["nm", "-D", "--defined-only", evk.LIB_PATH], text=True)
    exported = set(re.findall(r" T (evk_\w+)", out))
    assert set(header_symbols()) <= exported
    assert b"sm_100a" in L.evk_version()


def test_library_is_sm100a_only(evk):
    out = subprocess.run(["cuobjdump", "--list-elf", evk.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_struct_layouts(evk):
    assert evk.EVENT_DTYPE.itemsize == 16 and evk.EVENT_DTYPE.fields["t"][1] == 8
    assert C.sizeof(evk.DsParams) == 48 and C.sizeof(evk.KmParams) == 32
    assert C.sizeof(evk.SynthParams) == 56 and C.sizeof(evk.StageTimes) == 36
    assert C.sizeof(evk.AecParams) == 48 and C.sizeof(evk.DbscanParams) == 40
    assert evk.AEC_CLUSTER_DTYPE.itemsize == 40 and evk.AEC_FLOW_DTYPE.itemsize == 64


def test_no_cpu_fallback(evk):
    """without a CUDA device evk_create must fail with EVK_ERR_CUDA — never compute on the host"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(evk.EvkError) as e:
        evk.Evk(1024)
    assert e.value.status == -2


def test_product_does_not_reference_the_oracle():
    pkg = evk_loader.PKG_DIR
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".cu", ".cuh", ".py", ".hpp", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle/" not in txt and "liborc" not in txt and "import orc" not in txt, f


def test_null_handle_is_rejected_everywhere(evk):
    """no entry point dereferences a NULL handle (no GPU needed: nothing is computed)"""
    L = evk.lib()
    n = C.c_size_t(0)
    i = C.c_int(0)
    calls = [
        lambda: L.evk_aec_create(None, None), lambda: L.evk_aec_update(None, None, 0),
        lambda: L.evk_aec_update_voxels(None, 0.0, 0, 1, 0),
        lambda: L.evk_aec_get_clusters(None, None, 0, C.byref(n), C.byref(i)),
        lambda: L.evk_aec_get_points(None, 0, None, None, None, None, 0, C.byref(n)),
        lambda: L.evk_aec_report(None, None, 0, C.byref(n)),
        lambda: L.evk_ts_create(None, 1280, 720), lambda: L.evk_ts_corners(None, 1, C.byref(n)),
        lambda: L.evk_ts_get_corners(None, None, 0), lambda: L.evk_ts_get_surface(None, None, 0),
        lambda: L.evk_dbscan_points(None, None, 0, None, C.byref(n), C.byref(n)),
        lambda: L.evk_dbscan_voxels(None, None, C.byref(n), C.byref(n)),
        lambda: L.evk_dbscan_get(None, None, 0, None, None, 0, None, 0),
        lambda: L.evk_load_evt3(None, None, 0, C.byref(n)),
        lambda: L.evk_downsample_kmeans_submit(None, None, None, 1),
        lambda: L.evk_downsample_kmeans_wait(None, C.byref(n), C.byref(n), C.byref(i)),
        lambda: L.evk_downsample_kmeans_sharded_submit(None, None, None, 1, 0),
        lambda: L.evk_downsample_kmeans_sharded_wait(None, C.byref(n), C.byref(n), C.byref(i)),
        lambda: L.evk_window_config_events(None, None, None, 10),
        lambda: L.evk_num_voxels(None, C.byref(n), C.byref(n)),
    ]
    for k, f in enumerate(calls):
        assert f() < 0, k
    # the destroy calls are no-ops on NULL
    assert L.evk_aec_destroy(None) == 0 and L.evk_ts_destroy(None) == 0
    assert L.evk_dbscan_destroy(None) == 0
