"""ctypes view of oracle/_ref/libref.so: the REFERENCE's own OpenCL kernels (coordinate_processor.cl,
assign_to_centers.cl) compiled as C where they lie under /root/reference (oracle/Makefile, target
`ref`; oracle/cl_shim.h).  TEST INFRASTRUCTURE ONLY.  Built in the development container; the
prebuilt .so travels to the GPU box, the golden vectors made from it are committed under
tests/golden/ (tests/golden/make_ref_golden.py)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_ref", "libref.so")
REFERENCE_ROOT = "/root/reference"


def available():
    return os.path.exists(_LIB)


def build():
    """only where the reference sources are present; returns True when the library exists"""
    if os.path.isdir(REFERENCE_ROOT):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)
    return available()


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_LIB)
        vp, i32 = C.c_void_p, C.c_int
        L.ref_process_coordinates.argtypes = [vp, i32, vp, C.POINTER(i32), C.POINTER(i32)]
        L.ref_process_coordinates.restype = None
        L.ref_assign_to_centers.argtypes = [vp, vp, vp, i32]
        L.ref_assign_to_centers.restype = None
        L.ref_assign_data_cluster.argtypes = [vp, vp, vp, vp, i32]
        L.ref_assign_data_cluster.restype = None
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def process_coordinates(coords, unique_count=0, repeated_count=0):
    """the reference kernel, one work-item: coords int32 [x0,y0,x1,y1,...] ->
    (unique pairs in emission order, cumulative unique_count, cumulative repeated_count)"""
    coords = np.ascontiguousarray(coords, dtype=np.int32)
    n = len(coords) // 2
    uniq = np.zeros(2 * max(n, 1), dtype=np.int32)
    uc, rc = C.c_int(unique_count), C.c_int(repeated_count)
    lib().ref_process_coordinates(_p(coords), n, _p(uniq), C.byref(uc), C.byref(rc))
    emitted = uc.value - unique_count
    return uniq[: 2 * emitted].reshape(-1, 2).copy(), uc.value, rc.value


def assign_to_centers(data, centers):
    """the reference kernel per point: data float32 [x0,y0,...], centers float32[16] ->
    assignments int32 (2k, or 255 when no centre is within 50)"""
    data = np.ascontiguousarray(data, dtype=np.float32)
    centers = np.ascontiguousarray(centers, dtype=np.float32)
    assert centers.size == 16
    n = len(data) // 2
    out = np.zeros(n, dtype=np.int32)
    lib().ref_assign_to_centers(_p(data), _p(centers), _p(out), n)
    return out


def assign_data_cluster(data, assignments):
    """the reference scatter kernel -> (output float32[8*4096], cluster_index int32[8])"""
    data = np.ascontiguousarray(data, dtype=np.float32)
    assignments = np.ascontiguousarray(assignments, dtype=np.int32)
    out = np.zeros(8 * 4096, dtype=np.float32)
    ci = np.zeros(8, dtype=np.int32)
    lib().ref_assign_data_cluster(_p(data), _p(assignments), _p(ci), _p(out), len(assignments))
    return out, ci
