"""Host-side writer of RAW EVT 2.0 words (input generation for benchmarks and demos).

The reference replays recordings through Metavision::Camera::from_file
(event-cam-clustering-accel/event-cam-clustering-downsampling-accel/
metavision_sdk_get_started5_opencl_store.cpp:336); a recording is a stream of 32-bit words:
  CD event    type(4: 0 = OFF, 1 = ON) | t & 63 (6) | x (11) | y (11)
  TIME_HIGH   type(4: 8)               | t >> 6 (28)
with a TIME_HIGH word wherever the upper time bits change.  This module only WRITES the format
(vectorised numpy); decoding is done on the device by evk_load_evt2 (csrc/evk_evt2.cu).
"""
import numpy as np


def encode_evt2(events):
    """events: structured array with fields x, y, p, t (evk_event layout) -> uint32 words."""
    n = len(events)
    if n == 0:
        return np.zeros(0, np.uint32)
    x = events["x"].astype(np.uint32)
    y = events["y"].astype(np.uint32)
    t = events["t"].astype(np.int64)
    if (x >= 2048).any() or (y >= 2048).any() or (t < 0).any() or (t >= (1 << 34)).any():
        raise ValueError("event outside the EVT 2.0 field ranges")
    th = t >> 6
    new = np.empty(n, bool)
    new[0] = True
    np.not_equal(th[1:], th[:-1], out=new[1:])
    pos = np.arange(n, dtype=np.int64) + np.cumsum(new)   # slot of every CD word
    words = np.empty(n + int(new.sum()), np.uint32)
    words[pos] = ((events["p"] > 0).astype(np.uint32) << 28) | ((t & 63).astype(np.uint32) << 22) \
        | (x << 11) | y
    words[pos[new] - 1] = np.uint32(0x80000000) | (th[new] & 0x0FFFFFFF).astype(np.uint32)
    return words
