"""Generates tests/golden/codec_frames.npz by running the UNMODIFIED reference recorder codec
(tools/record.py compress_frame / decompress_frame, imported through oracle/refimport.py with a
`zstandard` shim over the system libzstd) on three small consecutive frames.  Authoring container only.

    python tests/golden/make_golden_codec.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refimport  # noqa: E402

rec = refimport.load_recorder()
import b200sim  # noqa: E402,F401
from b200sim import codec  # noqa: E402

rng = np.random.default_rng(7)
n = 600
frames = []
p = (rng.normal(size=(n, 3)) * 120).astype(np.float32)
c = rng.random((n, 3)).astype(np.float32)
frames.append((p, c))
for k in range(2):
    p = (p + rng.normal(size=(n, 3)).astype(np.float32) * (0.4 + 3.0 * k)).astype(np.float32)
    c = np.clip(c + rng.normal(size=(n, 3)).astype(np.float32) * 0.02, 0, 1).astype(np.float32)
    frames.append((p, c))
# a few deltas beyond the int16 range (wrap-around is part of the format's behaviour)
frames[2][0][:5, 0] += np.array([40.0, -40.0, 70.0, 33.0, -100.0], np.float32)

out = {"zstd_version": np.array(codec.zstd_version())}
prev = (None, None)
dec_prev = (None, None)
for k, (p, c) in enumerate(frames):
    data = rec.compress_frame(p, c, prev[0], prev[1])
    dp, dc = rec.decompress_frame(data, dec_prev[0], dec_prev[1])
    out[f"pos{k}"], out[f"col{k}"] = p, c
    out[f"bytes{k}"] = np.frombuffer(data, np.uint8)
    out[f"dec_pos{k}"], out[f"dec_col{k}"] = dp, dc
    prev = (p, c)            # the recorder deltas against the ORIGINAL previous frame (tools/record.py:470-490)
    dec_prev = (dp, dc)      # a reader decodes against the DECODED previous frame
path = os.path.join(ROOT, "tests", "golden", "codec_frames.npz")
np.savez_compressed(path, **out)
print("wrote", path, {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if k.startswith("bytes")})
