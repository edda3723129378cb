"""Frame codec against frames produced by the reference's own compress_frame
(tests/golden/codec_frames.npz, made by tests/golden/make_golden_codec.py).  CPU only."""
import os

import numpy as np
import pytest

import b200sim  # noqa: F401
from b200sim import codec


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "codec_frames.npz"))


def test_decodes_reference_frames_bit_exact(g):
    p, c = None, None
    for k in range(3):
        p, c = codec.decompress_frame(g[f"bytes{k}"].tobytes(), p, c)
        assert p.dtype == np.float32 and p.shape == g[f"pos{k}"].shape
        assert np.array_equal(p, g[f"dec_pos{k}"]) and np.array_equal(c, g[f"dec_col{k}"])
    # format 1 is lossless, format 2 quantises to 1e-3
    assert np.array_equal(g["dec_pos0"], g["pos0"])
    inr = np.abs(g["pos1"] - g["pos0"]) < 32.0
    assert np.abs(g["dec_pos1"] - g["pos1"])[inr].max() <= 1.001e-3


def test_payloads_and_layout_equal_the_references(g):
    prev = (None, None)
    for k in range(3):
        mine = codec.compress_frame(g[f"pos{k}"], g[f"col{k}"], prev[0], prev[1])
        ref = g[f"bytes{k}"].tobytes()
        assert mine[0] == ref[0] == (1 if k == 0 else 2)
        # the zstd payloads decode to identical bytes (the compressed bytes themselves depend on the libzstd build)
        def payloads(b):
            n0 = int.from_bytes(b[1:5], "little")
            n1 = int.from_bytes(b[5 + n0:9 + n0], "little")
            assert len(b) == 9 + n0 + n1
            return codec.zstd_decompress(b[5:5 + n0]), codec.zstd_decompress(b[9 + n0:9 + n0 + n1])
        assert payloads(mine) == payloads(ref)
        if str(g["zstd_version"]) == codec.zstd_version():
            assert mine == ref
        prev = (g[f"pos{k}"], g[f"col{k}"])


def test_delta_payload_is_the_recorders_arithmetic(g):
    d = codec.delta_payload(g["pos2"], g["pos1"])
    ref = ((g["pos2"] - g["pos1"]) * 1000).astype(np.int16)
    assert d.dtype == np.int16 and np.array_equal(d, ref)
    # out-of-range deltas wrap (the format's behaviour); in-range ones are truncated toward zero
    x = np.array([[0.0015, -0.0015, 0.0]], np.float32)
    assert codec.delta_payload(x, np.zeros_like(x)).tolist() == [[1, -1, 0]]


def test_load_frame_walks_back_to_a_base_frame(tmp_path, g):
    w = codec.FrameWriter(tmp_path, level=3, threads=2)
    w.submit_absolute(0, g["pos0"], g["col0"])
    w.submit_delta(1, codec.delta_payload(g["pos1"], g["pos0"]), codec.delta_payload(g["col1"], g["col0"]))
    w.submit_delta(2, codec.delta_payload(g["pos2"], g["pos1"]), codec.delta_payload(g["col2"], g["col1"]))
    w.close()
    assert w.bytes_out < w.bytes_in
    p2, c2 = codec.load_frame(tmp_path, 2)
    assert np.array_equal(p2, g["dec_pos2"]) and np.array_equal(c2, g["dec_col2"])
    p1, _ = codec.load_frame(tmp_path, 1)
    assert np.array_equal(p1, g["dec_pos1"])
    with pytest.raises(FileNotFoundError):
        codec.load_frame(tmp_path, 7)
    os.remove(codec.frame_path(tmp_path, 0))
    with pytest.raises(FileNotFoundError):
        codec.load_frame(tmp_path, 2)
    # an uncompressed .npz (what the recorder writes first) is a valid base
    np.savez(tmp_path / "frame_0000.npz", positions=g["pos0"], colors=g["col0"])
    p2b, _ = codec.load_frame(tmp_path, 2)
    assert np.array_equal(p2b, g["dec_pos2"])


def test_multithreaded_zstd_round_trips():
    rng = np.random.default_rng(0)
    a = (rng.normal(size=200_000) * 30).astype(np.int16)
    for workers in (0, 2):
        z = codec.zstd_compress(a, 5, workers)
        assert codec.zstd_decompress(z) == a.tobytes()
    with pytest.raises(ValueError):
        codec.compress_delta_frame(a.astype(np.int32), a)


def test_live_reference_codec_reads_our_frames(g):
    """When the reference tree is present (authoring container): its decompress_frame reads our bytes."""
    from oracle import refimport
    if not refimport.available():
        pytest.skip("reference tree not present")
    rec = refimport.load_recorder()
    b0 = codec.compress_frame(g["pos0"], g["col0"])
    b1 = codec.compress_delta_frame(codec.delta_payload(g["pos1"], g["pos0"]), codec.delta_payload(g["col1"], g["col0"]))
    p0, c0 = rec.decompress_frame(b0)
    p1, c1 = rec.decompress_frame(b1, p0, c0)
    assert np.array_equal(p0, g["pos0"]) and np.array_equal(p1, g["dec_pos1"]) and np.array_equal(c1, g["dec_col1"])
