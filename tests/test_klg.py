"""The on-disk input of the path: .klg logs, InstanceFusion src/utilities/RawLogReader.cpp.
The reference ships no log file and no test for its reader, so the reader is checked against logs written in the
documented layout (raw / zlib depth, raw / JPEG / absent image) and against hand-packed bytes."""
import struct
import zlib

import numpy as np

from instancefusion_b200.klg import KlgReader, KlgWriter


def _frames(n, w, h, seed=0):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        depth = rng.integers(0, 6000, (h, w), dtype=np.uint16)
        yy, xx = np.mgrid[0:h, 0:w]
        rgb = np.stack([(xx * 3 + i * 7) % 256, (yy * 5) % 256, ((xx + yy) * 2) % 256], -1).astype(np.uint8)
        out.append((1000 + 33 * i, depth, rgb))
    return out


def test_raw_and_zlib_round_trip(tmp_path):
    w, h = 64, 48
    frames = _frames(5, w, h)
    for compress in (False, True):
        p = str(tmp_path / f"log_{compress}.klg")
        with KlgWriter(p, w, h, compress_depth=compress, jpeg_quality=None) as wr:
            for f in frames:
                wr.write(*f)
        rd = KlgReader(p, w, h)
        assert len(rd) == 5
        for i, (ts, depth, rgb) in enumerate(rd):
            assert ts == frames[i][0] and np.array_equal(depth, frames[i][1]) and np.array_equal(rgb, frames[i][2])
        rd.close()


def test_hand_packed_bytes_and_missing_image(tmp_path):
    """byte layout of RawLogReader.cpp:33, :68-103 written by hand: zlib depth, no image (imageSize 0 -> zeros)"""
    w, h = 8, 4
    depth = np.arange(w * h, dtype=np.uint16).reshape(h, w) * 100
    z = zlib.compress(depth.tobytes())
    raw = struct.pack("<i", 1) + struct.pack("<q", 123456789012) + struct.pack("<i", len(z)) + struct.pack("<i", 0) + z
    p = tmp_path / "hand.klg"
    p.write_bytes(raw)
    rd = KlgReader(str(p), w, h)
    ts, d, rgb = rd.getNext()
    assert ts == 123456789012 and np.array_equal(d, depth) and not rgb.any() and rgb.shape == (h, w, 3)
    assert not rd.hasMore()
    rd.close()


def test_jpeg_flip_back_and_fast_forward(tmp_path):
    w, h = 64, 48
    frames = _frames(6, w, h, seed=3)
    p = str(tmp_path / "jpeg.klg")
    with KlgWriter(p, w, h, compress_depth=True, jpeg_quality=95) as wr:
        for f in frames:
            wr.write(*f)
    rd = KlgReader(p, w, h)
    ts, d, rgb = rd.getNext()
    assert np.array_equal(d, frames[0][1])
    assert np.abs(rgb.astype(int) - frames[0][2].astype(int)).mean() < 12            # JPEG is lossy
    rd.getNext()
    ts2, d2, _ = rd.getNext()
    assert ts2 == frames[2][0] and rd.currentFrame == 3
    tsb, db, _ = rd.getBack()                                                        # re-reads the frame just read (:48-57)
    assert tsb == frames[2][0] and np.array_equal(db, frames[2][1])
    rd.rewind()
    rd.fastForward(4)                                                                # :119-135
    ts4, d4, _ = rd.getNext()
    assert ts4 == frames[4][0] and np.array_equal(d4, frames[4][1])
    rd.close()
    flipped = KlgReader(p, w, h, flip_colors=True)
    _, _, rgbf = flipped.getNext()
    assert np.array_equal(rgbf, rgb[:, :, ::-1])                                     # :108-114
    flipped.close()
