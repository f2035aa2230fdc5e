"""Decode the reference's only real-data fixture -- elasticfusionpublic/GPUTest/{1c,1d,2c,2d}.png -- into
tests/golden/gputest_pair.npz, the way GPUTest.cpp reads it (GPUTest/src/GPUTest.cpp:30-60, :172-205):

  rgb1, rgb2   uint8 [480, 640, 3]   1c.png / 2c.png as RGB (pangolin::LoadImage -> GL_RGB upload)
  depth1_raw   uint16 [480, 640]     1d.png, raw 16-bit values (5000 units per metre)
  depth2_raw   uint16 [480, 640]     2d.png, raw

Everything GPUTest derives from these (depth2 / 5 -> millimetres, :49-55; model vertices / normals from 1d.png with
K = 528, 528, 320, 240 and forward differences, :62-130) is arithmetic and lives in tests/util.py::gputest_inputs so that
the fixture holds nothing but the decoded files.  Run here (the reference tree is not on the GPU box):

    python tests/golden/make_gputest_fixture.py
"""
import os
import sys

import cv2
import numpy as np

SRC = "/root/reference/elasticfusionpublic/GPUTest"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gputest_pair.npz")


def main():
    rgb = {}
    for name in ("1c", "2c"):
        bgr = cv2.imread(os.path.join(SRC, name + ".png"), cv2.IMREAD_COLOR)
        assert bgr is not None and bgr.shape == (480, 640, 3), name
        rgb[name] = np.ascontiguousarray(bgr[:, :, ::-1])
    dep = {}
    for name in ("1d", "2d"):
        d = cv2.imread(os.path.join(SRC, name + ".png"), cv2.IMREAD_UNCHANGED)
        assert d is not None and d.shape == (480, 640) and d.dtype == np.uint16, (name, None if d is None else (d.shape, d.dtype))
        dep[name] = d
    np.savez_compressed(OUT, rgb1=rgb["1c"], rgb2=rgb["2c"], depth1_raw=dep["1d"], depth2_raw=dep["2d"])
    print(OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    sys.exit(main())
