"""Reader / writer of ElasticFusion's .klg RGB-D logs -- the on-disk input of the path.

Format (InstanceFusion src/utilities/RawLogReader.cpp:21-115; the Logger2 tool writes it):
    int32 numFrames
    per frame:  int64 timestamp | int32 depthSize | int32 imageSize | depth bytes | image bytes
    depth  = width*height uint16 millimetres, raw when depthSize == width*height*2, else zlib (uncompress, :82-90)
    image  = width*height RGB8, raw when imageSize == width*height*3, JPEG when 0 < imageSize < that, absent when 0
             (all zeros, :92-103); flipColors swaps R and B (:108-114)

`KlgReader` mirrors RawLogReader's interface (getNext / getBack / fastForward / rewind / hasMore, currentFrame,
timestamp, depth, rgb); JPEG frames are decoded with OpenCV (the reference uses libjpeg through its JPEGLoader).
Host-side file I/O only: nothing here touches the GPU.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np


class KlgReader:
    def __init__(self, path: str, width: int = 640, height: int = 480, flip_colors: bool = False):
        self.width, self.height, self.flip_colors = width, height, flip_colors
        self.num_pixels = width * height
        self._fp = open(path, "rb")
        self.numFrames = struct.unpack("<i", self._read(4))[0]          # :33
        self.currentFrame = 0
        self._file_pointers = []                                        # the stack getBack() pops (:48-57)
        self.timestamp = 0
        self.depth = None
        self.rgb = None

    def _read(self, n: int) -> bytes:
        b = self._fp.read(n)
        if len(b) != n:
            raise EOFError("truncated .klg log")
        return b

    def close(self):
        self._fp.close()

    def __len__(self):
        return self.numFrames

    def hasMore(self) -> bool:
        return self.currentFrame + 1 < self.numFrames                  # RawLogReader::hasMore

    def rewind(self):
        self._fp.seek(4)
        self._file_pointers.clear()
        self.currentFrame = 0

    def getNext(self):
        self._file_pointers.append(self._fp.tell())                    # :61
        return self._get_core()

    def getBack(self):
        assert self._file_pointers, "getBack() without a previous getNext()"
        self._fp.seek(self._file_pointers.pop())                       # :52-54
        return self._get_core()

    def fastForward(self, frame: int):
        while self.currentFrame < frame and self.hasMore():            # :119-135: skip payloads without decoding
            self._file_pointers.append(self._fp.tell())
            self.timestamp, depth_size, image_size = struct.unpack("<qii", self._read(16))
            self._fp.seek(depth_size + image_size, 1)
            self.currentFrame += 1

    def _get_core(self):
        self.timestamp, depth_size, image_size = struct.unpack("<qii", self._read(16))   # :68-71
        depth_bytes = self._read(depth_size)
        image_bytes = self._read(image_size) if image_size > 0 else b""
        if depth_size == self.num_pixels * 2:                                              # :80
            raw = depth_bytes
        else:
            raw = zlib.decompress(depth_bytes)                                             # :86-89
        self.depth = np.frombuffer(raw, np.uint16, self.num_pixels).reshape(self.height, self.width).copy()
        if image_size == self.num_pixels * 3:                                              # :92
            rgb = np.frombuffer(image_bytes, np.uint8, self.num_pixels * 3).reshape(self.height, self.width, 3).copy()
        elif image_size > 0:                                                               # :96-99 JPEG
            import cv2
            bgr = cv2.imdecode(np.frombuffer(image_bytes, np.uint8), cv2.IMREAD_COLOR)
            if bgr is None or bgr.shape[:2] != (self.height, self.width):
                raise ValueError("cannot decode the JPEG image of frame %d" % self.currentFrame)
            rgb = bgr[:, :, ::-1].copy()                                                   # the reference's JPEGLoader yields RGB
        else:
            rgb = np.zeros((self.height, self.width, 3), np.uint8)                         # :100-103
        if self.flip_colors:
            rgb = rgb[:, :, ::-1].copy()                                                   # :108-114
        self.rgb = rgb
        self.currentFrame += 1
        return self.timestamp, self.depth, self.rgb

    def __iter__(self):
        self.rewind()
        for _ in range(self.numFrames):
            yield self.getNext()


class KlgWriter:
    """Writes the same format (what ElasticFusion's Logger2 produces): depth zlib-compressed or raw, image JPEG or raw."""

    def __init__(self, path: str, width: int = 640, height: int = 480, compress_depth: bool = True, jpeg_quality: int | None = 90):
        self.width, self.height = width, height
        self.compress_depth, self.jpeg_quality = compress_depth, jpeg_quality
        self._fp = open(path, "wb")
        self._fp.write(struct.pack("<i", 0))
        self.numFrames = 0

    def write(self, timestamp: int, depth: np.ndarray, rgb: np.ndarray | None):
        depth = np.ascontiguousarray(depth, np.uint16)
        assert depth.shape == (self.height, self.width)
        d = depth.tobytes()
        if self.compress_depth:
            d = zlib.compress(d)
        if rgb is None:
            img = b""
        else:
            rgb = np.ascontiguousarray(rgb, np.uint8)
            assert rgb.shape == (self.height, self.width, 3)
            if self.jpeg_quality is None:
                img = rgb.tobytes()
            else:
                import cv2
                ok, enc = cv2.imencode(".jpg", rgb[:, :, ::-1], [cv2.IMWRITE_JPEG_QUALITY, int(self.jpeg_quality)])
                assert ok
                img = enc.tobytes()
        self._fp.write(struct.pack("<qii", int(timestamp), len(d), len(img)))
        self._fp.write(d)
        self._fp.write(img)
        self.numFrames += 1

    def close(self):
        self._fp.seek(0)
        self._fp.write(struct.pack("<i", self.numFrames))
        self._fp.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
