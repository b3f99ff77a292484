"""Radiance .hdr (RGBE, RLE) reader -> float32 RGB, row 0 = top.

Stands in for the stb_image HDR decode inside nvvk::HdrIbl::loadEnvironment (external nvpro_core2;
reference call site src/renderer.cpp:1994-1996).  stbi's conversion is f = mantissa * 2^(e-136)
(no +0.5 bias), which is what is reproduced here.
"""
import numpy as np


def load_hdr(path):
    data = open(path, "rb").read()
    pos = 0
    w = h = None
    while True:
        end = data.index(b"\n", pos)
        line = data[pos:end].decode("latin-1").strip()
        pos = end + 1
        if line.startswith("-Y") or line.startswith("+Y"):
            parts = line.split()
            h, w = int(parts[1]), int(parts[3])
            flip_y = parts[0] == "+Y"
            break
    buf = np.frombuffer(data, np.uint8, offset=pos)
    rgbe = np.empty((h, w, 4), np.uint8)
    p = 0
    for y in range(h):
        if w < 8 or w > 32767 or buf[p] != 2 or buf[p + 1] != 2 or (buf[p + 2] & 0x80):
            # flat (non-RLE) file
            rgbe = np.array(buf[p - 0: p + (h - y) * w * 4]).reshape(-1, w, 4) if y == 0 else rgbe
            if y == 0:
                break
            raise ValueError("mixed RLE/flat scanlines not supported")
        p += 4
        for c in range(4):
            x = 0
            row = rgbe[y, :, c]
            while x < w:
                n = int(buf[p]); p += 1
                if n > 128:
                    n -= 128
                    row[x:x + n] = buf[p]; p += 1
                else:
                    row[x:x + n] = buf[p:p + n]; p += n
                x += n
    e = rgbe[..., 3].astype(np.int32)
    scale = np.where(e > 0, np.ldexp(1.0, e - 136), 0.0).astype(np.float32)
    rgb = rgbe[..., :3].astype(np.float32) * scale[..., None]
    if flip_y:
        rgb = rgb[::-1]
    return np.ascontiguousarray(rgb, np.float32)
