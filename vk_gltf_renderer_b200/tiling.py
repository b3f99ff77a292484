"""Framebuffer partition for multi-GPU rendering.

The reference is single-GPU.  Pixels are independent (seed = xxhash32(x, y, frame), running-mean accumulation per
pixel), so the frame is split into one contiguous row strip per rank, the scene is replicated, and the strips are
concatenated with one all-gather per frame.  Strips are padded to equal height so a single
all_gather_into_tensor works; `assemble` drops the padding.
"""


def strip_rows(height, world):
    """rows per rank (padded strip height)."""
    return (height + world - 1) // world


def partition_rows(height, world, rank):
    """(y0, rows) of `rank`'s strip; rows may be 0 for trailing ranks of tiny images."""
    per = strip_rows(height, world)
    y0 = min(rank * per, height)
    return y0, max(0, min(per, height - y0))


def assemble(gathered, height):
    """gathered: array/tensor [world * strip_rows, W, 4] -> [height, W, 4]."""
    return gathered[:height]
