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


# ---- interleaved bands: balances cheap (sky) and expensive (atrium floor) rows across ranks -------------------
def interleave_band(height, world, max_band=8):
    """largest band height <= max_band such that height is a multiple of band * world; None if there is none."""
    if world <= 1 or height % world:
        return None
    per = height // world
    for band in range(min(max_band, per), 0, -1):
        if per % band == 0:
            return band
    return None


def interleaved_rows(height, world, rank, band):
    """global row index of every local row of `rank` (length height // world)."""
    return [((l // band) * world + rank) * band + l % band for l in range(height // world)]


def deinterleave(gathered, height, world, band):
    """gathered: [world * (height // world), W, 4] in rank-major order -> [height, W, 4] in image order."""
    per = height // world
    nb = per // band
    g = gathered.reshape(world, nb, band, *gathered.shape[1:])
    perm = (1, 0, 2) + tuple(range(3, g.ndim))
    g = g.permute(*perm) if hasattr(g, "permute") else g.transpose(perm)
    return g.reshape(height, *gathered.shape[1:])
