"""CPU restatement (numpy, fp32) of the tone-mapping + 8-bit encode stage -- TEST INFRASTRUCTURE, never imported by the product.

Reference call site: GltfRenderer::tonemap, src/renderer.cpp:992-1054 (nvshaders::Tonemapper::runCompute on eImgRendered ->
eImgTonemapped).  The shader itself is nvpro_core2 code outside the reference tree, so the operators are restated from their
publications (Hejl / Burgess-Dawson filmic, Hable's Uncharted 2, Hill's ACES fit, Wrensch's minimal AgX, Khronos PBR neutral) and
**parity with the reference is unpinned** (no golden image, no test of the reference touches a pixel).  What pins this file are
the known answers in tests/test_tonemap.py (sRGB anchors, operator fixed points, monotonicity, the published PBR-neutral
properties).
"""
import numpy as np

F = np.float32
BINS, MIN_LOG, MAX_LOG = 256, -16.0, 16.0


def srgb(c):
    c = np.maximum(c, F(0)).astype(F)
    with np.errstate(invalid="ignore"):
        hi = F(1.055) * np.power(c, F(1.0 / 2.4), dtype=F) - F(0.055)
    return np.where(c <= F(0.0031308), c * F(12.92), hi).astype(F)


def _filmic(c):
    t = np.maximum(F(0), c - F(0.004))
    return (t * (F(6.2) * t + F(0.5))) / (t * (F(6.2) * t + F(1.7)) + F(0.06))


def _hable(x):
    a, b, c, d, e, f = F(0.15), F(0.50), F(0.10), F(0.20), F(0.02), F(0.30)
    return ((x * (a * x + c * b) + d * e) / (x * (a * x + b) + d * f)) - e / f


def _agx_contrast(x):
    x2 = x * x
    x4 = x2 * x2
    return F(15.5) * x4 * x2 - F(40.14) * x4 * x + F(31.96) * x4 - F(6.868) * x2 * x + F(0.4298) * x2 + F(0.1191) * x - F(0.00232)


def _mat(rows, c):
    """rows: 3x3 python floats (row-major); left-to-right sums in fp32"""
    out = []
    for r in rows:
        out.append(F(r[0]) * c[..., 0] + F(r[1]) * c[..., 1] + F(r[2]) * c[..., 2])
    return np.stack(out, -1).astype(F)


def operator(method, c):
    c = np.asarray(c, F)
    if method == 0:
        return _filmic(c).astype(F)
    if method == 1:
        ws = F(1.0) / _hable(F(11.2))
        return srgb(_hable(c * F(2.0)) * ws)
    if method == 3:
        v = _mat([[0.59719, 0.35458, 0.04823], [0.07600, 0.90834, 0.01566], [0.02840, 0.13383, 0.83777]], c)
        w = (v * (v + F(0.0245786)) - F(0.000090537)) / (v * (F(0.983729) * v + F(0.4329510)) + F(0.238081))
        return srgb(_mat([[1.60475, -0.53108, -0.07367], [-0.10208, 1.10813, -0.00605], [-0.00327, -0.07276, 1.07602]], w))
    if method == 4:
        mn, mx = F(-12.47393), F(4.026069)
        v = _mat([[0.842479062253094, 0.0784335999999992, 0.0792237451477643], [0.0423282422610123, 0.878468636469772, 0.0791661274605434],
                  [0.0423756549057051, 0.0784336, 0.879142973793104]], c)
        l = np.clip(np.log2(np.maximum(v, F(1e-10)), dtype=F), mn, mx)
        w = _agx_contrast(((l - mn) / (mx - mn)).astype(F)).astype(F)
        return _mat([[1.19687900512017, -0.0980208811401368, -0.0990297440797205], [-0.0528968517574562, 1.15190312990417, -0.0989611768448433],
                     [-0.0529716355144438, -0.0980434501171241, 1.15107367264116]], w)
    if method == 5:
        start, desat = F(0.8) - F(0.04), F(0.15)
        x = c.min(-1)
        offset = np.where(x < F(0.08), x - F(6.25) * x * x, F(0.04)).astype(F)
        k = (c - offset[..., None]).astype(F)
        peak = k.max(-1)
        d = F(1.0) - start
        with np.errstate(divide="ignore", invalid="ignore"):
            new_peak = (F(1.0) - d * d / (peak + d - start)).astype(F)
            s = (new_peak / peak).astype(F)
            k2 = (k * s[..., None]).astype(F)
            g = (F(1.0) - F(1.0) / (desat * (peak - new_peak) + F(1.0))).astype(F)
            comp = (k2 + (new_peak[..., None] - k2) * g[..., None]).astype(F)
        return srgb(np.where((peak < start)[..., None], k, comp))
    return srgb(c)  # clip


def auto_exposure(img, base):
    """base * 0.18 / log-average luminance, from the 256-bin log2 histogram (bin 0 = black and below 2^-16: not counted)"""
    c = np.asarray(img, F)[..., :3].reshape(-1, 3)
    lum = (F(0.2126) * c[:, 0] + F(0.7152) * c[:, 1] + F(0.0722) * c[:, 2]).astype(F)
    pos = lum > 0
    t = (np.log2(lum[pos], dtype=F) - F(MIN_LOG)) / F(MAX_LOG - MIN_LOG)
    b = np.clip((t * F(BINS)).astype(np.int32), 0, BINS - 1)
    hist = np.bincount(b, minlength=BINS).astype(np.float64)
    hist[0] = 0.0
    if hist.sum() == 0:
        return F(base), hist
    centres = MIN_LOG + (np.arange(BINS) + 0.5) * (MAX_LOG - MIN_LOG) / BINS
    avg = float((centres * hist).sum() / hist.sum())
    return F(base) * F(0.18 / 2.0 ** avg), hist


def tonemap(img, method=0, is_active=1, exposure=1.0, brightness=1.0, contrast=1.0, saturation=1.0, vignette=0.0, auto=0, y0=0, full_height=None):
    """RGBA32F [rows, width, 4] -> (RGBA8 [rows, width, 4], exposure used)"""
    img = np.asarray(img, F)
    rows, width = img.shape[:2]
    full_height = full_height or rows
    ex = F(exposure)
    c = img[..., :3]
    if is_active:
        if auto:
            ex, _ = auto_exposure(img, exposure)
        r = operator(method, (c * ex).astype(F))
        r = np.clip(F(0.5) + (r - F(0.5)) * F(contrast), F(0), F(1)).astype(F)
        r = np.power(r, F(1.0) / F(brightness), dtype=F)
        luma = (F(0.299) * r[..., 0] + F(0.587) * r[..., 1] + F(0.114) * r[..., 2]).astype(F)[..., None]
        r = (luma + (r - luma) * F(saturation)).astype(F)
        u = ((np.arange(width, dtype=F) + F(0.5)) / F(width) - F(0.5)) * F(2.0)
        v = ((np.arange(rows, dtype=F) + F(y0) + F(0.5)) / F(full_height) - F(0.5)) * F(2.0)
        vg = (F(1.0) - (u[None, :] * u[None, :] + v[:, None] * v[:, None]) * F(vignette)).astype(F)
        r = (r * vg[..., None]).astype(F)
    else:
        r = c
    out = np.concatenate([r, img[..., 3:4]], -1)
    out = np.where(np.isnan(out), F(0), out)
    q = (np.clip(out, F(0), F(1)) * F(255.0) + F(0.5)).astype(np.uint32)
    return q.astype(np.uint8), ex
