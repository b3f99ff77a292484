"""Packing of BSDF unit-test records for b200pt_bsdf_eval / b200pt_bsdf_sample (include/b200pt.h).

48 floats per record:
  0-2 baseColor | 3-4 roughness(alpha x,y) | 5 metallic | 6-8 N | 9-11 T | 12-14 B | 15-17 Ng |
  18 ior1 | 19 ior2 | 20 specular | 21-23 specularColor | 24 transmission | 25 thickness |
  26 clearcoat | 27 clearcoatRoughness | 28 iridescence | 29 iridescenceIor | 30 iridescenceThickness |
  31-33 sheenColor | 34 sheenRoughness | 35 diffuseTransmissionFactor | 36-38 diffuseTransmissionColor |
  39-41 k1 | 42-44 k2 | 45-47 xi
eval output  (8 floats): bsdf_diffuse.xyz, bsdf_glossy.xyz, pdf, 0
sample output(8 floats): k2.xyz, bsdf_over_pdf.xyz, pdf, event_type
"""
import numpy as np


def _unit(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def random_records(n, seed=1234, features="all"):
    """Random but physically sensible materials + directions (k1 in the upper hemisphere of N)."""
    rng = np.random.default_rng(seed)
    r = np.zeros((n, 48), np.float32)
    r[:, 0:3] = rng.random((n, 3))
    rough = np.maximum(rng.random((n, 1)), 0.0014142) ** 2
    aniso = rng.random((n, 1)) < 0.3
    r[:, 3:4] = np.where(aniso, rough + (1 - rough) * rng.random((n, 1)) ** 2, rough)
    r[:, 4:5] = rough
    r[:, 5] = np.where(rng.random(n) < 0.3, 1.0, np.where(rng.random(n) < 0.5, 0.0, rng.random(n)))
    N = _unit(rng.normal(size=(n, 3)))
    T = _unit(np.cross(N, _unit(rng.normal(size=(n, 3)))))
    B = np.cross(N, T)
    r[:, 6:9], r[:, 9:12], r[:, 12:15] = N, T, B
    r[:, 15:18] = N  # geometric normal == shading normal
    inside = rng.random(n) < 0.2
    r[:, 18] = np.where(inside, 1.5, 1.0)
    r[:, 19] = np.where(inside, 1.0, 1.0 + rng.random(n))
    r[:, 20] = np.where(rng.random(n) < 0.2, rng.random(n), 1.0)
    r[:, 21:24] = np.where(rng.random((n, 1)) < 0.3, rng.random((n, 3)), 1.0)
    r[:, 24] = np.where(rng.random(n) < 0.35, rng.random(n), 0.0)
    r[:, 25] = np.where(rng.random(n) < 0.5, 1.0, 0.0)
    r[:, 26] = np.where(rng.random(n) < 0.25, rng.random(n), 0.0)
    r[:, 27] = np.maximum(rng.random(n), 0.001)
    r[:, 28] = np.where(rng.random(n) < 0.25, rng.random(n), 0.0)
    r[:, 29] = 1.1 + rng.random(n)
    r[:, 30] = 100 + 500 * rng.random(n)
    r[:, 31:34] = np.where(rng.random((n, 1)) < 0.25, rng.random((n, 3)), 0.0)
    r[:, 34] = np.maximum(rng.random(n), 0.0014142)
    r[:, 35] = np.where(rng.random(n) < 0.2, rng.random(n), 0.0)
    r[:, 36:39] = rng.random((n, 3))
    if features == "basic":
        r[:, 24] = 0
        r[:, 26] = 0
        r[:, 28] = 0
        r[:, 31:34] = 0
        r[:, 35] = 0
    # k1 in the +N hemisphere, k2 anywhere
    d = _unit(rng.normal(size=(n, 3)))
    d = np.where(np.sum(d * N, -1, keepdims=True) < 0, -d, d)
    r[:, 39:42] = d
    r[:, 42:45] = _unit(rng.normal(size=(n, 3)))
    r[:, 45:48] = rng.random((n, 3))
    return r
