"""Shared scene helpers for the tests (world-space triangles through the oracle's flatten)."""
import os

import numpy as np

from conftest import MEDIA


def load_scene(gpurt, ctx, name):
    s = gpurt.Scene(ctx)
    if name == "cbox":
        s.load(os.path.join(MEDIA, "cbox", "cbox.gltf"))
    elif name == "mis_test":
        s.load(os.path.join(MEDIA, "mis_test", "mis_test.gltf"))
    elif name == "cube":
        s.load(os.path.join(MEDIA, "cube.gltf"))
    elif name == "sponza_standin":
        s.make_sponza_standin()
    else:
        raise KeyError(name)
    return s


def world_tris(orc, scene):
    """oracle-side flattening (N1) of a gpurt.Scene -> (n,9) f32 in global primitive order"""
    parts = [orc.flatten(*scene.object(i), np.array(d.model, np.float32)) for i, d in enumerate(scene.descs())]
    return np.concatenate(parts) if parts else np.zeros((0, 9), np.float32)


def soup(n, seed=1, ext=0.05):
    rng = np.random.default_rng(seed)
    c = rng.random((n, 1, 3), dtype=np.float32)
    return (c + (rng.random((n, 3, 3), dtype=np.float32) - 0.5) * ext).reshape(n, 9).astype(np.float32)


def same_bits(a, b):
    a = np.ascontiguousarray(a).reshape(-1).view(np.uint8)
    b = np.ascontiguousarray(b).reshape(-1).view(np.uint8)
    return a.size == b.size and bool((a == b).all())
