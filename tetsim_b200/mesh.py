"""Mesh inputs for the solver: the reference's Dragon, tiled copies of a body, the synthetic beam.

The reference's only input is src/Dragon.js (five JS array literals, src/Dragon.js:1,311,1080,1705,11640);
tools/extract_dragon.py turns it into tetsim_b200/assets/dragon_mesh.npz (a packaged asset), which is what is loaded here.
The tiler and the beam generator produce the BASELINE.json scale configs.
"""
from __future__ import annotations

import os

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRAGON_NPZ = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets", "dragon_mesh.npz")


def load_dragon(path: str = DRAGON_NPZ) -> dict:
    """dragonTetVerts / dragonTetIds / dragonTetEdgeIds / dragonAttachedVerts / dragonAttachedTriIds."""
    m = np.load(path)
    return {k: m[k] for k in m.files}


def tile_bodies(verts, tet_ids, nx: int, nz: int, pitch=(4.0, 0.0, 2.0), y_shift: float = 0.0):
    """nx*nz translated copies of one body on a grid in the xz-plane (BASELINE config 5).

    Copies are independent bodies (the reference has no body-body collision).  Translation is done in
    float32, so each copy's rest pose rounds differently -- copies are NOT bit-identical to each other.
    """
    v = np.asarray(verts, np.float32).reshape(-1, 3)
    t = np.asarray(tet_ids, np.int32).reshape(-1, 4)
    n = len(v)
    out_v, out_t = [], []
    k = 0
    for iz in range(nz):
        for ix in range(nx):
            off = np.array([(ix - (nx - 1) / 2.0) * pitch[0], y_shift, (iz - (nz - 1) / 2.0) * pitch[2]], np.float32)
            out_v.append(v + off)
            out_t.append(t + k * n)
            k += 1
    return np.concatenate(out_v).reshape(-1), np.concatenate(out_t).reshape(-1).astype(np.int32)


_KUHN = np.array([  # 6 tets per cube, all sharing the main diagonal 0-7 (corner index = x + 2y + 4z)
    [0, 1, 3, 7], [0, 3, 2, 7], [0, 2, 6, 7], [0, 6, 4, 7], [0, 4, 5, 7], [0, 5, 1, 7]], np.int64)


def make_beam(cells=(407, 64, 64), h: float = 0.01, y0: float = 1.0, jitter: float = 0.0, seed: int = 1234):
    """Kuhn-split grid beam, long axis x, centred in x and z, y in [y0, y0 + cells[1]*h] (BASELINE config 4).

    Default 407 x 64 x 64 cells -> 10,002,432 tets, 1,723,800 vertices.  Vertices are numbered
    x-fastest, tets cell-major (6 consecutive tets per cell, cells x-fastest).  All tets are positively
    oriented.  jitter (fraction of h) displaces interior vertices with numpy.random.default_rng(seed).
    """
    cx, cy, cz = cells
    nx, ny, nz = cx + 1, cy + 1, cz + 1
    xs = (np.arange(nx, dtype=np.float64) - cx / 2.0) * h
    ys = y0 + np.arange(ny, dtype=np.float64) * h
    zs = (np.arange(nz, dtype=np.float64) - cz / 2.0) * h
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    verts = np.stack([X, Y, Z], axis=-1).reshape(-1, 3)
    if jitter > 0.0:
        rng = np.random.default_rng(seed)
        d = rng.uniform(-jitter * h, jitter * h, size=verts.shape)
        iz, iy, ix = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
        interior = ((ix > 0) & (ix < cx) & (iy > 0) & (iy < cy) & (iz > 0) & (iz < cz)).reshape(-1)
        verts[interior] += d[interior]
    kz, ky, kx = np.meshgrid(np.arange(cz), np.arange(cy), np.arange(cx), indexing="ij")
    base = (kx + nx * (ky + ny * kz)).reshape(-1).astype(np.int64)
    corner = np.array([dx + nx * (dy + ny * dz) for dz in (0, 1) for dy in (0, 1) for dx in (0, 1)], np.int64)
    tets = base[:, None, None] + corner[_KUHN][None, :, :]
    return verts.astype(np.float32).reshape(-1), tets.reshape(-1).astype(np.int32)


def wide_bounds(extent: float = 64.0):
    """worldBounds wide enough for tiled scenes (the parameter is honoured per call, src/Softbody.js:215)."""
    return (-extent, -1.0, -extent, extent, 10.0, extent)


def shard_bodies(verts, tet_ids, rank: int, world_size: int):
    """The bodies (connected components) of a scene that rank `rank` of `world_size` simulates (BASELINE config 5 on N GPUs).

    Bodies never interact -- the reference has no body-body collision, its scene is a list of independent soft bodies
    (src/main.js:51,80-84) -- so the scene shards by bodies with NO exchange: body b goes to rank b * world_size // numBodies
    (contiguous runs, equal counts +-1).  Returns (verts, tet_ids, vert_ids, tet_index): the rank's sub-mesh with vertices
    renumbered 0..n-1 in ascending caller order, the caller's vertex id of each, and the caller's tet index of each local tet
    (ascending, so every body keeps the reference's sweep order).  Positions are the caller's float32 values untouched, so
    each body's result is bit-identical to the one it has in the unsharded scene."""
    from . import _capi
    v = np.asarray(verts, np.float32).reshape(-1, 3)
    t = np.asarray(tet_ids, np.int32).reshape(-1, 4)
    comp, nb = _capi.connected_components(t, len(v))
    owner = (comp.astype(np.int64) * world_size) // max(nb, 1)
    vert_ids = np.flatnonzero(owner == rank).astype(np.int32)
    tet_index = np.flatnonzero(owner[t[:, 0]] == rank).astype(np.int32)
    remap = np.full(len(v), -1, np.int32)
    remap[vert_ids] = np.arange(len(vert_ids), dtype=np.int32)
    return v[vert_ids].reshape(-1).copy(), remap[t[tet_index]].reshape(-1).astype(np.int32), vert_ids, tet_index
