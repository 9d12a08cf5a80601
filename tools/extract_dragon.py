#!/usr/bin/env python3
"""Extract the reference's only input mesh (src/Dragon.js) into a binary fixture.

The reference ships the dragon as five JS array literals (src/Dragon.js:1,311,1080,1705,11640).
There is no JS engine in this image, so the literals are parsed as text: decimal text ->
float32 for the two Float32Arrays (same rounding a JS `new Float32Array([...])` applies:
decimal -> float64 -> float32), ints for the three index arrays.

Run in the build container only (reads /root/reference); the output
tetsim_b200/assets/dragon_mesh.npz is committed and is what travels to the GPU box.
"""
import re
import sys
import numpy as np

SRC = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src/Dragon.js"
OUT = sys.argv[2] if len(sys.argv) > 2 else "tetsim_b200/assets/dragon_mesh.npz"

text = open(SRC).read()


def literal(name):
    m = re.search(r"export var %s\s*=\s*(?:new Float32Array\()?\[(.*?)\]\)?;" % name, text, re.S)
    assert m, name
    return [t for t in re.split(r"[\s,]+", m.group(1)) if t]


def f32(name):
    # JS: parse decimal to double, then Float32Array store rounds double -> float (RNE)
    return np.array([float(t) for t in literal(name)], dtype=np.float64).astype(np.float32)


def i32(name):
    return np.array([int(t) for t in literal(name)], dtype=np.int64)


verts = f32("dragonTetVerts")
tet_ids = i32("dragonTetIds")
edge_ids = i32("dragonTetEdgeIds")
vis_verts = f32("dragonAttachedVerts")
vis_tris = i32("dragonAttachedTriIds")

assert verts.size == 3 * 1234, verts.size
assert tet_ids.size == 4 * 3840, tet_ids.size
assert edge_ids.size == 2 * 6222, edge_ids.size
assert vis_verts.size == 4 * 29800, vis_verts.size
assert vis_tris.size == 3 * 59657, vis_tris.size
assert tet_ids.min() == 0 and tet_ids.max() == 1233
assert vis_tris.max() == 29799

np.savez_compressed(
    OUT,
    tet_verts=verts,
    tet_ids=tet_ids.astype(np.int32),
    tet_edge_ids=edge_ids.astype(np.int32),
    vis_verts=vis_verts,
    vis_tri_ids=vis_tris.astype(np.int32),
)
print("wrote", OUT, {k: v.shape for k, v in np.load(OUT).items()})
