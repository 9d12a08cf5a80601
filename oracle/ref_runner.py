"""Drive the mechanically transpiled reference classes (oracle/_ref/, made by tools/transpile_reference.py from
/root/reference/src/*.js) -- TEST INFRASTRUCTURE.  This is the reference's own text executing under JS number
semantics (oracle/jsrt.py); it is what pins oracle/softbody_oracle.c and, through the golden vectors
tools/make_ref_golden.py writes, the CUDA BITEXACT path on the GPU box (where /root/reference does not exist).
"""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np

from . import jsrt

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
REFERENCE = "/root/reference"


def reference_present() -> bool:
    return os.path.isfile(os.path.join(REFERENCE, "src", "Softbody.js"))


def ensure() -> bool:
    """(Re)generate oracle/_ref from the reference when it is present; True if oracle/_ref is importable."""
    if reference_present():
        sys.path.insert(0, os.path.join(_ROOT, "tools"))
        try:
            import transpile_reference
            transpile_reference.generate(REFERENCE, os.path.join(_HERE, "_ref"))
        finally:
            sys.path.pop(0)
    return os.path.isfile(os.path.join(_HERE, "_ref", "softbody_ref.py"))


def _mod(name):
    importlib.invalidate_caches()
    return importlib.import_module("oracle._ref." + name)


def js_params(p: dict):
    q = dict(p)
    q["worldBounds"] = jsrt.JSArray([float(x) for x in p["worldBounds"]])
    return jsrt.JSObject(**{k: (float(v) if isinstance(v, (int, float)) and not isinstance(v, bool) else v) for k, v in q.items()})


def _np(a):
    return np.frombuffer(a.tobytes(), dtype=np.float32).copy()


class RefSoftBody:
    """src/Softbody.js class SoftBody, transpiled; numpy views of its typed arrays."""

    def __init__(self, verts, tet_ids, params: dict, vis_verts=None, vis_tri=None, edge_ids=None):
        m = _mod("softbody_ref")
        self.params = js_params(params)
        self.vertices = jsrt.Float32Array(np.asarray(verts, np.float32).reshape(-1).tolist())
        tets = jsrt.JSArray(int(x) for x in np.asarray(tet_ids).reshape(-1))   # src/Dragon.js:311 is a plain Array
        vis = jsrt.Float32Array([] if vis_verts is None else np.asarray(vis_verts, np.float32).reshape(-1).tolist())
        tri = jsrt.JSArray([] if vis_tri is None else (int(x) for x in np.asarray(vis_tri).reshape(-1)))
        edges = jsrt.JSArray([] if edge_ids is None else (int(x) for x in np.asarray(edge_ids).reshape(-1)))
        self.js = m.SoftBody(self.vertices, tets, edges, self.params, vis, tri, None)

    def simulate(self, dt, params: dict | None = None):
        if params is not None:
            self.params = js_params(params)
        self.js.simulate(float(dt), self.params)

    def startGrab(self, p):
        self.js.startGrab(jsrt.JSObject(x=float(p[0]), y=float(p[1]), z=float(p[2])))

    def moveGrabbed(self, p):
        self.js.moveGrabbed(jsrt.JSObject(x=float(p[0]), y=float(p[1]), z=float(p[2])))

    def endGrab(self):
        self.js.endGrab()

    def endFrame(self):
        self.js.endFrame()

    pos = property(lambda s: _np(s.js.pos))
    prevPos = property(lambda s: _np(s.js.prevPos))
    vel = property(lambda s: _np(s.js.vel))
    invMass = property(lambda s: _np(s.js.invMass))
    invRestPose = property(lambda s: _np(s.js.invRestPose))
    invRestVolume = property(lambda s: _np(s.js.invRestVolume))
    volError = property(lambda s: float(s.js.volError))
    grabId = property(lambda s: int(s.js.grabId))
    visPositions = property(lambda s: _np(s.js.visMesh.geometry.attributes.position.array))
    edgePositions = property(lambda s: _np(s.js.edgeMesh.geometry.attributes.position.array))


class _Texture:
    """gpuCompute.createTexture(): a DataTexture whose image.data is a Float32Array(4 * W * H)
    (src/MultiTargetGPUComputationRenderer.js:406-413)."""

    def __init__(self, dim):
        self.image = jsrt.JSObject(data=jsrt.Float32Array(4 * dim * dim))
        self.needsUpdate = False


class RefSoftBodyGPUInit:
    """SoftBodyGPU.initPhysics (src/SoftbodyGPU.js:487-608) transpiled, on textures allocated the way the constructor
    does (:11-43; the constructor itself builds WebGL objects and GLSL strings and is not transpiled)."""

    def __init__(self, verts, tet_ids, density=1000.0):
        m = _mod("softbodygpu_ref")
        js = m.SoftBodyGPU.__new__(m.SoftBodyGPU)
        v = jsrt.Float32Array(np.asarray(verts, np.float32).reshape(-1).tolist())
        t = jsrt.JSArray(int(x) for x in np.asarray(tet_ids).reshape(-1))
        js.numParticles = jsrt._div(v.length, 3)                       # :11
        js.numElems = jsrt._div(t.length, 4)                           # :12
        js.texDim = jsrt.Math.ceil(jsrt.Math.sqrt(js.numElems))        # :13
        dim = int(js.texDim)
        js.inputPos = v.slice(0)                                       # :16
        js.grabPos = jsrt.Float32Array(3)
        js.grabId = -1
        for name in ("pos0", "vel0", "invMass", "invRestVolumeAndColor", "elemToParticlesTable", "quats0"):  # :24-28,42
            setattr(js, name, _Texture(dim))
        js.particleToElemVertsTable = jsrt.JSArray(_Texture(dim) for _ in range(9))   # :29-37
        js.elems0 = jsrt.JSArray(_Texture(dim) for _ in range(4))                     # :38-41
        js.tetIds = t                                                  # :45
        jsrt.console.lines.clear()
        js.initPhysics(float(density))                                 # :46
        self.js, self.texDim = js, dim
        self.biggestT = jsrt.console.lines[-1][0] if jsrt.console.lines else None

    def tex(self, name, k=None):
        t = getattr(self.js, name)
        if k is not None:
            t = t[k]
        return _np(t.image.data)


class RefSoftBodyGPU:
    """The WHOLE WebGL solver executed from the reference's text: SoftBodyGPU.initPhysics + simulate (src/SoftbodyGPU.js:487-641)
    and MultiTargetGPUComputationRenderer.addVariable / addPass / compute (src/MultiTargetGPUComputationRenderer.js:144-176,
    272-306) transpiled from JavaScript, the seven passes (src/SoftbodyGPU.js:59-376) transpiled from GLSL, wired exactly as the
    constructor wires them (the addVariable / addPass / uniform lines are extracted from its text: shaders_ref.VARIABLES,
    PASSES, UNIFORM_BINDINGS).  Hand-written here: the allocation of the textures (constructor :11-43), what init() does to
    the render targets and sampler uniforms (:192-257), and doRenderTarget = "run the fragment shader for every texel".
    Slow (pure Python, one fragment at a time): use for a few substeps."""

    def __init__(self, verts, tet_ids, params: dict):
        from . import glslrt
        self.glsl = glslrt
        sh = self.sh = _mod("shaders_ref")
        init = RefSoftBodyGPUInit(verts, tet_ids, float(params.get("density", 1000.0)))
        js = self.js = init.js
        self.W = W = init.texDim
        js.physicsParams = js_params(dict(params, dt=params.get("dt", 1.0 / 1200.0)))
        gpu = js.gpuCompute = _mod("gpucompute_ref").MultiTargetGPUComputationRenderer()
        gpu.variables, gpu.passes = jsrt.JSArray(), jsrt.JSArray()
        gpu.createShaderMaterial = lambda src: jsrt.JSObject(uniforms=jsrt.JSObject(), fragmentShader=src)
        gpu.doRenderTarget = self._render
        for member, texname, initmember, count in sh.VARIABLES:                      # :49-55
            setattr(js, member, gpu.addVariable(texname, getattr(js, initmember), count if count else jsrt.undefined))
        for pname, var, deps, outs, unis in sh.PASSES:                               # :59-376, in addPass order
            p = gpu.addPass(getattr(js, var), jsrt.JSArray(getattr(js, d) for d in deps), pname)
            p.material.run = getattr(sh, pname)
            setattr(js, pname, p)
        for pname, uname, expr in sh.UNIFORM_BINDINGS:                               # material.uniforms[...] = { value: ... }
            getattr(js, pname).material.uniforms[uname] = jsrt.JSObject(value=self._value(expr))
        # init(), src/MultiTargetGPUComputationRenderer.js:192-257: two render targets per variable, both holding the initial
        # texture; a sampler uniform per dependency (and prev_ for dependencies other than the written variable)
        for v in gpu.variables:
            for k in range(2):
                v.renderTargets[k] = jsrt.JSObject(texture=self._clone(v.initialValueTexture))
        for p in gpu.passes:
            for d in p.dependencies:
                p.material.uniforms[d.name] = jsrt.JSObject(value=None)
                if d.name != p.variable.name:
                    p.material.uniforms["prev_" + d.name] = jsrt.JSObject(value=None)

    def _value(self, expr):
        if expr.startswith("this."):
            o = self.js
            for part in expr[5:].split("."):
                o = getattr(o, part)
            return o
        if expr.startswith("new THREE.Vector3"):
            return jsrt.THREE.Vector3(*[float(x) for x in expr[expr.index("(") + 1:expr.rindex(")")].split(",")])
        return float(expr)

    def _sampler(self, t):
        if isinstance(t, self.glsl.Sampler):
            return t
        if isinstance(t, jsrt.JSArray):
            return [self._sampler(x) for x in t]
        if not hasattr(t, "_sampler"):                      # a DataTexture filled by initPhysics: image.data, W x W RGBA
            t._sampler = self.glsl.Sampler(_np(t.image.data), self.W, self.W)
        return t._sampler

    def _clone(self, t):
        if isinstance(t, jsrt.JSArray):
            return jsrt.JSArray(self._clone(x) for x in t)
        return self.glsl.Sampler(_np(t.image.data).copy(), self.W, self.W)

    def _render(self, material, output):
        g0 = {}
        for name, u in vars(material.uniforms).items():
            v = u.value
            if isinstance(v, (int, float)):
                v = self.glsl.F(v)
            elif isinstance(v, jsrt._Vector3):
                v = self.glsl.vec3(v.x, v.y, v.z)
            elif v is not None:
                v = self._sampler(v)
            g0[name] = v
        targets = output.texture if isinstance(output.texture, jsrt.JSArray) else [output.texture]
        W = self.W
        res = self.glsl.vec2(float(W), float(W))
        for y in range(W):
            for x in range(W):
                g = jsrt.JSObject(gl_FragCoord=self.glsl.vec4(x + 0.5, y + 0.5, 0.0, 1.0), resolution=res, **g0)
                outs = material.run(g)
                for t, o in zip(targets, outs):
                    t.data[y, x, :] = o.d

    def simulate(self, dt, params: dict | None = None):
        if params is not None:
            for k, v in params.items():
                if k != "worldBounds":
                    setattr(self.js.physicsParams, k, float(v) if isinstance(v, (int, float)) and not isinstance(v, bool) else v)
        self.js.simulate(float(dt), self.js.physicsParams)

    def _var(self, member, k=None):
        v = getattr(self.js, member)
        t = self.js.gpuCompute.getCurrentRenderTarget(v).texture
        if k is not None:
            t = t[k]
        return t.data.reshape(-1, 4)

    def _n(self):
        return int(self.js.numParticles), int(self.js.numElems)

    pos = property(lambda s: s._var("pos")[: s._n()[0], :3].reshape(-1).copy())
    prevPos = property(lambda s: s._var("prevPos")[: s._n()[0], :3].reshape(-1).copy())
    vel = property(lambda s: s._var("vel")[: s._n()[0], :3].reshape(-1).copy())
    quat = property(lambda s: s._var("quats")[: s._n()[1], :].reshape(-1).copy())
    rest = property(lambda s: np.stack([s._var("elems", k)[: s._n()[1], :3] for k in range(4)], axis=1).reshape(-1).copy())
