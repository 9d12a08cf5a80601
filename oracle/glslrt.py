"""GLSL ES 3.00 run-time semantics for the mechanically transpiled shader passes -- TEST INFRASTRUCTURE.

``tools/transpile_shaders.py`` re-emits the seven fragment-shader passes of ``src/SoftbodyGPU.js:59-376`` (template strings
in the reference) token for token as Python; this file is what their types and built-ins mean.  GLSL leaves the precision of
``highp float`` arithmetic and of the built-ins to the implementation; the FLOAT MODEL chosen here -- and stated so that it
can be argued with -- is the one ``oracle/polar_oracle.c`` and the CUDA BITEXACT flavour use:

* every ``+ - * /`` is an IEEE binary32 operation rounded on its own (numpy float32 arithmetic; no FMA contraction);
* ``sin(x)`` = float32(sin(float64(x)));  ``sqrt`` (inside ``length``) and ``/`` are correctly rounded;
* ``dot(a, b)`` = ((a.x*b.x + a.y*b.y) + a.z*b.z) (+ a.w*b.w), left to right;  ``length(v)`` = sqrt(dot(v, v));
  ``normalize(v)`` = v / length(v) component by component;  ``cross`` = (a.y*b.z - b.y*a.z, a.z*b.x - b.z*a.x, a.x*b.y - b.x*a.y);
* ``clamp(x, lo, hi)`` = min(max(x, lo), hi) and ``min/max`` as the GLSL ES specification defines them;
* ``texture(sampler, uv)`` with NEAREST filtering and CLAMP_TO_EDGE (src/MultiTargetGPUComputationRenderer.js:144-162:
  minFilter = magFilter = NearestFilter): texel = clamp(int(floor(uv * size)), 0, size - 1).

GLSL assignment and parameter passing are BY VALUE: the emitted code wraps every initialiser, right-hand side, argument and
return value in ``_v`` (a deep copy of vectors / matrices / arrays).
"""
from __future__ import annotations

import math

import numpy as np

np.seterr(all="ignore")   # unused texels legitimately compute 0 / 0 (src/SoftbodyGPU.js:320); this module is test infrastructure only
F = np.float32
_f32 = np.float32   # what the emitted code calls for float literals (a shader may name a variable F: src/SoftbodyGPU.js:352)
_SW = {"x": 0, "y": 1, "z": 2, "w": 3, "r": 0, "g": 1, "b": 2, "a": 3}


class vec:
    """vec2 / vec3 / vec4: float32 components, swizzles, component-wise arithmetic."""
    __slots__ = ("d",)

    def __init__(self, d):
        object.__setattr__(self, "d", np.asarray(d, dtype=np.float32).copy())

    def __len__(self):
        return self.d.size

    def __getitem__(self, i):
        return self.d[int(i)]

    def __setitem__(self, i, v):
        self.d[int(i)] = F(v)

    def __getattr__(self, name):
        try:
            idx = [_SW[c] for c in name]
        except KeyError:
            raise AttributeError(name)
        return self.d[idx[0]] if len(idx) == 1 else vec(self.d[idx])

    def __setattr__(self, name, value):
        idx = [_SW[c] for c in name]
        if len(idx) == 1:
            self.d[idx[0]] = F(value)
        else:
            self.d[idx] = value.d if isinstance(value, vec) else F(value)

    def _b(self, o, op):
        return vec(op(self.d, o.d if isinstance(o, vec) else F(o)))

    def __add__(self, o): return self._b(o, np.add)
    def __sub__(self, o): return self._b(o, np.subtract)
    def __mul__(self, o): return self._b(o, np.multiply)
    def __truediv__(self, o): return self._b(o, np.divide)
    def __radd__(self, o): return vec(np.add(F(o), self.d))
    def __rsub__(self, o): return vec(np.subtract(F(o), self.d))
    def __rmul__(self, o): return vec(np.multiply(F(o), self.d))
    def __rtruediv__(self, o): return vec(np.divide(F(o), self.d))
    def __neg__(self): return vec(-self.d)
    def copy(self): return vec(self.d)
    def __repr__(self): return "vec%d%s" % (self.d.size, tuple(float(x) for x in self.d))


class mat3:
    """mat3: three column vectors; m[c] is column c (a live view, so m[c][r] += x works), m[c][r] as in GLSL."""
    __slots__ = ("c",)

    def __init__(self, diag=0.0):
        self.c = [vec([F(diag) if r == k else F(0.0) for r in range(3)]) for k in range(3)]

    def __getitem__(self, k):
        return self.c[int(k)]

    def __setitem__(self, k, v):
        self.c[int(k)] = v.copy()

    def copy(self):
        m = mat3()
        m.c = [x.copy() for x in self.c]
        return m


def _v(x):
    """GLSL value semantics."""
    if isinstance(x, (vec, mat3)):
        return x.copy()
    if isinstance(x, list):
        return [_v(e) for e in x]
    return x


def _flat(args):
    out = []
    for a in args:
        if isinstance(a, vec):
            out.extend(a.d.tolist())
        else:
            out.append(float(a))
    return out


def _ctor(n):
    def make(*args):
        f = _flat(args)
        if len(f) == 1:
            f = f * n
        if len(f) != n:
            raise TypeError("vec%d constructed from %d components" % (n, len(f)))
        return vec(f)
    return make


vec2, vec3, vec4 = _ctor(2), _ctor(3), _ctor(4)


def float_(x):
    return F(x)


def int_(x):
    return int(x)   # truncation toward zero, like the GLSL constructor


def _div(a, b):
    if isinstance(a, int) and isinstance(b, int):
        q = abs(a) // abs(b)
        return q if (a >= 0) == (b >= 0) else -q
    with np.errstate(divide="ignore", invalid="ignore"):
        return a / b


def _mod(a, b):
    if isinstance(a, int) and isinstance(b, int):
        return a - b * _div(a, b)
    raise TypeError("% is an integer operator in GLSL")


def _arr(n, make):
    return [make() for _ in range(n)]


def dot(a, b):
    p = a.d * b.d
    s = p[0]
    for k in range(1, p.size):
        s = s + p[k]
    return s


def cross(a, b):
    return vec([a.d[1] * b.d[2] - b.d[1] * a.d[2], a.d[2] * b.d[0] - b.d[2] * a.d[0], a.d[0] * b.d[1] - b.d[0] * a.d[1]])


def length(v):
    return np.sqrt(dot(v, v))


def normalize(v):
    with np.errstate(divide="ignore", invalid="ignore"):
        return vec(v.d / length(v))


def _sin1(x):
    return F(math.sin(float(x))) if math.isfinite(float(x)) else F("nan")


def sin(x):
    return vec([_sin1(c) for c in x.d]) if isinstance(x, vec) else _sin1(x)


def _cw(f, a, b):
    if isinstance(a, vec) or isinstance(b, vec):
        n = len(a) if isinstance(a, vec) else len(b)
        ga = (lambda k: a.d[k]) if isinstance(a, vec) else (lambda k: F(a))
        gb = (lambda k: b.d[k]) if isinstance(b, vec) else (lambda k: F(b))
        return vec([f(ga(k), gb(k)) for k in range(n)])
    return f(F(a), F(b))


def min_(a, b):   # GLSL: y < x ? y : x
    return _cw(lambda x, y: y if y < x else x, a, b)


def max_(a, b):   # GLSL: x < y ? y : x
    return _cw(lambda x, y: y if x < y else x, a, b)


def clamp(x, lo, hi):
    return min_(max_(x, lo), hi)


def abs_(x):
    return vec(np.abs(x.d)) if isinstance(x, vec) else F(abs(x))


def floor(x):
    return vec(np.floor(x.d)) if isinstance(x, vec) else F(math.floor(x))


class Sampler:
    """A float RGBA texture of W x H texels, NEAREST + CLAMP_TO_EDGE."""

    def __init__(self, data, w, h):
        self.w, self.h = int(w), int(h)
        self.data = np.asarray(data, np.float32).reshape(self.h, self.w, 4)

    def fetch(self, uv):
        x = int(math.floor(float(uv.d[0] * F(self.w))))
        y = int(math.floor(float(uv.d[1] * F(self.h))))
        x = 0 if x < 0 else (self.w - 1 if x >= self.w else x)
        y = 0 if y < 0 else (self.h - 1 if y >= self.h else y)
        return vec(self.data[y, x])


def texture(sampler, uv):
    return sampler.fetch(uv)


texture2D = texture   # the renderer #defines texture2D as texture (src/MultiTargetGPUComputationRenderer.js:331)
