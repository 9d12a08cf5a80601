"""JavaScript run-time semantics for the mechanically transpiled reference -- TEST INFRASTRUCTURE.

``tools/transpile_reference.py`` rewrites classes of ``/root/reference/src/*.js`` into Python source
(``oracle/_ref/*.py``, git-ignored: derived from the reference's text, never committed).  The emitted
code is token-for-token the reference's; everything that makes JavaScript arithmetic different from
Python's lives here, so that the emitted text needs no per-line judgement:

* ``Float32Array``  -- reads widen f32 -> f64 (a Python float), writes round f64 -> f32 once (RNE, the C
  conversion ``array('f')`` performs), out-of-range reads give ``undefined`` (NaN in arithmetic),
  out-of-range writes are dropped, an index may be an integral float (every JS number is a double).
* ``JSArray``       -- a plain JS ``Array`` (``tetIds`` is one, src/Dragon.js:311): values unrounded.
* ``_div/_mod``     -- IEEE division (x/0 = +-Infinity, 0/0 = NaN) instead of ZeroDivisionError.
* ``Math, Number, console`` -- the members the reference uses, with JS NaN propagation in min/max.
* ``THREE``         -- inert stand-ins for the three.js scene objects the constructors touch
  (BufferGeometry / BufferAttribute / LineSegments / Mesh); ``computeVertexNormals`` and
  ``computeBoundingSphere`` do nothing (rendering side, not on the hot path).

Python floats are IEEE binary64 evaluated operation by operation with no contraction, which is
exactly the JS ``number`` arithmetic of SURVEY.md App. A.
"""
from __future__ import annotations

import array as _array
import math as _math
import sys as _sys

NaN = float("nan")
undefined = NaN  # only ever used arithmetically by the reference code in scope
null = None


def _div(a, b):
    try:
        return a / b
    except ZeroDivisionError:
        if a != a or a == 0:
            return NaN
        neg = (_math.copysign(1.0, a) < 0) != (_math.copysign(1.0, b) < 0)
        return -_math.inf if neg else _math.inf


def _mod(a, b):
    try:
        return _math.fmod(a, b)
    except (ZeroDivisionError, ValueError):
        return NaN


def _truthy(x):
    """JS ToBoolean for the values in play: false, 0, NaN (= undefined here), null and '' are falsy."""
    if x is None or x is False:
        return False
    if isinstance(x, float):
        return x == x and x != 0.0
    if isinstance(x, (int, str)):
        return bool(x)
    return True


class _Indexable:
    __slots__ = ("_a",)

    @property
    def length(self):
        return len(self._a)

    def __len__(self):
        return len(self._a)

    def _index(self, i):
        """JS property lookup by number: integral doubles address elements, anything else is a miss."""
        if i.__class__ is not int:
            j = int(i) if i == i and abs(i) != _math.inf else -1
            if j != i:
                return -1
            i = j
        return i if 0 <= i < len(self._a) else -1

    def __getitem__(self, i):
        if i.__class__ is int and 0 <= i:
            try:
                return self._a[i]
            except IndexError:
                return undefined
        j = self._index(i)
        return self._a[j] if j >= 0 else undefined

    def __iter__(self):
        return iter(self._a)


class Float32Array(_Indexable):
    """new Float32Array(length | array-like)."""
    __slots__ = ()

    def __init__(self, src=0):
        if isinstance(src, (int, float)):
            self._a = _array.array("f", bytes(4 * int(src)))
        else:
            self._a = _array.array("f", [float(x) for x in src])

    def __setitem__(self, i, v):
        if i.__class__ is int and 0 <= i:
            try:
                self._a[i] = v  # C double -> float conversion: one rounding, to nearest even; overflow -> inf
            except IndexError:
                pass
            return
        j = self._index(i)
        if j >= 0:
            self._a[j] = v

    def slice(self, begin=0, end=None):
        out = Float32Array(0)
        out._a = self._a[int(begin):] if end is None else self._a[int(begin):int(end)]
        return out

    def tobytes(self):
        return self._a.tobytes()


class JSArray(_Indexable):
    """A plain JS Array literal / Array of numbers (no rounding on store, grows on write past the end).
    Like every JS object it accepts ad-hoc properties (src/SoftbodyGPU.js:591 sets .needsUpdate on an Array)."""

    def __init__(self, src=()):
        self._a = list(src)

    def __setitem__(self, i, v):
        j = int(i)
        if j != i or j < 0:
            return
        while len(self._a) <= j:
            self._a.append(undefined)
        self._a[j] = v

    def slice(self, begin=0, end=None):
        return JSArray(self._a[int(begin):] if end is None else self._a[int(begin):int(end)])

    def push(self, *items):
        self._a.extend(items)
        return len(self._a)


class JSObject:
    """A plain JS object: property access by name, missing properties are undefined."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def __getattr__(self, name):  # only reached for missing attributes
        if name.startswith("__"):
            raise AttributeError(name)
        return undefined

    def __getitem__(self, key):   # obj['name'] and obj["a" + b]
        return getattr(self, str(key))

    def __setitem__(self, key, value):
        setattr(self, str(key), value)


class Math:
    PI = _math.pi

    @staticmethod
    def sqrt(x):
        return _math.sqrt(x) if x >= 0 else NaN  # NaN compares false -> NaN

    @staticmethod
    def min(*xs):
        r = _math.inf
        for x in xs:
            if x != x:
                return NaN
            if x < r or (x == 0 and r == 0 and _math.copysign(1.0, x) < 0):
                r = x
        return r

    @staticmethod
    def max(*xs):
        r = -_math.inf
        for x in xs:
            if x != x:
                return NaN
            if x > r or (x == 0 and r == 0 and _math.copysign(1.0, x) > 0):
                r = x
        return r

    @staticmethod
    def ceil(x):
        return float(_math.ceil(x)) if x == x and abs(x) != _math.inf else x

    @staticmethod
    def floor(x):
        return float(_math.floor(x)) if x == x and abs(x) != _math.inf else x

    @staticmethod
    def abs(x):
        return abs(x)


class Number:
    MAX_VALUE = _sys.float_info.max


class console:
    lines = []

    @staticmethod
    def log(*a):
        console.lines.append(a)

    error = log


# ---- inert three.js stand-ins (rendering side; only what the solver classes' constructors touch) ----
class _Layers:
    def enable(self, n):
        pass


class _BufferAttribute:
    def __init__(self, arr, itemSize):
        self.array = arr  # the reference aliases the caller's `vertices` here (src/Softbody.js:37)
        self.itemSize = itemSize
        self.needsUpdate = False


class _BufferGeometry:
    def __init__(self):
        self.attributes = JSObject()
        self.index = None

    def setAttribute(self, name, attr):
        setattr(self.attributes, name, attr)

    def setIndex(self, idx):
        self.index = idx

    def computeVertexNormals(self):
        pass

    def computeBoundingSphere(self):
        pass


class _Object3D:
    def __init__(self, geometry=None, material=None):
        self.geometry = geometry
        self.material = material
        self.userData = None
        self.visible = True
        self.castShadow = False
        self.layers = _Layers()


class _Vector3:
    def __init__(self, x=0.0, y=0.0, z=0.0):
        self.x, self.y, self.z = x, y, z


NearestFilter = 1003   # three.js constant; only stored by addVariable


class THREE:
    Vector3 = _Vector3
    BufferGeometry = _BufferGeometry
    BufferAttribute = _BufferAttribute
    LineSegments = _Object3D
    Mesh = _Object3D
