#!/usr/bin/env python3
"""Mechanical GLSL ES 3.00 -> Python transpiler for the reference's GPGPU passes (src/SoftbodyGPU.js:59-376).

The WebGL solver's arithmetic lives in seven fragment shaders, written as template strings inside the SoftBodyGPU constructor.
This tool pulls them out of the reference's text together with the wiring around them -- the `addVariable` calls (:49-55),
each `addPass(variable, [dependencies], glsl)` (:59-376) and the static `material.uniforms[...] = { value: ... }` bindings --
and re-emits every shader token for token as a Python function over the types and built-ins of oracle/glslrt.py (vec2/3/4,
mat3, swizzles, value semantics, NEAREST texture fetches; the float model GLSL leaves open is stated there).  Nothing in the
emitted text is hand-written; an unsupported construct aborts.  Together with tools/transpile_reference.py (initPhysics,
simulate, and MultiTargetGPUComputationRenderer's addVariable / addPass / compute, all plain JavaScript) this makes the WHOLE
polar substep executable from the reference's own text: oracle/ref_runner.py::RefSoftBodyGPU.

Output: oracle/_ref/shaders_ref.py (git-ignored).  Called by tools/transpile_reference.py::generate.
"""
from __future__ import annotations

import os
import re

TYPES = {"void", "int", "float", "vec2", "vec3", "vec4", "mat3", "sampler2D", "bool"}
QUALIFIERS = {"out", "in", "uniform", "highp", "mediump", "lowp", "const"}
RENAME = {"float": "float_", "int": "int_", "min": "min_", "max": "max_", "abs": "abs_"}

TOKEN_RE = re.compile(r"""
    (?P<ws>\s+) | (?P<lc>//[^\n]*) | (?P<bc>/\*.*?\*/)
  | (?P<num>(?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+|\d+)
  | (?P<name>[A-Za-z_][A-Za-z0-9_]*)
  | (?P<op>\+\+|--|\+=|-=|\*=|/=|==|!=|<=|>=|&&|\|\||[-+*/%=<>!?:;,.(){}\[\]])
""", re.X | re.S)


class Unsupported(Exception):
    pass


def tokenize(src):
    toks, pos, line = [], 0, 1
    while pos < len(src):
        m = TOKEN_RE.match(src, pos)
        if not m:
            raise Unsupported("GLSL line %d: cannot tokenise %r" % (line, src[pos:pos + 20]))
        if m.lastgroup in ("num", "name", "op"):
            toks.append((m.lastgroup, m.group(), line))
        line += m.group().count("\n")
        pos = m.end()
    toks.append(("eof", "", line))
    return toks


BINARY_PREC = [("||",), ("&&",), ("==", "!="), ("<", ">", "<=", ">="), ("+", "-"), ("*", "/", "%")]


class Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self, k=0):
        return self.t[self.i + k]

    def at(self, text):
        return self.t[self.i][1] == text and self.t[self.i][0] != "num"

    def eat(self, text=None, kind=None):
        tk = self.t[self.i]
        if (text is not None and tk[1] != text) or (kind is not None and tk[0] != kind):
            raise Unsupported("GLSL line %d: expected %r, found %r" % (tk[2], text or kind, tk[1]))
        self.i += 1
        return tk

    # ---- types:  [qualifiers] [layout(...)] type [ '[' n ']' ]
    def at_type(self):
        tk = self.peek()
        return tk[0] == "name" and (tk[1] in TYPES or tk[1] in QUALIFIERS or tk[1] == "layout")

    def parse_type(self):
        quals, layout = [], None
        while True:
            tk = self.peek()
            if tk[1] in QUALIFIERS:
                quals.append(self.eat()[1])
            elif tk[1] == "layout":
                self.eat(); self.eat("(")
                self.eat("location"); self.eat("=")
                layout = int(self.eat(kind="num")[1])
                self.eat(")")
            else:
                break
        base = self.eat(kind="name")[1]
        if base not in TYPES:
            raise Unsupported("GLSL line %d: type %r" % (self.peek()[2], base))
        n = None
        if self.at("["):
            self.eat("["); n = int(self.eat(kind="num")[1]); self.eat("]")
        return (base, n, tuple(quals), layout)

    # ---- translation unit
    def parse_unit(self):
        decls, funcs = [], []
        while self.peek()[0] != "eof":
            ty = self.parse_type()
            name = self.eat(kind="name")[1]
            if self.at("("):
                self.eat("(")
                params = []
                while not self.at(")"):
                    pt = self.parse_type()
                    params.append((pt, self.eat(kind="name")[1]))
                    if self.at(","):
                        self.eat(",")
                self.eat(")")
                funcs.append((ty, name, params, self.parse_block()))
            else:
                while True:
                    n = ty[1]
                    if self.at("["):
                        self.eat("["); n = int(self.eat(kind="num")[1]); self.eat("]")
                    init = None
                    if self.at("="):
                        self.eat("="); init = self.parse_assign()
                    decls.append(((ty[0], n, ty[2], ty[3]), name, init))
                    if self.at(","):
                        self.eat(","); name = self.eat(kind="name")[1]
                        continue
                    break
                self.eat(";")
        return decls, funcs

    # ---- statements
    def parse_block(self):
        self.eat("{")
        out = []
        while not self.at("}"):
            out.append(self.parse_stmt())
        self.eat("}")
        return ("block", out)

    def parse_stmt(self):
        tk = self.peek()
        if self.at("{"):
            return self.parse_block()
        if tk[0] == "name" and tk[1] == "if":
            self.eat(); self.eat("(")
            c = self.parse_expr(); self.eat(")")
            a = self.parse_stmt()
            b = None
            if self.at("else"):
                self.eat(); b = self.parse_stmt()
            return ("if", c, a, b)
        if tk[0] == "name" and tk[1] == "for":
            self.eat(); self.eat("(")
            init = self.parse_simple(); self.eat(";")
            cond = self.parse_expr(); self.eat(";")
            upd = self.parse_expr(); self.eat(")")
            return ("for", init, cond, upd, self.parse_stmt())
        if tk[0] == "name" and tk[1] == "return":
            self.eat()
            e = None if self.at(";") else self.parse_expr()
            self.eat(";")
            return ("return", e)
        if tk[0] == "name" and tk[1] == "break":
            self.eat(); self.eat(";")
            return ("break",)
        if tk[0] == "name" and tk[1] in ("while", "do", "switch", "continue", "discard"):
            raise Unsupported("GLSL line %d: statement %r" % (tk[2], tk[1]))
        s = self.parse_simple()
        self.eat(";")
        return s

    def parse_simple(self):
        if self.at_type() and self.peek(1)[1] != "(":     # a declaration (a constructor call has '(' after the type name)
            ty = self.parse_type()
            out = []
            while True:
                name = self.eat(kind="name")[1]
                n = ty[1]
                if self.at("["):
                    self.eat("["); n = int(self.eat(kind="num")[1]); self.eat("]")
                init = None
                if self.at("="):
                    self.eat("="); init = self.parse_assign()
                out.append(((ty[0], n), name, init))
                if self.at(","):
                    self.eat(",")
                    continue
                break
            return ("decl", out)
        return ("expr", self.parse_expr())

    # ---- expressions
    def parse_expr(self):
        return self.parse_assign()

    def parse_assign(self):
        left = self.parse_cond()
        if self.peek()[0] == "op" and self.peek()[1] in ("=", "+=", "-=", "*=", "/="):
            op = self.eat()[1]
            return ("assign", op, left, self.parse_assign())
        return left

    def parse_cond(self):
        c = self.parse_binary(0)
        if self.at("?"):
            self.eat(); a = self.parse_assign(); self.eat(":"); b = self.parse_assign()
            return ("cond", c, a, b)
        return c

    def parse_binary(self, level):
        if level == len(BINARY_PREC):
            return self.parse_unary()
        left = self.parse_binary(level + 1)
        while self.peek()[0] == "op" and self.peek()[1] in BINARY_PREC[level]:
            op = self.eat()[1]
            left = ("binary", op, left, self.parse_binary(level + 1))
        return left

    def parse_unary(self):
        tk = self.peek()
        if tk[0] == "op" and tk[1] in ("-", "+", "!"):
            self.eat()
            return ("unary", tk[1], self.parse_unary())
        return self.parse_postfix()

    def parse_postfix(self):
        tk = self.eat()
        if tk[0] == "num":
            e = ("num", tk[1])
        elif tk[0] == "name":
            e = ("name", tk[1])
        elif tk[1] == "(":
            e = ("paren", self.parse_expr()); self.eat(")")
        else:
            raise Unsupported("GLSL line %d: unexpected %r" % (tk[2], tk[1]))
        while True:
            if self.at("("):
                self.eat("(")
                args = []
                while not self.at(")"):
                    args.append(self.parse_assign())
                    if self.at(","):
                        self.eat(",")
                self.eat(")")
                e = ("call", e, args)
            elif self.at("["):
                self.eat("["); idx = self.parse_expr(); self.eat("]")
                e = ("index", e, idx)
            elif self.at("."):
                self.eat("."); e = ("member", e, self.eat(kind="name")[1])
            elif self.peek()[1] in ("++", "--") and self.peek()[0] == "op":
                e = ("postfix", self.eat()[1], e)
            else:
                return e


# ------------------------------------------------------------------------------------------ emitter
def default_value(base, n):
    one = {"int": "0", "float": "_f32(0.0)", "bool": "False", "vec2": "vec2(0.0)", "vec3": "vec3(0.0)", "vec4": "vec4(0.0)", "mat3": "mat3(0.0)"}.get(base)
    if one is None:
        raise Unsupported("no default value for type %s" % base)
    return one if n is None else "_arr(%d, lambda: %s)" % (n, one)


class Emitter:
    def __init__(self, globals_):
        self.g = set(globals_)
        self.lines = []

    def out(self, d, text):
        self.lines.append("    " * d + text)

    def ex(self, e, locals_):
        k = e[0]
        if k == "num":
            return e[1] if re.fullmatch(r"\d+", e[1]) else "_f32(%s)" % e[1]
        if k == "name":
            n = e[1]
            if n in locals_:
                return n
            if n in self.g:
                return "g." + n
            return RENAME.get(n, n)      # a function, a constructor or a built-in
        if k == "paren":
            return "(" + self.ex(e[1], locals_) + ")"
        if k == "member":
            return self.ex(e[1], locals_) + "." + e[2]
        if k == "index":
            return "%s[%s]" % (self.ex(e[1], locals_), self.ex(e[2], locals_))
        if k == "call":
            return "%s(%s)" % (self.ex(e[1], locals_), ", ".join("_v(%s)" % self.ex(a, locals_) for a in e[2]))
        if k == "unary":
            return "(not %s)" % self.ex(e[2], locals_) if e[1] == "!" else "(%s%s)" % (e[1], self.ex(e[2], locals_))
        if k == "binary":
            op, l, r = e[1], self.ex(e[2], locals_), self.ex(e[3], locals_)
            if op == "/":
                return "_div(%s, %s)" % (l, r)
            if op == "%":
                return "_mod(%s, %s)" % (l, r)
            if op in ("&&", "||"):
                return "(%s %s %s)" % (l, "and" if op == "&&" else "or", r)
            return "(%s %s %s)" % (l, op, r)
        if k == "cond":
            return "(%s if %s else %s)" % (self.ex(e[2], locals_), self.ex(e[1], locals_), self.ex(e[3], locals_))
        raise Unsupported("GLSL expression %r in value position" % (k,))

    def stmt(self, d, s, locals_):
        k = s[0]
        if k == "block":
            if not s[1]:
                self.out(d, "pass")
            for x in s[1]:
                self.stmt(d, x, locals_)
        elif k == "decl":
            for (base, n), name, init in s[1]:
                locals_.add(name)
                self.out(d, "%s = %s" % (name, "_v(%s)" % self.ex(init, locals_) if init is not None else default_value(base, n)))
        elif k == "expr":
            e = s[1]
            if e[0] == "assign":
                tgt, rhs = self.ex(e[2], locals_), self.ex(e[3], locals_)
                if e[1] == "=":
                    self.out(d, "%s = _v(%s)" % (tgt, rhs))
                elif e[1] == "/=":
                    self.out(d, "%s = _div(%s, %s)" % (tgt, tgt, rhs))
                else:
                    self.out(d, "%s = %s %s %s" % (tgt, tgt, e[1][0], rhs))
            elif e[0] == "postfix":
                t = self.ex(e[2], locals_)
                self.out(d, "%s = %s %s 1" % (t, t, e[1][0]))
            elif e[0] == "call":
                self.out(d, self.ex(e, locals_))
            else:
                raise Unsupported("GLSL expression statement %r" % (e[0],))
        elif k == "if":
            self.out(d, "if %s:" % self.ex(s[1], locals_))
            self.stmt(d + 1, s[2] if s[2][0] == "block" else ("block", [s[2]]), locals_)
            if s[3] is not None:
                self.out(d, "else:")
                self.stmt(d + 1, s[3] if s[3][0] == "block" else ("block", [s[3]]), locals_)
        elif k == "for":
            self.stmt(d, s[1], locals_)
            self.out(d, "while %s:" % self.ex(s[2], locals_))
            self.stmt(d + 1, s[4] if s[4][0] == "block" else ("block", [s[4]]), locals_)
            self.stmt(d + 1, ("expr", s[3]), locals_)
        elif k == "return":
            self.out(d, "return" if s[1] is None else "return _v(%s)" % self.ex(s[1], locals_))
        elif k == "break":
            self.out(d, "break")
        else:
            raise Unsupported("GLSL statement %r" % (k,))


def transpile_shader(name, glsl, injected):
    """One fragment shader -> `def NAME(g):` running main() for the fragment described by g (gl_FragCoord, resolution,
    samplers and uniforms as attributes).  `injected` = sampler names the renderer prepends (dependencies, prev_ ones)."""
    decls, funcs = Parser(tokenize(glsl)).parse_unit()
    globals_ = set(injected) | {"gl_FragCoord", "resolution"} | {n for _, n, _ in decls}
    em = Emitter(globals_)
    em.out(0, "def %s(g):" % name)
    outs = []
    for (base, n, quals, layout), dname, init in decls:
        if "uniform" in quals:
            continue                                   # bound by the harness (material.uniforms)
        if "out" in quals:
            outs.append((layout if layout is not None else 0, dname))
        em.out(1, "g.%s = %s" % (dname, "_v(%s)" % em.ex(init, set()) if init is not None else default_value(base, n)))
    for ty, fname, params, body in funcs:
        em.out(1, "def %s(%s):" % (RENAME.get(fname, fname), ", ".join(p for _, p in params)))
        em.stmt(2, body, {p for _, p in params})
    em.out(1, "main()")
    em.out(1, "return [%s]" % ", ".join("g." + n for _, n in sorted(outs)))
    return "\n".join(em.lines) + "\n", [n for _, n in sorted(outs)], [(n, t[1]) for t, n, _ in decls if "uniform" in t[2]]


# ------------------------------------------------------------------------------------------ extraction from SoftbodyGPU.js
def extract(src):
    variables = [(m.group(1), m.group(2), m.group(3), int(m.group(4)) if m.group(4) else None)
                 for m in re.finditer(r"this\.(\w+)\s*=\s*this\.gpuCompute\.addVariable\(\"(\w+)\"\s*,\s*this\.(\w+)(?:,\s*(\d+))?\)", src)]
    passes = [(m.group(1), m.group(2), [d.strip().replace("this.", "") for d in m.group(3).split(",") if d.strip()], m.group(4),
               src.count("\n", 0, m.start()) + 1)
              for m in re.finditer(r"this\.(\w+)\s*=\s*this\.gpuCompute\.addPass\(this\.(\w+),\s*\[([^\]]*)\],\s*`(.*?)`\)", src, re.S)]
    uniforms = [(m.group(1), m.group(2), m.group(3).strip())
                for m in re.finditer(r"this\.(\w+)\.material\.uniforms\['(\w+)'\s*\]\s*=\s*\{\s*value:\s*(.*?)\s*\}", src[:src.index("initPhysics(density)")])]
    if len(passes) != 7 or len(variables) != 5:
        raise Unsupported("expected 5 variables and 7 passes in SoftbodyGPU.js, found %d and %d" % (len(variables), len(passes)))
    return variables, passes, uniforms


def generate(ref, out):
    src = open(os.path.join(ref, "src", "SoftbodyGPU.js")).read()
    variables, passes, uniforms = extract(src)
    var_tex = {v[0]: (v[1], v[3]) for v in variables}          # JS member -> (texture name, MRT count)
    text = ['"""GENERATED by tools/transpile_shaders.py from src/SoftbodyGPU.js:49-376 -- do not edit, do not commit.',
            'Token-for-token re-emission of the reference\'s GLSL passes; types and built-ins live in oracle/glslrt.py."""',
            "from oracle.glslrt import *  # noqa: F401,F403", "from oracle.glslrt import _v, _div, _mod, _arr, _f32  # noqa: F401", ""]
    table = []
    for pname, var, deps, glsl, line in passes:
        injected = []
        for dvar in deps:                                      # what init() prepends, src/MultiTargetGPUComputationRenderer.js:235-257
            injected.append(var_tex[dvar][0])
            if dvar != var:
                injected.append("prev_" + var_tex[dvar][0])
        code, outs, unis = transpile_shader(pname, glsl, injected)
        text.append("# src/SoftbodyGPU.js:%d  %s -> %s, dependencies %s" % (line, pname, var, deps))
        text.append(code)
        table.append((pname, var, deps, outs, [u for u, _ in unis]))
    text.append("VARIABLES = %r   # (JS member, texture name, initial texture member, MRT count), src/SoftbodyGPU.js:49-55" % (variables,))
    text.append("PASSES = %r   # (pass member, written variable, dependencies, outputs by location, declared uniforms) in addPass order" % (table,))
    text.append("UNIFORM_BINDINGS = %r   # material.uniforms[name] = { value: <expr> } in the constructor" % (uniforms,))
    open(os.path.join(out, "shaders_ref.py"), "w").write("\n".join(text) + "\n")
    return {p[0]: p[4] for p in passes}
