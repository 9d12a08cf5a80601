#!/usr/bin/env python3
"""Mechanical JavaScript -> Python source-to-source transpiler for the reference's solver classes.

    python tools/transpile_reference.py [--ref /root/reference] [--out oracle/_ref]

Why: the reference (zalo/TetSim) is JavaScript and this image has no JS engine, so every parity claim used to rest
on a hand restatement (oracle/softbody_oracle.c).  This tool removes the human from the loop: it tokenises and parses
the reference's OWN text (a small ES subset: classes, methods, let/const, for/if/return/break, assignments, calls,
typed-array indexing, postfix ++) and re-emits it token for token as Python, with every JS-specific arithmetic rule
delegated to oracle/jsrt.py (f32 rounding on typed-array stores, f64 expressions, IEEE division, out-of-range
reads/writes, NaN-propagating Math.min/max).  Nothing in the emitted text is hand-written; an unsupported construct
aborts the run instead of being guessed at.  Output goes to oracle/_ref/ (git-ignored: it is derived from the
reference's source and is regenerated wherever /root/reference exists); the golden vectors made from it
(tools/make_ref_golden.py -> tests/golden/ref_*.npz) are what travels.

Emitted:  oracle/_ref/softbody_ref.py     class SoftBody     <- src/Softbody.js:3-412      (every method)
          oracle/_ref/softbodygpu_ref.py  class SoftBodyGPU  <- src/SoftbodyGPU.js          (initPhysics :487-608, updateVisMesh,
                                          the grab methods and the vec*/mat* helpers it has; the constructor and the GLSL passes
                                          are WebGL objects / shader strings, outside an ES-subset transpile)

Evaluation-order rules applied (ECMA-262 13.15.2 / 13.4.2), the only places where Python's order differs:
  a[i] = rhs        ->  _t = i ; a[_t] = rhs            (JS evaluates the target's subscript BEFORE rhs)
  a[i] op= rhs      ->  _t = i ; a[_t] = a[_t] op rhs   (old value read before rhs)
  v++ (in an expr)  ->  ((v := v + 1) - 1)
  for (init; c; u) body -> init ; while c: body ; u     (`continue` is rejected)
"""
from __future__ import annotations

import argparse
import keyword
import os
import re
import sys

# ------------------------------------------------------------------------------------------ tokenizer
TOKEN_RE = re.compile(r"""
    (?P<ws>\s+)
  | (?P<lc>//[^\n]*)
  | (?P<bc>/\*.*?\*/)
  | (?P<num>0[xX][0-9a-fA-F]+|(?:\d+\.\d*|\.\d+|\d+)(?:[eE][+-]?\d+)?)
  | (?P<name>[A-Za-z_$][A-Za-z0-9_$]*)
  | (?P<str>'(?:[^'\\\n]|\\.)*'|"(?:[^"\\\n]|\\.)*")
  | (?P<op>===|!==|\+\+|--|\+=|-=|\*=|/=|==|!=|<=|>=|&&|\|\||=>|[-+*/%=<>!?:;,.(){}\[\]])
""", re.X | re.S)


class Unsupported(Exception):
    pass


def tokenize(src, line0=1):
    toks, pos, line = [], 0, line0
    while pos < len(src):
        m = TOKEN_RE.match(src, pos)
        if not m:
            raise Unsupported("line %d: cannot tokenise %r" % (line, src[pos:pos + 20]))
        kind = m.lastgroup
        text = m.group()
        if kind in ("num", "name", "str", "op"):
            toks.append((kind, text, line))
        line += text.count("\n")
        pos = m.end()
    toks.append(("eof", "", line))
    return toks


# ------------------------------------------------------------------------------------------ parser
BINARY_PREC = [("||",), ("&&",), ("==", "!=", "===", "!=="), ("<", ">", "<=", ">="), ("+", "-"), ("*", "/", "%")]
ASSIGN_OPS = ("=", "+=", "-=", "*=", "/=")


class Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self, k=0):
        return self.t[self.i + k]

    def at(self, text):
        return self.t[self.i][1] == text and self.t[self.i][0] in ("op", "name")

    def eat(self, text=None, kind=None):
        tk = self.t[self.i]
        if (text is not None and tk[1] != text) or (kind is not None and tk[0] != kind):
            raise Unsupported("line %d: expected %r, found %r" % (tk[2], text or kind, tk[1]))
        self.i += 1
        return tk

    # ---- class ----
    def parse_class_body(self):
        """After `class X {`: a list of (name, params, body, line)."""
        methods = []
        while not self.at("}"):
            name = self.eat(kind="name")
            self.eat("(")
            params = []
            while not self.at(")"):
                p = self.eat(kind="name")[1]
                default = None
                if self.at("="):
                    self.eat("=")
                    default = self.parse_assign()
                params.append((p, default))
                if self.at(","):
                    self.eat(",")
            self.eat(")")
            body = self.parse_block()
            methods.append((name[1], params, body, name[2]))
        return methods

    # ---- statements ----
    def parse_block(self):
        self.eat("{")
        out = []
        while not self.at("}"):
            out.append(self.parse_stmt())
        self.eat("}")
        return ("block", out)

    def parse_stmt(self):
        tk = self.peek()
        if tk[1] == "{" and tk[0] == "op":
            return self.parse_block()
        if tk[0] == "name" and tk[1] in ("let", "const", "var"):
            s = self.parse_decl()
            self.semi()
            return s
        if tk[0] == "name" and tk[1] == "if":
            self.eat("if"); self.eat("(")
            c = self.parse_expr()
            self.eat(")")
            a = self.parse_stmt()
            b = None
            if self.at("else"):
                self.eat("else")
                b = self.parse_stmt()
            return ("if", c, a, b)
        if tk[0] == "name" and tk[1] == "for":
            self.eat("for"); self.eat("(")
            init = None
            if not self.at(";"):
                init = self.parse_decl() if self.peek()[1] in ("let", "const", "var") else ("expr", self.parse_expr())
            self.eat(";")
            cond = None if self.at(";") else self.parse_expr()
            self.eat(";")
            upd = None if self.at(")") else self.parse_expr()
            self.eat(")")
            return ("for", init, cond, upd, self.parse_stmt())
        if tk[0] == "name" and tk[1] == "return":
            self.eat("return")
            e = None if self.at(";") or self.at("}") else self.parse_expr()
            self.semi()
            return ("return", e)
        if tk[0] == "name" and tk[1] == "break":
            self.eat("break"); self.semi()
            return ("break",)
        if tk[0] == "name" and tk[1] in ("continue", "while", "do", "switch", "try", "throw", "function", "class"):
            raise Unsupported("line %d: statement %r is outside the supported subset" % (tk[2], tk[1]))
        if tk[1] == ";" and tk[0] == "op":
            self.eat(";")
            return ("block", [])
        e = self.parse_expr()
        self.semi()
        return ("expr", e)

    def semi(self):
        # automatic semicolon insertion is honoured only where the reference relies on it: before `}` or a new line
        if self.at(";"):
            self.eat(";")
        elif self.at("}") or self.peek()[2] > self.t[self.i - 1][2]:
            pass
        else:
            tk = self.peek()
            raise Unsupported("line %d: expected ';' before %r" % (tk[2], tk[1]))

    def parse_decl(self):
        self.eat(kind="name")  # let / const / var
        decls = []
        while True:
            n = self.eat(kind="name")[1]
            init = None
            if self.at("="):
                self.eat("=")
                init = self.parse_assign()
            decls.append((n, init))
            if self.at(","):
                self.eat(",")
                continue
            break
        return ("let", decls)

    # ---- expressions ----
    def parse_expr(self):
        e = self.parse_assign()
        if self.at(","):
            raise Unsupported("line %d: comma operator" % self.peek()[2])
        return e

    def parse_assign(self):
        left = self.parse_cond()
        if self.peek()[0] == "op" and self.peek()[1] in ASSIGN_OPS:
            op = self.eat()[1]
            right = self.parse_assign()
            if left[0] not in ("name", "member", "index"):
                raise Unsupported("assignment to a non-reference")
            return ("assign", op, left, right)
        return left

    def parse_cond(self):
        c = self.parse_binary(0)
        if self.at("?"):
            self.eat("?")
            a = self.parse_assign()
            self.eat(":")
            b = self.parse_assign()
            return ("cond", c, a, b)
        return c

    def parse_binary(self, level):
        if level == len(BINARY_PREC):
            return self.parse_unary()
        left = self.parse_binary(level + 1)
        while self.peek()[0] == "op" and self.peek()[1] in BINARY_PREC[level]:
            op = self.eat()[1]
            right = self.parse_binary(level + 1)
            left = ("binary", op, left, right)
        return left

    def parse_unary(self):
        tk = self.peek()
        if tk[0] == "op" and tk[1] in ("-", "+", "!"):
            self.eat()
            return ("unary", tk[1], self.parse_unary())
        if tk[0] == "op" and tk[1] in ("++", "--"):
            raise Unsupported("line %d: prefix %s" % (tk[2], tk[1]))
        return self.parse_postfix()

    def parse_postfix(self):
        e = self.parse_call()
        tk = self.peek()
        if tk[0] == "op" and tk[1] in ("++", "--") and tk[2] == self.t[self.i - 1][2]:
            self.eat()
            return ("postfix", tk[1], e)
        return e

    def parse_call(self):
        tk = self.peek()
        if tk[0] == "name" and tk[1] == "new":
            self.eat("new")
            callee = self.parse_member_only()
            args = self.parse_args() if self.at("(") else []
            e = ("new", callee, args)
        else:
            e = self.parse_primary()
        while True:
            if self.at("."):
                self.eat(".")
                e = ("member", e, self.eat(kind="name")[1])
            elif self.at("["):
                self.eat("[")
                idx = self.parse_expr()
                self.eat("]")
                e = ("index", e, idx)
            elif self.at("("):
                e = ("call", e, self.parse_args())
            else:
                return e

    def parse_member_only(self):
        e = ("name", self.eat(kind="name")[1])
        while self.at("."):
            self.eat(".")
            e = ("member", e, self.eat(kind="name")[1])
        return e

    def parse_args(self):
        self.eat("(")
        args = []
        while not self.at(")"):
            args.append(self.parse_assign())
            if self.at(","):
                self.eat(",")
        self.eat(")")
        return args

    def parse_primary(self):
        tk = self.eat()
        if tk[0] == "num":
            return ("num", tk[1])
        if tk[0] == "str":
            return ("str", tk[1])
        if tk[0] == "name":
            if tk[1] == "this":
                return ("this",)
            if tk[1] in ("true", "false", "null", "undefined"):
                return ("lit", tk[1])
            if tk[1] in ("function", "typeof", "delete", "void", "await", "yield"):
                raise Unsupported("line %d: %r" % (tk[2], tk[1]))
            return ("name", tk[1])
        if tk[1] == "(":
            e = self.parse_expr()
            self.eat(")")
            return ("paren", e)
        if tk[1] == "[":
            elems = []
            while not self.at("]"):
                elems.append(self.parse_assign())
                if self.at(","):
                    self.eat(",")
            self.eat("]")
            return ("array", elems)
        if tk[1] == "{":          # object literal { key: expr, ... } (only reachable in expression position)
            props = []
            while not self.at("}"):
                k = self.eat()
                if k[0] not in ("name", "str"):
                    raise Unsupported("line %d: object key %r" % (k[2], k[1]))
                self.eat(":")
                props.append((k[1] if k[0] == "name" else k[1][1:-1], self.parse_assign()))
                if self.at(","):
                    self.eat(",")
            self.eat("}")
            return ("object", props)
        raise Unsupported("line %d: unexpected %r" % (tk[2], tk[1]))


# ------------------------------------------------------------------------------------------ emitter
PY_RESERVED = set(keyword.kwlist) | {"self", "print"}
LITS = {"true": "True", "false": "False", "null": "None", "undefined": "undefined"}
CMP = {"==": "==", "===": "==", "!=": "!=", "!==": "!=", "<": "<", ">": ">", "<=": "<=", ">=": ">="}


def ident(n):
    n = n.replace("$", "_S_")
    return n + "_" if n in PY_RESERVED else n


def has_postfix(e):
    if not isinstance(e, tuple):
        return False
    if e[0] == "postfix":
        return True
    return any(has_postfix(x) if isinstance(x, tuple) else any(has_postfix(y) for y in x) if isinstance(x, list) else False
               for x in e[1:])


class Emitter:
    def __init__(self):
        self.lines = []
        self.tmp = 0

    def out(self, depth, text):
        self.lines.append("    " * depth + text)

    # ---- expressions ----
    def ex(self, e):
        k = e[0]
        if k == "num":
            return e[1]
        if k == "str":
            return repr(bytes(e[1][1:-1], "utf-8").decode("unicode_escape"))
        if k == "lit":
            return LITS[e[1]]
        if k == "name":
            return ident(e[1])
        if k == "this":
            return "self"
        if k == "paren":
            return "(" + self.ex(e[1]) + ")"
        if k == "member":
            return self.ex(e[1]) + "." + ident(e[2])
        if k == "index":
            return self.ex(e[1]) + "[" + self.ex(e[2]) + "]"
        if k == "call":
            return self.ex(e[1]) + "(" + ", ".join(self.ex(a) for a in e[2]) + ")"
        if k == "new":
            return self.ex(e[1]) + "(" + ", ".join(self.ex(a) for a in e[2]) + ")"
        if k == "array":
            return "JSArray([" + ", ".join(self.ex(a) for a in e[1]) + "])"
        if k == "object":
            return "JSObject(" + ", ".join("%s=%s" % (ident(n), self.ex(v)) for n, v in e[1]) + ")"
        if k == "unary":
            if e[1] == "!":
                return "(not _truthy(" + self.ex(e[2]) + "))"
            return "(" + e[1] + self.ex(e[2]) + ")"
        if k == "postfix":
            if e[2][0] != "name":
                raise Unsupported("postfix %s on a non-variable" % e[1])
            v, sign, inv = ident(e[2][1]), e[1][0], "-" if e[1] == "++" else "+"
            return "((%s := %s %s 1) %s 1)" % (v, v, sign, inv)
        if k == "binary":
            op, l, r = e[1], self.ex(e[2]), self.ex(e[3])
            if op == "/":
                return "_div(%s, %s)" % (l, r)
            if op == "%":
                return "_mod(%s, %s)" % (l, r)
            if op == "&&":   # JS truthiness (NaN / undefined are falsy), operand values returned like JS does
                t = self.new_tmp()
                return "(%s if _truthy(%s := %s) else %s)" % (r, t, l, t)
            if op == "||":
                t = self.new_tmp()
                return "(%s if _truthy(%s := %s) else %s)" % (t, t, l, r)
            if op in CMP:
                return "(%s %s %s)" % (l, CMP[op], r)
            return "(%s %s %s)" % (l, op, r)   # + - * : left-associative tree, parenthesised as parsed
        if k == "cond":
            return "(%s if _truthy(%s) else %s)" % (self.ex(e[2]), self.ex(e[1]), self.ex(e[3]))
        if k == "assign":
            raise Unsupported("assignment used as a value")
        raise Unsupported("expression kind %r" % (k,))

    def new_tmp(self):
        self.tmp += 1
        return "_t%d" % self.tmp

    def assign(self, d, e):
        _, op, target, value = e
        if target[0] == "index":
            if has_postfix(target[1]):
                raise Unsupported("side effect in the object of an indexed assignment")
            t = self.new_tmp()
            self.out(d, "%s = %s" % (t, self.ex(target[2])))       # subscript first (JS order)
            ref = "%s[%s]" % (self.ex(target[1]), t)
        elif target[0] == "member":
            if has_postfix(target[1]):
                raise Unsupported("side effect in the object of a member assignment")
            ref = self.ex(target)
        else:
            ref = ident(target[1])
        rhs = self.ex(value)
        if op == "=":
            self.out(d, "%s = %s" % (ref, rhs))
        elif op == "/=":
            self.out(d, "%s = _div(%s, %s)" % (ref, ref, rhs))
        else:
            self.out(d, "%s = %s %s %s" % (ref, ref, op[0], rhs))  # old value read before rhs is evaluated

    # ---- statements ----
    def stmt(self, d, s, scope):
        k = s[0]
        if k == "block":
            inner = dict(scope)
            if not s[1]:
                self.out(d, "pass")
            for x in s[1]:
                self.stmt(d, x, inner)
        elif k == "let":
            for n, init in s[1]:
                if scope.get(n) == "outer":
                    raise Unsupported("`let %s` shadows a variable of an enclosing block" % n)
                scope[n] = "here"
                if init is not None and init[0] == "assign":
                    raise Unsupported("chained assignment in a declaration")
                self.out(d, "%s = %s" % (ident(n), self.ex(init) if init is not None else "undefined"))
        elif k == "expr":
            e = s[1]
            if e[0] == "assign":
                self.assign(d, e)
            elif e[0] == "postfix":
                if e[2][0] != "name":
                    raise Unsupported("postfix on a non-variable")
                self.out(d, "%s = %s %s 1" % (ident(e[2][1]), ident(e[2][1]), e[1][0]))
            else:
                self.out(d, self.ex(e))
        elif k == "if":
            self.out(d, "if _truthy(%s):" % self.ex(s[1]))
            self.body(d + 1, s[2], scope)
            if s[3] is not None:
                self.out(d, "else:")
                self.body(d + 1, s[3], scope)
        elif k == "for":
            inner = {n: "outer" for n in scope}
            if s[1] is not None:
                self.stmt(d, s[1], inner)
            self.out(d, "while %s:" % ("_truthy(%s)" % self.ex(s[2]) if s[2] is not None else "True"))
            self.body(d + 1, s[4], inner)
            if s[3] is not None:
                self.stmt(d + 1, ("expr", s[3]), inner)
        elif k == "return":
            self.out(d, "return" if s[1] is None else "return " + self.ex(s[1]))
        elif k == "break":
            self.out(d, "break")
        else:
            raise Unsupported("statement kind %r" % (k,))

    def body(self, d, s, scope):
        inner = {n: "outer" for n in scope}
        if s[0] == "block":
            if not s[1]:
                self.out(d, "pass")
            for x in s[1]:
                self.stmt(d, x, inner)
        else:
            self.stmt(d, s, inner)

    def method(self, name, params, body):
        ps = ["self"]
        for p, default in params:
            ps.append(ident(p) if default is None else "%s=%s" % (ident(p), self.ex(default)))
        pyname = "__init__" if name == "constructor" else ident(name)
        self.out(1, "def %s(%s):" % (pyname, ", ".join(ps)))
        scope = {p: "here" for p, _ in params}
        n0 = len(self.lines)
        for x in body[1]:
            self.stmt(2, x, scope)
        if len(self.lines) == n0:
            self.out(2, "pass")
        self.out(0, "")


# ------------------------------------------------------------------------------------------ driver
def class_span(src, cls):
    """(text between the braces of `class cls { ... }`, line number of its first character)."""
    m = re.search(r"\bclass\s+%s\b[^{]*\{" % re.escape(cls), src)
    if not m:
        raise Unsupported("class %s not found" % cls)
    depth, i, n = 1, m.end(), len(src)
    in_str = None
    while i < n and depth:
        c = src[i]
        if in_str:
            if c == "\\":
                i += 1
            elif c == in_str:
                in_str = None
        elif c in "'\"`":
            in_str = c
        elif src.startswith("//", i):
            i = src.index("\n", i)
            continue
        elif src.startswith("/*", i):
            i = src.index("*/", i) + 1
        elif c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
        i += 1
    return src[m.end():i - 1], src.count("\n", 0, m.end()) + 1


def method_spans(body, line0):
    """Split a class body into {name: (text, first line)} without parsing method bodies (some hold GLSL strings)."""
    out, i, n = {}, 0, len(body)
    head = re.compile(r"\s*(?://[^\n]*\n\s*|/\*.*?\*/\s*)*([A-Za-z_$][\w$]*)\s*\(", re.S)
    while True:
        m = head.match(body, i)
        if not m:
            break
        name, start = m.group(1), m.start(1)
        # skip the parameter list (may contain defaults with parentheses)
        depth, p = 1, m.end()
        while depth:
            depth += {"(": 1, ")": -1}.get(body[p], 0)
            p += 1
        j = body.index("{", p)
        depth, k, in_str = 1, j + 1, None
        while depth:
            c = body[k]
            if in_str:
                if c == "\\":
                    k += 1
                elif c == in_str:
                    in_str = None
            elif c in "'\"`":
                in_str = c
            elif body.startswith("//", k):
                k = body.index("\n", k)
                continue
            elif body.startswith("/*", k):
                k = body.index("*/", k) + 1
            elif c == "{":
                depth += 1
            elif c == "}":
                depth -= 1
            k += 1
        out[name] = (body[start:k], line0 + body.count("\n", 0, start))
        i = k
    return out


def function_members(src, names):
    """`this.NAME = function (args) { body }` inside a constructor (src/MultiTargetGPUComputationRenderer.js) -> {NAME: (method text, line)}."""
    out = {}
    for n in names:
        m = re.search(r"this\.%s\s*=\s*function\s*\(([^)]*)\)\s*\{" % re.escape(n), src)
        if not m:
            raise Unsupported("function member %s not found" % n)
        depth, k = 1, m.end()
        while depth:
            depth += {"{": 1, "}": -1}.get(src[k], 0)
            k += 1
        out[n] = ("%s(%s) {%s" % (n, m.group(1), src[m.end():k]), src.count("\n", 0, m.start()) + 1)
    return out


def transpile_members(path, cls, names, rel=None):
    src = open(path).read()
    spans = function_members(src, names)
    em = Emitter()
    em.out(0, "class %s:" % cls)
    for n in names:
        text, ln = spans[n]
        methods = Parser(tokenize(text + "}", ln)).parse_class_body()
        em.out(1, "# %s:%d" % (rel or path, ln))
        em.method(*methods[0][:3])
    return "\n".join(em.lines), {n: spans[n][1] for n in names}


def transpile_class(path, cls, only=None, rel=None):
    src = open(path).read()
    body, line0 = class_span(src, cls)
    spans = method_spans(body, line0)
    names = [n for n in spans if only is None or n in only]
    if only:
        missing = [n for n in only if n not in spans]
        if missing:
            raise Unsupported("%s: methods not found: %s" % (cls, missing))
    em = Emitter()
    em.out(0, "class %s:" % cls)
    for n in names:
        text, ln = spans[n]
        methods = Parser(tokenize(text + "}", ln)).parse_class_body()
        assert len(methods) == 1 and methods[0][0] == n, (n, [m[0] for m in methods])
        em.out(1, "# %s:%d" % (rel or path, ln))
        em.method(*methods[0][:3])
    return "\n".join(em.lines), {n: spans[n][1] for n in names}


HEADER = '''"""GENERATED by tools/transpile_reference.py from %s -- do not edit, do not commit.
Token-for-token re-emission of the reference's own text; JS semantics live in oracle/jsrt.py."""
from oracle.jsrt import *  # noqa: F401,F403
from oracle.jsrt import _div, _mod, _truthy  # noqa: F401

'''

GPU_METHODS = ["initPhysics", "simulate", "updateVisMesh", "startGrab", "moveGrabbed", "endGrab", "vecSetZero", "vecCopy", "vecAdd",
               "vecSetDiff", "vecDistSquared", "matGetDeterminant"]


def generate(ref="/root/reference", out=None):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = out or os.path.join(root, "oracle", "_ref")
    os.makedirs(out, exist_ok=True)
    open(os.path.join(out, "__init__.py"), "w").write('"""Generated from the reference; git-ignored (see tools/transpile_reference.py)."""\n')
    report = {}
    text, lines = transpile_class(os.path.join(ref, "src", "Softbody.js"), "SoftBody", None, "src/Softbody.js")
    open(os.path.join(out, "softbody_ref.py"), "w").write(HEADER % "src/Softbody.js (class SoftBody, every method)" + text + "\n")
    report["SoftBody"] = lines
    text, lines = transpile_class(os.path.join(ref, "src", "SoftbodyGPU.js"), "SoftBodyGPU", GPU_METHODS, "src/SoftbodyGPU.js")
    open(os.path.join(out, "softbodygpu_ref.py"), "w").write(HEADER % "src/SoftbodyGPU.js (class SoftBodyGPU: initPhysics, grab, vec*/mat*)" + text + "\n")
    report["SoftBodyGPU"] = lines
    text, lines = transpile_members(os.path.join(ref, "src", "MultiTargetGPUComputationRenderer.js"), "MultiTargetGPUComputationRenderer",
                                    ["addVariable", "addPass", "compute", "getCurrentRenderTarget"], "src/MultiTargetGPUComputationRenderer.js")
    open(os.path.join(out, "gpucompute_ref.py"), "w").write(HEADER % "src/MultiTargetGPUComputationRenderer.js (addVariable, addPass, compute, getCurrentRenderTarget)" + text + "\n")
    report["MultiTargetGPUComputationRenderer"] = lines
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import transpile_shaders
    report["GLSL passes"] = transpile_shaders.generate(ref, out)
    return out, report


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    if not os.path.isdir(a.ref):
        sys.exit("transpile_reference: %s does not exist (the reference is only present in the build container)" % a.ref)
    out, report = generate(a.ref, a.out)
    for cls, lines in report.items():
        print("%s: %d methods -> %s" % (cls, len(lines), out))
        print("   " + ", ".join("%s:%d" % kv for kv in lines.items()))


if __name__ == "__main__":
    main()
