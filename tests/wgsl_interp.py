"""A small WGSL interpreter -- TEST INFRASTRUCTURE -- that executes the reference's shader SOURCE TEXT.

Why: the reference (Rust + wgpu) cannot run in this image (no rustc, no Vulkan / lavapipe, no naga), and it ships no
tests or golden vectors.  The C oracle (oracle/slime_oracle.c) is a hand restatement of
/root/reference/src/compute.wgsl; this module removes "the restatement misread the shader" as a risk by running the
shader file itself: tokenizer -> parser -> tree-walking evaluator for the subset of WGSL that compute.wgsl and
display.wgsl use, with WGSL's typing rules (abstract literals concretised to i32 / f32, f32 arithmetic in IEEE binary32
round-to-nearest through numpy scalars, truncating integer division and remainder, out-of-order module declarations).

Nothing in here knows what the shaders compute.  What the WGSL specification leaves open is passed in:
  * `sin`, `cos` (accuracy is implementation-defined in WGSL) and float `%` come from a `Builtins` object -- numpy's
    libm by default ("some conforming backend"), or the arithmetic spec of DESIGN.md section 2 (the oracle's sincos and
    exact fmod) when a bit-for-bit comparison is wanted;
  * the SCHEDULE of the invocations of a dispatch (the shaders race by design):
      "sequential": invocation i completes before i + 1 starts, storage is live (the order-fixed reference run);
      "lockstep":   every load of the dispatch sees the storage as it was when the dispatch started, stores go to the
                    live buffer (an infinitely wide SIMD machine -- the schedule the engine's phase_split semantics
                    reproduce when dep >= 1, and the Jacobi reading of diffuse_trail).

The interpreter is slow (about a millisecond per agent): golden vectors are small (tests/golden/make_wgsl_golden.py).
"""
from __future__ import annotations

import math
import re

import numpy as np

F32, I32, U32 = np.float32, np.int32, np.uint32


# ----------------------------------------------------------------------------------------------------------------------
# tokens
# ----------------------------------------------------------------------------------------------------------------------
_TOKEN = re.compile(r"""
    (?P<ws>\s+|//[^\n]*|/\*.*?\*/)
  | (?P<float>(?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?[fh]?|\d+[eE][+-]?\d+[fh]?|\d+f)
  | (?P<int>0[xX][0-9a-fA-F]+[iu]?|\d+[iu]?)
  | (?P<ident>[A-Za-z_][A-Za-z0-9_]*)
  | (?P<op>\+\+|--|\+=|-=|\*=|/=|%=|\|\||&&|<=|>=|==|!=|->|[-+*/%<>=!&|^~(){}\[\],;:.@])
""", re.X | re.S)


def tokenize(src):
    out, pos = [], 0
    while pos < len(src):
        m = _TOKEN.match(src, pos)
        if not m:
            raise SyntaxError(f"WGSL: cannot tokenise at {src[pos:pos + 30]!r}")
        pos = m.end()
        kind = m.lastgroup
        if kind != "ws":
            out.append((kind, m.group()))
    out.append(("eof", ""))
    return out


# ----------------------------------------------------------------------------------------------------------------------
# parser (AST = nested tuples)
# ----------------------------------------------------------------------------------------------------------------------
_TEMPLATED = {"vec2", "vec3", "vec4", "array", "texture_storage_2d", "ptr", "atomic"}
_BINARY_LEVELS = [("||",), ("&&",), ("|",), ("^",), ("&",), ("==", "!="), ("<", ">", "<=", ">="), ("+", "-"), ("*", "/", "%")]


class Parser:
    def __init__(self, src):
        self.t = tokenize(src)
        self.i = 0

    def peek(self, k=0):
        return self.t[self.i + k]

    def next(self):
        tok = self.t[self.i]
        self.i += 1
        return tok

    def accept(self, text):
        if self.t[self.i][1] == text and self.t[self.i][0] in ("op", "ident"):
            self.i += 1
            return True
        return False

    def expect(self, text):
        if not self.accept(text):
            raise SyntaxError(f"WGSL: expected {text!r}, found {self.t[self.i][1]!r} (token {self.i})")

    def ident(self):
        kind, text = self.next()
        if kind != "ident":
            raise SyntaxError(f"WGSL: expected an identifier, found {text!r}")
        return text

    # ---- module ----
    def module(self):
        decls = []
        while self.peek()[0] != "eof":
            if self.accept(";"):
                continue
            attrs = self.attributes()
            kind, text = self.peek()
            if text == "const":
                self.next()
                name = self.ident()
                ty = self.type() if self.accept(":") else None
                self.expect("=")
                decls.append(("const", name, ty, self.expr()))
                self.expect(";")
            elif text == "var":
                self.next()
                space = None
                if self.accept("<"):
                    space = [self.ident()]
                    while self.accept(","):
                        space.append(self.ident())
                    self.expect(">")
                name = self.ident()
                self.expect(":")
                decls.append(("gvar", name, space, self.type(), attrs))
                self.expect(";")
            elif text == "struct":
                self.next()
                name = self.ident()
                self.expect("{")
                members = []
                while not self.accept("}"):
                    self.attributes()
                    m = self.ident()
                    self.expect(":")
                    members.append((m, self.type()))
                    self.accept(",")
                decls.append(("struct", name, members))
            elif text == "fn":
                self.next()
                name = self.ident()
                self.expect("(")
                params = []
                while not self.accept(")"):
                    pattrs = self.attributes()
                    p = self.ident()
                    self.expect(":")
                    params.append((p, self.type(), pattrs))
                    self.accept(",")
                ret = None
                if self.accept("->"):
                    self.attributes()
                    ret = self.type()
                decls.append(("fn", name, params, ret, self.block(), attrs))
            else:
                raise SyntaxError(f"WGSL: unexpected {text!r} at module scope")
        return decls

    def attributes(self):
        attrs = {}
        while self.accept("@"):
            name = self.ident()
            args = []
            if self.accept("("):
                while not self.accept(")"):
                    args.append(self.expr())
                    self.accept(",")
            attrs[name] = args
        return attrs

    def type(self):
        name = self.ident()
        args = []
        if name in _TEMPLATED and self.accept("<"):
            while not self.accept(">"):
                args.append(self.type() if self.peek()[0] == "ident" else self.expr())
                self.accept(",")
        return (name, tuple(args))

    # ---- statements ----
    def block(self):
        self.expect("{")
        body = []
        while not self.accept("}"):
            body.append(self.statement())
        return ("block", body)

    def statement(self):
        kind, text = self.peek()
        if text == "{":
            return self.block()
        if text == ";":
            self.next()
            return ("block", [])
        if text in ("let", "var", "const") and kind == "ident":
            s = self.var_statement()
            self.expect(";")
            return s
        if text == "if":
            self.next()
            cond = self.expr()
            then = self.block()
            other = None
            if self.accept("else"):
                other = self.statement() if self.peek()[1] == "if" else self.block()
            return ("if", cond, then, other)
        if text == "for":
            self.next()
            self.expect("(")
            init = None if self.peek()[1] == ";" else self.simple_statement()
            self.expect(";")
            cond = None if self.peek()[1] == ";" else self.expr()
            self.expect(";")
            step = None if self.peek()[1] == ")" else self.simple_statement()
            self.expect(")")
            return ("for", init, cond, step, self.block())
        if text == "while":
            self.next()
            cond = self.expr()
            return ("for", None, cond, None, self.block())
        if text == "loop":
            self.next()
            return ("for", None, None, None, self.block())
        if text == "return":
            self.next()
            value = None if self.peek()[1] == ";" else self.expr()
            self.expect(";")
            return ("return", value)
        if text in ("break", "continue"):
            self.next()
            self.expect(";")
            return (text,)
        s = self.simple_statement()
        self.expect(";")
        return s

    def var_statement(self):
        kw = self.ident()
        name = self.ident()
        ty = self.type() if self.accept(":") else None
        init = self.expr() if self.accept("=") else None
        return ("decl", kw, name, ty, init)

    def simple_statement(self):
        if self.peek()[1] in ("let", "var", "const"):
            return self.var_statement()
        lhs = self.unary()
        kind, text = self.peek()
        if text in ("=", "+=", "-=", "*=", "/=", "%="):
            self.next()
            return ("assign", text[:-1], lhs, self.expr())
        if text in ("++", "--"):
            self.next()
            return ("assign", text[0], lhs, ("lit", "int", "1"))
        return ("expr", lhs)

    # ---- expressions ----
    def expr(self, level=0):
        if level == len(_BINARY_LEVELS):
            return self.unary()
        lhs = self.expr(level + 1)
        while self.peek()[0] == "op" and self.peek()[1] in _BINARY_LEVELS[level]:
            op = self.next()[1]
            lhs = ("bin", op, lhs, self.expr(level + 1))
        return lhs

    def unary(self):
        kind, text = self.peek()
        if kind == "op" and text in ("-", "!", "&", "*", "~"):
            self.next()
            return ("un", text, self.unary())
        return self.postfix(self.primary())

    def postfix(self, e):
        while True:
            if self.accept("."):
                e = ("member", e, self.ident())
            elif self.accept("["):
                e = ("index", e, self.expr())
                self.expect("]")
            else:
                return e

    def primary(self):
        kind, text = self.next()
        if kind in ("float", "int"):
            return ("lit", kind, text)
        if text == "(":
            e = self.expr()
            self.expect(")")
            return e
        if kind == "ident":
            if text in ("true", "false"):
                return ("lit", "bool", text)
            ty = None
            if text in _TEMPLATED and self.peek()[1] == "<":
                self.i -= 1
                ty = self.type()
            if self.accept("("):
                args = []
                while not self.accept(")"):
                    args.append(self.expr())
                    self.accept(",")
                return ("call", ty if ty else (text, ()), args)
            return ("name", text)
        raise SyntaxError(f"WGSL: unexpected {text!r} in an expression")


# ----------------------------------------------------------------------------------------------------------------------
# values
# ----------------------------------------------------------------------------------------------------------------------
class Vec:
    __slots__ = ("c",)

    def __init__(self, comps):
        self.c = list(comps)

    def __repr__(self):
        return f"Vec({self.c})"


_SWIZZLE = {"x": 0, "y": 1, "z": 2, "w": 3, "r": 0, "g": 1, "b": 2, "a": 3}


class Cell:
    """Storage of one `var` / `let`."""
    __slots__ = ("v",)

    def __init__(self, v):
        self.v = v


class StorageArray:
    """A storage buffer bound as array<f32> / array<u32> / array<vecN<f32>>.  `snapshot` (lockstep schedule): loads see
    the buffer as it was when the dispatch started."""

    def __init__(self, data):
        self.data = data            # numpy array, (n,) or (n, k)
        self.snapshot = None

    def load(self, i):
        src = self.data if self.snapshot is None else self.snapshot
        i = int(i)
        if not 0 <= i < src.shape[0]:
            raise IndexError(f"WGSL: storage index {i} out of bounds ({src.shape[0]})")   # the shaders never rely on clamping
        row = src[i]
        return Vec([row.dtype.type(v) for v in row]) if src.ndim == 2 else row.dtype.type(row)

    def store(self, i, v):
        i = int(i)
        if not 0 <= i < self.data.shape[0]:
            raise IndexError(f"WGSL: storage index {i} out of bounds ({self.data.shape[0]})")
        if self.data.ndim == 2:
            self.data[i, :] = [self.data.dtype.type(c) for c in v.c]
        else:
            self.data[i] = v


class StorageTexture:
    """texture_storage_2d<rgba8unorm, write>: textureStore converts f32 -> unorm8 (round(clamp(v, 0, 1) * 255))."""

    def __init__(self, width, height):
        self.data = np.zeros((height, width, 4), np.uint8)


class Struct:
    def __init__(self, fields):
        self.f = dict(fields)


class Builtins:
    """What WGSL leaves implementation-defined.  Default: numpy's libm in binary32."""

    def sin(self, x):
        return F32(np.sin(F32(x)))

    def cos(self, x):
        return F32(np.cos(F32(x)))

    def fmod(self, a, b):              # WGSL: a - b * trunc(a / b); every backend maps it to an exact remainder or close to it
        return F32(np.fmod(F32(a), F32(b)))

    # min / max with a NaN operand: "indeterminate" in WGSL.  IEEE-754 minNum / maxNum (the other operand wins) is what
    # GPU hardware and libm's fmin / fmax do, and what the arithmetic spec of DESIGN.md section 2 says.
    def min(self, a, b):
        if a != a:
            return b
        if b != b:
            return a
        return b if b < a else a

    def max(self, a, b):
        if a != a:
            return b
        if b != b:
            return a
        return b if b > a else a


class _Return(Exception):
    def __init__(self, value):
        self.value = value


class _Break(Exception):
    pass


class _Continue(Exception):
    pass


def _is_abstract(v):
    return type(v) in (int, float)


def _concretise(v):
    """let / var without a type: AbstractInt -> i32, AbstractFloat -> f32."""
    if type(v) is int:
        return I32(v)
    if type(v) is float:
        return F32(v)
    if isinstance(v, Vec):
        return Vec([_concretise(c) for c in v.c])
    return v


def _wrap_i32(v):
    v &= 0xFFFFFFFF
    return I32(v - (1 << 32) if v >= (1 << 31) else v)


def _convert(v, ty):
    """Value conversion T(v) / declaration with a type."""
    name = ty[0] if isinstance(ty, tuple) else ty
    if isinstance(v, Vec):
        return Vec([_convert(c, ty[1][0] if ty[1] else "f32") for c in v.c]) if name.startswith("vec") else v
    if name == "f32":
        return F32(v)
    if name == "i32":
        if type(v) in (F32, float):
            f = float(v)
            if math.isnan(f):
                return I32(0)
            f = max(-2147483648.0, min(2147483520.0, f))                     # WGSL: truncate, clamped to the representable range
            return I32(math.trunc(f))
        return _wrap_i32(int(v))
    if name == "u32":
        if type(v) in (F32, float):
            f = float(v)
            if math.isnan(f):
                return U32(0)
            return U32(math.trunc(max(0.0, min(4294967040.0, f))))
        return U32(int(v) & 0xFFFFFFFF)
    if name == "bool":
        return bool(v)
    return v


def _unify(a, b):
    """Operands of a binary operator: an abstract literal takes the concrete operand's type."""
    ta, tb = type(a), type(b)
    if ta is tb:
        return a, b
    if _is_abstract(a) and _is_abstract(b):
        return float(a), float(b)
    if _is_abstract(a):
        return tb(a), b
    if _is_abstract(b):
        return a, ta(b)
    raise TypeError(f"WGSL: operands of different types {ta.__name__} and {tb.__name__}")


class Interpreter:
    def __init__(self, source, builtins=None):
        self.src = source
        self.b = builtins or Builtins()
        self.structs, self.fns, self.gvars, self.consts = {}, {}, {}, {}
        self.bindings = {}
        decls = Parser(source).module()
        for d in decls:
            if d[0] == "struct":
                self.structs[d[1]] = d[2]
            elif d[0] == "fn":
                self.fns[d[1]] = d
            elif d[0] == "gvar":
                self.gvars[d[1]] = d
        for d in decls:                                   # module constants (may be used before they are declared)
            if d[0] == "const":
                v = self.eval(d[3], [{}])
                self.consts[d[1]] = _convert(v, d[2]) if d[2] else v

    # ---- host side ----
    def bind(self, name, obj):
        if name not in self.gvars:
            raise KeyError(f"shader has no resource called {name}")
        self.bindings[name] = obj

    def bind_uniform(self, name, **values):
        """Fill a uniform struct by member name; members the shader declares but the host does not set are an error."""
        ty = self.gvars[name][3][0]
        fields = {}
        for m, mty in self.structs[ty]:
            if m.startswith("_"):
                fields[m] = _convert(0, mty)
            else:
                fields[m] = _convert(values[m], mty)
        self.bindings[name] = Struct(fields)

    def workgroup_size(self, entry):
        ws = [self.eval(e, [{}]) for e in self.fns[entry][5]["workgroup_size"]]
        return tuple(int(v) for v in ws) + (1,) * (3 - len(ws))

    def dispatch(self, entry, invocations, schedule="sequential"):
        """Run `entry` once per global_invocation_id in `invocations` (an iterable of (x, y, z)), in that order."""
        storages = [o for o in self.bindings.values() if isinstance(o, StorageArray)]
        for s in storages:
            s.snapshot = s.data.copy() if schedule == "lockstep" else None
        fn = self.fns[entry]
        try:
            for gid in invocations:
                args = []
                for _p, _ty, pattrs in fn[2]:
                    which = pattrs["builtin"][0][1]
                    if which != "global_invocation_id":
                        raise NotImplementedError(which)
                    args.append(Vec([U32(g) for g in gid]))
                self.call_fn(fn, args)
        finally:
            for s in storages:
                s.snapshot = None

    # ---- evaluator ----
    def call_fn(self, fn, args):
        scope = [{p[0]: Cell(a) for p, a in zip(fn[2], args)}]
        try:
            self.exec(fn[4], scope)
        except _Return as r:
            return r.value
        return None

    def lookup(self, name, scope):
        for frame in reversed(scope):
            if name in frame:
                return frame[name]
        return None

    def exec(self, s, scope):
        kind = s[0]
        if kind == "block":
            scope.append({})
            try:
                for st in s[1]:
                    self.exec(st, scope)
            finally:
                scope.pop()
        elif kind == "decl":
            _, kw, name, ty, init = s
            if init is None:
                v = _convert(0, ty)                      # zero value
            else:
                v = self.eval(init, scope)
                v = _convert(v, ty) if ty else (_concretise(v) if kw != "const" else v)
            if isinstance(v, Vec):
                v = Vec(v.c)
            scope[-1][name] = Cell(v)
        elif kind == "assign":
            _, op, lhs, rhs = s
            v = self.eval(rhs, scope)
            if op:
                v = self.binary(op, self.eval(lhs, scope), v)
            self.store(lhs, v, scope)
        elif kind == "expr":
            self.eval(s[1], scope)
        elif kind == "if":
            if self.truth(self.eval(s[1], scope)):
                self.exec(s[2], scope)
            elif s[3] is not None:
                self.exec(s[3], scope)
        elif kind == "for":
            _, init, cond, step, body = s
            scope.append({})
            try:
                if init is not None:
                    self.exec(init, scope)
                while cond is None or self.truth(self.eval(cond, scope)):
                    try:
                        self.exec(body, scope)
                    except _Break:
                        break
                    except _Continue:
                        pass
                    if step is not None:
                        self.exec(step, scope)
            finally:
                scope.pop()
        elif kind == "return":
            raise _Return(None if s[1] is None else self.eval(s[1], scope))
        elif kind == "break":
            raise _Break()
        elif kind == "continue":
            raise _Continue()
        else:
            raise NotImplementedError(kind)

    @staticmethod
    def truth(v):
        if type(v) not in (bool, np.bool_):
            raise TypeError("WGSL: condition is not a bool")
        return bool(v)

    def store(self, lhs, v, scope):
        kind = lhs[0]
        if kind == "name":
            cell = self.lookup(lhs[1], scope)
            if cell is None:
                raise NameError(lhs[1])
            old = cell.v
            cell.v = Vec(v.c) if isinstance(v, Vec) else (type(old)(v) if _is_abstract(v) else v)
            if not isinstance(v, Vec) and type(cell.v) is not type(old):
                raise TypeError(f"WGSL: assigning {type(v).__name__} to a {type(old).__name__} variable {lhs[1]}")
        elif kind == "index":
            base = self.eval(lhs[1], scope)
            idx = self.eval(lhs[2], scope)
            if isinstance(base, StorageArray):
                if not isinstance(v, Vec):
                    v = base.data.dtype.type(v) if _is_abstract(v) else v
                    if type(v) is not base.data.dtype.type:
                        raise TypeError("WGSL: storing a value of the wrong type")
                base.store(idx, v)
            else:
                raise NotImplementedError("indexed store into a non-storage value")
        elif kind == "member":
            target = self.eval(lhs[1], scope)
            if isinstance(target, Vec) and lhs[2] in _SWIZZLE:
                target.c[_SWIZZLE[lhs[2]]] = type(target.c[0])(v) if _is_abstract(v) else v
            else:
                raise NotImplementedError("member store")
        else:
            raise NotImplementedError(kind)

    def binary(self, op, a, b):
        if op in ("&&", "||"):
            raise AssertionError("short-circuit operators are evaluated in eval")
        if isinstance(a, Vec) or isinstance(b, Vec):
            n = len(a.c) if isinstance(a, Vec) else len(b.c)
            ac = a.c if isinstance(a, Vec) else [a] * n
            bc = b.c if isinstance(b, Vec) else [b] * n
            return Vec([self.binary(op, x, y) for x, y in zip(ac, bc)])
        a, b = _unify(a, b)
        t = type(a)
        if op in ("<", ">", "<=", ">=", "==", "!="):
            return bool({"<": a < b, ">": a > b, "<=": a <= b, ">=": a >= b, "==": a == b, "!=": a != b}[op])
        if t in (F32, float):
            with np.errstate(all="ignore"):
                if op == "+":
                    return a + b
                if op == "-":
                    return a - b
                if op == "*":
                    return a * b
                if op == "/":
                    return a / b
                if op == "%":
                    return math.fmod(a, b) if t is float else self.b.fmod(a, b)
        if t in (I32, U32, int):
            x, y = int(a), int(b)
            if op == "+":
                r = x + y
            elif op == "-":
                r = x - y
            elif op == "*":
                r = x * y
            elif op in ("/", "%"):
                if y == 0:
                    r = x if op == "/" else 0               # WGSL: x / 0 = x, x % 0 = 0
                else:
                    q = abs(x) // abs(y) * (1 if (x >= 0) == (y >= 0) else -1)      # truncation toward zero
                    r = q if op == "/" else x - q * y
            elif op in ("&", "|", "^"):
                r = {"&": x & y, "|": x | y, "^": x ^ y}[op]
            else:
                raise NotImplementedError(op)
            if t is int:
                return r
            return _wrap_i32(r) if t is I32 else U32(r & 0xFFFFFFFF)
        if t in (bool, np.bool_) and op in ("&", "|"):
            return bool(a & b) if op == "&" else bool(a | b)
        raise NotImplementedError(f"{op} on {t.__name__}")

    def eval(self, e, scope):
        kind = e[0]
        if kind == "lit":
            _, lk, text = e
            if lk == "bool":
                return text == "true"
            if lk == "int":
                if text.endswith("u"):
                    return U32(int(text[:-1], 0))
                if text.endswith("i"):
                    return I32(int(text[:-1], 0))
                return int(text, 0)
            if text.endswith("f"):
                return F32(float(text[:-1]))
            return float(text)
        if kind == "name":
            name = e[1]
            cell = self.lookup(name, scope)
            if cell is not None:
                return cell.v
            if name in self.consts:
                return self.consts[name]
            if name in self.gvars:
                if name not in self.bindings:
                    raise NameError(f"resource {name} is not bound")
                return self.bindings[name]
            raise NameError(name)
        if kind == "bin":
            _, op, l, r = e
            if op == "&&":
                return self.truth(self.eval(l, scope)) and self.truth(self.eval(r, scope))
            if op == "||":
                return self.truth(self.eval(l, scope)) or self.truth(self.eval(r, scope))
            return self.binary(op, self.eval(l, scope), self.eval(r, scope))
        if kind == "un":
            _, op, x = e
            if op == "&":
                return self.eval(x, scope)               # pointers only feed arrayLength here
            v = self.eval(x, scope)
            if op == "-":
                if isinstance(v, Vec):
                    return Vec([-c for c in v.c])
                return _wrap_i32(-int(v)) if type(v) is I32 else -v
            if op == "!":
                return not self.truth(v)
            raise NotImplementedError(op)
        if kind == "member":
            v = self.eval(e[1], scope)
            m = e[2]
            if isinstance(v, Struct):
                return v.f[m]
            if isinstance(v, Vec):
                if len(m) == 1:
                    return v.c[_SWIZZLE[m]]
                return Vec([v.c[_SWIZZLE[ch]] for ch in m])
            raise TypeError(f"WGSL: .{m} on {type(v).__name__}")
        if kind == "index":
            base = self.eval(e[1], scope)
            idx = self.eval(e[2], scope)
            if isinstance(base, StorageArray):
                return base.load(idx)
            if isinstance(base, Vec):
                return base.c[int(idx)]
            raise TypeError("WGSL: indexing a non-array")
        if kind == "call":
            return self.call(e[1], [self.eval(a, scope) for a in e[2]])
        raise NotImplementedError(kind)

    # ---- calls: user functions, constructors / conversions, the builtin functions the two shaders use ----
    def call(self, ty, args):
        name = ty[0]
        if name in self.fns:
            fn = self.fns[name]
            conv = [_convert(a, p[1]) if _is_abstract(a) else a for a, p in zip(args, fn[2])]
            out = self.call_fn(fn, [Vec(a.c) if isinstance(a, Vec) else a for a in conv])
            return _convert(out, fn[3]) if _is_abstract(out) and fn[3] else out
        if name in ("f32", "i32", "u32", "bool"):
            return _convert(args[0], name)
        if name in ("vec2", "vec3", "vec4"):
            n = int(name[3])
            comps = []
            for a in args:
                comps.extend(a.c if isinstance(a, Vec) else [a])
            if len(comps) == 1:
                comps = comps * n
            if len(comps) != n:
                raise TypeError(f"WGSL: {name} built from {len(comps)} components")
            return Vec([_convert(c, ty[1][0]) for c in comps]) if ty[1] else Vec([_concretise(c) for c in comps])
        if name == "arrayLength":
            return U32(args[0].data.shape[0])
        if name == "textureDimensions":
            h, w = args[0].data.shape[:2]
            return Vec([U32(w), U32(h)])
        if name == "textureStore":
            tex, xy, color = args
            x, y = int(xy.c[0]), int(xy.c[1])
            h, w = tex.data.shape[:2]
            if 0 <= x < w and 0 <= y < h:                  # out-of-bounds texture writes are discarded
                for k in range(4):
                    c = float(F32(color.c[k]))
                    c = 0.0 if math.isnan(c) else min(max(c, 0.0), 1.0)
                    tex.data[y, x, k] = int(math.floor(c * 255.0 + 0.5))
            return None
        return self.builtin_math(name, args)

    def builtin_math(self, name, args):
        if any(isinstance(a, Vec) for a in args):
            n = max(len(a.c) for a in args if isinstance(a, Vec))
            cols = [a.c if isinstance(a, Vec) else [a] * n for a in args]
            return Vec([self.builtin_math(name, list(col)) for col in zip(*cols)])
        conc = [a for a in args if not _is_abstract(a)]
        if conc:
            t = type(conc[0])
            args = [t(a) if _is_abstract(a) else a for a in args]
            if any(type(a) is not t for a in args):
                raise TypeError(f"WGSL: {name} called with mixed types")
        else:
            args = [float(a) for a in args] if any(type(a) is float for a in args) else args
        a = args[0]
        isf = type(a) in (F32, float)
        with np.errstate(all="ignore"):
            if name == "floor":
                return F32(np.floor(a)) if type(a) is F32 else math.floor(a) * 1.0
            if name == "abs":
                return abs(a)
            if name == "min":
                return self.b.min(a, args[1])
            if name == "max":
                return self.b.max(a, args[1])
            if name == "clamp":                              # WGSL: min(max(e, low), high)
                return self.b.min(self.b.max(a, args[1]), args[2])
            if name == "sign":                               # WGSL: 1 if e > 0, -1 if e < 0, 0 otherwise
                one = type(a)(1)
                return one if a > 0 else (-one if a < 0 else type(a)(0))
            if name == "mix":                                # WGSL: e1 * (1 - e3) + e2 * e3
                e1, e2, e3 = args
                return e1 * (type(e3)(1) - e3) + e2 * e3
            if name == "fract":                              # WGSL: e - floor(e)
                return a - (F32(np.floor(a)) if type(a) is F32 else math.floor(a))
            if name == "sin" and isf:
                return self.b.sin(a) if type(a) is F32 else math.sin(a)
            if name == "cos" and isf:
                return self.b.cos(a) if type(a) is F32 else math.cos(a)
            if name == "sqrt" and isf:
                return F32(np.sqrt(a)) if type(a) is F32 else math.sqrt(a)
        raise NotImplementedError(f"WGSL builtin {name}")
