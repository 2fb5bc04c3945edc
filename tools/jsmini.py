"""jsmini -- a small ECMAScript-5 subset interpreter.

Purpose: there is no JavaScript engine in this image, so the reference (audiocogs/aac.js) cannot be
run as-is.  This module executes the reference's *own, unmodified source files* for the hot path
(src/fft.js, mdct.js, mdct_tables.js, filter_bank.js, tns.js and what they require) well enough to
produce golden vectors that pin the CPU oracle.  It is generic (nothing in here knows about AAC):

  * tokenizer + precedence-climbing parser -> AST tuples -> Python closures
  * values: Number = Python float (IEEE double), String, Boolean, null (None), undefined (UNDEF),
    objects with prototype chains, functions/closures with `this`, Arrays, typed arrays
    (Float32Array rounds on store exactly like the spec: numpy float32), Math.*
  * function-level `var` hoisting, `const` as `var`, switch/for/while/if/throw/return/break/continue
  * CommonJS `require('./x')`, `module.exports`, `exports.x`
  * the conversions the reference's behaviour depends on: ToNumber of a typed array is NaN
    ("0,0,...,0"), Math.max(0, NaN) is NaN, x[NaN] / out-of-range typed-array reads are undefined,
    out-of-range typed-array writes are dropped, NaN comparisons are false, ToInt32/ToUint32 shifts.

Not supported (and not used by those files): getters/setters, try/catch, regex literals, labels,
`with`, `in`/`delete`, ASI beyond "newline before }" (the reference terminates statements with ;
except after function-expression assignments, which the parser tolerates).
"""
from __future__ import annotations

import math
import os
import re

import numpy as np


class Undefined:
    __slots__ = ()

    def __repr__(self):
        return "undefined"

    def __bool__(self):
        return False


UNDEF = Undefined()


class JSThrow(Exception):
    def __init__(self, value):
        super().__init__(to_string(value.get("message")) if isinstance(value, JSObject) else to_string(value))
        self.value = value


class _Break(Exception):
    pass


class _Continue(Exception):
    pass


class _Return(Exception):
    def __init__(self, v):
        self.v = v


class JSObject:
    __slots__ = ("props", "proto", "cls")

    def __init__(self, proto=None, cls="Object"):
        self.props, self.proto, self.cls = {}, proto, cls

    def get(self, key):
        o = self
        while o is not None:
            if key in o.props:
                return o.props[key]
            o = o.proto
        return UNDEF

    def put(self, key, v):
        self.props[key] = v


class JSFunction(JSObject):
    __slots__ = ("params", "body", "env", "name", "native", "hoisted")

    def __init__(self, params=None, body=None, env=None, name="", native=None, hoisted=()):
        super().__init__(FUNCTION_PROTO, "Function")
        self.params, self.body, self.env, self.name, self.native, self.hoisted = params, body, env, name, native, hoisted
        if native is None:
            p = JSObject(OBJECT_PROTO)
            p.props["constructor"] = self
            self.props["prototype"] = p

    def call(self, this, args):
        if self.native is not None:
            return self.native(this, args)
        scope = {"this": this}
        for n in self.hoisted:
            scope[n] = UNDEF
        for i, p in enumerate(self.params):
            scope[p] = args[i] if i < len(args) else UNDEF
        scope["arguments"] = JSArray(list(args))
        env = (scope, self.env)
        try:
            self.body(env)
        except _Return as r:
            return r.v
        return UNDEF

    def construct(self, args):
        if self.native is not None:
            return self.native(None, args, True)
        proto = self.props.get("prototype")
        obj = JSObject(proto if isinstance(proto, JSObject) else OBJECT_PROTO)
        r = self.call(obj, args)
        return r if isinstance(r, JSObject) else obj


class JSArray(JSObject):
    __slots__ = ("items",)

    def __init__(self, items):
        super().__init__(ARRAY_PROTO, "Array")
        self.items = items


TYPED = {"Float32Array": np.float32, "Float64Array": np.float64, "Int32Array": np.int32, "Uint32Array": np.uint32,
         "Int16Array": np.int16, "Uint16Array": np.uint16, "Int8Array": np.int8, "Uint8Array": np.uint8}


class JSTyped(JSObject):
    __slots__ = ("a", "kind")

    def __init__(self, kind, a):
        super().__init__(OBJECT_PROTO, kind)
        self.kind, self.a = kind, a

    def get(self, key):
        if key == "set" and "set" not in self.props:
            return native(typed_set)
        return super().get(key)


OBJECT_PROTO = JSObject(None)
FUNCTION_PROTO = JSObject(OBJECT_PROTO)
ARRAY_PROTO = JSObject(OBJECT_PROTO)


# ------------------------------------------------------------------ conversions
def to_number(v):
    if isinstance(v, float):
        return v
    if isinstance(v, bool):
        return 1.0 if v else 0.0
    if isinstance(v, int):
        return float(v)
    if v is None:
        return 0.0
    if v is UNDEF:
        return math.nan
    if isinstance(v, str):
        s = v.strip()
        if s == "":
            return 0.0
        try:
            return float(int(s, 16)) if s[:2].lower() == "0x" else float(s)
        except ValueError:
            return math.nan
    if isinstance(v, (JSArray, JSTyped)):  # ToPrimitive -> toString -> join(",")
        return to_number(to_string(v))
    return math.nan


def num_to_string(x):
    if x != x:
        return "NaN"
    if x in (math.inf, -math.inf):
        return "Infinity" if x > 0 else "-Infinity"
    if x == int(x) and abs(x) < 1e21:
        return str(int(x))
    return repr(x)


def to_string(v):
    if isinstance(v, str):
        return v
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, (float, int)):
        return num_to_string(float(v))
    if v is None:
        return "null"
    if v is UNDEF:
        return "undefined"
    if isinstance(v, JSArray):
        return ",".join("" if (e is None or e is UNDEF) else to_string(e) for e in v.items)
    if isinstance(v, JSTyped):
        return ",".join(num_to_string(float(e)) for e in v.a)
    if isinstance(v, JSFunction):
        return "function"
    return "[object Object]"


def to_bool(v):
    if isinstance(v, bool):
        return v
    if isinstance(v, float):
        return not (v == 0.0 or v != v)
    if isinstance(v, str):
        return v != ""
    if v is None or v is UNDEF:
        return False
    return True


def to_int32(v):
    x = to_number(v)
    if x != x or x in (math.inf, -math.inf):
        return 0
    n = int(x) & 0xFFFFFFFF
    return n - 0x100000000 if n >= 0x80000000 else n


def to_uint32(v):
    return to_int32(v) & 0xFFFFFFFF


def _index(key):
    """Canonical array index of a property key, or None."""
    if isinstance(key, float):
        if key >= 0 and key == int(key):
            return int(key)
        return None
    if isinstance(key, int) and not isinstance(key, bool):
        return key if key >= 0 else None
    if isinstance(key, str) and key.isdigit():
        return int(key)
    return None


def get_member(obj, key):
    if isinstance(obj, JSTyped):
        i = _index(key)
        if i is not None:
            return float(obj.a[i]) if i < obj.a.size else UNDEF
        k = to_string(key)
        if k == "length":
            return float(obj.a.size)
        return obj.get(k)
    if isinstance(obj, JSArray):
        i = _index(key)
        if i is not None:
            return obj.items[i] if i < len(obj.items) else UNDEF
        k = to_string(key)
        if k == "length":
            return float(len(obj.items))
        return obj.get(k)
    if isinstance(obj, JSObject):
        return obj.get(key if isinstance(key, str) else to_string(key))
    if isinstance(obj, str):
        if key == "length":
            return float(len(obj))
        i = _index(key)
        return obj[i] if i is not None and i < len(obj) else UNDEF
    if obj is None or obj is UNDEF:
        raise JSThrow(make_error(f"TypeError: cannot read property '{to_string(key)}' of {to_string(obj)}"))
    return UNDEF


def set_member(obj, key, v):
    if isinstance(obj, JSTyped):
        i = _index(key)
        if i is not None:
            if i < obj.a.size:  # out-of-range writes are dropped
                if obj.a.dtype.kind == "f":
                    obj.a[i] = to_number(v)  # numpy rounds double -> float32 to nearest-even, like the spec
                else:
                    n = to_int32(v)
                    bits = obj.a.dtype.itemsize * 8
                    n &= (1 << bits) - 1
                    if obj.a.dtype.kind == "i" and n >= 1 << (bits - 1):
                        n -= 1 << bits
                    obj.a[i] = n
            return
        if isinstance(key, float):  # NaN / fractional / negative keys on typed arrays: ignored
            return
        obj.put(to_string(key), v)
        return
    if isinstance(obj, JSArray):
        i = _index(key)
        if i is not None:
            if i >= len(obj.items):
                obj.items.extend([UNDEF] * (i + 1 - len(obj.items)))
            obj.items[i] = v
            return
        k = to_string(key)
        if k == "length":
            n = int(to_number(v))
            del obj.items[n:]
            obj.items.extend([UNDEF] * (n - len(obj.items)))
            return
        obj.put(k, v)
        return
    if isinstance(obj, JSObject):
        obj.put(key if isinstance(key, str) else to_string(key), v)
        return
    raise JSThrow(make_error(f"TypeError: cannot set property of {to_string(obj)}"))


def make_error(msg):
    e = JSObject(ERROR_PROTO, "Error")
    e.put("message", msg)
    return e


ERROR_PROTO = JSObject(OBJECT_PROTO)


# ------------------------------------------------------------------ tokenizer
TOKEN_RE = re.compile(r"""
    (?P<ws>\s+|//[^\n]*|/\*.*?\*/)
  | (?P<num>0[xX][0-9a-fA-F]+|(?:\d+\.?\d*(?:[eE][+-]?\d+)?|\.\d+(?:[eE][+-]?\d+)?))
  | (?P<id>[A-Za-z_$][A-Za-z0-9_$]*)
  | (?P<str>"(?:[^"\\\n]|\\.)*"|'(?:[^'\\\n]|\\.)*')
  | (?P<op>>>>=|===|!==|>>>|<<=|>>=|\+\+|--|&&|\|\||==|!=|<=|>=|\+=|-=|\*=|/=|%=|&=|\|=|\^=|<<|>>|[-+*/%=<>!~&|^?:;,.(){}\[\]])
""", re.S | re.X)

KEYWORDS = {"var", "const", "function", "return", "if", "else", "for", "while", "do", "break", "continue", "switch",
            "case", "default", "new", "this", "throw", "typeof", "null", "true", "false", "instanceof", "undefined",
            "try", "catch", "finally"}


def tokenize(src):
    out, pos = [], 0
    while pos < len(src):
        m = TOKEN_RE.match(src, pos)
        if not m:
            raise SyntaxError(f"jsmini: cannot tokenize at {pos}: {src[pos:pos + 40]!r}")
        pos = m.end()
        kind = m.lastgroup
        if kind == "ws":
            continue
        text = m.group(kind)
        if kind == "num":
            out.append(("num", float(int(text, 16)) if text[:2] in ("0x", "0X") else float(text)))
        elif kind == "str":
            out.append(("str", bytes(text[1:-1], "utf8").decode("unicode_escape")))
        elif kind == "id":
            out.append(("kw" if text in KEYWORDS else "id", text))
        else:
            out.append(("op", text))
    out.append(("eof", None))
    return out


# ------------------------------------------------------------------ parser -> AST tuples
BINARY_PREC = {"||": 1, "&&": 2, "|": 3, "^": 4, "&": 5, "==": 6, "!=": 6, "===": 6, "!==": 6, "<": 7, ">": 7, "<=": 7,
               ">=": 7, "instanceof": 7, "<<": 8, ">>": 8, ">>>": 8, "+": 9, "-": 9, "*": 10, "/": 10, "%": 10}
ASSIGN_OPS = {"=", "+=", "-=", "*=", "/=", "%=", "<<=", ">>=", ">>>=", "&=", "|=", "^="}


class Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self, k=0):
        return self.t[self.i + k]

    def next(self):
        tok = self.t[self.i]
        self.i += 1
        return tok

    def at(self, kind, val=None):
        k, v = self.t[self.i]
        return k == kind and (val is None or v == val)

    def eat(self, kind, val=None):
        if self.at(kind, val):
            return self.next()
        return None

    def expect(self, kind, val=None):
        tok = self.eat(kind, val)
        if tok is None:
            raise SyntaxError(f"jsmini: expected {val or kind}, got {self.peek()} at token {self.i}")
        return tok

    def program(self):
        body = []
        while not self.at("eof"):
            body.append(self.statement())
        return ("block", body)

    def block(self):
        self.expect("op", "{")
        body = []
        while not self.at("op", "}"):
            body.append(self.statement())
        self.expect("op", "}")
        return ("block", body)

    def semi(self):
        if not self.eat("op", ";"):
            if not (self.at("op", "}") or self.at("eof")):
                pass  # tolerated: the reference omits ';' after some function-expression assignments

    def var_decl(self):
        self.next()
        decls = []
        while True:
            name = self.expect("id")[1]
            init = self.assign() if self.eat("op", "=") else None
            decls.append((name, init))
            if not self.eat("op", ","):
                break
        return ("var", decls)

    def statement(self):
        k, v = self.peek()
        if k == "op" and v == "{":
            return self.block()
        if k == "op" and v == ";":
            self.next()
            return ("block", [])
        if k == "kw":
            if v in ("var", "const"):
                d = self.var_decl()
                self.semi()
                return d
            if v == "function":
                self.next()
                name = self.expect("id")[1]
                return ("funcdecl", name, self.function_rest(name))
            if v == "return":
                self.next()
                e = None if (self.at("op", ";") or self.at("op", "}")) else self.expression()
                self.semi()
                return ("return", e)
            if v == "if":
                self.next()
                self.expect("op", "(")
                c = self.expression()
                self.expect("op", ")")
                a = self.statement()
                b = self.statement() if self.eat("kw", "else") else None
                return ("if", c, a, b)
            if v == "for":
                self.next()
                self.expect("op", "(")
                init = None
                if not self.at("op", ";"):
                    init = self.var_decl() if (self.at("kw", "var") or self.at("kw", "const")) else ("expr", self.expression())
                self.expect("op", ";")
                cond = None if self.at("op", ";") else self.expression()
                self.expect("op", ";")
                step = None if self.at("op", ")") else self.expression()
                self.expect("op", ")")
                return ("for", init, cond, step, self.statement())
            if v == "while":
                self.next()
                self.expect("op", "(")
                c = self.expression()
                self.expect("op", ")")
                return ("while", c, self.statement())
            if v == "break":
                self.next()
                self.semi()
                return ("break",)
            if v == "continue":
                self.next()
                self.semi()
                return ("continue",)
            if v == "throw":
                self.next()
                e = self.expression()
                self.semi()
                return ("throw", e)
            if v == "try":
                self.next()
                body, name, handler, final = self.block(), None, None, None
                if self.eat("kw", "catch"):
                    self.expect("op", "(")
                    name = self.expect("id")[1]
                    self.expect("op", ")")
                    handler = self.block()
                if self.eat("kw", "finally"):
                    final = self.block()
                return ("try", body, name, handler, final)
            if v == "switch":
                self.next()
                self.expect("op", "(")
                disc = self.expression()
                self.expect("op", ")")
                self.expect("op", "{")
                cases = []
                while not self.eat("op", "}"):
                    if self.eat("kw", "default"):
                        test = None
                    else:
                        self.expect("kw", "case")
                        test = self.expression()
                    self.expect("op", ":")
                    body = []
                    while not (self.at("kw", "case") or self.at("kw", "default") or self.at("op", "}")):
                        body.append(self.statement())
                    cases.append((test, body))
                return ("switch", disc, cases)
        e = self.expression()
        self.semi()
        return ("expr", e)

    def function_rest(self, name=""):
        self.expect("op", "(")
        params = []
        while not self.at("op", ")"):
            params.append(self.expect("id")[1])
            self.eat("op", ",")
        self.expect("op", ")")
        return ("function", name, params, self.block())

    def expression(self):
        e = self.assign()
        while self.eat("op", ","):
            e = ("comma", e, self.assign())
        return e

    def assign(self):
        left = self.ternary()
        if self.at("op") and self.peek()[1] in ASSIGN_OPS:
            op = self.next()[1]
            return ("assign", op, left, self.assign())
        return left

    def ternary(self):
        c = self.binary(1)
        if self.eat("op", "?"):
            a = self.assign()
            self.expect("op", ":")
            return ("cond", c, a, self.assign())
        return c

    def binary(self, prec):
        left = self.unary()
        while True:
            k, v = self.peek()
            if not ((k == "op" or (k == "kw" and v == "instanceof")) and v in BINARY_PREC and BINARY_PREC[v] >= prec):
                return left
            self.next()
            right = self.binary(BINARY_PREC[v] + 1)
            left = ("logical" if v in ("&&", "||") else "binary", v, left, right)

    def unary(self):
        k, v = self.peek()
        if k == "op" and v in ("-", "+", "!", "~"):
            self.next()
            return ("unary", v, self.unary())
        if k == "op" and v in ("++", "--"):
            self.next()
            return ("update", v, True, self.unary())
        if k == "kw" and v == "typeof":
            self.next()
            return ("unary", "typeof", self.unary())
        e = self.postfix()
        if self.at("op", "++") or self.at("op", "--"):
            return ("update", self.next()[1], False, e)
        return e

    def postfix(self):
        if self.eat("kw", "new"):
            callee = self.member_only()
            args = self.arguments() if self.at("op", "(") else []
            e = ("new", callee, args)
        else:
            e = self.primary()
        while True:
            if self.eat("op", "."):
                e = ("member", e, ("lit", self.next()[1]))
            elif self.eat("op", "["):
                idx = self.expression()
                self.expect("op", "]")
                e = ("member", e, idx)
            elif self.at("op", "("):
                e = ("call", e, self.arguments())
            else:
                return e

    def member_only(self):
        e = self.primary()
        while True:
            if self.eat("op", "."):
                e = ("member", e, ("lit", self.next()[1]))
            elif self.eat("op", "["):
                idx = self.expression()
                self.expect("op", "]")
                e = ("member", e, idx)
            else:
                return e

    def arguments(self):
        self.expect("op", "(")
        args = []
        while not self.at("op", ")"):
            args.append(self.assign())
            self.eat("op", ",")
        self.expect("op", ")")
        return args

    def primary(self):
        k, v = self.next()
        if k == "num" or k == "str":
            return ("lit", v)
        if k == "id":
            return ("name", v)
        if k == "kw":
            if v == "this":
                return ("name", "this")
            if v == "null":
                return ("lit", None)
            if v == "undefined":
                return ("lit", UNDEF)
            if v == "true":
                return ("lit", True)
            if v == "false":
                return ("lit", False)
            if v == "function":
                name = self.eat("id")
                return self.function_rest(name[1] if name else "")
        if k == "op":
            if v == "(":
                e = self.expression()
                self.expect("op", ")")
                return e
            if v == "[":
                items = []
                while not self.at("op", "]"):
                    items.append(self.assign())
                    self.eat("op", ",")
                self.expect("op", "]")
                return ("array", items)
            if v == "{":
                props = []
                while not self.at("op", "}"):
                    key = self.next()[1]
                    self.expect("op", ":")
                    props.append((to_string(key) if not isinstance(key, str) else key, self.assign()))
                    self.eat("op", ",")
                self.expect("op", "}")
                return ("object", props)
        raise SyntaxError(f"jsmini: unexpected token {(k, v)} at {self.i}")


# ------------------------------------------------------------------ AST -> closures
def lookup(env, name):
    e = env
    while e is not None:
        scope, e2 = e
        if name in scope:
            return scope
        e = e2
    return None


def js_add(a, b):
    if isinstance(a, float) and isinstance(b, float):
        return a + b
    if isinstance(a, JSObject):
        a = to_string(a)
    if isinstance(b, JSObject):
        b = to_string(b)
    if isinstance(a, str) or isinstance(b, str):
        return to_string(a) + to_string(b)
    return to_number(a) + to_number(b)


def js_div(a, b):
    a, b = to_number(a), to_number(b)
    if b == 0.0:
        if a != a or a == 0.0:
            return math.nan
        return math.copysign(math.inf, a) * math.copysign(1.0, b)
    return a / b


def strict_eq(a, b):
    if isinstance(a, bool) or isinstance(b, bool):
        return isinstance(a, bool) and isinstance(b, bool) and a == b
    if isinstance(a, float) and isinstance(b, float):
        return a == b
    if isinstance(a, str) and isinstance(b, str):
        return a == b
    return a is b


def loose_eq(a, b):
    if (a is None or a is UNDEF) and (b is None or b is UNDEF):
        return True
    if isinstance(a, JSObject) or isinstance(b, JSObject) or a is None or b is None or a is UNDEF or b is UNDEF:
        return a is b
    if type(a) is type(b):
        return strict_eq(a, b)
    return to_number(a) == to_number(b)


def compare(op, a, b):
    if isinstance(a, str) and isinstance(b, str):
        return {"<": a < b, ">": a > b, "<=": a <= b, ">=": a >= b}[op]
    x, y = to_number(a), to_number(b)
    return {"<": x < y, ">": x > y, "<=": x <= y, ">=": x >= y}[op]


BINOPS = {
    "+": js_add,
    "-": lambda a, b: to_number(a) - to_number(b),
    "*": lambda a, b: to_number(a) * to_number(b),
    "/": js_div,
    "%": lambda a, b: math.fmod(to_number(a), to_number(b)) if to_number(b) != 0 else math.nan,
    "<<": lambda a, b: float(to_int32(to_int32(a) << (to_uint32(b) & 31))),
    ">>": lambda a, b: float(to_int32(a) >> (to_uint32(b) & 31)),
    ">>>": lambda a, b: float(to_uint32(a) >> (to_uint32(b) & 31)),
    "&": lambda a, b: float(to_int32(a) & to_int32(b)),
    "|": lambda a, b: float(to_int32(to_int32(a) | to_int32(b))),
    "^": lambda a, b: float(to_int32(to_int32(a) ^ to_int32(b))),
    "===": strict_eq, "!==": lambda a, b: not strict_eq(a, b),
    "==": loose_eq, "!=": lambda a, b: not loose_eq(a, b),
    "<": lambda a, b: compare("<", a, b), ">": lambda a, b: compare(">", a, b),
    "<=": lambda a, b: compare("<=", a, b), ">=": lambda a, b: compare(">=", a, b),
}


def instance_of(a, f):
    proto = f.get("prototype") if isinstance(f, JSObject) else None
    o = a.proto if isinstance(a, JSObject) else None
    while o is not None:
        if o is proto:
            return True
        o = o.proto
    return False


def hoisted_names(node, out):
    """var / function-declaration names of a function body (not descending into nested functions)."""
    if not isinstance(node, tuple):
        if isinstance(node, list):
            for n in node:
                hoisted_names(n, out)
        return
    tag = node[0]
    if tag == "var":
        for name, _ in node[1]:
            out.add(name)
        return
    if tag == "funcdecl":
        out.add(node[1])
        return
    if tag == "function":
        return
    if tag == "switch":
        for _, body in node[2]:
            hoisted_names(body, out)
        return
    for child in node[1:]:
        if isinstance(child, (tuple, list)):
            hoisted_names(child, out)


def compile_node(n):  # noqa: C901  (a flat dispatcher)
    tag = n[0]
    if tag == "lit":
        v = n[1]
        return lambda env: v
    if tag == "name":
        name = n[1]

        def f_name(env):
            e = env
            while e is not None:
                scope = e[0]
                if name in scope:
                    return scope[name]
                e = e[1]
            raise JSThrow(make_error(f"ReferenceError: {name} is not defined"))
        return f_name
    if tag == "member":
        fo, fk = compile_node(n[1]), compile_node(n[2])
        return lambda env: get_member(fo(env), fk(env))
    if tag == "binary":
        op, fa, fb = n[1], compile_node(n[2]), compile_node(n[3])
        if op == "instanceof":
            return lambda env: instance_of(fa(env), fb(env))
        fn = BINOPS[op]
        if op in ("-", "*"):
            def f_fast(env, fa=fa, fb=fb, fn=fn, sub=(op == "-")):
                a, b = fa(env), fb(env)
                if isinstance(a, float) and isinstance(b, float):
                    return a - b if sub else a * b
                return fn(a, b)
            return f_fast
        return lambda env: fn(fa(env), fb(env))
    if tag == "logical":
        op, fa, fb = n[1], compile_node(n[2]), compile_node(n[3])
        if op == "&&":
            def f_and(env):
                a = fa(env)
                return fb(env) if to_bool(a) else a
            return f_and

        def f_or(env):
            a = fa(env)
            return a if to_bool(a) else fb(env)
        return f_or
    if tag == "unary":
        op, fa = n[1], compile_node(n[2])
        if op == "-":
            return lambda env: -to_number(fa(env))
        if op == "+":
            return lambda env: to_number(fa(env))
        if op == "!":
            return lambda env: not to_bool(fa(env))
        if op == "~":
            return lambda env: float(~to_int32(fa(env)))
        if op == "typeof":
            def f_typeof(env):
                try:
                    v = fa(env)
                except JSThrow:
                    return "undefined"
                if v is UNDEF:
                    return "undefined"
                if isinstance(v, bool):
                    return "boolean"
                if isinstance(v, float):
                    return "number"
                if isinstance(v, str):
                    return "string"
                return "function" if isinstance(v, JSFunction) else "object"
            return f_typeof
    if tag == "cond":
        fc, fa, fb = compile_node(n[1]), compile_node(n[2]), compile_node(n[3])
        return lambda env: fa(env) if to_bool(fc(env)) else fb(env)
    if tag == "comma":
        fa, fb = compile_node(n[1]), compile_node(n[2])

        def f_comma(env):
            fa(env)
            return fb(env)
        return f_comma
    if tag in ("assign", "update"):
        if tag == "assign":
            op, target, fv = n[1], n[2], compile_node(n[3])
            binop = None if op == "=" else BINOPS[op[:-1]]
            prefix = True
        else:
            op, prefix, target = n[1], n[2], n[3]
            fv = None
            binop = None
        if target[0] == "name":
            name = target[1]

            def f_assign_name(env):
                scope = lookup(env, name)
                if scope is None:  # implicit global
                    e = env
                    while e[1] is not None:
                        e = e[1]
                    scope = e[0]
                    if tag == "update" or binop is not None:
                        raise JSThrow(make_error(f"ReferenceError: {name} is not defined"))
                if tag == "update":
                    old = to_number(scope[name])
                    new = old + 1.0 if op == "++" else old - 1.0
                    scope[name] = new
                    return new if prefix else old
                v = fv(env)
                if binop is not None:
                    v = binop(scope[name], v)
                scope[name] = v
                return v
            return f_assign_name
        if target[0] == "member":
            fo, fk = compile_node(target[1]), compile_node(target[2])

            def f_assign_member(env):
                obj, key = fo(env), fk(env)
                if tag == "update":
                    old = to_number(get_member(obj, key))
                    new = old + 1.0 if op == "++" else old - 1.0
                    set_member(obj, key, new)
                    return new if prefix else old
                if binop is not None:
                    cur = get_member(obj, key)
                    v = binop(cur, fv(env))
                else:
                    v = fv(env)
                set_member(obj, key, v)
                return v
            return f_assign_member
        raise SyntaxError("jsmini: bad assignment target")
    if tag == "call":
        callee, fargs = n[1], [compile_node(a) for a in n[2]]
        if callee[0] == "member":
            fo, fk = compile_node(callee[1]), compile_node(callee[2])

            def f_mcall(env):
                this = fo(env)
                fn = get_member(this, fk(env))
                if not isinstance(fn, JSFunction):
                    raise JSThrow(make_error(f"TypeError: {to_string(fk(env))} is not a function"))
                return fn.call(this, [a(env) for a in fargs])
            return f_mcall
        ff = compile_node(callee)

        def f_call(env):
            fn = ff(env)
            if not isinstance(fn, JSFunction):
                raise JSThrow(make_error("TypeError: not a function"))
            return fn.call(UNDEF, [a(env) for a in fargs])
        return f_call
    if tag == "new":
        ff, fargs = compile_node(n[1]), [compile_node(a) for a in n[2]]

        def f_new(env):
            fn = ff(env)
            if not isinstance(fn, JSFunction):
                raise JSThrow(make_error("TypeError: not a constructor"))
            return fn.construct([a(env) for a in fargs])
        return f_new
    if tag == "function":
        name, params, body = n[1], n[2], n[3]
        names = set()
        hoisted_names(body, names)
        fbody = compile_node(body)
        hoisted = tuple(names - set(params))
        return lambda env: JSFunction(params, fbody, env, name, hoisted=hoisted)
    if tag == "array":
        fitems = [compile_node(a) for a in n[1]]
        return lambda env: JSArray([f(env) for f in fitems])
    if tag == "object":
        fprops = [(k, compile_node(v)) for k, v in n[1]]

        def f_object(env):
            o = JSObject(OBJECT_PROTO)
            for k, f in fprops:
                o.props[k] = f(env)
            return o
        return f_object
    # ---- statements
    if tag == "block":
        fs = [compile_node(s) for s in n[1]]
        funcs = [s for s in n[1] if s[0] == "funcdecl"]
        ffuncs = [(s[1], compile_node(s[2])) for s in funcs]

        def f_block(env):
            for name, ff in ffuncs:  # function declarations are hoisted with their value
                scope = lookup(env, name) or env[0]
                scope[name] = ff(env)
            for f in fs:
                f(env)
        return f_block
    if tag == "funcdecl":
        return lambda env: None
    if tag == "var":
        decls = [(name, compile_node(init) if init is not None else None) for name, init in n[1]]

        def f_var(env):
            for name, fi in decls:
                if fi is not None:
                    scope = lookup(env, name) or env[0]
                    scope[name] = fi(env)
                elif lookup(env, name) is None:
                    env[0][name] = UNDEF
        return f_var
    if tag == "expr":
        return compile_node(n[1])
    if tag == "return":
        fe = compile_node(n[1]) if n[1] is not None else None

        def f_return(env):
            raise _Return(fe(env) if fe else UNDEF)
        return f_return
    if tag == "if":
        fc, fa = compile_node(n[1]), compile_node(n[2])
        fb = compile_node(n[3]) if n[3] is not None else None

        def f_if(env):
            if to_bool(fc(env)):
                fa(env)
            elif fb is not None:
                fb(env)
        return f_if
    if tag == "for":
        fi = compile_node(n[1]) if n[1] is not None else None
        fc = compile_node(n[2]) if n[2] is not None else None
        fs = compile_node(n[3]) if n[3] is not None else None
        fb = compile_node(n[4])

        def f_for(env):
            if fi:
                fi(env)
            while fc is None or to_bool(fc(env)):
                try:
                    fb(env)
                except _Continue:
                    pass
                except _Break:
                    break
                if fs:
                    fs(env)
        return f_for
    if tag == "while":
        fc, fb = compile_node(n[1]), compile_node(n[2])

        def f_while(env):
            while to_bool(fc(env)):
                try:
                    fb(env)
                except _Continue:
                    continue
                except _Break:
                    break
        return f_while
    if tag == "break":
        def f_break(env):
            raise _Break()
        return f_break
    if tag == "continue":
        def f_continue(env):
            raise _Continue()
        return f_continue
    if tag == "throw":
        fe = compile_node(n[1])

        def f_throw(env):
            raise JSThrow(fe(env))
        return f_throw
    if tag == "try":
        fb = compile_node(n[1])
        name = n[2]
        fh = compile_node(n[3]) if n[3] is not None else None
        ff = compile_node(n[4]) if n[4] is not None else None

        def f_try(env):
            try:
                try:
                    fb(env)
                except JSThrow as t:
                    if fh is None:
                        raise
                    fh(({name: t.value}, env))   # the catch parameter lives in its own scope
            finally:
                if ff is not None:
                    ff(env)
        return f_try
    if tag == "switch":
        fd = compile_node(n[1])
        cases = [(compile_node(t) if t is not None else None, [compile_node(s) for s in body]) for t, body in n[2]]

        def f_switch(env):
            d = fd(env)
            start = None
            for i, (ft, _) in enumerate(cases):
                if ft is not None and strict_eq(d, ft(env)):
                    start = i
                    break
            if start is None:
                for i, (ft, _) in enumerate(cases):
                    if ft is None:
                        start = i
                        break
            if start is None:
                return
            try:
                for _, body in cases[start:]:
                    for f in body:
                        f(env)
            except _Break:
                pass
        return f_switch
    raise SyntaxError(f"jsmini: cannot compile {tag}")


# ------------------------------------------------------------------ runtime library
def native(fn):
    return JSFunction(native=lambda this, args, new=False: fn(this, args))


def _fn_call(this, args):      # Function.prototype.call(thisArg, ...args)
    return this.call(args[0] if args else UNDEF, list(args[1:]))


def _fn_apply(this, args):     # Function.prototype.apply(thisArg, argsArray)
    arr = args[1] if len(args) > 1 else UNDEF
    return this.call(args[0] if args else UNDEF, list(arr.items) if isinstance(arr, JSArray) else [])


def _array_push(this, args):   # Array.prototype.push
    this.items.extend(args)
    return float(len(this.items))


ARRAY_PROTO.props["push"] = native(_array_push)
FUNCTION_PROTO.props["call"] = native(_fn_call)
FUNCTION_PROTO.props["apply"] = native(_fn_apply)


def js_max(this, args):
    r = -math.inf
    for a in args:
        x = to_number(a)
        if x != x:
            return math.nan
        r = max(r, x)
    return r


def js_min(this, args):
    r = math.inf
    for a in args:
        x = to_number(a)
        if x != x:
            return math.nan
        r = min(r, x)
    return r


def js_pow(this, args):
    x, y = to_number(args[0]), to_number(args[1])
    try:
        return math.pow(x, y)
    except (ValueError, OverflowError):
        return math.nan


class JSArrayBuffer(JSObject):
    """ArrayBuffer: raw bytes; typed-array views and DataViews alias it (numpy views)."""
    __slots__ = ("b",)

    def __init__(self, nbytes):
        super().__init__(OBJECT_PROTO, "ArrayBuffer")
        self.b = np.zeros(int(nbytes), np.uint8)
        self.props["byteLength"] = float(int(nbytes))


def arraybuffer_ctor(this, args, new=False):
    return JSArrayBuffer(to_number(args[0]) if args else 0)


def dataview_ctor(this, args, new=False):
    buf = args[0]
    assert isinstance(buf, JSArrayBuffer), "DataView needs an ArrayBuffer"
    base = int(to_number(args[1])) if len(args) > 1 and args[1] is not UNDEF else 0
    view = JSObject(OBJECT_PROTO, "DataView")
    view.props["buffer"] = buf

    def accessor(fmt, size, setter):
        def fn(this_, a):
            import struct

            at = base + int(to_number(a[0]))
            little = to_bool(a[2 if setter else 1]) if len(a) > (2 if setter else 1) else False
            code = ("<" if little else ">") + fmt
            if setter:
                v = to_number(a[1])
                if fmt in "bBhHiI":
                    v = to_int32(v) & ((1 << (8 * size)) - 1) if fmt in "BHI" else int(np.array(to_int32(v)).astype(
                        {1: np.int8, 2: np.int16, 4: np.int32}[size]))
                buf.b[at:at + size] = np.frombuffer(struct.pack(code, v), np.uint8)
                return UNDEF
            return float(struct.unpack(code, bytes(buf.b[at:at + size]))[0])
        return native(fn)

    for name, fmt, size in (("Float32", "f", 4), ("Float64", "d", 8), ("Uint8", "B", 1), ("Int8", "b", 1), ("Uint16", "H", 2),
                            ("Int16", "h", 2), ("Uint32", "I", 4), ("Int32", "i", 4)):
        view.props["set" + name] = accessor(fmt, size, True)
        view.props["get" + name] = accessor(fmt, size, False)
    return view


def typed_set(this, args):
    """TypedArray.prototype.set(source[, offset])"""
    src, off = args[0], int(to_number(args[1])) if len(args) > 1 and args[1] is not UNDEF else 0
    if isinstance(src, JSTyped):
        this.a[off:off + src.a.size] = src.a.astype(this.a.dtype)
    else:
        for i, v in enumerate(src.items):
            set_member(this, float(off + i), v)
    return UNDEF


def make_typed_ctor(kind):
    dt = TYPED[kind]

    def ctor(this, args, new=False):
        a0 = args[0] if args else 0.0
        if isinstance(a0, JSArrayBuffer):   # a view: new T(buffer[, byteOffset[, length]])
            off = int(to_number(args[1])) if len(args) > 1 and args[1] is not UNDEF else 0
            isz = np.dtype(dt).itemsize
            n = int(to_number(args[2])) if len(args) > 2 and args[2] is not UNDEF else (a0.b.size - off) // isz
            assert off % isz == 0, "typed-array view must be aligned"
            t = JSTyped(kind, a0.b[off:off + n * isz].view(dt))
            t.props["buffer"] = a0
            t.props["set"] = native(typed_set)
            return t
        if isinstance(a0, JSArray):
            t = JSTyped(kind, np.zeros(len(a0.items), dt))
            for i, v in enumerate(a0.items):
                set_member(t, float(i), v)
            return t
        if isinstance(a0, JSTyped):
            return JSTyped(kind, a0.a.astype(dt))
        return JSTyped(kind, np.zeros(int(to_number(a0)), dt))
    return JSFunction(native=ctor)


def array_ctor(this, args, new=False):
    if len(args) == 1 and isinstance(args[0], float):
        return JSArray([UNDEF] * int(args[0]))
    return JSArray(list(args))


def error_ctor(this, args, new=False):
    return make_error(to_string(args[0]) if args else "")


def make_globals():
    g = {}
    m = JSObject(OBJECT_PROTO)
    for name, fn in (("sin", math.sin), ("cos", math.cos), ("sqrt", lambda x: math.sqrt(x) if x >= 0 else math.nan),
                     ("floor", lambda x: float(math.floor(x)) if x == x and abs(x) != math.inf else x),
                     ("abs", abs), ("log", lambda x: math.log(x) if x > 0 else (-math.inf if x == 0 else math.nan)),
                     ("exp", math.exp), ("round", lambda x: float(math.floor(x + 0.5))),
                     ("ceil", lambda x: float(math.ceil(x)))):
        m.props[name] = native(lambda this, args, fn=fn: fn(to_number(args[0]) if args else math.nan))
    m.props["max"], m.props["min"], m.props["pow"] = native(js_max), native(js_min), native(js_pow)
    m.props["PI"] = math.pi
    g["Math"] = m
    for kind in TYPED:
        g[kind] = make_typed_ctor(kind)
    g["Array"] = JSFunction(native=array_ctor)
    g["ArrayBuffer"] = JSFunction(native=arraybuffer_ctor)
    g["DataView"] = JSFunction(native=dataview_ctor)
    g["Error"] = JSFunction(native=error_ctor)
    g["Error"].props["prototype"] = ERROR_PROTO
    g["NaN"], g["Infinity"], g["undefined"] = math.nan, math.inf, UNDEF
    g["isNaN"] = native(lambda this, args: to_number(args[0]) != to_number(args[0]))
    return g


class Runtime:
    """A CommonJS world rooted at one source directory."""

    def __init__(self, src_dir, stubs=None):
        self.src_dir, self.modules, self.stubs = src_dir, {}, dict(stubs or {})
        self.globals = make_globals()

    def require(self, name):
        if name in self.stubs:
            return self.stubs[name]
        path = os.path.normpath(os.path.join(self.src_dir, name + ("" if name.endswith(".js") else ".js")))
        if path in self.modules:
            return self.modules[path].get("exports")
        module = JSObject(OBJECT_PROTO)
        exports = JSObject(OBJECT_PROTO)
        module.put("exports", exports)
        self.modules[path] = module
        ast = Parser(tokenize(open(path).read())).program()
        names = set()
        hoisted_names(ast, names)
        scope = {n: UNDEF for n in names}
        scope.update(module=module, exports=exports, this=exports,
                     require=native(lambda this, args: self.require(to_string(args[0]))))
        compile_node(ast)((scope, (self.globals, None)))
        return module.get("exports")

    def run(self, source, extra=None):
        """Evaluate a snippet in a fresh scope on top of the globals; returns that scope."""
        ast = Parser(tokenize(source)).program()
        names = set()
        hoisted_names(ast, names)
        scope = {n: UNDEF for n in names}
        scope["require"] = native(lambda this, args: self.require(to_string(args[0])))
        scope.update(extra or {})
        compile_node(ast)((scope, (self.globals, None)))
        return scope


def float32array(values):
    return JSTyped("Float32Array", np.asarray(values, np.float32).copy())


def int32array(values):
    return JSTyped("Int32Array", np.asarray(values, np.int32).copy())


def obj(**props):
    o = JSObject(OBJECT_PROTO)
    for k, v in props.items():
        o.props[k] = float(v) if isinstance(v, int) and not isinstance(v, bool) else v
    return o
