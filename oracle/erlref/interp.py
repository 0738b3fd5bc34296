"""A small Erlang evaluator: enough of the language to RUN THE REFERENCE'S OWN SOURCE for the hot path
(erlamsa_main:fuzzer/1 with paths=[direct], output=return, and the modules below it) where the sources lie
under /root/reference/src.  TEST INFRASTRUCTURE (oracle/): it generates reference vectors (tests/golden/) and
validates the C++ oracle; nothing in the product imports it.

Method: AST -> Python closures.  env = dict(var -> value).  Tail calls are trampolined (TailCall)."""
import os
import sys

from .terms import (NIL, Cons, Bits, Pid, Ref, ErlError, from_py, to_py, erl_cmp, exact_eq, is_list, fmt_term)
from .lexer import Preprocessor
from .parser import parse_function, parse_expr_string

T, F = "true", "false"


class BudgetExceeded(Exception):
    pass


class TailCall(object):
    __slots__ = ("fun", "args")

    def __init__(self, fun, args):
        self.fun = fun
        self.args = args


class Fun(object):
    __slots__ = ("clauses", "arity", "env", "name", "module", "selfname")

    def __init__(self, clauses, arity, env, name, module, selfname=None):
        self.clauses = clauses
        self.arity = arity
        self.env = env
        self.name = name
        self.module = module
        self.selfname = selfname

    def __repr__(self):
        return "#Fun<%s.%s/%d>" % (self.module, self.name, self.arity)


class Bif(object):
    __slots__ = ("fn", "arity", "name")

    def __init__(self, fn, arity, name):
        self.fn = fn
        self.arity = arity
        self.name = name

    def __repr__(self):
        return "#Fun<%s/%d>" % (self.name, self.arity)


class Process(object):
    def __init__(self):
        self.pid = Pid()
        self.dict = {}
        self.mailbox = []


class Module(object):
    def __init__(self, name):
        self.name = name
        self.funs = {}
        self.records = {}


def erl_int_to_float(n):
    """BEAM integer -> float: smalls are a hardware conversion; bignums go digit by digit
    (erts big_to_double: d = d * 2^64 + digit), which is NOT always the correctly rounded value."""
    if -(1 << 59) <= n < (1 << 59):
        return float(n)
    neg = n < 0
    a = -n if neg else n
    digits = []
    while a:
        digits.append(a & 0xFFFFFFFFFFFFFFFF)
        a >>= 64
    d = 0.0
    for dg in reversed(digits):
        d = d * 18446744073709551616.0 + float(dg)
        if d == float("inf"):
            raise ErlError("error", "badarith")
    return -d if neg else d


def num_to_float(x):
    return x if type(x) is float else erl_int_to_float(x)


def arith(op, a, b):
    ta, tb = type(a), type(b)
    if not ((ta is int or ta is float) and (tb is int or tb is float)):
        raise ErlError("error", "badarith")
    try:
        if op == "+":
            if ta is int and tb is int:
                return a + b
            r = num_to_float(a) + num_to_float(b)
        elif op == "-":
            if ta is int and tb is int:
                return a - b
            r = num_to_float(a) - num_to_float(b)
        elif op == "*":
            if ta is int and tb is int:
                return a * b
            r = num_to_float(a) * num_to_float(b)
        else:   # "/"
            r = num_to_float(a) / num_to_float(b)
    except (ZeroDivisionError, OverflowError):
        raise ErlError("error", "badarith")
    if r != r or r in (float("inf"), float("-inf")):
        raise ErlError("error", "badarith")
    return r


def int_div(a, b):
    if type(a) is not int or type(b) is not int or b == 0:
        raise ErlError("error", "badarith")
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def int_rem(a, b):
    if type(a) is not int or type(b) is not int or b == 0:
        raise ErlError("error", "badarith")
    r = abs(a) % abs(b)
    return r if a >= 0 else -r


def list_append(a, b):
    if a is NIL:
        return b
    return from_py(to_py(a), b)


def list_subtract(a, b):
    la = to_py(a)
    for x in to_py(b):
        for i, y in enumerate(la):
            if exact_eq(x, y):
                del la[i]
                break
    return from_py(la)


def _ints(a, b):
    if type(a) is not int or type(b) is not int:
        raise ErlError("error", "badarith")


def _bools(a, b):
    if a not in (T, F) or b not in (T, F):
        raise ErlError("error", "badarg")


BINOP_IMPL = {
    "+": lambda a, b: a + b if (type(a) is int and type(b) is int) else arith("+", a, b),
    "-": lambda a, b: a - b if (type(a) is int and type(b) is int) else arith("-", a, b),
    "*": lambda a, b: a * b if (type(a) is int and type(b) is int) else arith("*", a, b),
    "/": lambda a, b: arith("/", a, b),
    "div": int_div,
    "rem": int_rem,
    "band": lambda a, b: (_ints(a, b), a & b)[1],
    "bor": lambda a, b: (_ints(a, b), a | b)[1],
    "bxor": lambda a, b: (_ints(a, b), a ^ b)[1],
    "bsl": lambda a, b: (_ints(a, b), a << b if b >= 0 else a >> -b)[1],
    "bsr": lambda a, b: (_ints(a, b), a >> b if b >= 0 else a << -b)[1],
    "and": lambda a, b: (_bools(a, b), T if (a == T and b == T) else F)[1],
    "or": lambda a, b: (_bools(a, b), T if (a == T or b == T) else F)[1],
    "xor": lambda a, b: (_bools(a, b), T if ((a == T) != (b == T)) else F)[1],
    "==": lambda a, b: T if erl_cmp(a, b) == 0 else F,
    "/=": lambda a, b: T if erl_cmp(a, b) != 0 else F,
    "<": lambda a, b: T if erl_cmp(a, b) < 0 else F,
    ">": lambda a, b: T if erl_cmp(a, b) > 0 else F,
    "=<": lambda a, b: T if erl_cmp(a, b) <= 0 else F,
    ">=": lambda a, b: T if erl_cmp(a, b) >= 0 else F,
    "=:=": lambda a, b: T if exact_eq(a, b) else F,
    "=/=": lambda a, b: F if exact_eq(a, b) else T,
    "++": list_append,
    "--": list_subtract,
}


# --------------------------------------------------------------------------- bit syntax
def _seg_spec(specs):
    typ, signed, endian, unit = "integer", False, "big", None
    for a, n in specs:
        if a in ("integer", "float", "binary", "bytes", "bitstring", "bits", "utf8", "utf16", "utf32"):
            typ = {"bytes": "binary", "bits": "bitstring"}.get(a, a)
        elif a == "signed":
            signed = True
        elif a == "unsigned":
            signed = False
        elif a in ("big", "little", "native"):
            endian = "little" if a in ("little", "native") else "big"
        elif a == "unit":
            unit = n
    if unit is None:
        unit = 8 if typ == "binary" else 1
    return typ, signed, endian, unit


def bits_of(v):
    """value -> (int, nbits) for bytes / Bits"""
    if type(v) is bytes:
        return int.from_bytes(v, "big"), len(v) * 8
    if type(v) is Bits:
        return v.val, v.nbits
    raise ErlError("error", "badarg")


def make_bits(val, nbits):
    if nbits % 8 == 0:
        return val.to_bytes(nbits // 8, "big")
    return Bits(val, nbits)


import struct


def build_binary(parts):
    """parts: list of (kind, payload): ('b', bytes) | ('i', value, nbits, little)"""
    if all((p[0] == "b" and type(p[1]) is bytes) or (p[0] == "i" and p[2] % 8 == 0) for p in parts):
        out = []
        for p in parts:
            if p[0] == "b":
                out.append(p[1])
            else:
                nb = p[2] // 8
                out.append((p[1] & ((1 << p[2]) - 1)).to_bytes(nb, "little" if p[3] else "big"))
        return b"".join(out)
    acc, n = 0, 0
    for p in parts:
        if p[0] == "b":
            v, nb = bits_of(p[1])
        else:
            nb = p[2]
            v = p[1] & ((1 << nb) - 1)
            if p[3] and nb % 8 == 0:
                v = int.from_bytes(v.to_bytes(nb // 8, "big"), "little")
        acc = (acc << nb) | v
        n += nb
    return make_bits(acc, n)


def pattern_vars(node, out):
    """names of the variables a pattern can bind"""
    if isinstance(node, tuple):
        if node and node[0] == "var" and len(node) == 2 and isinstance(node[1], str):
            if node[1] != "_":
                out.add(node[1])
            return out
        for x in node:
            pattern_vars(x, out)
    elif isinstance(node, list):
        for x in node:
            pattern_vars(x, out)
    return out


# --------------------------------------------------------------------------- compiler
class Compiler(object):
    def __init__(self, rt, module):
        self.rt = rt
        self.module = module      # Module

    # ---- patterns: closure m(value, env, trail) -> bool
    def pattern(self, node):
        k = node[0]
        if k == "var":
            name = node[1]
            if name == "_":
                return lambda v, env, trail: True

            def m_var(v, env, trail):
                if name in env:
                    return exact_eq(env[name], v)
                env[name] = v
                trail.append(name)
                return True
            return m_var
        if k in ("int", "atom"):
            c = node[1]
            return lambda v, env, trail: type(v) is type(c) and v == c
        if k == "float":
            c = node[1]
            return lambda v, env, trail: type(v) is float and v == c
        if k == "nil":
            return lambda v, env, trail: v is NIL
        if k == "str":
            want = from_py(node[1])
            return lambda v, env, trail: (v is NIL and want is NIL) or (type(v) is Cons and type(want) is Cons and exact_eq(v, want))
        if k == "paren":
            return self.pattern(node[1])
        if k == "cons":
            mh, mt = self.pattern(node[1]), self.pattern(node[2])

            def m_cons(v, env, trail):
                return type(v) is Cons and mh(v.h, env, trail) and mt(v.t, env, trail)
            return m_cons
        if k == "tuple":
            ms = [self.pattern(e) for e in node[1]]
            n = len(ms)

            def m_tuple(v, env, trail):
                if type(v) is not tuple or len(v) != n:
                    return False
                for i in range(n):
                    if not ms[i](v[i], env, trail):
                        return False
                return True
            return m_tuple
        if k == "match":
            ma, mb = self.pattern(node[1]), self.pattern(node[2])
            return lambda v, env, trail: ma(v, env, trail) and mb(v, env, trail)
        if k == "bin":
            return self.bin_pattern(node[1])
        if k == "op" and node[1] == "++":
            # "prefix" ++ Rest
            left = node[2]
            if left[0] == "paren":
                left = left[1]
            if left[0] == "str":
                pre = left[1]
            elif left[0] == "nil":
                pre = []
            else:
                raise NotImplementedError("++ pattern with non-literal prefix")
            mr = self.pattern(node[3])

            def m_pp(v, env, trail):
                for c in pre:
                    if type(v) is not Cons or v.h != c:
                        return False
                    v = v.t
                return mr(v, env, trail)
            return m_pp
        if k in ("op", "unop"):
            f = self.expr(node)     # constant arithmetic in a pattern
            return lambda v, env, trail: exact_eq(v, f(env))
        if k == "map":
            items = [(self.expr(kk), self.pattern(vv)) for kk, vv, _ in node[1]]

            def m_map(v, env, trail):
                if type(v) is not dict:
                    return False
                for kf, mv in items:
                    key = kf(env)
                    if key not in v or not mv(v[key], env, trail):
                        return False
                return True
            return m_map
        if k == "record":
            rname = node[2]
            fields = self.module.records[rname]
            idx = {f[0]: i + 1 for i, f in enumerate(fields)}
            ms = [(idx[fn], self.pattern(fe)) for fn, fe in node[3]]
            n = len(fields) + 1

            def m_rec(v, env, trail):
                if type(v) is not tuple or len(v) != n or v[0] != rname:
                    return False
                for i, mm in ms:
                    if not mm(v[i], env, trail):
                        return False
                return True
            return m_rec
        raise NotImplementedError("pattern kind %s" % k)

    def bin_pattern(self, segs):
        comp = []
        for (v, size, specs) in segs:
            typ, signed, endian, unit = _seg_spec(specs)
            if v[0] == "str":
                for c in v[1]:
                    comp.append(("lit", c, None, typ, signed, endian, unit))
                continue
            sizef = self.expr(size) if size is not None else None
            if v[0] == "var":
                comp.append(("var", v[1], sizef, typ, signed, endian, unit))
            elif v[0] in ("int", "float"):
                comp.append(("lit", v[1], sizef, typ, signed, endian, unit))
            else:
                comp.append(("pat", self.pattern(v), sizef, typ, signed, endian, unit))
        nseg = len(comp)

        def m_bin(val, env, trail):
            tv = type(val)
            if tv is bytes:
                data, total = val, len(val) * 8
                ival = None
            elif tv is Bits:
                data, total = None, val.nbits
                ival = val.val
            else:
                return False
            pos = 0
            for si in range(nseg):
                kind, x, sizef, typ, signed, endian, unit = comp[si]
                if typ == "integer" or typ == "float":
                    if sizef is None:
                        nb = 8 if typ == "integer" else 64
                    else:
                        s = sizef(env)
                        if type(s) is not int or s < 0:
                            return False
                        nb = s * unit
                    if pos + nb > total:
                        return False
                    if data is not None and pos % 8 == 0 and nb % 8 == 0:
                        if nb == 8:
                            iv = data[pos >> 3]
                        else:
                            iv = int.from_bytes(data[pos >> 3:(pos + nb) >> 3], endian)
                    else:
                        if ival is None:
                            ival = int.from_bytes(data, "big")
                        iv = (ival >> (total - pos - nb)) & ((1 << nb) - 1)
                        if endian == "little" and nb % 8 == 0 and nb > 8:
                            iv = int.from_bytes(iv.to_bytes(nb // 8, "big"), "little")
                    if signed and nb > 0 and iv >> (nb - 1):
                        iv -= 1 << nb
                    if typ == "float":
                        if nb == 64:
                            iv = struct.unpack(">d", (iv & ((1 << 64) - 1)).to_bytes(8, "big"))[0]
                        elif nb == 32:
                            iv = struct.unpack(">f", (iv & ((1 << 32) - 1)).to_bytes(4, "big"))[0]
                        else:
                            return False
                    pos += nb
                    got = iv
                elif typ == "binary" or typ == "bitstring":
                    if sizef is None:
                        nb = total - pos
                        if typ == "binary" and nb % 8 != 0:
                            return False
                    else:
                        s = sizef(env)
                        if type(s) is not int or s < 0:
                            return False
                        nb = s * unit
                    if pos + nb > total:
                        return False
                    if data is not None and pos % 8 == 0 and nb % 8 == 0:
                        got = data[pos >> 3:(pos + nb) >> 3]
                    else:
                        if ival is None:
                            ival = int.from_bytes(data, "big")
                        got = make_bits((ival >> (total - pos - nb)) & ((1 << nb) - 1), nb)
                    pos += nb
                elif typ == "utf8":
                    if data is None or pos % 8:
                        return False
                    b0 = data[pos >> 3] if (pos >> 3) < len(data) else None
                    if b0 is None:
                        return False
                    ln = 1 if b0 < 0x80 else 2 if b0 >> 5 == 6 else 3 if b0 >> 4 == 14 else 4 if b0 >> 3 == 30 else 0
                    try:
                        got = ord(data[pos >> 3:(pos >> 3) + ln].decode("utf8"))
                    except Exception:
                        return False
                    pos += ln * 8
                else:
                    raise NotImplementedError("binary segment type %s" % typ)
                if kind == "var":
                    if x == "_":
                        continue
                    if x in env:
                        if not exact_eq(env[x], got):
                            return False
                    else:
                        env[x] = got
                        trail.append(x)
                elif kind == "lit":
                    if got != x or (type(x) is float) != (type(got) is float):
                        return False
                else:
                    if not x(got, env, trail):
                        return False
            return pos == total
        return m_bin

    # ---- guards
    def guards(self, alts):
        if not alts:
            return None
        comp = [[self.expr(g) for g in conj] for conj in alts]

        def g_fn(env):
            for conj in comp:
                try:
                    ok = True
                    for g in conj:
                        if g(env) != T:
                            ok = False
                            break
                    if ok:
                        return True
                except ErlError:
                    pass
            return False
        return g_fn

    # ---- clauses (function / fun / case / receive): list of (matchers, guard, body)
    def clause(self, c, tail):
        pats, guards, body = c
        return ([self.pattern(p) for p in pats], self.guards(guards), self.body(body, tail))

    def body(self, exprs, tail):
        fs = [self.expr(e) for e in exprs[:-1]]
        last = self.expr(exprs[-1], tail)
        if not fs:
            return last
        if len(fs) == 1:
            f0 = fs[0]

            def b2(env):
                f0(env)
                return last(env)
            return b2

        def bn(env):
            for f in fs:
                f(env)
            return last(env)
        return bn

    # ---- expressions: closure f(env) -> value
    def expr(self, node, tail=False):
        k = node[0]
        m = getattr(self, "e_" + k, None)
        if m is None:
            raise NotImplementedError("expression kind %s" % k)
        return m(node, tail)

    def e_int(self, node, tail):
        c = node[1]
        return lambda env: c
    e_float = e_int
    e_atom = e_int

    def e_var(self, node, tail):
        name = node[1]

        def f(env):
            try:
                return env[name]
            except KeyError:
                raise ErlError("error", ("unbound_var", name))
        return f

    def e_str(self, node, tail):
        chars = node[1]
        return lambda env: from_py(chars)

    def e_nil(self, node, tail):
        return lambda env: NIL

    def e_paren(self, node, tail):
        return self.expr(node[1], tail)

    def e_cons(self, node, tail):
        fh, ft = self.expr(node[1]), self.expr(node[2])
        return lambda env: Cons(fh(env), ft(env))

    def e_tuple(self, node, tail):
        fs = [self.expr(e) for e in node[1]]
        if len(fs) == 2:
            a, b = fs
            return lambda env: (a(env), b(env))
        return lambda env: tuple([f(env) for f in fs])

    def e_block(self, node, tail):
        return self.body(node[1], tail)

    def e_match(self, node, tail):
        pat = self.pattern(node[1])
        val = self.expr(node[2])

        def f(env):
            v = val(env)
            trail = []
            if not pat(v, env, trail):
                for n in trail:
                    del env[n]
                raise ErlError("error", ("badmatch", v))
            return v
        return f

    def e_op(self, node, tail):
        op = node[1]
        impl = BINOP_IMPL[op]
        fa, fb = self.expr(node[2]), self.expr(node[3])
        return lambda env: impl(fa(env), fb(env))

    def e_unop(self, node, tail):
        op = node[1]
        fe = self.expr(node[2])
        if op == "-":
            def f(env):
                v = fe(env)
                if type(v) is not int and type(v) is not float:
                    raise ErlError("error", "badarith")
                return -v
            return f
        if op == "+":
            return fe
        if op == "not":
            def f(env):
                v = fe(env)
                if v == T:
                    return F
                if v == F:
                    return T
                raise ErlError("error", "badarg")
            return f
        if op == "bnot":
            def f(env):
                v = fe(env)
                if type(v) is not int:
                    raise ErlError("error", "badarith")
                return ~v
            return f
        raise NotImplementedError(op)

    def e_andalso(self, node, tail):
        fa, fb = self.expr(node[1]), self.expr(node[2])

        def f(env):
            a = fa(env)
            if a == F:
                return F
            if a != T:
                raise ErlError("error", ("badarg", a))
            return fb(env)
        return f

    def e_orelse(self, node, tail):
        fa, fb = self.expr(node[1]), self.expr(node[2])

        def f(env):
            a = fa(env)
            if a == T:
                return T
            if a != F:
                raise ErlError("error", ("badarg", a))
            return fb(env)
        return f

    def e_send(self, node, tail):
        fa, fb = self.expr(node[1]), self.expr(node[2])
        rt = self.rt

        def f(env):
            pid = fa(env)
            msg = fb(env)
            rt.send(pid, msg)
            return msg
        return f

    def e_catch(self, node, tail):
        fe = self.expr(node[1])

        def f(env):
            try:
                return fe(env)
            except ErlError as e:
                if e.cls == "throw":
                    return e.reason
                if e.cls == "exit":
                    return ("EXIT", e.reason)
                return ("EXIT", (e.reason, NIL))
        return f

    def e_if(self, node, tail):
        clauses = [(self.guards(g), self.body(b, tail)) for g, b in node[1]]

        def f(env):
            for g, b in clauses:
                if g(env):
                    return b(env)
            raise ErlError("error", "if_clause")
        return f

    def e_case(self, node, tail):
        fe = self.expr(node[1])
        clauses = [self.clause(c, tail) for c in node[2]]
        simple = [(ms[0], g, b) for ms, g, b in clauses]

        def f(env):
            v = fe(env)
            for m, g, b in simple:
                trail = []
                if m(v, env, trail) and (g is None or g(env)):
                    return b(env)
                for n in trail:
                    del env[n]
            raise ErlError("error", ("case_clause", v))
        return f

    def e_receive(self, node, tail):
        clauses = [self.clause(c, False) for c in node[1]]
        ftime = self.expr(node[2]) if node[2] is not None else None
        fafter = self.body(node[3], False) if node[3] is not None else None
        rt = self.rt

        def f(env):
            mb = rt.current.mailbox
            for i, msg in enumerate(mb):
                for ms, g, b in clauses:
                    trail = []
                    if ms[0](msg, env, trail) and (g is None or g(env)):
                        del mb[i]
                        return b(env)
                    for n in trail:
                        del env[n]
            if ftime is not None:
                ftime(env)
                return fafter(env)
            raise RuntimeError("receive would block forever (no matching message, no `after`)")
        return f

    def e_lc(self, node, tail):
        fe = self.expr(node[1])
        quals = self._quals(node[2])

        def f(env):
            out = []
            self._run_quals(quals, 0, env, lambda e: out.append(fe(e)))
            return from_py(out)
        return f

    def e_bc(self, node, tail):
        seg = node[1]
        fe = self.e_bin(("bin", [seg]), False)
        quals = self._quals(node[2])

        def f(env):
            out = []
            self._run_quals(quals, 0, env, lambda e: out.append(fe(e)))
            return build_binary([("b", o) for o in out])
        return f

    def _quals(self, quals):
        comp = []
        for q in quals:
            if q[0] == "gen":
                comp.append(("gen", self.pattern(q[1]), self.expr(q[2]), tuple(pattern_vars(q[1], set()))))
            elif q[0] == "bgen":
                comp.append(("bgen", self.pattern(q[1]), self.expr(q[2]), q[1]))
            else:
                comp.append(("filter", self.expr(q[1])))
        return comp

    def _run_quals(self, quals, i, env, emit):
        if i == len(quals):
            emit(env)
            return
        q = quals[i]
        if q[0] == "gen":
            lst = q[2](env)
            while type(lst) is Cons:
                e2 = dict(env)
                for sv in q[3]:
                    e2.pop(sv, None)
                if q[1](lst.h, e2, []):
                    self._run_quals(quals, i + 1, e2, emit)
                lst = lst.t
            if lst is not NIL:
                raise ErlError("error", ("bad_generator", lst))
        elif q[0] == "bgen":
            # << Pattern >> <= Binary : the pattern is a binary pattern matched repeatedly against the front
            data = q[2](env)
            segs = q[3][1] if q[3][0] == "bin" else None
            if segs is None:
                raise NotImplementedError("binary generator with non-binary pattern")
            restvar = "$bgen_rest%d" % i
            pat = self.bin_pattern(segs + [(("var", restvar), None, [("bitstring", None)])])
            while True:
                e2 = dict(env)
                e2.pop(restvar, None)
                if not pat(data, e2, []):
                    break
                data = e2.pop(restvar)
                self._run_quals(quals, i + 1, e2, emit)
        else:
            v = q[1](env)
            if v == T:
                self._run_quals(quals, i + 1, env, emit)
            elif v != F:
                raise ErlError("error", ("bad_filter", v))

    def e_bin(self, node, tail):
        comp = []
        for (v, size, specs) in node[1]:
            typ, signed, endian, unit = _seg_spec(specs)
            default_type = not any(a in ("integer", "float", "binary", "bytes", "bitstring", "bits", "utf8", "utf16", "utf32") for a, _ in specs)
            if v[0] == "str":
                sizef = self.expr(size) if size is not None else None
                for c in v[1]:
                    comp.append(((lambda cc: (lambda env: cc))(c), sizef, typ, endian, unit))
                continue
            comp.append((self.expr(v), self.expr(size) if size is not None else None, typ, endian, unit))

        def f(env):
            parts = []
            for fv, fs, typ, endian, unit in comp:
                val = fv(env)
                if typ == "integer":
                    nb = 8 if fs is None else fs(env) * unit
                    if type(val) is not int or type(nb) is not int or nb < 0:
                        raise ErlError("error", "badarg")
                    parts.append(("i", val, nb, endian == "little"))
                elif typ == "binary" or typ == "bitstring":
                    if type(val) is bytes:
                        if fs is not None:
                            n = fs(env) * unit
                            if n % 8 or n // 8 > len(val):
                                raise ErlError("error", "badarg")
                            val = val[:n // 8]
                        parts.append(("b", val))
                    elif type(val) is Bits:
                        if typ == "binary":
                            raise ErlError("error", "badarg")
                        parts.append(("b", val))
                    else:
                        raise ErlError("error", "badarg")
                elif typ == "float":
                    nb = 64 if fs is None else fs(env) * unit
                    x = num_to_float(val) if type(val) in (int, float) else None
                    if x is None:
                        raise ErlError("error", "badarg")
                    raw = struct.pack(">d" if nb == 64 else ">f", x)
                    parts.append(("i", int.from_bytes(raw, "big"), nb, endian == "little"))
                elif typ == "utf8":
                    if type(val) is not int:
                        raise ErlError("error", "badarg")
                    try:
                        parts.append(("b", chr(val).encode("utf8")))
                    except Exception:
                        raise ErlError("error", "badarg")
                else:
                    raise NotImplementedError("binary construction type %s" % typ)
            return build_binary(parts)
        return f

    def e_map(self, node, tail):
        items = [(self.expr(k), self.expr(v)) for k, v, _ in node[1]]
        return lambda env: {k(env): v(env) for k, v in items}

    def e_mapupd(self, node, tail):
        fb = self.expr(node[1])
        items = [(self.expr(k), self.expr(v), op) for k, v, op in node[2]]

        def f(env):
            base = fb(env)
            if type(base) is not dict:
                raise ErlError("error", ("badmap", base))
            d = dict(base)
            for k, v, op in items:
                kk = k(env)
                if op == ":=" and kk not in d:
                    raise ErlError("error", ("badkey", kk))
                d[kk] = v(env)
            return d
        return f

    def e_record(self, node, tail):
        base, rname, fields = node[1], node[2], node[3]
        rdef = self.module.records[rname]
        idx = {f[0]: i + 1 for i, f in enumerate(rdef)}
        sets = [(idx[fn], self.expr(fe)) for fn, fe in fields]
        if base is None:
            defaults = []
            for fname, dtoks in rdef:
                if dtoks is None:
                    defaults.append(lambda env: "undefined")
                else:
                    from .parser import Parser
                    defaults.append(self.expr(Parser(dtoks, self.module.records, "record default").expr()))

            def f(env):
                vals = [rname] + [d(env) for d in defaults]
                for i, fe in sets:
                    vals[i] = fe(env)
                return tuple(vals)
            return f
        fb = self.expr(base)

        def f2(env):
            vals = list(fb(env))
            for i, fe in sets:
                vals[i] = fe(env)
            return tuple(vals)
        return f2

    def e_recfield(self, node, tail):
        fb = self.expr(node[1])
        rdef = self.module.records[node[2]]
        i = [f[0] for f in rdef].index(node[3]) + 1
        return lambda env: fb(env)[i]

    def e_fun(self, node, tail):
        selfname, clauses = node[1], node[2]
        # variables in a fun head are fresh: they shadow the enclosing bindings
        comp = [self.clause(c, True) + (tuple(pattern_vars(c[0], set())),) for c in clauses]
        arity = len(clauses[0][0])
        mod = self.module.name

        def f(env):
            fn = Fun(comp, arity, dict(env), "-fun-", mod, selfname)
            if selfname is not None:
                fn.env[selfname] = fn
            return fn
        return f

    def e_funref(self, node, tail):
        m, fn, ar = node[1], node[2], node[3]
        rt = self.rt
        if m is None:
            name, arity = fn[1], ar[1]
            modname = self.module.name

            def f(env):
                return rt.resolve(modname, name, arity, local=True)
            return f

        def val(x):
            if x[0] == "var":
                return lambda env: env[x[1]]
            return lambda env: x[1]
        fm, ff, fa = val(m), val(fn), val(ar)
        return lambda env: rt.resolve(fm(env), ff(env), fa(env))

    def e_remote(self, node, tail):
        raise NotImplementedError("bare remote reference")

    def e_call(self, node, tail):
        fexpr, args = node[1], node[2]
        fargs = [self.expr(a) for a in args]
        n = len(fargs)
        rt = self.rt
        apply_fun = rt.apply_fun
        if fexpr[0] == "atom":
            name = fexpr[1]
            modname = self.module.name
            cell = [None]

            def f(env):
                fn = cell[0]
                if fn is None:
                    fn = cell[0] = rt.resolve(modname, name, n, local=True)
                a = [fa(env) for fa in fargs]
                if tail and type(fn) is Fun:
                    return TailCall(fn, a)
                return apply_fun(fn, a)
            return f
        ff = self.expr(fexpr)

        def f2(env):
            fn = ff(env)
            a = [fa(env) for fa in fargs]
            if tail and type(fn) is Fun:
                if fn.arity != n:
                    raise ErlError("error", ("badarity", fn))
                return TailCall(fn, a)
            return apply_fun(fn, a)
        return f2

    def e_rcall(self, node, tail):
        m, fn, args = node[1], node[2], node[3]
        fargs = [self.expr(a) for a in args]
        n = len(fargs)
        rt = self.rt
        apply_fun = rt.apply_fun
        if m[0] == "atom" and fn[0] == "atom":
            mod, name = m[1], fn[1]
            cell = [None]

            def f(env):
                target = cell[0]
                if target is None:
                    target = cell[0] = rt.resolve(mod, name, n)
                a = [fa(env) for fa in fargs]
                if tail and type(target) is Fun:
                    return TailCall(target, a)
                return apply_fun(target, a)
            return f
        fm, ff = self.expr(m), self.expr(fn)

        def f2(env):
            target = rt.resolve(fm(env), ff(env), n)
            return apply_fun(target, [fa(env) for fa in fargs])
        return f2

    def e_try(self, node, tail):
        body = self.body(node[1], False)
        of_clauses = [self.clause(c, False) for c in node[2]] if node[2] is not None else None
        catches = []
        for cls, pat, g, b in node[3]:
            catches.append((self.pattern(cls), self.pattern(pat), self.guards(g), self.body(b, False)))
        after = self.body(node[4], False) if node[4] is not None else None

        def f(env):
            try:
                try:
                    v = body(env)
                except ErlError as e:
                    for mc, mp, g, b in catches:
                        trail = []
                        if mc(e.cls, env, trail) and mp(e.reason, env, trail) and (g is None or g(env)):
                            return b(env)
                        for nme in trail:
                            del env[nme]
                    raise
                if of_clauses is None:
                    return v
                for ms, g, b in of_clauses:
                    trail = []
                    if ms[0](v, env, trail) and (g is None or g(env)):
                        return b(env)
                    for nme in trail:
                        del env[nme]
                raise ErlError("error", ("try_clause", v))
            finally:
                if after is not None:
                    after(env)
        return f


# --------------------------------------------------------------------------- runtime
class Runtime(object):
    def __init__(self, src_dirs, budget=None):
        self.src_dirs = list(src_dirs)
        self.modules = {}
        self.bifs = {}            # (mod, name, arity) -> python callable(rt-independent, *args)
        self.current = Process()
        self.procs = {id(self.current.pid): self.current}
        self.steps = 0
        self.budget = budget
        self.ets = {}
        self.trace = None
        self.recent = []
        from . import bifs
        bifs.install(self)

    # ---- modules
    def load_module(self, name):
        if name in self.modules:
            return self.modules[name]
        path = None
        for d in self.src_dirs:
            p = os.path.join(d, name + ".erl")
            if os.path.exists(p):
                path = p
                break
        if path is None:
            raise ErlError("error", ("undef_module", name))
        pp = Preprocessor(path)
        mod = Module(name)
        mod.records = pp.records
        self.modules[name] = mod
        comp = Compiler(self, mod)
        parsed = {}
        for form in pp.forms:
            fname, arity, clauses = parse_function(form, pp.records, "%s.erl:%d" % (name, form[0].line))
            parsed[(fname, arity)] = clauses
        # compile lazily: keep the parsed clauses, compile at first use
        mod.parsed = parsed
        mod.compiler = comp
        return mod

    def module_fun(self, mod, name, arity):
        key = (name, arity)
        fn = mod.funs.get(key)
        if fn is not None:
            return fn
        clauses = mod.parsed.get(key)
        if clauses is None:
            return None
        fn = Fun(None, arity, None, name, mod.name)
        mod.funs[key] = fn
        fn.clauses = [mod.compiler.clause(c, True) for c in clauses]
        return fn

    def resolve(self, modname, name, arity, local=False):
        if local:
            mod = self.modules[modname]
            fn = self.module_fun(mod, name, arity)
            if fn is not None:
                return fn
            b = self.bifs.get(("erlang", name, arity))
            if b is not None:
                return b
            raise ErlError("error", ("undef", (modname, name, arity)))
        b = self.bifs.get((modname, name, arity))
        if b is not None:
            return b
        if modname in self.python_only_modules:
            raise ErlError("error", ("undef", (modname, name, arity)))
        mod = self.load_module(modname)
        fn = self.module_fun(mod, name, arity)
        if fn is None:
            raise ErlError("error", ("undef", (modname, name, arity)))
        return fn

    python_only_modules = set()

    def register(self, mod, name, arity, fn):
        self.bifs[(mod, name, arity)] = Bif(fn, arity, "%s:%s" % (mod, name))

    # ---- application with trampolined tail calls
    def apply_fun(self, fn, args):
        while True:
            tf = type(fn)
            if tf is Bif:
                if fn.arity != len(args) and fn.arity >= 0:
                    raise ErlError("error", ("badarity", fn))
                try:
                    return fn.fn(*args)
                except ErlError:
                    raise
                except (RecursionError, MemoryError, KeyboardInterrupt, BudgetExceeded, RuntimeError, NotImplementedError):
                    raise
                except Exception as e:
                    if self.trace:
                        print("bif %s raised %r" % (fn.name, e), file=sys.stderr)
                    raise ErlError("error", "badarg")
            if tf is not Fun:
                raise ErlError("error", ("badfun", fn))
            if fn.arity != len(args):
                raise ErlError("error", ("badarity", (fn, from_py(args))))
            self.steps += 1
            if self.budget is not None and self.steps > self.budget:
                raise BudgetExceeded()
            if self.trace:
                self.recent.append("%s:%s/%d%s" % (fn.module, fn.name, fn.arity, (" " + " ".join(fmt_term(a)[:60] for a in args)) if self.trace == 2 else ""))
                if len(self.recent) > 4000:
                    del self.recent[:2000]
            base = fn.env
            n = fn.arity
            res = None
            matched = False
            for cl in fn.clauses:
                ms, g, b = cl[0], cl[1], cl[2]
                env = dict(base) if base else {}
                if len(cl) > 3:
                    for sv in cl[3]:
                        env.pop(sv, None)
                ok = True
                trail = []
                for i in range(n):
                    if not ms[i](args[i], env, trail):
                        ok = False
                        break
                if ok and (g is None or g(env)):
                    matched = True
                    res = b(env)
                    break
            if not matched:
                raise ErlError("error", ("function_clause", (fn.module, fn.name, from_py(args))))
            if type(res) is TailCall:
                fn, args = res.fun, res.args
                continue
            return res

    def call(self, mod, name, *args):
        return self.apply_fun(self.resolve(mod, name, len(args)), list(args))

    # ---- processes
    def spawn(self, fn):
        p = Process()
        self.procs[id(p.pid)] = p
        saved = self.current
        self.current = p
        d0 = self.draws
        try:
            self.apply_fun(fn, [])
        except ErlError as e:
            self.last_crash = e
            if self.trace:
                print("spawned process died: %s; last calls: %s" % (e, " ".join(self.recent[-12:])), file=sys.stderr)
        finally:
            self.current = saved
            if hasattr(self, "child_draws"):
                self.child_draws.append(self.draws - d0)
        return p.pid

    def send(self, pid, msg):
        p = self.procs.get(id(pid))
        if p is not None:
            p.mailbox.append(msg)

    def eval(self, src, bindings=None):
        """evaluate an expression sequence given as Erlang source text (for tests and tools)"""
        mod = self.modules.get("$shell")
        if mod is None:
            mod = Module("$shell")
            mod.parsed = {}
            mod.compiler = Compiler(self, mod)
            self.modules["$shell"] = mod
        body = mod.compiler.body(parse_expr_string(src), False)
        env = dict(bindings or {})
        r = body(env)
        return r


def run_with_big_stack(fn, *args):
    """run fn(*args) in a thread with a large stack and recursion limit (deep non-tail recursion over long lists)"""
    import threading
    sys.setrecursionlimit(3000000)
    threading.stack_size(1 << 30)
    box = {}

    def runner():
        try:
            box["r"] = fn(*args)
        except BaseException as e:
            box["e"] = e
    t = threading.Thread(target=runner)
    t.start()
    t.join()
    if "e" in box:
        raise box["e"]
    return box.get("r")
