"""OTP functions the reference's hot path calls, for the reference runner (oracle/erlref, TEST INFRASTRUCTURE).

`random` is OTP stdlib's deprecated module (AS183 Wichmann-Hill): not part of /root/reference, restated from the
published algorithm.  `lists:sort/2` and friends whose exact algorithm is observable are run from Erlang source text
(otp/*.erl in this directory = a restatement of OTP's lists.erl merge sort); everything else here has results that do
not depend on the implementation."""
import math
import os
import zlib as _zlib
import base64 as _b64

from .terms import (NIL, Cons, Bits, Pid, Ref, ErlError, from_py, to_py, erl_cmp, exact_eq, is_list, fmt_term, cmp_key, float_to_str)

T, F = "true", "false"


def badarg():
    return ErlError("error", "badarg")


def iolist_to_bytes(x, out=None):
    top = out is None
    if top:
        out = bytearray()
    stack = [x]
    while stack:
        v = stack.pop()
        tv = type(v)
        if tv is int:
            if not 0 <= v <= 255:
                raise badarg()
            out.append(v)
        elif tv is bytes:
            out += v
        elif tv is Cons:
            # iterate the spine, pushing in reverse so order is preserved
            items = []
            c = v
            while type(c) is Cons:
                items.append(c.h)
                c = c.t
            if c is not NIL:
                items.append(c)      # improper tail must be a binary
                if type(c) is not bytes:
                    raise badarg()
            for it in reversed(items):
                stack.append(it)
        elif v is NIL:
            pass
        else:
            raise badarg()
    return bytes(out) if top else out


def flatten(l):
    out = []
    stack = [l]
    # depth-first, left to right
    def walk(x):
        while type(x) is Cons:
            h = x.h
            if type(h) is Cons:
                walk(h)
            elif h is not NIL:
                out.append(h)
            x = x.t
        if x is not NIL:
            raise badarg()
    walk(l)
    return from_py(out)


def chars_to_str(l):
    """flat-or-deep char list / binary -> python str (latin1 code points)"""
    if type(l) is bytes:
        return l.decode("latin1")
    if type(l) is str:
        return l
    return "".join(chr(c) for c in to_py(flatten(l)))


def str_to_chars(s):
    return from_py([ord(c) for c in s])


# --------------------------------------------------------------------------- io_lib:format
def format_impl(fmt, args):
    """returns a deep char list like io_lib:format (literal chars inline, one nested list per directive)"""
    f = to_py(flatten(fmt)) if type(fmt) is not bytes else list(fmt)
    if type(f) and f and type(f[0]) is not int:
        raise badarg()
    a = to_py(args)
    out = []
    i, n, ai = 0, len(f), 0
    while i < n:
        c = f[i]
        if c != 126:
            out.append(c)
            i += 1
            continue
        i += 1
        # ~F.P.PadModC
        width = prec = None
        pad = " "
        j = i
        def num(j):
            if j < n and f[j] == 42:
                return "*", j + 1
            k = j
            if k < n and f[k] == 45:
                k += 1
            while k < n and 48 <= f[k] <= 57:
                k += 1
            return ("".join(chr(x) for x in f[j:k]) or None), k
        w, j = num(j)
        if w == "*":
            w = str(a[ai]); ai += 1
        width = int(w) if w else None
        if j < n and f[j] == 46:
            p, j = num(j + 1)
            if p == "*":
                p = str(a[ai]); ai += 1
            prec = int(p) if p else None
            if j < n and f[j] == 46:
                pad = chr(f[j + 1])
                j += 2
        while j < n and chr(f[j]) in "tl":
            j += 1
        d = chr(f[j])
        i = j + 1
        if d == "n":
            out.append(10)
            continue
        if d == "~":
            out.append(126)
            continue
        arg = a[ai]
        ai += 1
        if d == "s":
            if type(arg) is str:
                s = arg
            else:
                s = chars_to_str(arg)
        elif d in "pw":
            s = fmt_term(arg) if d == "p" else fmt_term_w(arg)
        elif d in "Bb":
            base = prec if prec else 10
            prec = None
            s = int_to_base(arg, base, d == "B")
        elif d == "c":
            s = chr(arg)
        elif d in "fe g".replace(" ", ""):
            s = ("%." + str(prec if prec is not None else 6) + d) % arg
            prec = None
        elif d in "Xx#+":
            base = a[ai - 1]
            s = int_to_base(arg, 10, True)
        elif d == "i":
            continue
        else:
            raise badarg()
        if prec is not None and d == "s":
            s = s[:prec]
        if width is not None:
            if width < 0:
                s = s.ljust(-width, pad)
            else:
                s = s.rjust(width, pad)
        out.append(str_to_chars(s))
    return from_py(out)


def fmt_term_w(x):
    t = type(x)
    if t is Cons or x is NIL:
        try:
            items = to_py(x)
        except ErlError:
            return fmt_term(x)
        return "[" + ",".join(fmt_term_w(i) for i in items) + "]"
    if t is tuple:
        return "{" + ",".join(fmt_term_w(i) for i in x) + "}"
    if t is bytes:
        return "<<" + ",".join(str(c) for c in x) + ">>"
    return fmt_term(x)


def int_to_base(v, base, upper):
    if type(v) is not int:
        raise badarg()
    digs = "0123456789ABCDEFGHIJKLMNOPQRSTUVWXYZ" if upper else "0123456789abcdefghijklmnopqrstuvwxyz"
    if v == 0:
        return "0"
    neg = v < 0
    v = abs(v)
    out = []
    while v:
        out.append(digs[v % base])
        v //= base
    return ("-" if neg else "") + "".join(reversed(out))


# --------------------------------------------------------------------------- base64 (OTP stdlib base64.erl semantics)
_B64 = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/"
_B64D = {ord(c): i for i, c in enumerate(_B64)}
_WS = (9, 10, 13, 32)


def b64_decode(data):
    """base64:decode/1: whitespace (tab, LF, CR, space) is skipped anywhere; any other illegal character, a lone
    trailing group or misplaced padding raises (function_clause / badarg in OTP) -> ErlError error:badarg here."""
    if type(data) is bytes:
        chars = list(data)
    else:
        chars = to_py(data)
        for c in chars:
            if type(c) is not int:
                raise ErlError("error", "function_clause")
    cs = [c for c in chars if c not in _WS]
    out = bytearray()
    i, n = 0, len(cs)
    while i < n:
        grp = cs[i:i + 4]
        if len(grp) < 4:
            raise ErlError("error", "function_clause")
        c1, c2, c3, c4 = grp
        if c1 not in _B64D or c2 not in _B64D:
            raise ErlError("error", "function_clause")
        b1, b2 = _B64D[c1], _B64D[c2]
        if c3 == 61:
            # "xx==" must end the data
            if c4 != 61 or i + 4 != n:
                raise ErlError("error", "function_clause")
            out.append(((b1 << 2) | (b2 >> 4)) & 255)
            return bytes(out)
        if c3 not in _B64D:
            raise ErlError("error", "function_clause")
        b3 = _B64D[c3]
        if c4 == 61:
            if i + 4 != n:
                raise ErlError("error", "function_clause")
            out.append(((b1 << 2) | (b2 >> 4)) & 255)
            out.append(((b2 << 4) | (b3 >> 2)) & 255)
            return bytes(out)
        if c4 not in _B64D:
            raise ErlError("error", "function_clause")
        b4 = _B64D[c4]
        out.append(((b1 << 2) | (b2 >> 4)) & 255)
        out.append(((b2 << 4) | (b3 >> 2)) & 255)
        out.append(((b3 << 6) | b4) & 255)
        i += 4
    return bytes(out)


# --------------------------------------------------------------------------- gb_trees as a sorted association list
class GbTree(object):
    __slots__ = ("items",)

    def __init__(self, items=None):
        self.items = items or []      # sorted list of (key, value)

    def find(self, key):
        lo, hi = 0, len(self.items)
        while lo < hi:
            mid = (lo + hi) // 2
            c = erl_cmp(self.items[mid][0], key)
            if c == 0:
                return mid, True
            if c < 0:
                lo = mid + 1
            else:
                hi = mid
        return lo, False

    def enter(self, key, val):
        i, found = self.find(key)
        items = list(self.items)
        if found:
            items[i] = (key, val)
        else:
            items.insert(i, (key, val))
        return GbTree(items)


def install(rt):
    from . import interp
    from .interp import erl_int_to_float, num_to_float, Fun, Bif, make_bits, bits_of

    R = rt.register
    apply_fun = rt.apply_fun

    def reg(mod, name, arity):
        def deco(fn):
            R(mod, name, arity, fn)
            return fn
        return deco

    def boolfn(mod, name, pred):
        R(mod, name, 1, lambda x: T if pred(x) else F)

    # ---------------------------------------------------------------- erlang: type tests
    boolfn("erlang", "is_integer", lambda x: type(x) is int)
    boolfn("erlang", "is_float", lambda x: type(x) is float)
    boolfn("erlang", "is_number", lambda x: type(x) in (int, float))
    boolfn("erlang", "is_atom", lambda x: type(x) is str)
    boolfn("erlang", "is_boolean", lambda x: x == T or x == F)
    boolfn("erlang", "is_list", is_list)
    boolfn("erlang", "is_tuple", lambda x: type(x) is tuple)
    boolfn("erlang", "is_binary", lambda x: type(x) is bytes)
    boolfn("erlang", "is_bitstring", lambda x: type(x) in (bytes, Bits))
    boolfn("erlang", "is_function", lambda x: type(x) in (Fun, Bif))
    boolfn("erlang", "is_map", lambda x: type(x) is dict)
    boolfn("erlang", "is_pid", lambda x: type(x) is Pid)
    boolfn("erlang", "is_reference", lambda x: type(x) is Ref)
    R("erlang", "is_function", 2, lambda x, n: T if type(x) in (Fun, Bif) and x.arity == n else F)

    # ---------------------------------------------------------------- erlang: numbers
    def e_abs(x):
        if type(x) not in (int, float):
            raise badarg()
        return abs(x)
    R("erlang", "abs", 1, e_abs)

    def e_trunc(x):
        if type(x) is int:
            return x
        if type(x) is float:
            return int(x)
        raise badarg()
    R("erlang", "trunc", 1, e_trunc)

    def e_round(x):
        if type(x) is int:
            return x
        if type(x) is float:
            return int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))
        raise badarg()
    R("erlang", "round", 1, e_round)
    R("erlang", "float", 1, lambda x: num_to_float(x) if type(x) in (int, float) else (_ for _ in ()).throw(badarg()))
    R("erlang", "max", 2, lambda a, b: a if erl_cmp(a, b) >= 0 else b)
    R("erlang", "min", 2, lambda a, b: a if erl_cmp(a, b) <= 0 else b)

    # ---------------------------------------------------------------- erlang: lists / tuples
    def e_length(l):
        n = 0
        while type(l) is Cons:
            n += 1
            l = l.t
        if l is not NIL:
            raise badarg()
        return n
    R("erlang", "length", 1, e_length)

    def e_hd(l):
        if type(l) is not Cons:
            raise badarg()
        return l.h
    R("erlang", "hd", 1, e_hd)

    def e_tl(l):
        if type(l) is not Cons:
            raise badarg()
        return l.t
    R("erlang", "tl", 1, e_tl)

    def e_element(i, t):
        if type(t) is not tuple or type(i) is not int or not 1 <= i <= len(t):
            raise badarg()
        return t[i - 1]
    R("erlang", "element", 2, e_element)

    def e_setelement(i, t, v):
        if type(t) is not tuple or type(i) is not int or not 1 <= i <= len(t):
            raise badarg()
        return t[:i - 1] + (v,) + t[i:]
    R("erlang", "setelement", 3, e_setelement)

    def e_tuple_size(t):
        if type(t) is not tuple:
            raise badarg()
        return len(t)
    R("erlang", "tuple_size", 1, e_tuple_size)

    def e_size(x):
        if type(x) is tuple or type(x) is bytes:
            return len(x)
        if type(x) is Bits:
            return x.nbits // 8
        raise badarg()
    R("erlang", "size", 1, e_size)
    R("erlang", "list_to_tuple", 1, lambda l: tuple(to_py(l)))

    def e_tuple_to_list(t):
        if type(t) is not tuple:
            raise badarg()
        return from_py(t)
    R("erlang", "tuple_to_list", 1, e_tuple_to_list)

    # ---------------------------------------------------------------- erlang: binaries
    def e_byte_size(b):
        if type(b) is bytes:
            return len(b)
        if type(b) is Bits:
            return (b.nbits + 7) // 8
        raise badarg()
    R("erlang", "byte_size", 1, e_byte_size)

    def e_bit_size(b):
        if type(b) is bytes:
            return len(b) * 8
        if type(b) is Bits:
            return b.nbits
        raise badarg()
    R("erlang", "bit_size", 1, e_bit_size)

    def e_list_to_binary(l):
        if not is_list(l):
            raise badarg()
        return iolist_to_bytes(l)
    R("erlang", "list_to_binary", 1, e_list_to_binary)

    def e_iolist_to_binary(l):
        if type(l) is bytes:
            return l
        if not is_list(l):
            raise badarg()
        return iolist_to_bytes(l)
    R("erlang", "iolist_to_binary", 1, e_iolist_to_binary)
    R("erlang", "iolist_size", 1, lambda l: len(e_iolist_to_binary(l)))

    def e_list_to_bitstring(l):
        # iolist whose leaves may also be bitstrings (the last one unaligned)
        acc, n = 0, 0
        stack = [l]
        while stack:
            v = stack.pop()
            tv = type(v)
            if tv is int:
                if not 0 <= v <= 255:
                    raise badarg()
                acc = (acc << 8) | v
                n += 8
            elif tv is bytes or tv is Bits:
                val, nb = bits_of(v)
                acc = (acc << nb) | val
                n += nb
            elif tv is Cons:
                items = []
                c = v
                while type(c) is Cons:
                    items.append(c.h)
                    c = c.t
                if c is not NIL:
                    items.append(c)
                for it in reversed(items):
                    stack.append(it)
            elif v is NIL:
                pass
            else:
                raise badarg()
        return make_bits(acc, n)
    R("erlang", "list_to_bitstring", 1, e_list_to_bitstring)

    def e_binary_to_list(b):
        if type(b) is not bytes:
            raise badarg()
        return from_py(b)
    R("erlang", "binary_to_list", 1, e_binary_to_list)

    def e_binary_to_list3(b, s, e):
        if type(b) is not bytes or s < 1 or e > len(b) or s > e + 1:
            raise badarg()
        return from_py(b[s - 1:e])
    R("erlang", "binary_to_list", 3, e_binary_to_list3)

    def e_bitstring_to_list(b):
        if type(b) is bytes:
            return from_py(b)
        full = b.nbits // 8
        rest = b.nbits - full * 8
        head = (b.val >> rest).to_bytes(full, "big")
        return from_py(list(head) + [Bits(b.val & ((1 << rest) - 1), rest)])
    R("erlang", "bitstring_to_list", 1, e_bitstring_to_list)

    def e_binary_part(b, pos, ln):
        if type(b) is not bytes:
            raise badarg()
        if ln < 0:
            pos, ln = pos + ln, -ln
        if pos < 0 or pos + ln > len(b):
            raise badarg()
        return b[pos:pos + ln]
    R("erlang", "binary_part", 3, e_binary_part)
    R("erlang", "binary_part", 2, lambda b, pl: e_binary_part(b, pl[0], pl[1]))

    def e_split_binary(b, n):
        if type(b) is not bytes or not 0 <= n <= len(b):
            raise badarg()
        return (b[:n], b[n:])
    R("erlang", "split_binary", 2, e_split_binary)
    R("erlang", "crc32", 1, lambda b: _zlib.crc32(e_iolist_to_binary(b)) & 0xFFFFFFFF)
    R("erlang", "crc32", 2, lambda c, b: _zlib.crc32(e_iolist_to_binary(b), c) & 0xFFFFFFFF)
    R("erlang", "adler32", 1, lambda b: _zlib.adler32(e_iolist_to_binary(b)) & 0xFFFFFFFF)

    # ---------------------------------------------------------------- erlang: conversions
    def e_integer_to_list(i):
        if type(i) is not int:
            raise badarg()
        return str_to_chars(str(i))
    R("erlang", "integer_to_list", 1, e_integer_to_list)
    R("erlang", "integer_to_list", 2, lambda i, b: str_to_chars(int_to_base(i, b, True)))
    R("erlang", "integer_to_binary", 1, lambda i: str(i).encode() if type(i) is int else (_ for _ in ()).throw(badarg()))

    def e_list_to_integer(l, base=10):
        s = chars_to_str(l)
        if not s or s in "+-" or any(c.isspace() or c == "_" for c in s):
            raise badarg()
        try:
            return int(s, base)
        except ValueError:
            raise badarg()
    R("erlang", "list_to_integer", 1, e_list_to_integer)
    R("erlang", "list_to_integer", 2, e_list_to_integer)
    R("erlang", "binary_to_integer", 1, lambda b: e_list_to_integer(b))

    def e_list_to_float(l):
        s = chars_to_str(l)
        import re
        if not re.match(r"^[+-]?\d+\.\d+([eE][+-]?\d+)?$", s):
            raise badarg()
        return float(s)
    R("erlang", "list_to_float", 1, e_list_to_float)
    R("erlang", "binary_to_float", 1, e_list_to_float)

    def e_float_to_list(f):
        if type(f) is not float:
            raise badarg()
        return str_to_chars("%.20e" % f)
    R("erlang", "float_to_list", 1, e_float_to_list)

    def e_atom_to_list(a):
        if type(a) is not str:
            raise badarg()
        return str_to_chars(a)
    R("erlang", "atom_to_list", 1, e_atom_to_list)
    R("erlang", "list_to_atom", 1, lambda l: chars_to_str(l))
    R("erlang", "list_to_existing_atom", 1, lambda l: chars_to_str(l))
    R("erlang", "atom_to_binary", 2, lambda a, _e: a.encode("latin1"))
    R("erlang", "binary_to_atom", 2, lambda b, _e: b.decode("latin1"))

    # ---------------------------------------------------------------- erlang: process things
    R("erlang", "self", 0, lambda: rt.current.pid)

    def e_put(k, v):
        d = rt.current.dict
        key = _dkey(k)
        old = d.get(key, (None, "undefined"))[1]
        d[key] = (k, v)
        return old

    def _dkey(k):
        return k if type(k) in (str, int, bytes, tuple) else fmt_term(k)
    R("erlang", "put", 2, e_put)
    R("erlang", "get", 1, lambda k: rt.current.dict.get(_dkey(k), (None, "undefined"))[1])
    R("erlang", "erase", 1, lambda k: rt.current.dict.pop(_dkey(k), (None, "undefined"))[1])
    R("erlang", "spawn", 1, lambda fn: rt.spawn(fn))
    R("erlang", "spawn_link", 1, lambda fn: rt.spawn(fn))

    def e_throw(x):
        raise ErlError("throw", x)
    R("erlang", "throw", 1, e_throw)

    def e_error(x):
        raise ErlError("error", x)
    R("erlang", "error", 1, e_error)
    R("erlang", "error", 2, lambda x, _a: e_error(x))

    def e_exit(x):
        raise ErlError("exit", x)
    R("erlang", "exit", 1, e_exit)
    R("erlang", "exit", 2, lambda _p, _r: T)
    clock = [1500000000, 0, 0]

    def e_now():
        clock[2] += 1
        return (clock[0] // 1000000, clock[0] % 1000000, clock[2])
    R("erlang", "now", 0, e_now)
    R("os", "timestamp", 0, e_now)
    R("erlang", "timestamp", 0, e_now)
    R("erlang", "make_ref", 0, lambda: Ref())
    R("erlang", "node", 0, lambda: "nonode@nohost")
    R("erlang", "apply", 3, lambda m, f, a: apply_fun(rt.resolve(m, f, len(to_py(a))), to_py(a)))
    R("erlang", "apply", 2, lambda f, a: apply_fun(f, to_py(a)))
    R("erlang", "process_info", 2, lambda _p, what: (what, 100000) if what == "memory" else (what, NIL))
    R("erlang", "garbage_collect", 0, lambda: T)
    R("erlang", "send", 2, lambda p, m: (rt.send(p, m), m)[1])
    R("erlang", "whereis", 1, lambda _n: "undefined")
    R("erlang", "register", 2, lambda _n, _p: T)
    R("erlang", "map_size", 1, lambda m: len(m))
    R("erlang", "system_time", 0, lambda: 1500000000000000000)
    R("erlang", "system_time", 1, lambda _u: 1500000000)
    R("erlang", "monotonic_time", 0, lambda: 0)
    R("erlang", "unique_integer", 0, lambda: 1)
    R("erlang", "term_to_binary", 1, lambda t: fmt_term(t).encode("latin1"))
    R("erlang", "phash2", 2, lambda t, n: hash(fmt_term(t)) % n)

    # ---------------------------------------------------------------- random (AS183), OTP stdlib `random`
    def seed_get():
        s = rt.current.dict.get("random_seed")
        if s is None:
            return (3172, 9814, 20125)
        return s[1]

    def seed_put(s):
        rt.current.dict["random_seed"] = ("random_seed", s)

    def r_seed3(a1, a2, a3):
        old = rt.current.dict.get("random_seed")
        seed_put((abs(a1) % (30269 - 1) + 1, abs(a2) % (30307 - 1) + 1, abs(a3) % (30323 - 1) + 1))
        return "undefined" if old is None else old[1]

    def r_uniform0():
        a1, a2, a3 = seed_get()
        b1 = (a1 * 171) % 30269
        b2 = (a2 * 172) % 30307
        b3 = (a3 * 170) % 30323
        seed_put((b1, b2, b3))
        rt.draws += 1
        r = b1 / 30269 + b2 / 30307 + b3 / 30323
        return r - int(r)

    def r_uniform1(n):
        if type(n) is not int or n < 1:
            raise ErlError("error", "function_clause")
        return int(r_uniform0() * num_to_float(n)) + 1
    rt.draws = 0
    R("random", "seed", 1, lambda s: r_seed3(*s))
    R("random", "seed", 3, r_seed3)
    R("random", "seed", 0, lambda: (rt.current.dict.pop("random_seed", None), "undefined")[1])
    R("random", "uniform", 0, r_uniform0)
    R("random", "uniform", 1, r_uniform1)

    # ---------------------------------------------------------------- lists (order-insensitive parts; sort/2 runs from Erlang text)
    def l_reverse(l, tail=NIL):
        out = tail
        while type(l) is Cons:
            out = Cons(l.h, out)
            l = l.t
        if l is not NIL:
            raise badarg()
        return out
    R("lists", "reverse", 1, l_reverse)
    R("lists", "reverse", 2, l_reverse)
    R("lists", "flatten", 1, flatten)
    R("lists", "flatten", 2, lambda l, t: interp.list_append(flatten(l), t))
    R("lists", "append", 2, interp.list_append)

    def l_append1(ll):
        out = NIL
        for l in reversed(to_py(ll)):
            out = interp.list_append(l, out)
        return out
    R("lists", "append", 1, l_append1)
    R("lists", "concat", 1, lambda ll: str_to_chars("".join(x if type(x) is str else str(x) if type(x) in (int, float) else chars_to_str(x) for x in to_py(ll))))

    def l_foldl(f, acc, l):
        while type(l) is Cons:
            acc = apply_fun(f, [l.h, acc])
            l = l.t
        if l is not NIL:
            raise ErlError("error", "function_clause")
        return acc
    R("lists", "foldl", 3, l_foldl)

    def l_foldr(f, acc, l):
        for x in reversed(to_py(l)):
            acc = apply_fun(f, [x, acc])
        return acc
    R("lists", "foldr", 3, l_foldr)

    def l_map(f, l):
        out = []
        while type(l) is Cons:
            out.append(apply_fun(f, [l.h]))
            l = l.t
        if l is not NIL:
            raise ErlError("error", "function_clause")
        return from_py(out)
    R("lists", "map", 2, l_map)

    def l_foreach(f, l):
        for x in to_py(l):
            apply_fun(f, [x])
        return "ok"
    R("lists", "foreach", 2, l_foreach)

    def l_filter(f, l):
        return from_py([x for x in to_py(l) if apply_fun(f, [x]) == T])
    R("lists", "filter", 2, l_filter)

    def l_filtermap(f, l):
        out = []
        for x in to_py(l):
            r = apply_fun(f, [x])
            if r == T:
                out.append(x)
            elif type(r) is tuple and r[0] == T:
                out.append(r[1])
        return from_py(out)
    R("lists", "filtermap", 2, l_filtermap)

    def l_flatmap(f, l):
        out = NIL
        for x in reversed(to_py(l)):
            out = interp.list_append(apply_fun(f, [x]), out)
        return out
    R("lists", "flatmap", 2, l_flatmap)

    def l_mapfoldl(f, acc, l):
        out = []
        while type(l) is Cons:
            r = apply_fun(f, [l.h, acc])
            if type(r) is not tuple or len(r) != 2:
                raise ErlError("error", ("badmatch", r))
            out.append(r[0])
            acc = r[1]
            l = l.t
        return (from_py(out), acc)
    R("lists", "mapfoldl", 3, l_mapfoldl)

    def l_seq(a, b, step=1):
        if type(a) is not int or type(b) is not int:
            raise ErlError("error", "function_clause")
        if step == 1 and b < a - 1:
            raise ErlError("error", "function_clause")
        return from_py(list(range(a, b + (1 if step > 0 else -1), step)))
    R("lists", "seq", 2, l_seq)
    R("lists", "seq", 3, l_seq)

    def l_nth(n, l):
        if type(n) is not int or n < 1:
            raise ErlError("error", "function_clause")
        while n > 1:
            if type(l) is not Cons:
                raise ErlError("error", "function_clause")
            l = l.t
            n -= 1
        if type(l) is not Cons:
            raise ErlError("error", "function_clause")
        return l.h
    R("lists", "nth", 2, l_nth)

    def l_nthtail(n, l):
        if type(n) is not int or n < 0:
            raise ErlError("error", "function_clause")
        while n > 0:
            if type(l) is not Cons:
                raise ErlError("error", "function_clause")
            l = l.t
            n -= 1
        return l
    R("lists", "nthtail", 2, l_nthtail)

    def l_sublist2(l, n):
        if type(n) is not int or n < 0:
            raise ErlError("error", "function_clause")
        out = []
        while n > 0 and type(l) is Cons:
            out.append(l.h)
            l = l.t
            n -= 1
        return from_py(out)
    R("lists", "sublist", 2, l_sublist2)

    def l_sublist3(l, s, n):
        if type(s) is not int or s < 1:
            raise ErlError("error", "function_clause")
        return l_sublist2(l_nthtail_lenient(s - 1, l), n)

    def l_nthtail_lenient(n, l):
        # lists:sublist/3 accepts Start = length+1 .. and raises beyond
        while n > 0:
            if type(l) is not Cons:
                if n == 0:
                    return l
                raise ErlError("error", "function_clause")
            l = l.t
            n -= 1
        return l
    R("lists", "sublist", 3, l_sublist3)

    def l_split(n, l):
        if type(n) is not int or n < 0:
            raise badarg()
        out = []
        while n > 0:
            if type(l) is not Cons:
                raise badarg()
            out.append(l.h)
            l = l.t
            n -= 1
        return (from_py(out), l)
    R("lists", "split", 2, l_split)
    R("lists", "member", 2, lambda x, l: T if any(exact_eq(x, y) for y in to_py(l)) else F)

    def l_last(l):
        if type(l) is not Cons:
            raise ErlError("error", "function_clause")
        while type(l.t) is Cons:
            l = l.t
        return l.h
    R("lists", "last", 1, l_last)
    R("lists", "sum", 1, lambda l: sum(to_py(l)))
    R("lists", "max", 1, lambda l: max(to_py(l), key=cmp_key))
    R("lists", "min", 1, lambda l: min(to_py(l), key=cmp_key))
    R("lists", "duplicate", 2, lambda n, x: from_py([x] * n))
    R("lists", "zip", 2, lambda a, b: from_py(list(zip(to_py(a), to_py(b)))))
    R("lists", "unzip", 1, lambda l: (from_py([x[0] for x in to_py(l)]), from_py([x[1] for x in to_py(l)])))
    R("lists", "all", 2, lambda f, l: T if all(apply_fun(f, [x]) == T for x in to_py(l)) else F)
    R("lists", "any", 2, lambda f, l: T if any(apply_fun(f, [x]) == T for x in to_py(l)) else F)
    R("lists", "droplast", 1, lambda l: from_py(to_py(l)[:-1]))

    def l_delete(x, l):
        out = to_py(l)
        for i, y in enumerate(out):
            if exact_eq(x, y):
                del out[i]
                break
        return from_py(out)
    R("lists", "delete", 2, l_delete)

    def l_keyfind(k, n, l):
        for t in to_py(l):
            if type(t) is tuple and len(t) >= n and erl_cmp(t[n - 1], k) == 0:
                return t
        return F
    R("lists", "keyfind", 3, l_keyfind)
    R("lists", "keymember", 3, lambda k, n, l: F if l_keyfind(k, n, l) == F else T)

    def l_keysearch(k, n, l):
        r = l_keyfind(k, n, l)
        return F if r == F else ("value", r)
    R("lists", "keysearch", 3, l_keysearch)

    def l_sort1(l):
        # term order is total, equal terms are indistinguishable: any correct sort gives OTP's result
        return from_py(sorted(to_py(l), key=cmp_key))
    R("lists", "sort", 1, l_sort1)

    def l_usort1(l):
        out = []
        for x in sorted(to_py(l), key=cmp_key):
            if not out or erl_cmp(out[-1], x) != 0:
                out.append(x)
        return from_py(out)
    R("lists", "usort", 1, l_usort1)
    R("lists", "sort", 2, lambda f, l: rt.call("otp_lists", "sort", f, l))
    R("lists", "takewhile", 2, lambda f, l: from_py(_takewhile(f, l)))

    def _takewhile(f, l):
        out = []
        for x in to_py(l):
            if apply_fun(f, [x]) != T:
                break
            out.append(x)
        return out

    def l_dropwhile(f, l):
        while type(l) is Cons and apply_fun(f, [l.h]) == T:
            l = l.t
        return l
    R("lists", "dropwhile", 2, l_dropwhile)

    def l_partition(f, l):
        a, b = [], []
        for x in to_py(l):
            (a if apply_fun(f, [x]) == T else b).append(x)
        return (from_py(a), from_py(b))
    R("lists", "partition", 2, l_partition)

    # ---------------------------------------------------------------- maps
    def hk(k):
        return k

    def m_get2(k, m):
        if type(m) is not dict:
            raise ErlError("error", ("badmap", m))
        if k not in m:
            raise ErlError("error", ("badkey", k))
        return m[k]
    R("maps", "get", 2, m_get2)

    def m_get3(k, m, d):
        if type(m) is not dict:
            raise ErlError("error", ("badmap", m))
        return m.get(k, d)
    R("maps", "get", 3, m_get3)

    def m_put(k, v, m):
        if type(m) is not dict:
            raise ErlError("error", ("badmap", m))
        d = dict(m)
        d[k] = v
        return d
    R("maps", "put", 3, m_put)
    R("maps", "new", 0, lambda: {})
    R("maps", "from_list", 1, lambda l: {t[0]: t[1] for t in to_py(l)})
    R("maps", "to_list", 1, lambda m: from_py(sorted(m.items(), key=lambda kv: cmp_key(kv[0]))))
    R("maps", "is_key", 2, lambda k, m: T if k in m else F)
    R("maps", "keys", 1, lambda m: from_py(sorted(m.keys(), key=cmp_key)))
    R("maps", "values", 1, lambda m: from_py([m[k] for k in sorted(m.keys(), key=cmp_key)]))
    R("maps", "size", 1, lambda m: len(m))
    R("maps", "find", 2, lambda k, m: ("ok", m[k]) if k in m else "error")

    def m_remove(k, m):
        d = dict(m)
        d.pop(k, None)
        return d
    R("maps", "remove", 2, m_remove)

    def m_merge(a, b):
        d = dict(a)
        d.update(b)
        return d
    R("maps", "merge", 2, m_merge)

    def m_fold(f, acc, m):
        for k in sorted(m.keys(), key=cmp_key):
            acc = apply_fun(f, [k, m[k], acc])
        return acc
    R("maps", "fold", 3, m_fold)

    # ---------------------------------------------------------------- gb_trees
    R("gb_trees", "empty", 0, lambda: GbTree())
    R("gb_trees", "enter", 3, lambda k, v, t: t.enter(k, v))

    def g_insert(k, v, t):
        _, found = t.find(k)
        if found:
            raise ErlError("error", ("key_exists", k))
        return t.enter(k, v)
    R("gb_trees", "insert", 3, g_insert)

    def g_lookup(k, t):
        i, found = t.find(k)
        return ("value", t.items[i][1]) if found else "none"
    R("gb_trees", "lookup", 2, g_lookup)

    def g_get(k, t):
        i, found = t.find(k)
        if not found:
            raise ErlError("error", "function_clause")
        return t.items[i][1]
    R("gb_trees", "get", 2, g_get)
    R("gb_trees", "is_defined", 2, lambda k, t: T if t.find(k)[1] else F)
    R("gb_trees", "to_list", 1, lambda t: from_py(list(t.items)))
    R("gb_trees", "keys", 1, lambda t: from_py([k for k, _ in t.items]))
    R("gb_trees", "values", 1, lambda t: from_py([v for _, v in t.items]))
    R("gb_trees", "size", 1, lambda t: len(t.items))
    R("gb_trees", "is_empty", 1, lambda t: T if not t.items else F)

    def g_update(k, v, t):
        _, found = t.find(k)
        if not found:
            raise ErlError("error", "function_clause")
        return t.enter(k, v)
    R("gb_trees", "update", 3, g_update)

    def g_delete_any(k, t):
        i, found = t.find(k)
        if not found:
            return t
        items = list(t.items)
        del items[i]
        return GbTree(items)
    R("gb_trees", "delete_any", 2, g_delete_any)
    R("gb_trees", "delete", 2, g_delete_any)
    R("gb_trees", "from_orddict", 1, lambda l: GbTree(list(to_py(l))))

    # ---------------------------------------------------------------- io / io_lib / file / timer / misc side-effect modules
    R("io_lib", "format", 2, format_impl)
    R("io_lib", "fwrite", 2, format_impl)

    def io_format(*args):
        if rt.trace:
            try:
                fmt, a = (args[-2], args[-1]) if len(args) >= 2 else (args[0], NIL)
                import sys
                sys.stderr.write(chars_to_str(format_impl(fmt, a)))
            except Exception:
                pass
        return "ok"
    R("io", "format", 1, io_format)
    R("io", "format", 2, io_format)
    R("io", "format", 3, io_format)
    R("io", "fwrite", 1, io_format)
    R("io", "fwrite", 2, io_format)
    R("io", "fwrite", 3, io_format)
    R("io", "write", 1, lambda _x: "ok")
    R("io", "put_chars", 1, lambda _x: "ok")
    R("io", "put_chars", 2, lambda _d, _x: "ok")
    R("io", "setopts", 1, lambda _o: "ok")
    R("io", "setopts", 2, lambda _d, _o: "ok")
    # files: only virtual ones (rt.vfs: path string -> bytes) and standard_io (rt.stdin) exist -- enough for the file and stdin
    # generators (src/erlamsa_gen.erl:59-121); opening anything else fails like a missing file
    class FileHandle(object):
        def __init__(self, data):
            self.data = data
            self.pos = 0
    rt.vfs = {}
    rt.stdin = FileHandle(b"")

    rt.vfs_out = {}                       # files written by the run (erlamsa_out:file_writer/1, -o "dir/%n"): name -> bytes

    class OutHandle(object):
        def __init__(self, name):
            self.name = name
            rt.vfs_out[name] = b""

    def f_open(path, modes):
        name = chars_to_str(path)
        if "write" in [m for m in to_py(modes) if isinstance(m, str)]:
            return ("ok", OutHandle(name))
        if name not in rt.vfs:
            return ("error", "enoent")
        return ("ok", FileHandle(rt.vfs[name]))

    def f_write(fd, data):
        if isinstance(fd, OutHandle):
            rt.vfs_out[fd.name] += e_iolist_to_binary(data)
        return "ok"

    def f_read(fd, n):
        h = rt.stdin if fd == "standard_io" else fd
        if not isinstance(h, FileHandle):
            return ("error", "badarg")
        if h.pos >= len(h.data):
            return "eof"
        out = h.data[h.pos:h.pos + n]
        h.pos += len(out)
        return ("ok", out)
    R("file", "open", 2, f_open)
    R("file", "read", 2, f_read)
    R("file", "close", 1, lambda _fd: "ok")
    R("file", "write", 2, f_write)
    R("file", "write_file", 2, lambda _p, _d: "ok")
    R("file", "write_file", 3, lambda _p, _d, _m: "ok")
    R("timer", "sleep", 1, lambda _t: "ok")
    R("timer", "now_diff", 2, lambda a, b: ((a[0] - b[0]) * 1000000 + (a[1] - b[1])) * 1000000 + (a[2] - b[2]))
    R("crypto", "strong_rand_bytes", 1, lambda n: os.urandom(n))
    R("math", "pow", 2, lambda a, b: _fl(lambda: math.pow(num_to_float(a), num_to_float(b))))
    R("math", "exp", 1, lambda a: _fl(lambda: math.exp(num_to_float(a))))
    R("math", "log", 1, lambda a: _fl(lambda: math.log(num_to_float(a))))
    R("math", "log2", 1, lambda a: _fl(lambda: math.log2(num_to_float(a))))
    R("math", "log10", 1, lambda a: _fl(lambda: math.log10(num_to_float(a))))
    R("math", "sqrt", 1, lambda a: _fl(lambda: math.sqrt(num_to_float(a))))
    R("math", "floor", 1, lambda a: float(math.floor(a)))
    R("math", "ceil", 1, lambda a: float(math.ceil(a)))

    def _fl(thunk):
        try:
            r = thunk()
        except (OverflowError, ValueError, ZeroDivisionError):
            raise ErlError("error", "badarith")
        if r != r or r in (float("inf"), float("-inf")):
            raise ErlError("error", "badarith")
        return r

    # ets: only what get_ssrf_ep_unsafe reads (table global_config); no table -> badarg like the real thing
    def ets_match(tab, pat):
        t = rt.ets.get(tab)
        if t is None:
            raise badarg()
        key = pat[0]
        if key in t:
            return from_py([from_py([t[key]])])
        return NIL
    R("ets", "match", 2, ets_match)

    def ets_insert(tab, obj):
        rt.ets.setdefault(tab, {})[obj[0]] = obj[1]
        return T
    R("ets", "insert", 2, ets_insert)
    R("ets", "new", 2, lambda name, _o: (rt.ets.setdefault(name, {}), name)[1])
    R("inet", "ntoa", 1, lambda ip: str_to_chars(".".join(str(x) for x in ip)))

    # erlamsa_logger is a process-based logger outside the hot path: calls are no-ops here
    rt.logged_data = []
    rt.logged_cases = {}                   # case number I -> written bytes (the [I, N] argument list of the same call)

    def log_data(*a):
        if a and type(a[-1]) is bytes:
            rt.logged_data.append(a[-1])       # erlamsa_main logs every written test case (:199): the harness reads it back here
            if len(a) >= 2:
                try:
                    rt.logged_cases[to_py(a[-2])[0]] = a[-1]
                except Exception:
                    pass
        return "ok"
    for ar in (2, 3, 4, 5):
        R("erlamsa_logger", "log", ar, lambda *a: "ok")
        R("erlamsa_logger", "log_data", ar, log_data)

    # ---------------------------------------------------------------- string
    def s_tokens(s, seps):
        sp = set(to_py(seps))
        out, cur = [], []
        for c in to_py(s):
            if c in sp:
                if cur:
                    out.append(from_py(cur))
                    cur = []
            else:
                cur.append(c)
        if cur:
            out.append(from_py(cur))
        return from_py(out)
    R("string", "tokens", 2, s_tokens)

    def s_join(l, sep):
        items = to_py(l)
        if not items:
            return NIL
        out = []
        sp = to_py(sep)
        for i, it in enumerate(items):
            if i:
                out.extend(sp)
            out.extend(to_py(it))
        return from_py(out)
    R("string", "join", 2, s_join)

    def s_to_lower(s):
        if type(s) is int:
            return s + 32 if (65 <= s <= 90 or 192 <= s <= 214 or 216 <= s <= 222) else s
        return from_py([s_to_lower(c) if type(c) is int else c for c in to_py(s)])
    R("string", "to_lower", 1, s_to_lower)

    def s_to_upper(s):
        if type(s) is int:
            return s - 32 if (97 <= s <= 122 or 224 <= s <= 246 or 248 <= s <= 254) else s
        return from_py([s_to_upper(c) if type(c) is int else c for c in to_py(s)])
    R("string", "to_upper", 1, s_to_upper)
    R("string", "len", 1, lambda s: len(to_py(s)))
    R("string", "concat", 2, interp.list_append)

    def s_str(s, sub):
        a, b = to_py(s), to_py(sub)
        for i in range(len(a) - len(b) + 1):
            if a[i:i + len(b)] == b:
                return i + 1
        return 0
    R("string", "str", 2, s_str)

    def s_chr(s, c):
        for i, x in enumerate(to_py(s)):
            if x == c:
                return i + 1
        return 0
    R("string", "chr", 2, s_chr)
    R("string", "substr", 2, lambda s, st: from_py(to_py(s)[st - 1:]))
    R("string", "substr", 3, lambda s, st, ln: from_py(to_py(s)[st - 1:st - 1 + ln]))
    R("string", "strip", 1, lambda s: str_to_chars(chars_to_str(s).strip(" ")))
    R("string", "to_integer", 1, lambda s: _to_integer(s))

    def _to_integer(s):
        import re
        st = chars_to_str(s)
        m = re.match(r"^[+-]?\d+", st)
        if not m:
            return ("error", "no_integer")
        return (int(m.group(0)), str_to_chars(st[m.end():]))

    # ---------------------------------------------------------------- base64 / zlib / zip
    R("base64", "decode", 1, b64_decode)
    R("base64", "decode_to_string", 1, lambda d: from_py(b64_decode(d)))
    R("base64", "encode", 1, lambda d: _b64.b64encode(e_iolist_to_binary(d)))
    R("base64", "encode_to_string", 1, lambda d: from_py(_b64.b64encode(e_iolist_to_binary(d))))

    def z_err(thunk):
        try:
            return thunk()
        except Exception:
            raise ErlError("error", "data_error")
    R("zlib", "gunzip", 1, lambda b: z_err(lambda: _zlib.decompress(e_iolist_to_binary(b), 31)))
    R("zlib", "gzip", 1, lambda b: rt.unsupported("zlib:gzip"))
    R("zlib", "uncompress", 1, lambda b: z_err(lambda: _zlib.decompress(e_iolist_to_binary(b))))
    R("zlib", "compress", 1, lambda b: rt.unsupported("zlib:compress"))
    R("zlib", "open", 0, lambda: Ref())
    R("zlib", "close", 1, lambda _z: "ok")
    # inflateInit/1 is the ZLIB format (window bits 15: CMF/FLG header + adler32), not raw deflate; /2 takes the window bits
    # (negative = raw). inflate/2 is a streaming call: output so far comes back without an error when the input just ends,
    # invalid data raises data_error.
    z_wbits = {}

    def z_inflate_init(z, w=15):
        z_wbits[id(z)] = w
        return "ok"

    def z_inflate(z, b):
        d = _zlib.decompressobj(z_wbits.get(id(z), 15))
        out = z_err(lambda: d.decompress(e_iolist_to_binary(b)))
        return from_py([out] if out else [])
    R("zlib", "inflateInit", 1, z_inflate_init)
    R("zlib", "inflateInit", 2, z_inflate_init)
    R("zlib", "inflate", 2, z_inflate)
    R("zlib", "inflateEnd", 1, lambda _z: "ok")
    R("zlib", "deflateInit", 1, lambda _z: "ok")
    R("zlib", "deflateInit", 2, lambda _z, _l: "ok")
    def z_deflate(_z, b, flush):
        # only the call erlamsa_patterns makes (deflateInit(Z, default); deflate(Z, Data, finish)): one complete zlib stream at the
        # default level. Small inputs compress to the same bytes under every zlib release; larger ones are not relied upon (the oracle
        # flags really compressed inputs, so they are never compared).
        data = e_iolist_to_binary(b)
        if flush != "finish" or len(data) > 64:
            rt.unsupported("zlib:deflate")
        return from_py([_zlib.compress(data)])
    R("zlib", "deflate", 3, z_deflate)
    R("zlib", "deflateEnd", 1, lambda _z: "ok")

    def zip_foldl(_f, _acc, spec):
        data = spec[1] if type(spec) is tuple else None
        if type(data) is bytes and b"PK\x05\x06" in data:
            rt.unsupported("zip:foldl on a real archive")
        return ("error", "bad_eocd")      # zip:foldl on data without an end-of-central-directory record
    R("zip", "foldl", 3, zip_foldl)
    R("zip", "create", 3, lambda *_a: rt.unsupported("zip:create"))
    R("zip", "unzip", 2, lambda data, _o: rt.unsupported("zip:unzip") if (type(data) is bytes and b"PK\x05\x06" in data) else ("error", "bad_eocd"))
    R("zip", "zip", 3, lambda *_a: rt.unsupported("zip:zip"))

    # re: only what erlamsa_out:file_writer/1 does with its "%n" template (compile/1, split/2: binaries, empty trailing piece kept)
    import re as _re
    R("re", "compile", 1, lambda p: ("ok", ("re_pattern", _re.compile(_re.escape(e_iolist_to_binary(p)) if e_iolist_to_binary(p) == b"%n" else e_iolist_to_binary(p)))))
    R("re", "split", 2, lambda subj, mp: from_py(list(mp[1].split(e_iolist_to_binary(subj)))))

    class Unsupported(RuntimeError):
        pass
    rt.Unsupported = Unsupported

    def unsupported(what):
        raise Unsupported(what)
    rt.unsupported = unsupported

    rt.python_only_modules = {"erlang", "lists", "maps", "random", "gb_trees", "io_lib", "io", "file", "timer", "crypto", "math", "ets",
                              "inet", "erlamsa_logger", "string", "base64", "zlib", "zip", "os", "re"}
