"""Erlang terms for the reference runner (oracle/erlref): TEST INFRASTRUCTURE, never imported by the product.

Representation:  integer -> int, float -> float, atom -> str, tuple -> tuple, binary -> bytes,
unaligned bitstring -> Bits, list -> Cons / NIL, fun -> Fun (interp.py) or Bif, map -> dict, pid -> Pid.
"""


class Nil(object):
    __slots__ = ()

    def __repr__(self):
        return "[]"

    def __iter__(self):
        return iter(())


NIL = Nil()


class Cons(object):
    __slots__ = ("h", "t")

    def __init__(self, h, t):
        self.h = h
        self.t = t

    def __iter__(self):
        c = self
        while type(c) is Cons:
            yield c.h
            c = c.t
        if c is not NIL:
            raise ErlError("error", "badarg")

    def __eq__(self, o):
        return type(o) is Cons and exact_eq(self, o)

    def __ne__(self, o):
        return not self.__eq__(o)

    def __hash__(self):
        h = 7
        c = self
        n = 0
        while type(c) is Cons and n < 64:
            x = c.h
            h = (h * 1000003) ^ (hash(x) if type(x) is not Cons else 17)
            c = c.t
            n += 1
        return h & 0xFFFFFFFFFFFF

    def __repr__(self):
        return fmt_term(self)


class Bits(object):
    """a bitstring whose length is not a multiple of 8"""
    __slots__ = ("val", "nbits")

    def __init__(self, val, nbits):
        self.val = val
        self.nbits = nbits

    def __eq__(self, o):
        return type(o) is Bits and o.val == self.val and o.nbits == self.nbits

    def __hash__(self):
        return hash((self.val, self.nbits))

    def __repr__(self):
        return "<<%d:%d>>" % (self.val, self.nbits)


class Pid(object):
    _n = 0

    def __init__(self):
        Pid._n += 1
        self.n = Pid._n

    def __repr__(self):
        return "<0.%d.0>" % self.n


class Ref(object):
    _n = 0

    def __init__(self):
        Ref._n += 1
        self.n = Ref._n

    def __repr__(self):
        return "#Ref<%d>" % self.n


class ErlError(Exception):
    def __init__(self, cls, reason):
        Exception.__init__(self, cls, reason)
        self.cls = cls
        self.reason = reason

    def __str__(self):
        return "%s:%s" % (self.cls, fmt_term(self.reason))


def badarg():
    return ErlError("error", "badarg")


def from_py(seq, tail=NIL):
    """python sequence -> erlang list"""
    l = tail
    if not isinstance(seq, (list, tuple, bytes, bytearray)):
        seq = list(seq)
    for x in reversed(seq):
        l = Cons(x, l)
    return l


def to_py(l):
    """proper erlang list -> python list (badarg when improper)"""
    out = []
    ap = out.append
    while type(l) is Cons:
        ap(l.h)
        l = l.t
    if l is not NIL:
        raise badarg()
    return out


def is_list(x):
    return x is NIL or type(x) is Cons


def erl_bool(b):
    return "true" if b else "false"


def is_number(x):
    t = type(x)
    return t is int or t is float


def type_rank(x):
    t = type(x)
    if t is int or t is float:
        return 0
    if t is str:
        return 1
    if t is Ref:
        return 2
    if t is tuple:
        return 6
    if t is dict:
        return 7
    if x is NIL:
        return 8
    if t is Cons:
        return 9
    if t is bytes or t is Bits:
        return 10
    if t is Pid:
        return 5
    return 3   # funs


def erl_cmp(a, b):
    """Erlang term order: number < atom < reference < fun < port < pid < tuple < map < nil < list < bitstring"""
    ta, tb = type(a), type(b)
    if (ta is int or ta is float) and (tb is int or tb is float):
        return -1 if a < b else (1 if a > b else 0)
    if ta is tb:
        if ta is str:
            return -1 if a < b else (1 if a > b else 0)
        if ta is bytes:
            return -1 if a < b else (1 if a > b else 0)
        if ta is tuple:
            if len(a) != len(b):
                return -1 if len(a) < len(b) else 1
            for x, y in zip(a, b):
                c = erl_cmp(x, y)
                if c:
                    return c
            return 0
        if ta is Cons:
            while True:
                c = erl_cmp(a.h, b.h)
                if c:
                    return c
                a, b = a.t, b.t
                if type(a) is not Cons or type(b) is not Cons:
                    return erl_cmp(a, b)
        if a is NIL:
            return 0
        if ta is Bits:
            # compare as bit sequences
            sa, sb = a.val << (8 - a.nbits % 8), b.val << (8 - b.nbits % 8)
            ba = sa.to_bytes((a.nbits + 8) // 8, "big")
            bb = sb.to_bytes((b.nbits + 8) // 8, "big")
            if ba != bb:
                return -1 if ba < bb else 1
            return -1 if a.nbits < b.nbits else (1 if a.nbits > b.nbits else 0)
        if ta is dict:
            if len(a) != len(b):
                return -1 if len(a) < len(b) else 1
            ka = sorted(a.keys(), key=cmp_key)
            kb = sorted(b.keys(), key=cmp_key)
            for x, y in zip(ka, kb):
                c = erl_cmp(x, y)
                if c:
                    return c
            for x, y in zip(ka, kb):
                c = erl_cmp(a[x], b[y])
                if c:
                    return c
            return 0
        ia, ib = id(a), id(b)
        if hasattr(a, "n") and hasattr(b, "n"):
            ia, ib = a.n, b.n
        return -1 if ia < ib else (1 if ia > ib else 0)
    ra, rb = type_rank(a), type_rank(b)
    if ra != rb:
        return -1 if ra < rb else 1
    # bytes vs Bits
    if ra == 10:
        xa = a if ta is Bits else Bits(int.from_bytes(a, "big"), len(a) * 8 + 0)
        xb = b if tb is Bits else Bits(int.from_bytes(b, "big"), len(b) * 8 + 0)
        na, nb = xa.nbits, xb.nbits
        m = min(na, nb)
        pa, pb = xa.val >> (na - m), xb.val >> (nb - m)
        if pa != pb:
            return -1 if pa < pb else 1
        return -1 if na < nb else (1 if na > nb else 0)
    return 0


class cmp_key(object):
    __slots__ = ("v",)

    def __init__(self, v):
        self.v = v

    def __lt__(self, o):
        return erl_cmp(self.v, o.v) < 0


def exact_eq(a, b):
    """=:="""
    ta = type(a)
    if ta is not type(b):
        return False
    if ta is Cons:
        while True:
            if not exact_eq(a.h, b.h):
                return False
            a, b = a.t, b.t
            if type(a) is not Cons or type(b) is not Cons:
                return exact_eq(a, b)
    if ta is tuple:
        if len(a) != len(b):
            return False
        for x, y in zip(a, b):
            if not exact_eq(x, y):
                return False
        return True
    if ta is dict:
        return erl_cmp(a, b) == 0 and all(exact_eq(a[k], b[k]) for k in a)
    if ta is Nil:
        return True
    return a == b if ta in (int, float, str, bytes, Bits) else a is b


def fmt_term(x, depth=0):
    """~p / ~w style printing (enough for integers, atoms, strings, tuples, lists, binaries)"""
    t = type(x)
    if t is int:
        return str(x)
    if t is float:
        return float_to_str(x)
    if t is str:
        return x if x and x[0].islower() and all(c.isalnum() or c in "_@" for c in x) else "'%s'" % x
    if t is tuple:
        return "{" + ",".join(fmt_term(e, depth + 1) for e in x) + "}"
    if t is bytes:
        if x and all(32 <= c < 127 for c in x):
            return '<<"%s">>' % x.decode("latin1")
        return "<<" + ",".join(str(c) for c in x) + ">>"
    if x is NIL:
        return "[]"
    if t is Cons:
        items = []
        c = x
        n = 0
        while type(c) is Cons and n < 10000:
            items.append(c.h)
            c = c.t
            n += 1
        if c is NIL and items and all(type(i) is int and (32 <= i < 127 or i in (9, 10, 13)) for i in items):
            s = "".join(chr(i) for i in items)
            return '"' + s.replace("\\", "\\\\").replace('"', '\\"').replace("\n", "\\n").replace("\r", "\\r").replace("\t", "\\t") + '"'
        s = "[" + ",".join(fmt_term(i, depth + 1) for i in items)
        if c is not NIL:
            s += "|" + fmt_term(c, depth + 1)
        return s + "]"
    if t is dict:
        return "#{" + ",".join("%s => %s" % (fmt_term(k), fmt_term(v)) for k, v in x.items()) + "}"
    return repr(x)


def float_to_str(f):
    """shortest round-trip representation the way io_lib:format ~p prints floats"""
    r = repr(float(f))
    if "e" in r or "E" in r:
        m, e = r.lower().split("e")
        if "." not in m:
            m += ".0"
        e = int(e)
        return "%se%s%d" % (m, "+" if e >= 0 else "-", abs(e)) if False else "%se%d" % (m, e)
    if "." not in r and "inf" not in r and "nan" not in r:
        r += ".0"
    return r
