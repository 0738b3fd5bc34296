"""Erlang tokenizer + preprocessor (-define / -include / -ifdef / ?MACRO) for the reference runner.
TEST INFRASTRUCTURE (oracle/): reads the reference's .erl sources where they lie, never copies them."""
import os

KEYWORDS = {"after", "and", "andalso", "band", "begin", "bnot", "bor", "bsl", "bsr", "bxor", "case", "catch", "cond", "div",
            "end", "fun", "if", "let", "not", "of", "or", "orelse", "receive", "rem", "try", "when", "xor"}

PUNCT3 = ("=:=", "=/=", "...")
PUNCT2 = ("->", "<-", "<=", "=>", ":=", "<<", ">>", "||", "++", "--", "==", "/=", "=<", ">=", "::", "..")
PUNCT1 = "(){}[],;:#|!?=<>+-*/.@"


class Tok(object):
    __slots__ = ("t", "v", "line")

    def __init__(self, t, v, line):
        self.t = t
        self.v = v
        self.line = line

    def __repr__(self):
        return "%s:%r@%d" % (self.t, self.v, self.line)


class LexError(Exception):
    pass


def _escape(s, i):
    """s[i] is the char after a backslash; returns (codepoint, next index)"""
    c = s[i]
    simple = {"n": 10, "r": 13, "t": 9, "s": 32, "b": 8, "e": 27, "f": 12, "v": 11, "d": 127, "\\": 92, '"': 34, "'": 39}
    if c in simple:
        return simple[c], i + 1
    if c == "x":
        if s[i + 1] == "{":
            j = s.index("}", i)
            return int(s[i + 2:j], 16), j + 1
        return int(s[i + 1:i + 3], 16), i + 3
    if c == "^":
        return ord(s[i + 1]) & 31, i + 2
    if c in "01234567":
        j = i
        while j < len(s) and j < i + 3 and s[j] in "01234567":
            j += 1
        return int(s[i:j], 8), j
    return ord(c), i + 1


def tokenize(src):
    toks = []
    i, n, line = 0, len(src), 1
    while i < n:
        c = src[i]
        if c == "\n":
            line += 1
            i += 1
            continue
        if c in " \t\r\f":
            i += 1
            continue
        if c == "%":
            while i < n and src[i] != "\n":
                i += 1
            continue
        if c.isdigit():
            j = i
            while j < n and src[j].isdigit():
                j += 1
            if j < n and src[j] == "#":
                base = int(src[i:j])
                k = j + 1
                while k < n and (src[k].isalnum()):
                    k += 1
                toks.append(Tok("int", int(src[j + 1:k], base), line))
                i = k
                continue
            if j + 1 < n and src[j] == "." and src[j + 1].isdigit():
                k = j + 1
                while k < n and src[k].isdigit():
                    k += 1
                if k < n and src[k] in "eE":
                    k2 = k + 1
                    if k2 < n and src[k2] in "+-":
                        k2 += 1
                    if k2 < n and src[k2].isdigit():
                        while k2 < n and src[k2].isdigit():
                            k2 += 1
                        k = k2
                toks.append(Tok("float", float(src[i:k]), line))
                i = k
                continue
            toks.append(Tok("int", int(src[i:j]), line))
            i = j
            continue
        if c.isalpha() or c == "_":
            j = i
            while j < n and (src[j].isalnum() or src[j] in "_@"):
                j += 1
            w = src[i:j]
            if c.isupper() or c == "_":
                toks.append(Tok("var", w, line))
            elif w in KEYWORDS:
                toks.append(Tok(w, w, line))
            else:
                toks.append(Tok("atom", w, line))
            i = j
            continue
        if c == "'":
            j = i + 1
            out = []
            while src[j] != "'":
                if src[j] == "\\":
                    cp, j = _escape(src, j + 1)
                    out.append(chr(cp))
                else:
                    if src[j] == "\n":
                        line += 1
                    out.append(src[j])
                    j += 1
            toks.append(Tok("atom", "".join(out), line))
            i = j + 1
            continue
        if c == '"':
            j = i + 1
            out = []
            while src[j] != '"':
                if src[j] == "\\":
                    cp, j = _escape(src, j + 1)
                    out.append(cp)
                else:
                    if src[j] == "\n":
                        line += 1
                    out.append(ord(src[j]))
                    j += 1
            if toks and toks[-1].t == "str":
                toks[-1].v = toks[-1].v + out     # adjacent literals concatenate
            else:
                toks.append(Tok("str", out, line))
            i = j + 1
            continue
        if c == "$":
            if src[i + 1] == "\\":
                cp, j = _escape(src, i + 2)
                toks.append(Tok("int", cp, line))
                i = j
            else:
                toks.append(Tok("int", ord(src[i + 1]), line))
                i += 2
            continue
        if src[i:i + 3] in PUNCT3:
            toks.append(Tok(src[i:i + 3], None, line))
            i += 3
            continue
        if src[i:i + 2] in PUNCT2:
            toks.append(Tok(src[i:i + 2], None, line))
            i += 2
            continue
        if c == ".":
            if i + 1 >= n or src[i + 1] in " \t\r\n%":
                toks.append(Tok("dot", None, line))
                i += 1
                continue
        if c in PUNCT1:
            toks.append(Tok(c, None, line))
            i += 1
            continue
        raise LexError("bad character %r at line %d" % (c, line))
    return toks


class Preprocessor(object):
    """splits a source file into forms (token lists), handling attributes that matter and expanding macros"""

    def __init__(self, path, predefined=None):
        self.dir = os.path.dirname(path)
        self.path = path
        self.macros = dict(predefined or {})     # name -> (params or None, tokens)
        self.module = None
        self.records = {}
        self.forms = []                          # function forms: lists of tokens (macro-expanded), without the dot
        self.exports = set()
        self.cond = []                           # stack of booleans
        with open(path, "r", encoding="latin1") as f:
            self._run(tokenize(f.read()))

    def _split(self, toks):
        cur = []
        for t in toks:
            if t.t == "dot":
                yield cur
                cur = []
            else:
                cur.append(t)
        if cur:
            yield cur

    def _active(self):
        return all(self.cond)

    def _run(self, toks):
        for form in self._split(toks):
            if not form:
                continue
            if form[0].t == "-" and len(form) > 1 and form[1].t in ("atom", "if"):
                self._attribute(form)
            elif self._active():
                self.forms.append(self.expand(form))

    def _attribute(self, form):
        name = form[1].v
        if name == "ifdef" or name == "ifndef":
            m = form[3].v
            d = m in self.macros
            self.cond.append(d if name == "ifdef" else not d)
            return
        if name == "else":
            self.cond[-1] = not self.cond[-1]
            return
        if name == "endif":
            self.cond.pop()
            return
        if not self._active():
            return
        if name == "module":
            self.module = form[3].v
            self.macros["MODULE"] = (None, [Tok("atom", self.module, form[3].line)])
        elif name == "define":
            self._define(form)
        elif name == "include":
            p = os.path.join(self.dir, "".join(chr(c) for c in form[3].v))
            if os.path.exists(p):
                with open(p, "r", encoding="latin1") as f:
                    self._run(tokenize(f.read()))
        elif name == "record":
            self._record(form)
        elif name == "export":
            i = 0
            while i < len(form):
                if form[i].t == "atom" and i + 2 < len(form) and form[i + 1].t == "/" and form[i + 2].t == "int":
                    self.exports.add((form[i].v, form[i + 2].v))
                i += 1
        # everything else (-spec, -type, -author, -compile, -include_lib ...) carries nothing the runner needs

    def _define(self, form):
        # - define ( NAME [ (params) ] , body... )
        i = 3
        mname = form[i].v
        i += 1
        params = None
        if form[i].t == "(":
            params = []
            i += 1
            while form[i].t != ")":
                if form[i].t == "var":
                    params.append(form[i].v)
                i += 1
            i += 1
        assert form[i].t == ",", "bad -define of %s" % mname
        body = form[i + 1:-1]      # drop the closing paren of -define(
        self.macros[mname] = (params, body)

    def _record(self, form):
        # - record ( name , { f1 [= default] , ... } )
        rname = form[3].v
        fields = []
        i = 6
        depth = 0
        cur = []
        body = form[6:-2]
        for t in body:
            if t.t in ("(", "{", "[", "<<"):
                depth += 1
            elif t.t in (")", "}", "]", ">>"):
                depth -= 1
            if t.t == "," and depth == 0:
                fields.append(cur)
                cur = []
            else:
                cur.append(t)
        if cur:
            fields.append(cur)
        out = []
        for f in fields:
            fname = f[0].v
            default = None
            for k, t in enumerate(f):
                if t.t == "=":
                    default = self.expand(f[k + 1:])
                    break
                if t.t == "::":
                    break
            out.append((fname, default))
        self.records[rname] = out

    def expand(self, toks, depth=0):
        out = []
        i, n = 0, len(toks)
        while i < n:
            t = toks[i]
            if t.t == "?" and i + 1 < n and toks[i + 1].t in ("var", "atom"):
                mname = toks[i + 1].v
                if mname == "LINE":
                    out.append(Tok("int", t.line, t.line))
                    i += 2
                    continue
                if mname not in self.macros:
                    raise LexError("undefined macro ?%s in %s line %d" % (mname, self.path, t.line))
                params, body = self.macros[mname]
                i += 2
                if params is None:
                    out.extend(self.expand(body, depth + 1))
                    continue
                # macro call: collect arguments
                assert toks[i].t == "(", "macro ?%s needs arguments (line %d)" % (mname, t.line)
                i += 1
                args, cur, d = [], [], 0
                while True:
                    x = toks[i]
                    if x.t in ("(", "{", "[", "<<"):
                        d += 1
                    elif x.t in ("}", "]", ">>"):
                        d -= 1
                    elif x.t == ")":
                        if d == 0:
                            break
                        d -= 1
                    if x.t == "," and d == 0:
                        args.append(cur)
                        cur = []
                    else:
                        cur.append(x)
                    i += 1
                i += 1
                if cur or args:
                    args.append(cur)
                amap = dict(zip(params, args))
                sub = []
                for b in body:
                    if b.t == "var" and b.v in amap:
                        sub.extend(amap[b.v])
                    else:
                        sub.append(b)
                out.extend(self.expand(sub, depth + 1))
                continue
            out.append(t)
            i += 1
        return out
