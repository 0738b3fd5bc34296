"""Erlang expression / function parser for the reference runner (oracle/erlref, TEST INFRASTRUCTURE).

AST nodes are tuples: (kind, ...).  Clauses are (patterns, guards, body) with guards a list (`;`) of lists (`,`)."""
from .lexer import Tok


class ParseError(Exception):
    pass


BINOPS = {
    # op -> (precedence, right-assoc)
    "=": (100, True), "!": (100, True),
    "orelse": (150, True), "andalso": (160, True),
    "==": (200, False), "/=": (200, False), "=<": (200, False), "<": (200, False), ">=": (200, False), ">": (200, False),
    "=:=": (200, False), "=/=": (200, False),
    "++": (300, True), "--": (300, True),
    "+": (400, False), "-": (400, False), "bor": (400, False), "bxor": (400, False), "bsl": (400, False), "bsr": (400, False),
    "or": (400, False), "xor": (400, False),
    "*": (500, False), "/": (500, False), "div": (500, False), "rem": (500, False), "band": (500, False), "and": (500, False),
}
UNOPS = {"+", "-", "bnot", "not"}


class Parser(object):
    def __init__(self, toks, records=None, where=""):
        self.toks = toks
        self.i = 0
        self.records = records or {}
        self.where = where

    # ---- token helpers
    def peek(self, k=0):
        j = self.i + k
        return self.toks[j].t if j < len(self.toks) else None

    def next(self):
        t = self.toks[self.i]
        self.i += 1
        return t

    def expect(self, t):
        if self.peek() != t:
            got = self.toks[self.i] if self.i < len(self.toks) else None
            raise ParseError("%s: expected %r, got %r" % (self.where, t, got))
        return self.next()

    def accept(self, t):
        if self.peek() == t:
            self.i += 1
            return True
        return False

    # ---- forms
    def function(self):
        """name(Args) [when G] -> Body ; name(...) -> ... (the final dot is already stripped)"""
        clauses = []
        name = None
        while True:
            t = self.expect("atom")
            name = t.v
            clauses.append(self.clause_rest())
            if not self.accept(";"):
                break
        if self.i != len(self.toks):
            raise ParseError("%s: trailing tokens after function %s: %r" % (self.where, name, self.toks[self.i]))
        return name, len(clauses[0][0]), clauses

    def clause_rest(self):
        self.expect("(")
        pats = self.expr_list(")")
        guards = self.guard_opt()
        self.expect("->")
        body = self.body()
        return (pats, guards, body)

    def guard_opt(self):
        if not self.accept("when"):
            return []
        alts = []
        cur = [self.expr()]
        while True:
            if self.accept(","):
                cur.append(self.expr())
            elif self.accept(";"):
                alts.append(cur)
                cur = [self.expr()]
            else:
                break
        alts.append(cur)
        return alts

    def body(self):
        es = [self.expr()]
        while self.accept(","):
            es.append(self.expr())
        return es

    def expr_list(self, close):
        es = []
        if self.accept(close):
            return es
        es.append(self.expr())
        while self.accept(","):
            es.append(self.expr())
        self.expect(close)
        return es

    # ---- expressions
    def expr(self):
        if self.peek() == "catch":
            self.next()
            return ("catch", self.expr())
        return self.binop(0)

    def binop(self, minprec):
        left = self.unary()
        while True:
            op = self.peek()
            if op not in BINOPS:
                return left
            prec, right = BINOPS[op]
            if prec < minprec:
                return left
            self.next()
            rhs = self.binop(prec if right else prec + 1)
            if op == "=":
                left = ("match", left, rhs)
            elif op == "!":
                left = ("send", left, rhs)
            elif op == "andalso" or op == "orelse":
                left = (op, left, rhs)
            else:
                left = ("op", op, left, rhs)

    def unary(self):
        t = self.peek()
        if t in UNOPS:
            self.next()
            e = self.unary()
            if t == "-" and e[0] in ("int", "float"):
                return (e[0], -e[1])
            if t == "+" and e[0] in ("int", "float"):
                return e
            return ("unop", t, e)
        return self.postfix()

    def postfix(self):
        e = self.primary()
        while True:
            t = self.peek()
            if t == "(":
                self.next()
                args = self.expr_list(")")
                e = ("call", e, args)
            elif t == ":" and self.peek(1) != ":":
                self.next()
                f = self.primary()
                if self.peek() == "(":
                    self.next()
                    args = self.expr_list(")")
                    e = ("rcall", e, f, args)
                else:
                    e = ("remote", e, f)
            elif t == "#":
                e = self.record_or_map(e)
            else:
                return e

    def record_or_map(self, base):
        self.expect("#")
        if self.peek() == "{":
            self.next()
            return ("mapupd", base, self.map_fields()) if base is not None else ("map", self.map_fields())
        name = self.expect("atom").v
        if self.accept("."):
            f = self.expect("atom").v
            return ("recfield", base, name, f)
        self.expect("{")
        fields = []
        if not self.accept("}"):
            while True:
                ft = self.next()
                self.expect("=")
                fields.append((ft.v, self.expr()))
                if not self.accept(","):
                    break
            self.expect("}")
        return ("record", base, name, fields)

    def map_fields(self):
        fields = []
        if self.accept("}"):
            return fields
        while True:
            k = self.expr()
            if self.accept("=>"):
                op = "=>"
            else:
                self.expect(":=")
                op = ":="
            v = self.expr()
            fields.append((k, v, op))
            if not self.accept(","):
                break
        self.expect("}")
        return fields

    def primary(self):
        t = self.peek()
        if t == "int" or t == "float" or t == "atom" or t == "var":
            tok = self.next()
            return (t, tok.v)
        if t == "str":
            return ("str", self.next().v)
        if t == "(":
            self.next()
            e = self.expr()
            self.expect(")")
            return ("paren", e)
        if t == "{":
            self.next()
            return ("tuple", self.expr_list("}"))
        if t == "[":
            return self.list_expr()
        if t == "<<":
            return self.binary_expr()
        if t == "#":
            return self.record_or_map(None)
        if t == "begin":
            self.next()
            b = self.body()
            self.expect("end")
            return ("block", b)
        if t == "if":
            self.next()
            clauses = []
            while True:
                g = self.guard_seq()
                self.expect("->")
                clauses.append((g, self.body()))
                if not self.accept(";"):
                    break
            self.expect("end")
            return ("if", clauses)
        if t == "case":
            self.next()
            e = self.expr()
            self.expect("of")
            clauses = self.cr_clauses()
            self.expect("end")
            return ("case", e, clauses)
        if t == "receive":
            self.next()
            clauses = []
            if self.peek() != "after":
                clauses = self.cr_clauses()
            timeout = None
            tbody = None
            if self.accept("after"):
                timeout = self.expr()
                self.expect("->")
                tbody = self.body()
            self.expect("end")
            return ("receive", clauses, timeout, tbody)
        if t == "fun":
            return self.fun_expr()
        if t == "try":
            return self.try_expr()
        tok = self.toks[self.i] if self.i < len(self.toks) else None
        raise ParseError("%s: unexpected token %r" % (self.where, tok))

    def guard_seq(self):
        alts = []
        cur = [self.expr()]
        while True:
            if self.accept(","):
                cur.append(self.expr())
            elif self.accept(";"):
                alts.append(cur)
                cur = [self.expr()]
            else:
                break
        alts.append(cur)
        return alts

    def cr_clauses(self):
        clauses = []
        while True:
            p = self.expr()
            g = self.guard_opt()
            self.expect("->")
            clauses.append(([p], g, self.body()))
            if not self.accept(";"):
                break
        return clauses

    def list_expr(self):
        self.expect("[")
        if self.accept("]"):
            return ("nil",)
        first = self.expr()
        if self.accept("||"):
            quals = self.qualifiers()
            self.expect("]")
            return ("lc", first, quals)
        elems = [first]
        tail = ("nil",)
        while True:
            if self.accept(","):
                elems.append(self.expr())
            elif self.accept("|"):
                tail = self.expr()
                self.expect("]")
                break
            else:
                self.expect("]")
                break
        node = tail
        for e in reversed(elems):
            node = ("cons", e, node)
        return node

    def qualifiers(self):
        quals = []
        while True:
            save = self.i
            e = self.expr()
            if self.accept("<-"):
                quals.append(("gen", e, self.expr()))
            elif self.accept("<="):
                quals.append(("bgen", e, self.expr()))
            else:
                quals.append(("filter", e))
            if not self.accept(","):
                break
        return quals

    def binary_expr(self):
        self.expect("<<")
        if self.accept(">>"):
            return ("bin", [])
        # binary comprehension?
        segs = [self.bin_segment()]
        if self.accept("||"):
            quals = self.qualifiers()
            self.expect(">>")
            return ("bc", segs[0], quals)
        while self.accept(","):
            segs.append(self.bin_segment())
        self.expect(">>")
        return ("bin", segs)

    def bin_segment(self):
        # Value[:Size][/TypeSpecifierList]; value is a restricted expression (no binary operators unless parenthesised)
        t = self.peek()
        if t in UNOPS:
            self.next()
            v = self.primary_for_bin()
            if t == "-" and v[0] in ("int", "float"):
                v = (v[0], -v[1])
            elif t != "+":
                v = ("unop", t, v)
        else:
            v = self.primary_for_bin()
        size = None
        if self.peek() == ":" :
            self.next()
            size = self.primary_for_bin()
        specs = []
        if self.accept("/"):
            while True:
                a = self.expect("atom").v
                if self.accept(":"):
                    specs.append((a, self.expect("int").v))
                else:
                    specs.append((a, None))
                if not self.accept("-"):
                    break
        return (v, size, specs)

    def primary_for_bin(self):
        e = self.primary()
        # allow calls / remote calls as segment values when written directly (rare)
        while self.peek() == "(":
            self.next()
            e = ("call", e, self.expr_list(")"))
        return e

    def fun_expr(self):
        self.expect("fun")
        t = self.peek()
        if t == "(":
            clauses = []
            while True:
                clauses.append(self.clause_rest())
                if not self.accept(";"):
                    break
            self.expect("end")
            return ("fun", None, clauses)
        if t == "var" and self.peek(1) == "(":
            name = self.toks[self.i].v
            clauses = []
            while True:
                self.expect("var")
                clauses.append(self.clause_rest())
                if not self.accept(";"):
                    break
            self.expect("end")
            return ("fun", name, clauses)
        # fun name/arity | fun mod:name/arity
        a = self.next()
        if self.accept(":"):
            f = self.next()
            self.expect("/")
            ar = self.next()
            return ("funref", (a.t, a.v), (f.t, f.v), (ar.t, ar.v))
        self.expect("/")
        ar = self.expect("int").v
        return ("funref", None, ("atom", a.v), ("int", ar))

    def try_expr(self):
        self.expect("try")
        body = self.body()
        of_clauses = None
        if self.accept("of"):
            of_clauses = self.cr_clauses()
        catch_clauses = []
        after = None
        if self.accept("catch"):
            while True:
                # [Class:]Pattern[:Stack] [when G] -> Body
                first = self.expr_no_colon()
                cls = ("atom", "throw")
                pat = first
                if self.accept(":"):
                    cls = first
                    pat = self.expr_no_colon()
                    if self.accept(":"):
                        self.expect("var")
                g = self.guard_opt()
                self.expect("->")
                catch_clauses.append((cls, pat, g, self.body()))
                if not self.accept(";"):
                    break
        if self.accept("after"):
            after = self.body()
        self.expect("end")
        return ("try", body, of_clauses, catch_clauses, after)

    def expr_no_colon(self):
        """an expression in a catch-clause head, where `:` separates class from pattern"""
        # parse a binop expression but stop the postfix loop at ':'
        save = Parser.postfix
        try:
            Parser.postfix = Parser._postfix_no_colon
            return self.expr()
        finally:
            Parser.postfix = save

    def _postfix_no_colon(self):
        e = self.primary()
        while True:
            t = self.peek()
            if t == "(":
                self.next()
                e = ("call", e, self.expr_list(")"))
            elif t == "#":
                e = self.record_or_map(e)
            else:
                return e


def parse_function(toks, records=None, where=""):
    return Parser(toks, records, where).function()


def parse_expr_string(src):
    from .lexer import tokenize
    toks = [t for t in tokenize(src) if t.t != "dot"]
    p = Parser(toks, where="<expr>")
    return p.body()
