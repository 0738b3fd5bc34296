"""CPU checks of two re-formulations the CUDA engine uses, restated in Python lane for lane and compared with the plain
restatement of the reference's algorithm on many random inputs (no GPU needed; the GPU parity tests then check the CUDA
transcription itself):

 * the string lexer as an automaton that hops between events found in position masks (erlamsa_b200/csrc/eb_mut_text.cuh)
   against the byte-by-byte walk of erlamsa_strlex:lex/1 (reference src/erlamsa_strlex.erl:46-143);
 * one level of the fuse search taken in steps -- runs of one-suffix nodes, runs of small nodes packed one suffix per lane
   with (node, byte) group keys, single nodes by ascending class -- (erlamsa_b200/csrc/eb_mut_fuse.cuh) against the list
   semantics of char_suffixes/1 and split/2 (reference src/erlamsa_fuse.erl:62-100), node order and suffix order included.
"""
import random


# ------------------------------------------------------------------ string lexer
def texty(b):
    if b < 9 or b > 126:
        return False
    if b > 31:
        return True
    return b in (9, 10, 13)


BYTE, TEXT, DELIM = 0, 1, 2


def lex_serial(d):
    """lex/1 :75-143 byte by byte: [(type, start)]"""
    n = len(d); out = []; p = 0; have_raw = False; raw_start = 0
    while p < n:
        te = True
        for k in range(6):                                   # texty_enough/2 :54-64
            if p + k >= n:
                break
            if not texty(d[p + k]):
                te = False; break
        if not te:
            if not have_raw:
                have_raw = True; raw_start = p
            p += 1; continue
        if have_raw:
            out.append((BYTE, raw_start)); have_raw = False
        seen = p
        while True:                                          # step_text :95-107
            if p >= n:
                out.append((TEXT, seen)); break
            h = d[p]
            if h in (34, 39):                                # step_delimited :114-143
                q = p; p += 1; closed = False
                while True:
                    if p >= n:
                        break
                    c = d[p]
                    if c == h:
                        closed = True; break
                    if c == 92 and p + 1 >= n:
                        p += 1; continue
                    if c == 92:
                        p += 2 if texty(d[p + 1]) else 1; continue
                    if texty(c):
                        p += 1; continue
                    break
                if closed:
                    if q > seen:
                        out.append((TEXT, seen))
                    out.append((DELIM, q)); p += 1
                else:
                    out.append((TEXT, seen))
                break
            if texty(h):
                p += 1; continue
            out.append((TEXT, seen)); break
    if have_raw:
        out.append((BYTE, raw_start))
    return out


def lex_events(d):
    """the device's formulation: masks E / MT / MD1 / MD2 and find-next hops"""
    n = len(d)
    T = [texty(c) for c in d]
    E = [all(T[p + k] for k in range(6) if p + k < n) for p in range(n)]
    MT = [(not T[p]) or d[p] in (34, 39) for p in range(n)]
    MD = {34: [d[p] == 34 or (not T[p]) or d[p] == 92 for p in range(n)], 39: [d[p] == 39 or (not T[p]) or d[p] == 92 for p in range(n)]}

    def nxt(M, p):
        while p < n and not M[p]:
            p += 1
        return p
    out = []; p = 0
    while p < n:
        p2 = nxt(E, p)
        if p2 > p:
            out.append((BYTE, p))
        if p2 >= n:
            break
        seen = p2; x = nxt(MT, p2)
        if x >= n:
            out.append((TEXT, seen)); break
        h = d[x]
        if h not in (34, 39):
            out.append((TEXT, seen)); p = x; continue
        q = x; p = x + 1
        while True:
            y = nxt(MD[h], p)
            if y >= n:
                out.append((TEXT, seen)); p = n; break
            c = d[y]
            if c == h:
                if q > seen:
                    out.append((TEXT, seen))
                out.append((DELIM, q)); p = y + 1; break
            if c == 92:
                if y + 1 >= n:
                    p = y + 1; continue
                p = y + 2 if T[y + 1] else y + 1; continue
            out.append((TEXT, seen)); p = y; break
    return out


def test_lexer_event_formulation_matches_the_byte_walk():
    r = random.Random(3)
    alph = [b'"', b"'", b"\\", b"a", b"b", b" ", b"\x00", b"\xff", b"\n", b"\x01"]
    for it in range(40000):
        n = r.randrange(0, 40)
        d = bytes(r.randrange(256) for _ in range(n)) if it % 3 == 0 else b"".join(r.choice(alph) for _ in range(n))
        assert lex_serial(d) == lex_events(d), d


# ------------------------------------------------------------------ fuse search, one level
def char_suffixes(S, data, n):
    subs = {}
    for p in S:
        if p >= n:
            continue                      # ([], Subs) -> Subs
        h = data[p]
        el = [p + 1] + subs.get(h, [])
        if el == [n]:
            el = []                       # fix_empty_list([[]]) -> []
        subs[h] = el
    return subs


def level_ref(nodes, a, b):
    """lists:foldl(fun split/2, [], Nodes) :85-100 (a node [[[]], []] is one empty suffix on each side)"""
    acc = []
    for F, T in nodes:
        sas = char_suffixes(F, a, len(a)); sbs = char_suffixes(T, b, len(b))
        for ch in sorted(sas):
            if sas[ch] == []:
                acc = [([len(a)], [len(b)])] + acc
            elif ch in sbs:
                acc = [(sas[ch], sbs[ch])] + acc
    return acc


def node_by_classes(Fl, Tl, a, b):
    """fuse_mid: first bytes in registers, classes of A ascending by min-reduction, newest suffix first"""
    na, nb = len(a), len(b)
    ca = [a[p] if p < na else 0x1ff for p in Fl]
    cb = [b[p] if p < nb else 0x1ff for p in Tl]

    def special_drop(lst, cls, n):
        for i, p in enumerate(lst):
            if p < n and p + 1 == n:
                return i if not any(cls[j] == cls[i] for j in range(i)) else -1
        return -1
    da = special_drop(Fl, ca, na); db = special_drop(Tl, cb, nb)
    out = []; last = -1
    while True:
        m = min([v for v in ca if v > last and v != 0x1ff], default=0x1ff)
        if m == 0x1ff:
            break
        last = c = m
        selA = [i for i in range(len(Fl)) if ca[i] == c and i != da]
        selB = [i for i in range(len(Tl)) if cb[i] == c and i != db]
        if not selA:
            out.append(([na], [nb])); continue
        if not any(v == c for v in cb):
            continue
        out.append(([Fl[i] + 1 for i in reversed(selA)], [Tl[i] + 1 for i in reversed(selB)]))
    return out


def one_suffix_node(Fl, Tl, a, b):
    """fuse_tiny"""
    na, nb = len(a), len(b)
    pa = Fl[0] if Fl else None; pb = Tl[0] if Tl else None
    if pa is None or pa >= na:
        return []
    if pa + 1 == na:
        return [([na], [nb])]
    if pb is not None and pb < nb and a[pa] == b[pb]:
        return [([pa + 1], [] if pb + 1 == nb else [pb + 1])]
    return []


def packed_run(nodes, a, b):
    """fuse_packed: lane j = j-th suffix of the run per side, key = node << 9 | byte, ranks by all-pairs counts"""
    na, nb = len(a), len(b)
    A = [(i, p) for i, (Fl, _) in enumerate(nodes) for p in Fl]
    B = [(i, p) for i, (_, Tl) in enumerate(nodes) for p in Tl]
    assert len(A) <= 32 and len(B) <= 32
    ka = [((i << 9) | a[p]) if p < na else 0x80000000 | j for j, (i, p) in enumerate(A)]
    kb = [((i << 9) | b[p]) if p < nb else 0xc0000000 | j for j, (i, p) in enumerate(B)]
    la = [p < na for _, p in A]; lb = [p < nb for _, p in B]
    fa = [la[j] and ka[j] not in ka[:j] for j in range(len(A))]
    fb = [lb[j] and kb[j] not in kb[:j] for j in range(len(B))]
    pla = [la[j] and not (fa[j] and A[j][1] + 1 == na) for j in range(len(A))]
    plb = [lb[j] and not (fb[j] and B[j][1] + 1 == nb) for j in range(len(B))]
    Fout = {}; Tout = {}; kids = []
    for j in range(len(A)):
        grp = [t for t in range(len(A)) if ka[t] == ka[j] and pla[t]]
        less = sum(1 for t in range(len(A)) if pla[t] and ka[t] < ka[j])
        if pla[j]:
            Fout[less + len(grp) - 1 - sum(1 for t in grp if t < j)] = A[j][1] + 1
        if fa[j]:
            if not grp:
                kids.append((ka[j], None))
            elif any(lb[t] and kb[t] == ka[j] for t in range(len(B))):
                kids.append((ka[j], (less, len(grp), sum(1 for t in range(len(B)) if plb[t] and kb[t] < ka[j]), sum(1 for t in range(len(B)) if plb[t] and kb[t] == ka[j]))))
    for j in range(len(B)):
        grp = [t for t in range(len(B)) if kb[t] == kb[j] and plb[t]]
        less = sum(1 for t in range(len(B)) if plb[t] and kb[t] < kb[j])
        if plb[j]:
            Tout[less + len(grp) - 1 - sum(1 for t in grp if t < j)] = B[j][1] + 1
    out = []
    for _, v in sorted(kids, key=lambda x: x[0]):
        out.append(([na], [nb]) if v is None else ([Fout[v[0] + i] for i in range(v[1])], [Tout[v[2] + i] for i in range(v[3])]))
    return out


def level_device(nodes_em, a, b):
    """fuse_step over one level: nodes in emission order, taken from the top (= the reference's list order)"""
    out = []; todo = list(reversed(nodes_em)); i = 0
    while i < len(todo):
        tiny = 0
        while i + tiny < len(todo) and tiny < 32 and len(todo[i + tiny][0]) <= 1 and len(todo[i + tiny][1]) <= 1:
            tiny += 1
        if tiny >= 16 or (tiny > 0 and tiny == len(todo) - i):
            for F, T in todo[i:i + tiny]:
                out += one_suffix_node(F, T, a, b)
            i += tiny; continue
        G = 0; sa = sb = 0
        while i + G < len(todo) and G < 32 and sa + len(todo[i + G][0]) <= 32 and sb + len(todo[i + G][1]) <= 32:
            sa += len(todo[i + G][0]); sb += len(todo[i + G][1]); G += 1
        if G == 0:
            out += node_by_classes(todo[i][0], todo[i][1], a, b); i += 1
        else:
            out += packed_run(todo[i:i + G], a, b); i += G
    return out


def test_fuse_level_paths_match_the_list_semantics():
    r = random.Random(2)
    levels = 0
    for it in range(1500):
        alpha = r.choice([2, 3, 4, 16, 256])
        na = r.randrange(1, 90); nb = r.randrange(1, 90)
        a = bytes(r.randrange(alpha) for _ in range(na))
        mode = r.randrange(4)
        if mode == 0:
            b = a
        elif mode == 1:
            b = bytes(r.randrange(alpha) for _ in range(nb))
        elif mode == 2:
            b = a[r.randrange(na):] + a[:r.randrange(na)]
        else:
            u = bytes(r.randrange(alpha) for _ in range(r.randrange(1, 6)))
            a = u * r.randrange(1, 12); b = u * r.randrange(1, 9) + bytes([r.randrange(alpha)])
        ref = [(list(range(len(a))), list(range(len(b))))]
        dev = [(list(range(len(a))), list(range(len(b))))]
        for _ in range(40):
            ref = level_ref(ref, a, b); dev = level_device(dev, a, b)
            assert ref == list(reversed(dev)), (a, b)
            levels += 1
            if not ref:
                break
    assert levels > 10000
