// ORACLE (test infrastructure only -- never linked into the product path).
// Restatement of erlamsa_sgml (src/erlamsa_sgml.erl): the tolerant SGML/XML tokenizer (:65-176), the
// AST builder with its unpaired / closed-earlier tag recovery (:187-279), the folder (:290-331), the
// numbered walks (:344-470) and the twelve-way mutation (:478-757). Included at the end of mutations.hpp.
//
// Lists: an Erlang accumulator that is built by prepending and reversed at the end is kept here as a
// std::vector in CHRONOLOGICAL order (push_back == prepend, vector order == the reversed list).
#pragma once
#include <memory>
#include <tuple>

namespace eo {
namespace sgml {

struct Param { std::string name, value, quote; };   // quote: "'" | "\"" | "" (:152-165)
inline bool operator==(const Param& a, const Param& b) { return a.name == b.name && a.value == b.value && a.quote == b.quote; }

struct Node;
using NodeP = std::shared_ptr<const Node>;
using List = std::vector<NodeP>;
struct Node {
    enum K { TAG, TEXT, SC, QUE, BANG, COMMENT, OPEN, CLOSE, TAGCLOSE } k;
    std::string a, b;             // TAG: open name a, close name b | TEXT/QUE/BANG/COMMENT: payload a | SC/OPEN/CLOSE/TAGCLOSE: name a
    std::vector<Param> params;    // TAG / SC / OPEN
    List kids;                    // TAG
};
inline NodeP mk(Node::K k, const std::string& a, const std::string& b = "", const std::vector<Param>& p = {}, const List& kids = {}) {
    auto n = std::make_shared<Node>(); n->k = k; n->a = a; n->b = b; n->params = p; n->kids = kids; return n;
}
inline bool equal(const List& x, const List& y);
inline bool equal(const NodeP& x, const NodeP& y) {
    if (x == y) return true;
    return x->k == y->k && x->a == y->a && x->b == y->b && x->params == y->params && equal(x->kids, y->kids);
}
inline bool equal(const List& x, const List& y) {
    if (x.size() != y.size()) return false;
    for (size_t i = 0; i < x.size(); i++) if (!equal(x[i], y[i])) return false;
    return true;
}

// ---------------------------------------------------------------- tokenizer :65-176
struct Token { enum K { OPEN, SC, CLOSE, TEXT, BANG, COMMENT, QUE, EOFTEXT } k; std::string a, lower; std::vector<Param> params; };
struct TokError { bool is_throw; };   // throw(incorrect_sgml) vs an Erlang error (function_clause): only the former is caught by sgml_mutate

inline bool is_ws(uint8_t c) { return c == ' ' || c == '\r' || c == '\n' || c == '\t'; }            // ?ws :58
inline bool is_ev(uint8_t c) { return is_ws(c) || c == '>' || c == '='; }                             // ?ev :64
inline std::string to_lower_latin1(const std::string& s) {   // string:to_lower/1 (OTP, ISO-8859-1 rules)
    std::string o = s;
    for (char& ch : o) { uint8_t c = (uint8_t)ch; if ((c >= 'A' && c <= 'Z') || (c >= 0xC0 && c <= 0xD6) || (c >= 0xD8 && c <= 0xDE)) ch = (char)(c + 32); }
    return o;
}

// tz/2 from state {tag,""} at S[i..] until a {text,Str,Token} result: returns the token and the index of Str
struct LastPos { size_t gt, cend, qend, sq, dq; };   // last occurrence of '>', "-->", "?>", "'", '"': a scan past it cannot terminate
inline LastPos last_positions(const Bin& S) { LastPos l; l.gt = S.rfind('>'); l.cend = S.rfind("-->"); l.qend = S.rfind("?>"); l.sq = S.rfind('\''); l.dq = S.rfind('"'); return l; }
inline Token scan_tag(const Bin& S, size_t i, size_t& next, const LastPos& last) {
    const size_t n = S.size();
    auto ws = [&](size_t p) { while (p < n && is_ws((uint8_t)S[p])) p++; return p; };
    auto starts = [&](size_t p, const char* lit) { size_t l = strlen(lit); return p + l <= n && S.compare(p, l, lit) == 0; };
    enum St { TAGN, BANG, COMMENT, QUE, ETAG, ENDTAG, ENDTAG_GT, ATTR, EATT, VAL, SQVAL, DQVAL, UQVAL } st = TAGN;
    std::string tag, dt, a, v; std::vector<Param> attrs;
    for (;;) {
        switch (st) {
        case TAGN:
            if (tag.empty()) {                                                      // :86-89
                if (starts(i, "!--")) { st = COMMENT; dt.clear(); i += 3; continue; }
                if (starts(i, "!")) { st = BANG; dt.clear(); i = ws(i + 1); continue; }
                if (starts(i, "?")) { st = QUE; dt.clear(); i = ws(i + 1); continue; }
                if (starts(i, "/")) { st = ENDTAG; tag.clear(); i = ws(i + 1); continue; }
            }
            if (starts(i, "/>")) { next = i + 2; Token t; t.k = Token::SC; t.a = tag; return t; }              // :90
            if (i < n && is_ev((uint8_t)S[i])) { st = ATTR; a.clear(); attrs.clear(); i = ws(i); continue; }   // :91
            if (i < n) { tag.push_back(S[i]); i++; continue; }                                                 // :92 (?ok is always true)
            throw TokError{true};                                                                              // :93
        case BANG:                                                                                             // :95-97
            { size_t q = (last.gt == Bin::npos || i > last.gt) ? Bin::npos : S.find('>', i);   // the first '>' ends it
              if (q == Bin::npos) throw TokError{true};
              next = q + 1; Token t; t.k = Token::BANG; t.a = S.substr(i, q - i); return t; }
        case COMMENT:                                                                                          // :99-100 (no clause for <<>>: function_clause)
            { size_t q = (last.cend == Bin::npos || i > last.cend) ? Bin::npos : S.find("-->", i);
              if (q == Bin::npos) throw TokError{false};
              next = q + 3; Token t; t.k = Token::COMMENT; t.a = S.substr(i, q - i); return t; }
        case QUE:                                                                                              // :102-104
            { size_t q = (last.qend == Bin::npos || i > last.qend) ? Bin::npos : S.find("?>", i);
              if (q == Bin::npos) throw TokError{true};
              next = q + 2; Token t; t.k = Token::QUE; t.a = S.substr(i, q - i); return t; }
        case ETAG:                                                                                             // :106-108
            if (starts(i, "/>")) { next = i + 2; Token t; t.k = Token::SC; t.a = tag; t.params = attrs; return t; }
            if (starts(i, ">")) { next = i + 1; Token t; t.k = Token::OPEN; t.a = tag; t.params = attrs; return t; }
            throw TokError{true};
        case ENDTAG:                                                                                           // :110,112,114
            if (i < n && is_ev((uint8_t)S[i])) { st = ENDTAG_GT; i = ws(i); continue; }
            if (i < n) { tag.push_back(S[i]); i++; continue; }
            throw TokError{true};
        case ENDTAG_GT:                                                                                        // :111,113
            if (starts(i, ">")) { next = i + 1; Token t; t.k = Token::CLOSE; t.a = tag; t.lower = to_lower_latin1(tag); return t; }
            throw TokError{true};
        case ATTR:                                                                                             // :116-121
            if (a.empty() && i < n && is_ev((uint8_t)S[i])) { st = ETAG; continue; }
            if (a.empty() && starts(i, "/>")) { st = ETAG; continue; }
            if (i < n && is_ev((uint8_t)S[i])) { st = EATT; i = ws(i); continue; }
            if (starts(i, "/>")) { st = EATT; i = ws(i); continue; }
            if (i < n) { a.push_back(S[i]); i++; continue; }
            throw TokError{true};
        case EATT:                                                                                             // :123-124
            if (starts(i, "=")) { st = VAL; i = ws(i + 1); continue; }
            attrs.push_back({a, "", ""}); a.clear(); st = ATTR; i = ws(i); continue;
        case VAL:                                                                                              // :126-128
            v.clear();
            if (starts(i, "'")) { st = SQVAL; i++; continue; }
            if (starts(i, "\"")) { st = DQVAL; i++; continue; }
            st = UQVAL; continue;
        case SQVAL:                                                                                            // :130-132
            { size_t q = (last.sq == Bin::npos || i > last.sq) ? Bin::npos : S.find('\'', i);
              if (q == Bin::npos) throw TokError{true};
              attrs.push_back({a, S.substr(i, q - i), "'"}); a.clear(); st = ATTR; i = ws(q + 1); continue; }
        case DQVAL:                                                                                            // :134-136
            { size_t q = (last.dq == Bin::npos || i > last.dq) ? Bin::npos : S.find('"', i);
              if (q == Bin::npos) throw TokError{true};
              attrs.push_back({a, S.substr(i, q - i), "\""}); a.clear(); st = ATTR; i = ws(q + 1); continue; }
        case UQVAL:                                                                                            // :139-142
            if ((i < n && is_ev((uint8_t)S[i])) || starts(i, "/>")) { attrs.push_back({a, v, ""}); a.clear(); st = ATTR; i = ws(i); continue; }
            if (i < n) { v.push_back(S[i]); i++; continue; }
            throw TokError{true};
        }
    }
}

// tokenize/1 :65-96. Bytes before the first '<' are dropped (tz(nil,_) :82); a tag that fails to scan after the
// first one is folded into the surrounding text together with its '<' (the catch-all try at :74-79 / bad_text :83-91;
// the white space after that '<' is lost because EStr is already ws-skipped).
inline std::vector<Token> tokenize(const Bin& S) {
    const size_t n = S.size();
    auto ws = [&](size_t p) { while (p < n && is_ws((uint8_t)S[p])) p++; return p; };
    size_t lt = S.find('<');
    if (lt == Bin::npos) throw TokError{true};                         // tz(nil, <<>>) :83
    std::vector<Token> out;
    size_t p = 0;
    const LastPos last = last_positions(S);
    Token cur = scan_tag(S, ws(lt + 1), p, last);                      // outside any try: failures propagate
    for (;;) {
        std::string prefix;
        for (;;) {
            size_t q = S.find('<', p);
            if (q == Bin::npos) { out.push_back(cur); Token e; e.k = Token::EOFTEXT; e.a = prefix + S.substr(p); out.push_back(e); return out; }
            std::string txt = S.substr(p, q - p);
            size_t e = ws(q + 1), nx = 0;
            try {
                Token t2 = scan_tag(S, e, nx, last);
                out.push_back(cur); Token tx; tx.k = Token::TEXT; tx.a = prefix + txt; out.push_back(tx);
                cur = t2; p = nx; break;
            } catch (const TokError&) {
                prefix += txt; prefix.push_back('<'); p = e;
            }
        }
    }
}

// ---------------------------------------------------------------- AST builder :187-279
struct BuildRes {
    enum K { OK, NO_PAIR, CLOSED_EARLIER } k;
    List list;                 // OK: forward list (nested: kids; the {tagclose,Tag} head is carried in close_tag) | errors: accumulator
    bool has_tagclose = false; std::string close_tag;
    size_t rest = 0;           // index of the remaining tokens
    std::vector<std::string> tags;
    std::string ce_name, ce_close;
    long n = 0, nt = 0;
};
inline BuildRes build_ast2(const std::vector<Token>& tk, size_t i, List acc, std::vector<std::string> tags, long n, long nt) {
    // tags: index 0 = head of the Erlang list
    for (;;) {
        const Token& t = tk.at(i);
        switch (t.k) {
        case Token::OPEN: {
            std::string lower = to_lower_latin1(t.a);
            std::vector<std::string> t2 = tags; t2.insert(t2.begin(), lower);
            BuildRes r = build_ast2(tk, i + 1, List(), t2, 0, 0);
            if (r.k == BuildRes::OK) {                                                             // :194-197
                if (!r.has_tagclose) throw CaseDied("sgml build_ast2: case_clause");
                acc.push_back(mk(Node::TAG, t.a, r.close_tag, t.params, r.list));
                i = r.rest; n += r.n + 1; nt += r.nt + 1; continue;
            }
            if (r.k == BuildRes::NO_PAIR) {                                                        // :204-209
                acc.push_back(mk(Node::OPEN, t.a, "", t.params)); acc.insert(acc.end(), r.list.begin(), r.list.end());
                i = r.rest; tags = r.tags; n += r.n + 1; nt += r.nt; continue;
            }
            if (r.ce_name == lower) {                                                              // :210-213
                acc.push_back(mk(Node::TAG, t.a, r.ce_close, t.params, r.list));
                i = r.rest; tags = r.tags; n += r.n + 1; nt += r.nt + 1; continue;
            }
            acc.push_back(mk(Node::OPEN, t.a, "", t.params)); acc.insert(acc.end(), r.list.begin(), r.list.end());   // :214-220
            r.list = acc; r.n = n + r.n + 1; r.nt = nt + r.nt; return r;
        }
        case Token::CLOSE:
            if (!tags.empty() && tags[0] == t.lower) {                                             // :222-224
                BuildRes r; r.k = BuildRes::OK; r.list = acc; r.has_tagclose = true; r.close_tag = t.a; r.rest = i + 1; r.n = n; r.nt = nt; return r;
            }
            if (!tags.empty()) {                                                                   // :225-234
                size_t j = 1; while (j < tags.size() && tags[j] != t.lower) j++;
                if (j == tags.size()) { acc.push_back(mk(Node::CLOSE, t.a)); n++; i++; continue; }
                BuildRes r; r.k = BuildRes::CLOSED_EARLIER; r.ce_name = t.lower; r.ce_close = t.a; r.rest = i + 1; r.list = acc;
                r.tags.assign(tags.begin() + j + 1, tags.end()); r.n = n; r.nt = nt; return r;    // push_till :181-183
            }
            acc.push_back(mk(Node::CLOSE, t.a)); n++; i++; continue;                               // :235-237
        case Token::TEXT:
            if (!t.a.empty()) { acc.push_back(mk(Node::TEXT, t.a)); n++; }                         // :238-243
            i++; continue;
        case Token::BANG: acc.push_back(mk(Node::BANG, t.a)); n++; i++; continue;                  // :244-246
        case Token::COMMENT: acc.push_back(mk(Node::COMMENT, t.a)); n++; i++; continue;            // :247-249
        case Token::QUE: acc.push_back(mk(Node::QUE, t.a)); n++; i++; continue;                    // :250-252
        case Token::SC: acc.push_back(mk(Node::SC, t.a, "", t.params)); n++; i++; continue;        // :253-255
        case Token::EOFTEXT: {
            BuildRes r;
            if (tags.empty()) {                                                                    // :256-261
                if (!t.a.empty()) { acc.push_back(mk(Node::TEXT, t.a)); n++; }
                r.k = BuildRes::OK; r.list = acc; r.rest = tk.size(); r.n = n; r.nt = nt; return r;
            }
            r.k = BuildRes::NO_PAIR; r.rest = i; r.list = acc; r.tags.assign(tags.begin() + 1, tags.end()); r.n = n; r.nt = nt; return r;   // :262-264
        }
        }
    }
}

// The same builder without recursion (documents with tens of thousands of unclosed tags made the recursive form
// quadratic). What build_ast2/4 computes is a stack discipline over the token stream, in document order:
//   * a close tag that matches the innermost open tag closes it (:222-224);
//   * one that matches a tag further out closes THAT tag and turns every open tag in between into a plain
//     {open,_,_} element followed by what would have been its children (:225-234 with :210-220);
//   * one that matches nothing is a {close,_} element (:230-231, :235-237);
//   * at the end of input every tag still open becomes a plain {open,_,_} element (:204-209, :262-264).
// N counts elements, NT paired tags. build_ast2() above is kept as the executable statement of the reference's
// clauses; eo_sgml_selfcheck() (driver.cpp) compares the two.
struct BuildOut { List list; long n = 0, nt = 0; };
inline BuildOut build_ast_iter(const std::vector<Token>& tk) {
    struct Frame { const Token* open; std::string lower; List kids; long n = 0, nt = 0; };
    std::vector<Frame> st; st.emplace_back(); st.back().open = nullptr;
    auto flatten_top = [&]() {   // the innermost open tag stays unpaired
        Frame f = std::move(st.back()); st.pop_back(); Frame& p = st.back();
        p.kids.push_back(mk(Node::OPEN, f.open->a, "", f.open->params));
        p.kids.insert(p.kids.end(), f.kids.begin(), f.kids.end());
        p.n += f.n + 1; p.nt += f.nt;
    };
    for (const Token& t : tk) {
        switch (t.k) {
        case Token::OPEN: { Frame f; f.open = &t; f.lower = to_lower_latin1(t.a); st.push_back(std::move(f)); break; }
        case Token::CLOSE: {
            size_t j = st.size();
            for (size_t q = st.size(); q-- > 1;) if (st[q].lower == t.lower) { j = q; break; }
            if (j == st.size()) { st.back().kids.push_back(mk(Node::CLOSE, t.a)); st.back().n++; break; }
            while (st.size() - 1 > j) flatten_top();
            Frame f = std::move(st.back()); st.pop_back(); Frame& p = st.back();
            p.kids.push_back(mk(Node::TAG, f.open->a, t.a, f.open->params, f.kids));
            p.n += f.n + 1; p.nt += f.nt + 1;
            break;
        }
        case Token::TEXT: if (!t.a.empty()) { st.back().kids.push_back(mk(Node::TEXT, t.a)); st.back().n++; } break;
        case Token::BANG: st.back().kids.push_back(mk(Node::BANG, t.a)); st.back().n++; break;
        case Token::COMMENT: st.back().kids.push_back(mk(Node::COMMENT, t.a)); st.back().n++; break;
        case Token::QUE: st.back().kids.push_back(mk(Node::QUE, t.a)); st.back().n++; break;
        case Token::SC: st.back().kids.push_back(mk(Node::SC, t.a, "", t.params)); st.back().n++; break;
        case Token::EOFTEXT:
            while (st.size() > 1) flatten_top();
            if (!t.a.empty()) { st.back().kids.push_back(mk(Node::TEXT, t.a)); st.back().n++; }
            break;
        }
    }
    while (st.size() > 1) flatten_top();
    BuildOut o; o.list = std::move(st.back().kids); o.n = st.back().n; o.nt = st.back().nt; return o;
}

// ---------------------------------------------------------------- folder :290-331
inline void fold_params(const std::vector<Param>& ps, Bin& o) {
    for (const Param& p : ps) {
        if (p.value.empty()) { o += " "; o += p.name; }                       // :297-298 (after the final reverse: " Name")
        else { o += " "; o += p.name; o += "="; o += p.quote; o += p.value; o += p.quote; }
    }
}
inline void fold_ast(const List& l, Bin& o) {
    for (const NodeP& e : l) {
        switch (e->k) {
        case Node::TAG: o += "<"; o += e->a; fold_params(e->params, o); o += ">"; fold_ast(e->kids, o); o += "</"; o += e->b; o += ">"; break;
        case Node::TEXT: o += e->a; break;
        case Node::SC: o += "<"; o += e->a; fold_params(e->params, o); o += " />"; break;
        case Node::QUE: o += "<?"; o += e->a; o += "?>"; break;
        case Node::BANG: o += "<!"; o += e->a; o += ">"; break;
        case Node::COMMENT: o += "<!--"; o += e->a; o += "-->"; break;
        case Node::OPEN: o += "<"; o += e->a; fold_params(e->params, o); o += ">"; break;
        case Node::CLOSE: o += "</"; o += e->a; o += ">"; break;
        case Node::TAGCLOSE: break;
        }
    }
}

// ---------------------------------------------------------------- numbered walks :344-470
// walk/3: pre-order numbering (elements from 1, paired tags from 1); Fun sees a tag AFTER its children, rebuilt.
// fun(elem, acc, tag_no (0 when not a tag... the running tag count), elem_no) appends to acc (push_back == prepend).
template <class F>
inline void walk_rec(const List& ast, F& fun, List& acc, long& tc, long& c) {
    for (const NodeP& e : ast) {
        if (e->k == Node::TAG) {
            long mytc = ++tc, myc = ++c;
            List child; walk_rec(e->kids, fun, child, tc, c);
            fun(mk(Node::TAG, e->a, e->b, e->params, child), acc, mytc, myc);
        } else { ++c; fun(e, acc, tc, c); }
    }
}
template <class F>
inline List walk(const List& ast, F fun) { List acc; long tc = 0, c = 0; walk_rec(ast, fun, acc, tc, c); return acc; }
inline long count_elems(const List& ast) { long c = 0; for (const NodeP& e : ast) { c++; if (e->k == Node::TAG) c += count_elems(e->kids); } return c; }   // ResN of count/1 :420-431

struct Sel { NodeP elem; long t = 0, i = 0; };
inline bool select_rec(const List& ast, long want_tag, long want_elem, long& tc, long& c, Sel& out) {   // select/2 :382-404
    for (const NodeP& e : ast) {
        if (e->k == Node::TAG) {
            long mytc = ++tc, myc = ++c;
            if ((want_tag && mytc == want_tag) || (want_elem && myc == want_elem)) { out.elem = e; out.t = mytc; out.i = myc; return true; }
            if (select_rec(e->kids, want_tag, want_elem, tc, c, out)) return true;
        } else {
            ++c;
            if (want_elem && c == want_elem) { out.elem = e; out.t = tc; out.i = c; return true; }
        }
    }
    return false;
}
inline Sel select_tag(const List& ast, long n) { Sel s; long tc = 0, c = 0; if (!select_rec(ast, n, 0, tc, c, s)) throw CaseDied("sgml select_tag: badmatch"); return s; }
inline Sel select_elem(const List& ast, long n) { Sel s; long tc = 0, c = 0; if (!select_rec(ast, 0, n, tc, c, s)) throw CaseDied("sgml select_elem: badmatch"); return s; }

inline List replace_elem(const List& ast, long r, const NodeP& el) {   // :453-462
    return walk(ast, [&](const NodeP& e, List& acc, long, long i) { acc.push_back(i == r ? el : e); });
}
inline List repeat_elem(const List& ast, long r, long times) {         // :464-476
    return walk(ast, [&](const NodeP& e, List& acc, long, long i) { acc.push_back(e); if (i == r) for (long k = 0; k < times; k++) acc.push_back(e); });
}
inline List insert_elem(const List& ast, long r, const NodeP& ne) {    // :478-485
    return walk(ast, [&](const NodeP& e, List& acc, long, long i) { acc.push_back(e); if (i == r) acc.push_back(ne); });
}
inline NodeP pump_path(NodeP start, long end, long n) {                // :496-508
    for (; n > 0; n--) {
        List one{start};
        List pumped = walk(one, [&](const NodeP& e, List& acc, long, long i) { acc.push_back(i == end ? start : e); });
        start = pumped.at(0); end = end * 2 - 1;
    }
    return start;
}

}  // namespace sgml

// ---------------------------------------------------------------- mutation :478-757
inline std::string Mutations::ssrf_uri() const { return "://" + opts.ssrf_host + ":" + std::to_string(opts.ssrf_port) + "/"; }   // get_ssrf_uri/0, src/erlamsa_mutations.erl:728-731

// inner_mutations/1 + mutators_mutator/1 (src/erlamsa_mutations.erl:1341-1356,1387-1395): a fresh table (2 draws),
// the listed codes at their default priorities, scores drawn walking the table backwards
inline void Mutations::inner_muta(const std::vector<int>& ids, std::unique_ptr<Mutations>& m, Opts& o2, std::vector<MutNode>& nodes) {
    o2 = opts;
    for (int i = 0; i < M_COUNT; i++) o2.muta_pri[i] = -1;
    for (int id : ids) o2.muta_pri[id] = MUT_DEFAULT_PRI[id];
    m.reset(new Mutations(rng, o2)); m->depth = depth + 1;
    m->build_table_draws();
    nodes = m->make_mutator_nodes();
}
// Muta([Binary], []) -> hd(NewLl): one scheduler round on a copy of the closure (the caller never keeps NewMuta)
inline Bin Mutations::inner_round(Mutations& m, const std::vector<MutNode>& nodes, const Bin& b) {
    std::vector<MutNode> fs = nodes; Blocks in1{b};
    Blocks nl = m.mux_fuzzers(fs, in1, nullptr);
    return nl.at(0);
}

inline MutRes Mutations::sgml_mutate(const Blocks& ll) {
    using namespace sgml;
    MutRes r; r.ll = ll; r.delta = -1;
    const Bin& h = ll[0];
    if (binarish(h)) return r;                                                       // parse/2 :185-186
    std::vector<Token> tokens;
    try { tokens = tokenize(h); }
    catch (const TokError& e) { if (e.is_throw) return r; throw CaseDied("sgml tokenizer: function_clause"); }
    BuildOut br = build_ast_iter(tokens);
    const List& ast = br.list; const long N = br.n, NT = br.nt;
    List res = ast; double d = 1;
    uint64_t which = rng.rand(12);                                                   // :698
    switch (which) {
    case 0: {                                                                        // sgml_swap :540-553
        long r1 = (long)rng.erand(N), r2 = (long)rng.erand(N);
        NodeP e1 = select_elem(ast, r1).elem, e2 = select_elem(ast, r2).elem;
        res = walk(ast, [&](const NodeP& e, List& acc, long, long i) { acc.push_back(i == r1 ? e2 : i == r2 ? e1 : e); });
        break;
    }
    case 1: { long rr = (long)rng.erand(N); res = repeat_elem(ast, rr, 1); break; }  // sgml_dup :532-534
    case 2: {                                                                        // sgml_pump :511-530
        d = -2;
        if (NT == 0) break;
        long rr = (long)rng.erand(NT);
        Sel s = select_tag(ast, rr);
        long sub = count_elems(List{s.elem});
        long e = (long)rng.erand((uint64_t)(sub - 1)) + 1;
        long cnt = (long)rng.erand((uint64_t)std::trunc(1000.0 / (100.0 + (double)sub)));
        res = replace_elem(ast, s.i, pump_path(s.elem, e, cnt));
        break;
    }
    case 3: { long rr = (long)rng.erand(N); long times = (long)rng.erand(100); res = repeat_elem(ast, rr, times); break; }   // sgml_repeat :536-538
    case 4: {                                                                        // sgml_insert2 :574-578
        long r1 = (long)rng.erand(N), r2 = (long)rng.erand(N);
        res = insert_elem(ast, r2, select_elem(ast, r1).elem);
        break;
    }
    case 5: {                                                                        // sgml_permparams :580-588
        long rr = (long)rng.erand(NT);
        res = walk(ast, [&](const NodeP& e, List& acc, long t, long) {
            if (e->k == Node::TAG && t == rr) {
                auto less = [](const Param& x, const Param& y) { return std::tie(x.name, x.value, x.quote) < std::tie(y.name, y.value, y.quote); };
                acc.push_back(mk(Node::TAG, e->a, e->b, rng.random_permutation(e->params, less), e->kids));
            } else acc.push_back(e);
        });
        break;
    }
    case 6: {                                                                        // sgml_breaktag :590-601
        long rr = (long)rng.erand(NT);
        res = walk(ast, [&](const NodeP& e, List& acc, long t, long) {
            if (e->k == Node::TAG && t == rr) {
                NodeP brk = rng.rand(1) == 0 ? mk(Node::OPEN, e->a, "", e->params) : mk(Node::CLOSE, e->b);
                acc.push_back(brk);                                                  // Internals ++ [Broken | Tree]: the children land after it, reversed
                for (size_t k = e->kids.size(); k-- > 0;) acc.push_back(e->kids[k]);
            } else acc.push_back(e);
        });
        break;
    }
    case 7: {                                                                        // sgml_insert :557-571
        long r1 = (long)rng.erand(N), r2 = (long)rng.erand(N);
        NodeP ne = select_elem(ast, r1).elem;
        if (ne->k == Node::TAG) res = walk(ast, [&](const NodeP& e, List& acc, long, long i) { acc.push_back(i == r2 ? mk(Node::TAG, ne->a, ne->b, ne->params, List{e}) : e); });
        else res = insert_elem(ast, r2, ne);
        break;
    }
    case 8: {                                                                        // sgml_xmlfeatures(_, NT, 1) :664-677
        d = -1;
        if (NT <= 0) break;
        res = walk(ast, [&](const NodeP& e, List& acc, long t, long) {
            if (e->k != Node::TAG) { acc.push_back(e); return; }
            if (rng.erand((uint64_t)std::trunc((double)t * 1.5)) != 1) { acc.push_back(e); return; }      // xmlns_modify :629-636
            std::vector<Param> np; bool changed = false;                                                  // xmlns_modify_params :603-627
            for (const Param& p : e->params) {
                if (p.name.compare(0, 5, "xmlns") == 0) {
                    std::string nu = rng.erand(2) == 1 ? p.value + " http" + ssrf_uri() : "http" + ssrf_uri();
                    if (nu != p.value) changed = true;
                    np.push_back({p.name, nu, p.quote});
                } else np.push_back(p);
            }
            if (!changed) {
                std::string uri = "http" + ssrf_uri();
                std::vector<Param> pre{{"xmlns", uri, "\""}, {"xmlns:xsi", uri, "\""}, {"xsi:schemaLocation", uri, "\""}};
                pre.insert(pre.end(), e->params.begin(), e->params.end()); np = pre;
            }
            acc.push_back(mk(Node::TAG, e->a, e->b, np, e->kids));
        });
        d = equal(ast, res) ? -1 : 1;
        break;
    }
    default: {                                                                       // inner text :721-733, 679-694
        std::unique_ptr<Mutations> im; Opts o2; std::vector<MutNode> nodes;
        inner_muta({M_AB, M_AD, M_BD, M_B64, M_LD, M_LP, M_LRI, M_LR, M_NUM, M_SD, M_URI}, im, o2, nodes);
        auto innertext = [&](const Bin& b, long nt) -> Bin {                         // mutate_innertext/3 :683-690
            long nw = 0; for (uint8_t c : b) if (c != 0 && c != 10 && c != 13 && c != 32) nw++;
            if (nw > 0 && nt > 0) {
                double rnd = rng.rand_float();
                if (rnd > 3.0 / (double)nt) return b;
                return inner_round(*im, nodes, b);
            }
            return b;
        };
        // walk2acc :361-380: children first, then the tag's own attribute values; text nodes as they come
        std::function<List(const List&)> w2 = [&](const List& l) -> List {
            List acc;
            for (const NodeP& e : l) {
                if (e->k == Node::TAG) {
                    List kids = w2(e->kids);
                    std::vector<Param> np; long npn = (long)e->params.size();
                    for (const Param& p : e->params) np.push_back({p.name, innertext(p.value, NT + npn), p.quote});
                    acc.push_back(mk(Node::TAG, e->a, e->b, np, kids));
                } else if (e->k == Node::TEXT) acc.push_back(mk(Node::TEXT, innertext(e->a, NT)));
                else acc.push_back(e);
            }
            return acc;
        };
        res = w2(ast);
        break;
    }
    }
    Bin nb; fold_ast(res, nb);
    if (nb == h) { r.delta = -1; return r; }                                         // :745-747
    r.ll[0] = nb; r.delta = d + std::trunc((double)nb.size() / (double)(AVG_BLOCK_SIZE * 10));
    return r;
}

}  // namespace eo
