// ORACLE (test infrastructure only -- never linked into the product path).
// CPU restatement of erlamsa's per-case pipeline: erlamsa_main:fuzzer/1
// (src/erlamsa_main.erl:124-252), erlamsa_gen (src/erlamsa_gen.erl:43-56,152-199),
// erlamsa_patterns (src/erlamsa_patterns.erl:45-443) and the `return` output
// (src/erlamsa_out.erl:46-77,642-676), over the mutators in mutations.hpp.
//
// PARITY STATUS: the reference cannot run in the build container (no Erlang/OTP)
// and its own tests hold no byte-level golden vectors, so this
// oracle is pinned by: the public AS183 known-answer values, the one fixed-seed
// reference test (st_line_ins_test, src/erlamsa_mutations_test.erl:223-230) and
// the 31 property tests mirrored in tests/.  Beyond that: "parity unpinned".
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load this library.
#include <cstring>
#include <functional>
#include <thread>
#include <pthread.h>
#include "mutations.hpp"

namespace eo {

static int64_t erl_round(double x) { return (int64_t)(x >= 0 ? std::floor(x + 0.5) : -std::floor(-x + 0.5)); }

struct CaseRunner {
    Rng& rng; const Opts& opts; Mutations& muts; std::vector<MutNode> fs; Meta& meta;
    const std::vector<std::pair<int, int>>& sorted_pats;   // {pri, PatId} after sort_by_priority
    int pat_sum;
    using Cont = std::function<Blocks(const Blocks&)>;
    // file / stdin generators hand the pattern an UNFORCED stream (port_stream/2 returns a fun, src/erlamsa_gen.erl:59-61): the
    // blocks -- one rand_block_size draw each -- and the finish/1 tail come into being at the pattern's first uncons/2, i.e.
    // after the pattern's own first draws (SURVEY.md appendix A, W1'')
    std::function<Blocks()> lazy;
    Blocks force(const Blocks& ll) { if (!lazy) return ll; auto f = lazy; lazy = nullptr; return f(); }

    // split/1 :45-60 -- oversize head blocks are cut up (with draws)
    void split_big(Bin& th, Blocks& rest) {
        if (th.size() <= ABSMAX_BINARY_BLOCK) return;
        std::vector<Bin> pieces; Bin cur = th;
        while (cur.size() > ABSMAX_BINARY_BLOCK) {
            uint64_t s = ABSMAXHALF_BINARY_BLOCK;
            uint64_t as = s + rng.rand(s) - 1;
            pieces.push_back(cur.substr(0, as)); cur = cur.substr(as);
        }
        pieces.push_back(cur);
        th = pieces[0];
        rest.insert(rest.begin(), pieces.begin() + 1, pieces.end());
    }
    // mutate_once_loop/6 :283-296
    Blocks mutate_once_loop(const Cont& cont, uint64_t ip, Bin th, Blocks ll) {
        Blocks out;
        for (;;) {
            uint64_t n = rng.rand(ip);
            if (n == 0 || ll.empty()) {
                Blocks arg; arg.push_back(th); arg.insert(arg.end(), ll.begin(), ll.end());
                Blocks l = muts.mux_fuzzers(fs, arg, &meta);
                Blocks rest = cont(l);
                out.insert(out.end(), rest.begin(), rest.end());
                return out;
            }
            out.push_back(th); th = ll[0]; ll.erase(ll.begin());
        }
    }
    // mutate_once/4 :267-278
    Blocks mutate_once(const Blocks& ll0, const Cont& cont) {
        if (!lazy && ll0.size() == 1 && ll0[0].empty()) return Blocks();   // {Mutator, Meta}: nothing is written
        uint64_t ip = rng.rand(INITIAL_IP);
        Blocks ll = force(ll0);
        if (ll.empty()) return cont(Blocks());
        Bin th = ll[0]; Blocks rest(ll.begin() + 1, ll.end());
        split_big(th, rest);
        return mutate_once_loop(cont, ip, th, rest);
    }
    Cont cont_od() { return [](const Blocks& l) { return l; }; }
    Cont cont_nd() {   // pat_many_dec_cont :315-321
        return [this](const Blocks& l) -> Blocks {
            if (rng.rand_occurs_fixed(4, 5)) return mutate_once(l, cont_nd());
            return l;
        };
    }
    Cont cont_bu() {   // pat_burst_cont :332-343
        return [this](const Blocks& l0) -> Blocks {
            Blocks l = l0;
            for (int n = 1;; n++) {
                bool p = rng.rand_occurs_fixed(4, 5);
                if (!(p || n < 2)) return l;
                l = muts.mux_fuzzers(fs, l, &meta);
            }
        };
    }
    Cont cont_pat(int pat) { return [this, pat](const Blocks& l) { return run_pattern(pat, l); }; }

    static uint32_t crc32_of(const uint8_t* p, size_t n) { return (uint32_t)::crc32(0L, p, (uInt)n); }
    struct Csum { bool crc; uint64_t plen, blen; };
    std::vector<Csum> get_possible_csum_locations(const Bin& bin) {   // src/erlamsa_field_predict.erl:154-161
        std::vector<Csum> out; uint64_t len = bin.size();
        if (len == 0) return out;
        uint64_t maxa = std::min<uint64_t>((uint64_t)std::trunc(2.0 * (double)len / 3.0), 30 * PREAMBLE_MAX_BYTES);
        const uint8_t* d = (const uint8_t*)bin.data();
        // xor8 of [A, Len-1) via suffix xors
        std::vector<uint8_t> sx(len + 1, 0);
        for (uint64_t i = len - 1; i-- > 0;) sx[i] = sx[i + 1] ^ d[i];   // xor over [i, len-1)
        for (uint64_t a = 0; a <= maxa; a++) if (sx[a] == d[len - 1]) out.push_back({false, a, len - a - 1});
        for (uint64_t a = 0; a <= maxa; a++) {
            if (len - a < 4 || a > len) continue;
            uint32_t c = ((uint32_t)d[len - 4] << 24) | ((uint32_t)d[len - 3] << 16) | ((uint32_t)d[len - 2] << 8) | d[len - 1];
            if (crc32_of(d + a, len - a - 4) == c) out.push_back({true, a, len - a - 4});
        }
        return out;
    }
    static bool looks_like_zip(const Bin& b) { return b.find(std::string("PK\x05\x06", 4)) != Bin::npos; }
    static bool maybe_compressed(const Bin& b) {
        if (b.size() >= 2 && (uint8_t)b[0] == 0x1f && (uint8_t)b[1] == 0x8b) return true;   // gzip magic
        if (b.size() >= 2) { unsigned cmf = (uint8_t)b[0], flg = (uint8_t)b[1]; if ((cmf & 15) == 8 && (cmf >> 4) <= 7 && ((cmf << 8) | flg) % 31 == 0) return true; }
        return false;
    }

    Blocks run_pattern(int pat, const Blocks& ll) {
        switch (pat) {
        case P_OD: return mutate_once(ll, cont_od());
        case P_ND: return mutate_once(ll, cont_nd());
        case P_BU: return mutate_once(ll, cont_bu());
        case P_CO: if (rng.erand(2) == 1) return run_pattern(P_NU, ll); return run_pattern(P_OD, ll);
        case P_NU: { Blocks o = force(ll); if (!o.empty()) { Bin th = o[0]; Blocks rest(o.begin() + 1, o.end()); split_big(th, rest); o.clear(); o.push_back(th); o.insert(o.end(), rest.begin(), rest.end()); } return o; }
        default: break;
        }
        // make_complex_pat :352-357: the continuation is drawn uniformly from all ten patterns
        int next = (int)rng.rand_elem_idx(P_COUNT);
        uint64_t ip = rng.rand(INITIAL_IP);
        Blocks llf = force(ll);
        if (llf.empty()) throw CaseDied("complex pattern on an empty block list");
        Bin bin = llf[0]; Blocks rest(llf.begin() + 1, llf.end());
        if (pat == P_SK) {   // mutate_once_skipper :148-161
            uint64_t len = rng.rand((uint64_t)std::trunc((double)bin.size() / 2.0));
            Bin head = bin.substr(0, len), th = bin.substr(len);
            split_big(th, rest);
            Blocks res = mutate_once_loop(cont_pat(next), ip, th, rest);
            res.insert(res.begin(), head); return res;
        }
        if (pat == P_SZ) {   // mutate_once_sizer :83-111
            auto els = muts.get_possible_simple_lens(bin);
            int64_t ei = rng.rand_elem_idx(els.size());
            if (ei < 0) { split_big(bin, rest); return mutate_once_loop(cont_pat(next), ip, bin, rest); }
            auto e = els[ei]; int fb = e.size / 8;
            Bin h = bin.substr(0, e.a), blob = bin.substr(e.a + fb, e.len), tail = bin.substr(e.a + fb + e.len);
            split_big(blob, rest);
            Blocks inner = mutate_once_loop(cont_pat(next), ip, blob, rest);
            Bin nb; for (auto& b : inner) nb += b;
            Blocks o; o.push_back(h + Mutations::enc(nb.size(), e.size, e.big) + nb); o.push_back(tail); return o;
        }
        if (pat == P_CS) {   // mutate_once_csum :117-144
            auto els = get_possible_csum_locations(bin);
            int64_t ei = rng.rand_elem_idx(els.size());
            if (ei < 0) { split_big(bin, rest); return mutate_once_loop(cont_pat(next), ip, bin, rest); }
            auto e = els[ei];
            Bin p = bin.substr(0, e.plen), blob = bin.substr(e.plen, e.blen);
            split_big(blob, rest);
            Blocks inner = mutate_once_loop(cont_pat(next), ip, blob, rest);
            Bin nb; for (auto& b : inner) nb += b;
            Bin c;
            if (e.crc) { uint32_t v = crc32_of((const uint8_t*)nb.data(), nb.size()); c = Mutations::enc(v, 32, true); }
            else { uint8_t x = 0; for (char ch : nb) x ^= (uint8_t)ch; c.push_back((char)x); }
            Blocks o; o.push_back(p + nb + c); return o;
        }
        if (pat == P_AR) {   // mutate_once_archiver :167-214, non-archive input only
            Bin all = bin; for (auto& b : rest) all += b;
            if (looks_like_zip(all)) throw Unsupported("ar pattern on ZIP-looking data");
            Blocks none; split_big(all, none);
            return mutate_once_loop(cont_pat(next), ip, all, none);
        }
        // P_CP: mutate_once_compressed :217-260, non-compressed input only
        if (maybe_compressed(bin)) throw Unsupported("cp pattern on compressed-looking data");
        // a block too short to hold the two header bytes: zlib:inflate/2 is a streaming call and answers [] instead of raising, so the
        // reference takes the compressed branch (mutates <<>> and deflates it) -- not restated, and the engine passes the block through
        if (bin.size() < 2) throw Unsupported("cp pattern on a block shorter than a zlib header");
        split_big(bin, rest);
        return mutate_once_loop(cont_pat(next), ip, bin, rest);
    }
};

struct Fuzzer {
    Opts opts; Rng parent; Mutations muts; std::vector<MutNode> fs0;
    std::vector<std::pair<int, int>> sorted_pats; int pat_sum = 0;
    int generator = 0;   // 0 direct, 1 random, 2 file, 3 stdin (n == 1), 4 jump
    const uint8_t* corpus_data = nullptr; const uint64_t* corpus_off = nullptr; uint64_t corpus_n = 0;   // for the file generator's path choice

    explicit Fuzzer(const Opts& o) : opts(o), muts(parent, opts) {
        parent.seed(opts.seed[0], opts.seed[1], opts.seed[2]);   // :134
        muts.build_table_draws();                                  // make_mutator -> mutations/1
        fs0 = muts.make_mutator_nodes();                           // mutators_mutator/1
        // make_generator :233-236 + mux_generators :194-199 for paths == [direct]
        std::vector<std::pair<int, int>> gs;                       // {pri, kind}; option order: random, ..., direct
        if (opts.gen_random_pri >= 0) gs.push_back({opts.gen_random_pri, 1});
        if (opts.gen_jump_pri >= 0) gs.push_back({opts.gen_jump_pri, 4});       // erlamsa_gen:generators/0 :239-246 lists jump between random and direct
        if (opts.gen_direct_pri >= 0) gs.push_back({opts.gen_direct_pri, 0});
        if (opts.gen_file_pri >= 0) gs.push_back({opts.gen_file_pri, 2});
        if (opts.gen_stdin_pri >= 0) gs.push_back({opts.gen_stdin_pri, 3});
        if (gs.empty()) throw Unsupported("no generators");
        generator = choose_pri(sort_by_priority(gs), parent, true);
        // make_pattern :409-422: foldl prepends -> reversed table order, then sort_by_priority
        std::vector<std::pair<int, int>> ps;
        for (int i = P_COUNT; i-- > 0;) if (opts.pat_pri[i] >= 0) ps.push_back({opts.pat_pri[i], i});
        sorted_pats = sort_by_priority(ps);
        for (auto& p : sorted_pats) pat_sum += p.first;
    }
    static std::vector<std::pair<int, int>> sort_by_priority(const std::vector<std::pair<int, int>>& l) {   // src/erlamsa_utils.erl:114-117
        ErlSort<std::pair<int, int>> s([](const std::pair<int, int>& a, const std::pair<int, int>& b) { return a.first > b.first; });
        return s.sort(l);
    }
    static int choose_pri(const std::vector<std::pair<int, int>>& sorted, Rng& rng, bool) {   // :155-160
        int sum = 0; for (auto& p : sorted) sum += p.first;
        int64_t n = (int64_t)rng.rand((uint64_t)sum);
        for (auto& p : sorted) { if (n == 0) return p.second; if (n < p.first) return p.second; n -= p.first; }
        throw CaseDied("choose_pri: function_clause");
    }
    // finish/1 :43-51
    void finish(Rng& rng, uint64_t len, Blocks& out) {
        uint64_t n = rng.rand(len + 1);
        if (n != len) return;
        uint64_t bits = (uint64_t)rng.rand_range(1, 16);
        uint64_t nlen = rng.rand(1ull << bits);
        Bin t = rng.random_numbers_256(nlen);
        if (!t.empty()) out.push_back(t);
    }
    uint64_t rand_block_size(Rng& rng) {   // :55-56
        return std::max<uint64_t>(rng.rand((uint64_t)erl_round(MAX_BLOCK_SIZE * opts.blockscale)), (uint64_t)erl_round(MIN_BLOCK_SIZE * opts.blockscale));
    }
    // stream_port/5 :65-88 over an in-memory file: full blocks of the wanted size (a new size is drawn after each), the short
    // rest as the last block, then finish/1 of the total length
    Blocks stream_blocks(Rng& rng, const Bin& input) {
        Blocks ll; uint64_t pos = 0, n = input.size();
        uint64_t wanted = rand_block_size(rng);
        while (n - pos >= wanted) { ll.push_back(input.substr(pos, wanted)); pos += wanted; wanted = rand_block_size(rng); }
        if (pos < n) ll.push_back(input.substr(pos));
        finish(rng, n, ll);
        return ll;
    }
    // jump_somewhere/2 :123-132, forced at the pattern's first uncons/2 (every pattern starts with Ip = rand(24), then uncons/2, whose
    // function clause calls the fun and whose binary clause makes the result a one-block list, src/erlamsa_utils.erl:89-93): both files
    // are streamed to the end -- one rand_block_size draw per block and the finish/1 draws, stream_port/5 is not lazy -- and one block
    // of each is picked; then S1, S2, L1, L2 in that order. rand_elem([]) is [] and size([]) raises: the case's process dies.
    Bin jump_block(Rng& rng, const Bin& f1, const Bin& f2) {
        Blocks l1 = stream_blocks(rng, f1);
        if (l1.empty()) throw CaseDied("jump_somewhere: size([])");
        const Bin d1 = l1[rng.erand(l1.size()) - 1];
        Blocks l2 = stream_blocks(rng, f2);
        // (Data1 is bound before Ll2() runs, but size(Data1) is only evaluated after both: an empty second list dies all the same)
        if (l2.empty()) throw CaseDied("jump_somewhere: size([])");
        const Bin d2 = l2[rng.erand(l2.size()) - 1];
        uint64_t s1 = rng.rand(d1.size()), s2 = rng.rand(d2.size());
        uint64_t n1 = rng.erand(d1.size() - s1), n2 = rng.erand(d2.size() - s2);
        return d1.substr(s1, n1) + d2.substr(s2, n2);
    }
    Blocks generate(Rng& rng, const Bin& input) {
        Blocks ll;
        if (generator == 0) {   // direct_generator :161-164 (split_binary's first clause never matches)
            (void)rand_block_size(rng);
            ll.push_back(input); finish(rng, input.size(), ll);
        } else {                // random_stream :168-178
            for (;;) {
                uint64_t n = (uint64_t)rng.rand_range(32, erl_round(MAX_BLOCK_SIZE * opts.blockscale));
                ll.push_back(rng.random_block(n));
                uint64_t ip = (uint64_t)rng.rand_range(1, 100);
                if (rng.rand(ip) == 0) break;
            }
        }
        return ll;
    }
    // one iteration of FuzzingLoopFun :166-243; the parent stream advances by three draws
    Bin run_case(const Bin& input, Meta& meta) {
        int64_t ts[3]; parent.gen_predictable_seed(ts);
        return run_case_seeded(input, ts, meta);
    }
    Bin run_case_seeded(const Bin& input, const int64_t ts[3], Meta& meta) {
        Rng rng; rng.seed(ts[0], ts[1], ts[2]);
        meta.thread_seed[0] = ts[0]; meta.thread_seed[1] = ts[1]; meta.thread_seed[2] = ts[2];
        meta.generator = generator;
        Mutations m(rng, opts); m.snand_kind = muts.snand_kind;
        m.thread_seed[0] = ts[0]; m.thread_seed[1] = ts[1]; m.thread_seed[2] = ts[2];
        Bin out;
        try {
            Blocks ll;
            CaseRunner cr{rng, opts, m, fs0, meta, sorted_pats, pat_sum};
            if (generator == 4) {          // jump_streamer :135-150: Path1, Path2 = rand_elem(Paths), the block itself comes later
                if (!corpus_data || corpus_n < 2) throw Unsupported("jump generator needs two or more files");
                uint64_t p1 = rng.erand(corpus_n) - 1, p2 = rng.erand(corpus_n) - 1;
                Bin in1((const char*)corpus_data + corpus_off[p1], corpus_off[p1 + 1] - corpus_off[p1]);
                Bin in2((const char*)corpus_data + corpus_off[p2], corpus_off[p2 + 1] - corpus_off[p2]);
                Rng* rp = &rng;
                cr.lazy = [this, rp, in1, in2]() { Blocks one; one.push_back(jump_block(*rp, in1, in2)); return one; };
            } else if (generator == 2 || generator == 3) {
                Bin in = input;
                if (generator == 2) {      // file_streamer :106-121: P = erand(N) picks the path
                    if (!corpus_data || !corpus_n) throw Unsupported("file generator without a corpus");
                    uint64_t p = rng.erand(corpus_n) - 1;
                    in = Bin((const char*)corpus_data + corpus_off[p], corpus_off[p + 1] - corpus_off[p]);
                }
                Rng* rp = &rng;
                cr.lazy = [this, rp, in]() { return stream_blocks(*rp, in); };
            } else ll = generate(rng, input);
            int pat = choose_pri(sorted_pats, rng, false);
            meta.pattern = pat;
            Blocks res = cr.run_pattern(pat, ll);
            for (auto& b : res) out += b;
        } catch (const Unsupported&) { meta.status = 1; out.clear(); }
        catch (const CaseDied&) { meta.status = 2; out.clear(); }
        catch (const std::overflow_error&) { meta.status = 2; out.clear(); }   // badarith inside random:uniform/1 (rnd.hpp rand_big)
        catch (const CaseOverflow&) { meta.status = 3; out = input; }
        meta.draws = rng.draws;
        return out;
    }
};

}  // namespace eo

// ------------------------------------------------------------------ C API (ctypes)
extern "C" {

struct eo_opts_c {
    int64_t seed[3];
    double blockscale;
    int32_t muta_pri[41];
    int32_t pat_pri[10];
    int32_t gen_direct_pri, gen_random_pri;
    char ssrf_host[64];
    int32_t ssrf_port;
    uint64_t max_case_out;
    const uint8_t* donor_pool; const uint32_t* donor_len; uint64_t n_donors; uint32_t donor_stride; uint32_t pad;
    int32_t gen_file_pri, gen_stdin_pri;
    int32_t gen_jump_pri, pad2;
    int64_t case_stream_seed[3]; uint64_t case_stream_first;   // --workers: see eo_fuzzer
};
struct eo_meta_c {
    int32_t pattern, generator, n_used, n_failed;
    int32_t used[16];
    uint64_t draws;
    int32_t status; int32_t pad;
    int64_t thread_seed[3];
};

static eo::Opts conv(const eo_opts_c* c) {
    eo::Opts o; for (int i = 0; i < 3; i++) o.seed[i] = c->seed[i];
    o.blockscale = c->blockscale;
    for (int i = 0; i < eo::M_COUNT; i++) o.muta_pri[i] = c->muta_pri[i];
    for (int i = 0; i < eo::P_COUNT; i++) o.pat_pri[i] = c->pat_pri[i];
    o.gen_direct_pri = c->gen_direct_pri; o.gen_random_pri = c->gen_random_pri;
    o.ssrf_host = std::string(c->ssrf_host, strnlen(c->ssrf_host, 64)); o.ssrf_port = c->ssrf_port;
    if (c->max_case_out) o.max_case_out = c->max_case_out;
    o.gen_file_pri = c->gen_file_pri; o.gen_stdin_pri = c->gen_stdin_pri; o.gen_jump_pri = c->gen_jump_pri;
    o.donor_pool = c->donor_pool; o.donor_len = c->donor_len; o.n_donors = c->n_donors; o.donor_stride = c->donor_stride;
    return o;
}
static void conv_meta(const eo::Meta& m, eo_meta_c* c) {
    c->pattern = m.pattern; c->generator = m.generator; c->n_used = m.n_used; c->n_failed = m.n_failed;
    for (int i = 0; i < 16; i++) c->used[i] = m.used[i];
    c->draws = m.draws; c->status = m.status; c->pad = 0;
    for (int i = 0; i < 3; i++) c->thread_seed[i] = m.thread_seed[i];
}

void eo_default_opts(eo_opts_c* c) {
    eo::Opts o; memset(c, 0, sizeof(*c));
    for (int i = 0; i < 3; i++) c->seed[i] = o.seed[i];
    c->blockscale = 1.0;
    for (int i = 0; i < eo::M_COUNT; i++) c->muta_pri[i] = o.muta_pri[i];
    for (int i = 0; i < eo::P_COUNT; i++) c->pat_pri[i] = o.pat_pri[i];
    c->gen_direct_pri = 500; c->gen_random_pri = 1; c->gen_file_pri = -1; c->gen_stdin_pri = -1; c->gen_jump_pri = -1;
    strcpy(c->ssrf_host, "localhost"); c->ssrf_port = 51234;
}

// Runs cases first_case .. first_case+n_cases-1 (1-based, as the I of FuzzingLoopFun) of ONE
// fuzzer/1 call; case I reads corpus blob (I-1) mod n_blobs. Outputs are packed; out_off has n_cases+1 entries.
// Everything runs on a thread with a large stack (tree mutators recurse by nesting depth).
struct RunArgs {
    const eo_opts_c* opts; const uint8_t* data; const uint64_t* off; uint64_t n_blobs; uint64_t first_case, n_cases;
    std::string* out; uint64_t* out_off; eo_meta_c* meta; int rc;
};
static void* run_thread(void* p) {
    RunArgs* a = (RunArgs*)p;
    try {
        eo::Fuzzer f(conv(a->opts));
        f.corpus_data = a->data; f.corpus_off = a->off; f.corpus_n = a->n_blobs;
        eo::Meta skip;
        uint64_t stream_first = 1;
        if (a->opts->case_stream_first) {
            // multi-threaded mode, run_fuzzing_loop/7 (src/erlamsa_main.erl:254-280): a worker process is re-seeded with its own seed S and
            // FuzzingLoop draws the thread seed of its first case (number A) as the FIRST draws of that stream; the mutator table, the
            // generator and the pattern list were made by the parent before and are shared
            f.parent.seed(a->opts->case_stream_seed[0], a->opts->case_stream_seed[1], a->opts->case_stream_seed[2]);
            stream_first = a->opts->case_stream_first;
            if (a->first_case < stream_first) throw std::runtime_error("first_case before the worker's first case");
        }
        for (uint64_t i = stream_first; i < a->first_case; i++) { int64_t ts[3]; f.parent.gen_predictable_seed(ts); }
        a->out_off[0] = 0;
        for (uint64_t k = 0; k < a->n_cases; k++) {
            uint64_t b = (a->first_case - 1 + k) % a->n_blobs;
            eo::Bin in((const char*)a->data + a->off[b], a->off[b + 1] - a->off[b]);
            eo::Meta m; eo::Bin o = f.run_case(in, m);
            a->out->append(o); a->out_off[k + 1] = a->out->size();
            if (a->meta) conv_meta(m, &a->meta[k]);
        }
        a->rc = 0;
    } catch (const std::exception&) { a->rc = -1; }
    return nullptr;
}
static int on_big_stack(void* (*fn)(void*), void* arg) {
    pthread_attr_t at; pthread_attr_init(&at); pthread_attr_setstacksize(&at, (size_t)4 << 30);
    pthread_t th; if (pthread_create(&th, &at, fn, arg)) return -1;
    pthread_join(th, nullptr); pthread_attr_destroy(&at); return 0;
}

static thread_local std::string g_out;
int eo_fuzzer(const eo_opts_c* opts, const uint8_t* data, const uint64_t* off, uint64_t n_blobs,
              uint64_t first_case, uint64_t n_cases, const uint8_t** out_data, uint64_t* out_off, eo_meta_c* meta) {
    g_out.clear();
    RunArgs a{opts, data, off, n_blobs, first_case, n_cases, &g_out, out_off, meta, -1};
    if (on_big_stack(run_thread, &a)) return -2;
    *out_data = (const uint8_t*)g_out.data();
    return a.rc;
}

// ---- unit hooks (mirror the reference's eunit style: seed, run one mutator on one block)
struct UnitArgs { const eo_opts_c* opts; int muta; const int64_t* seed; const uint8_t* in; uint64_t len; const uint8_t* next; uint64_t next_len; int rounds; std::string* out; double delta; int rc; };
static void* unit_thread(void* p) {
    UnitArgs* a = (UnitArgs*)p;
    try {
        eo::Opts o = conv(a->opts); eo::Rng rng; rng.seed(a->seed[0], a->seed[1], a->seed[2]);
        eo::Mutations m(rng, o); m.snand_kind = 0;
        if (a->muta == eo::M_SNAND || a->muta == eo::M_SRND) m.build_table_draws();
        eo::MutNode node; node.name = node.fn = a->muta; node.pri = 1; node.score = 10;
        eo::Blocks ll; ll.push_back(eo::Bin((const char*)a->in, a->len));
        if (a->next) ll.push_back(eo::Bin((const char*)a->next, a->next_len));
        eo::MutRes r;
        for (int i = 0; i < a->rounds; i++) { r = m.apply(node, ll); ll = r.ll; }
        a->out->clear(); for (auto& b : r.ll) a->out->append(b);
        a->delta = r.delta; a->rc = 0;
    } catch (const eo::Unsupported&) { a->rc = 1; } catch (const eo::CaseDied&) { a->rc = 2; } catch (...) { a->rc = -1; }
    return nullptr;
}
static thread_local std::string g_unit;
// Applies mutator `muta` `rounds` times (closure state carried) to [in | next?]; returns concatenated blocks.
int eo_run_mutator(const eo_opts_c* opts, int muta, const int64_t seed[3], const uint8_t* in, uint64_t len,
                   const uint8_t* next, uint64_t next_len, int rounds, const uint8_t** out, uint64_t* out_len, double* delta) {
    UnitArgs a{opts, muta, seed, in, len, next, next_len, rounds, &g_unit, 0, -1};
    if (on_big_stack(unit_thread, &a)) return -2;
    *out = (const uint8_t*)g_unit.data(); *out_len = g_unit.size(); if (delta) *delta = a.delta;
    return a.rc;
}

// raw RNG access for known-answer tests
static thread_local eo::Rng g_rng;
void eo_rnd_seed(int64_t a, int64_t b, int64_t c) { g_rng.seed(a, b, c); g_rng.draws = 0; }
void eo_rnd_seed0(void) { g_rng = eo::Rng(); }
double eo_rnd_uniform(void) { return g_rng.uniform(); }
uint64_t eo_rnd_rand(uint64_t n) { return g_rng.rand(n); }
uint64_t eo_rnd_erand(uint64_t n) { return g_rng.erand(n); }
void eo_rnd_state(int64_t s[3]) { s[0] = g_rng.a1; s[1] = g_rng.a2; s[2] = g_rng.a3; }

// lists:sort/2 restatement on integer keys: cmp 0 => A =< B ... see tests. `kind`: 0 strict '>' on key, 1 '>=' on key
void eo_lists_sort(const int32_t* keys, int32_t n, int32_t kind, int32_t* perm_out) {
    std::vector<std::pair<int, int>> v; for (int i = 0; i < n; i++) v.push_back({keys[i], i});
    eo::ErlSort<std::pair<int, int>> s(kind == 0
        ? eo::ErlSort<std::pair<int, int>>::Fun([](const std::pair<int, int>& a, const std::pair<int, int>& b) { return a.first > b.first; })
        : eo::ErlSort<std::pair<int, int>>::Fun([](const std::pair<int, int>& a, const std::pair<int, int>& b) { return a.first >= b.first; }));
    auto r = s.sort(v); for (int i = 0; i < n; i++) perm_out[i] = r[i].second;
}

// strlex round trip
int eo_lex_unlex(const uint8_t* in, uint64_t len, const uint8_t** out, uint64_t* out_len, int32_t* n_chunks) {
    eo::Chunks cs = eo::lex(eo::Bin((const char*)in, len)); g_unit = eo::unlex(cs);
    *out = (const uint8_t*)g_unit.data(); *out_len = g_unit.size(); if (n_chunks) *n_chunks = (int32_t)cs.size(); return 0;
}
// sgml: the stack-discipline builder against the clause-by-clause one (returns 0 when they agree or both refuse)
int eo_sgml_selfcheck(const uint8_t* in, uint64_t len) {
    using namespace eo::sgml;
    eo::Bin s((const char*)in, len);
    std::vector<Token> tk;
    try { tk = tokenize(s); } catch (const TokError&) { return 0; }
    BuildRes a = build_ast2(tk, 0, List(), {}, 0, 0);
    BuildOut b = build_ast_iter(tk);
    if (a.k != BuildRes::OK) return 1;
    eo::Bin fa, fb; fold_ast(a.list, fa); fold_ast(b.list, fb);
    if (fa != fb) return 2;
    if (a.n != b.n || a.nt != b.nt) return 3;
    return equal(a.list, b.list) ? 0 : 4;
}
int eo_tree_selfcheck(const uint8_t* in, uint64_t len) {
    eo::TreeParser a{in, (size_t)len}, b{in, (size_t)len};
    return a.parse() == b.parse_ref() ? 0 : 1;
}
int eo_funny_unicode_count(void) { return (int)eo::Mutations::funny_unicode().size(); }
const char* eo_mutator_code(int i) { return (i >= 0 && i < eo::M_COUNT) ? eo::MUT_CODES[i] : nullptr; }
const char* eo_pattern_code(int i) { return (i >= 0 && i < eo::P_COUNT) ? eo::PAT_CODES[i] : nullptr; }

}  // extern "C"
