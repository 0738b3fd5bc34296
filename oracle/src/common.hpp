// ORACLE (test infrastructure only -- never linked into the product path).
// Shared types for the CPU restatement of erlamsa's mutation hot path.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include <stdexcept>
#include "rnd.hpp"

namespace eo {

using Bin = std::string;              // an Erlang binary
using Blocks = std::vector<Bin>;      // "Ll": list of blocks of one test case

// src/erlamsa.hrl:44-58
constexpr int INITIAL_IP = 24;
constexpr uint64_t AVG_BLOCK_SIZE = 2048;
constexpr uint64_t MIN_BLOCK_SIZE = 256;
constexpr uint64_t MAX_BLOCK_SIZE = 2 * AVG_BLOCK_SIZE;
constexpr uint64_t ABSMAXHALF_BINARY_BLOCK = 500000;
constexpr uint64_t ABSMAX_BINARY_BLOCK = 2 * ABSMAXHALF_BINARY_BLOCK;
constexpr uint64_t SIZER_MAX_FIRST_BYTES = 512;
constexpr uint64_t PREAMBLE_MAX_BYTES = 32;

// Mutator table order, src/erlamsa_mutations.erl:1291-1331
enum MutId {
    M_SGM = 0, M_JS, M_UW, M_UI, M_AB, M_AD, M_TR2, M_TD, M_NUM, M_TS1, M_TR, M_TS2,
    M_BD, M_BEI, M_BED, M_BF, M_BI, M_BER, M_BR, M_SP, M_SR, M_SD, M_SNAND, M_SRND,
    M_LD, M_LDS, M_LR2, M_LRI, M_LR, M_LS, M_LP, M_LIS, M_LRS, M_FT, M_FN, M_FO,
    M_LEN, M_B64, M_URI, M_ZIP, M_NIL, M_COUNT
};
static const char* const MUT_CODES[M_COUNT] = {
    "sgm", "js", "uw", "ui", "ab", "ad", "tr2", "td", "num", "ts1", "tr", "ts2",
    "bd", "bei", "bed", "bf", "bi", "ber", "br", "sp", "sr", "sd", "snand", "srnd",
    "ld", "lds", "lr2", "lri", "lr", "ls", "lp", "lis", "lrs", "ft", "fn", "fo",
    "len", "b64", "uri", "zip", "nil"};
static const int MUT_DEFAULT_PRI[M_COUNT] = {
    10, 3, 1, 2, 1, 1, 1, 1, 3, 2, 2, 2,
    1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
    1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 1, 2,
    2, 7, 1, 1, 0};

// Pattern table order, src/erlamsa_patterns.erl:395-404
enum PatId { P_OD = 0, P_ND, P_BU, P_SK, P_SZ, P_CS, P_AR, P_CP, P_CO, P_NU, P_COUNT };
static const char* const PAT_CODES[P_COUNT] = {"od", "nd", "bu", "sk", "sz", "cs", "ar", "cp", "co", "nu"};
static const int PAT_DEFAULT_PRI[P_COUNT] = {1, 2, 1, 2, 2, 1, 1, 1, 0, 0};

struct Opts {
    int64_t seed[3] = {1, 2, 3};
    double blockscale = 1.0;
    int muta_pri[M_COUNT];   // -1: not selected
    int pat_pri[P_COUNT];    // -1: not selected
    int gen_direct_pri = 500;   // src/erlamsa_gen.erl:246-253 (-1: not selected)
    int gen_random_pri = 1;
    int gen_file_pri = -1;      // paths = files (src/erlamsa_gen.erl:104-121): case picks blob erand(N), lazily split into random-size blocks
    int gen_stdin_pri = -1;     // paths = ["-"] with n == 1 (:92-102): the blob is the stdin data, split lazily like a file
    int gen_jump_pri = -1;      // >= 2 paths (:123-150): the case's one block = a random slice of a random block of one file ++ the same of another
    std::string ssrf_host = "localhost";   // get_ssrf_ep/0 default, src/erlamsa_mutations.erl:697-702
    int ssrf_port = 51234;
    // cross-seed donor pool for sed_fuse_old (BASELINE config C5; not a reference option, see fuse_old in mutations.hpp)
    const uint8_t* donor_pool = nullptr; const uint32_t* donor_len = nullptr; uint64_t n_donors = 0; uint32_t donor_stride = 0;
    uint64_t max_case_out = 64ull << 20;   // not a reference option: the harness' guard against runaway repeats (cf. maxrunningtime)
    Opts() {
        for (int i = 0; i < M_COUNT; i++) muta_pri[i] = MUT_DEFAULT_PRI[i];
        for (int i = 0; i < P_COUNT; i++) pat_pri[i] = PAT_DEFAULT_PRI[i];
    }
};

// thrown when a path needs something the oracle does not restate (zip, zlib, sgml, json...)
struct Unsupported : std::runtime_error { using std::runtime_error::runtime_error; };
// thrown where the reference's worker process would crash (case yields empty output)
struct CaseDied : std::runtime_error { using std::runtime_error::runtime_error; };
// the case grew past the harness' output cap (the reference would run into its maxrunningtime watchdog instead)
struct CaseOverflow : std::runtime_error { using std::runtime_error::runtime_error; };

// per-case record, compared 1:1 with the engine's eb200_meta
struct Meta {
    int pattern = -1;          // PatId chosen first
    int generator = 0;         // 0 direct, 1 random
    int n_used = 0;            // successful mutator applications
    int n_failed = 0;          // mutator attempts that left hd(Ll) unchanged
    int used[16];              // first 16 used mutator ids
    uint64_t draws = 0;        // RNG draws consumed by the worker stream
    int status = 0;            // 0 ok, 1 unsupported, 2 died
    int64_t thread_seed[3] = {0, 0, 0};
    Meta() { for (int& u : used) u = -1; }
};

// erlamsa_utils:binarish/1, src/erlamsa_utils.erl:238-247
inline bool binarish(const Bin& b) {
    const unsigned char* p = (const unsigned char*)b.data(); size_t n = b.size();
    for (size_t i = 0;; i++) {
        size_t left = n - i;
        if (left >= 3 && p[i] == 0xEF && p[i + 1] == 0xBB && p[i + 2] == 0xBF) return false;
        if (left >= 2 && p[i] == 0xFE && p[i + 1] == 0x0F) return false;
        if (i == 8) return false;
        if (left == 0) return false;
        if (p[i] == 0) return true;
        if (p[i] & 128) return true;
    }
}

// erlamsa_utils:flush_bvecs/2, src/erlamsa_utils.erl:169-175
inline void flush_bvecs(const Bin& bin, Blocks& out) {
    size_t pos = 0, len = bin.size();
    while (len >= AVG_BLOCK_SIZE) { out.push_back(bin.substr(pos, AVG_BLOCK_SIZE)); pos += AVG_BLOCK_SIZE; len -= AVG_BLOCK_SIZE; }
    out.push_back(bin.substr(pos));
}

}  // namespace eo
