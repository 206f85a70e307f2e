// C++ host-side mirror exercised the way the reference's own unit tests exercise the Rust API.
// Exit codes: 0 all checks passed, 3 no CUDA device (the library refuses to compute on the CPU), 1 a check failed.
#include <cstdio>
#include <string>
#include <vector>

#include "kmers_b200.hpp"

using namespace kmers_b200;
using encoding::Naive;
using encoding::Xor10;

static int failures = 0;
#define CHECK(cond)                                                        \
    do {                                                                   \
        if (!(cond)) { std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); ++failures; } \
    } while (0)

// naive_impl::Kmer::from(&str): base i at bits 2i+1:2i, A0 C1 G2 T3 (naive_impl/kmer.rs:209-232)
static uint64_t kmer_word(const std::string& s) {
    uint64_t w = 0;
    for (size_t i = 0; i < s.size(); ++i) {
        char c = s[i] | 0x20;
        uint64_t code = c == 'a' ? 0 : c == 'c' ? 1 : c == 'g' ? 2 : 3;
        w |= code << (2 * i);
    }
    return w;
}

int main() {
    if (kmb_device_count() == 0) {
        try {
            Context ctx(0);
        } catch (const Error& e) {
            std::printf("no device: %s\n", e.what());
            return e.code == KMB_ERR_NO_DEVICE ? 3 : 1;
        }
        return 1;
    }
    Context ctx(0);
    const std::string read =
        "TTTTGGCCATTTTTCCTGTTCTTCAAGAAAACAGGAGATAACTAGAAGGACTAGAGAATGGGGCTGCCAGAACTAGTGGGAAGCTCCCTAGAAATGGTGACATCGCCCACCAAACAGACC";

    {  // canonical_kmer_iterator.rs:123-134 test_iter_init / :137-148 test_iter_inc / :192-206 exhaustion
        auto res = ctx.upload({read}).canonical_kmers(31, /*want_fw_rc=*/true);
        CHECK(res.canon.size() == read.size() - 30);
        CHECK(res.fw[0] == kmer_word(read.substr(0, 31)));
        CHECK(res.fw[1] == kmer_word(read.substr(1, 31)));
        CHECK(res.fw[10] == kmer_word(read.substr(10, 31)));
        CHECK(res.digest.n_valid == read.size() - 30);
        auto rc = ctx.reverse_complement({res.fw[0]}, 31);
        CHECK(res.rc[0] == rc[0]);
        CHECK(res.canon[0] == (res.fw[0] < res.rc[0] ? res.fw[0] : res.rc[0]));
    }
    {  // canonical_kmer_iterator.rs:165-175 test_iter_init_invalid : N at index 4 -> first k-mer at pos 5
        std::string r = read.substr(0, 4) + "N" + read.substr(4);
        auto res = ctx.upload({r}).canonical_kmers(31, true);
        for (int p = 0; p < 5; ++p) CHECK(res.canon[p] == KMB_SENTINEL);
        CHECK(res.fw[5] == kmer_word(r.substr(5, 31)));
    }
    {  // naive_impl/kmer.rs:387-424 test_rc
        const char* pairs[][2] = {{"a", "t"}, {"aaa", "ttt"}, {"ta", "ta"}, {"ccg", "cgg"}, {"gatacataggatgg", "ccatcctatgtatc"}};
        for (auto& p : pairs) {
            std::string s = p[0];
            CHECK(ctx.reverse_complement({kmer_word(s)}, (uint32_t)s.size())[0] == kmer_word(p[1]));
        }
    }
    {  // naive_impl/kmer.rs:293-317 test_into_canon / test_is_canon
        std::vector<uint8_t> flag;
        auto c = ctx.to_canonical({kmer_word("tta"), kmer_word("taa")}, 3, &flag);
        CHECK(c[0] == kmer_word("taa") && c[1] == kmer_word("taa") && flag[0] == 0 && flag[1] == 1);
    }
    {  // naive_impl/hash.rs:84-104 lex_order
        auto h = ctx.lex_hash({kmer_word("aaa"), kmer_word("aac"), kmer_word("caa"), kmer_word("cac")}, 3);
        CHECK(h[0] == 0 && h[1] == 1 && h[2] == 0b010000 && h[3] == 0b010001);
    }
    {  // naive_impl/kmer.rs:196-207 String::from(Kmer): lower case, base 0 first
        auto t = ctx.to_strings({kmer_word("gatacataggatgg"), kmer_word("ccatcctatgtatc")}, 14);
        CHECK(t[0] == "gatacataggatgg" && t[1] == "ccatcctatgtatc");
    }
    {  // naive_impl/canonical_kmer.rs:283-297 test_equivalency
        auto m = ctx.get_word_equivalency({kmer_word("acttg"), kmer_word("acttg"), kmer_word("acttg")},
                                          {kmer_word("caagt"), kmer_word("acttg"), kmer_word("cttgc")}, 5);
        CHECK(m[0] == MatchType::TwinMatch && m[1] == MatchType::IdentityMatch && m[2] == MatchType::NoMatch);
    }
    {  // encoding/naive.rs:388-416 k45pu64 and :297-313 k15pu8
        const std::string s45 = "TAAGGATTCTAATCATAAGGATTCTAATCATAAGGATTCTAATCA";
        auto a = ctx.encode<uint64_t, 45>(Naive::ACGT, reinterpret_cast<const uint8_t*>(s45.data()), 1);
        CHECK(a[0][0] == 3585846758293238403ull && a[0][1] == 7397160ull);
        CHECK(ctx.decode(Naive::ACGT, a)[0] == s45 + std::string(19, 'A'));
        auto rc = ctx.rev_comp<45>(Naive::ACGT, a);
        CHECK(ctx.decode(Naive::ACGT, rc)[0] == "TGATTAGAATCCTTATGATTAGAATCCTTATGATTAGAATCCTTA" + std::string(19, 'A'));
        const std::string s15 = "TAAGGATTCTAATCA";
        auto b = ctx.encode<uint8_t, 15>(Naive::ACGT, reinterpret_cast<const uint8_t*>(s15.data()), 1);
        CHECK(b[0][0] == 131 && b[0][1] == 242 && b[0][2] == 13 && b[0][3] == 7);
        auto x = ctx.encode<uint64_t, 45>(Xor10{}, reinterpret_cast<const uint8_t*>(s45.data()), 1);  // xor10.rs:247-275
        CHECK(x[0][0] == 2414607732474225602ull && x[0][1] == 6330940ull);
        static_assert(word_for_k<uint64_t, 32>() == 1 && word_for_k<uint64_t, 64>() == 2 && word_for_k<uint8_t, 5>() == 2, "kmer.rs:97-118");
    }
    {  // naive_impl/kmer.rs:476-480 too_long -> Panic
        bool panicked = false;
        try { ctx.upload({std::string(40, 'A')}).canonical_kmers(33); } catch (const Panic&) { panicked = true; }
        CHECK(panicked);
    }
    {  // K = 63 two-word extension + histogram smoke
        auto batch = ctx.generate(42, 100, 150);
        Digest d;
        auto wide = batch.canonical_kmers_wide(63, Naive::ACGT, &d);
        CHECK(wide.size() == 100 * 88 && d.n_valid == 100 * 88);
        auto hist = batch.histogram(31, 8, &d);
        uint64_t total = 0;
        for (auto v : hist) total += v;
        CHECK(total == d.n_valid && total == 100 * 120);
    }
    {  // canonical_kmer_iterator.rs:178-189 : N at index 35 -> the iterator's positions are 0..4, 36..
        std::string r = read.substr(0, 35) + "N" + read.substr(35);
        auto c = ctx.upload({r, std::string("ACGT")}).canonical_kmer_positions(31);
        CHECK(c.pos.size() == 60 && c.pos[4] == 4 && c.pos[5] == 36 && c.emit_offsets[1] == 60 && c.emit_offsets[2] == 60);
        CHECK(c.canon.size() == 60 && c.hash.size() == 60);
    }
    {  // seq_vector/minimizers.rs:251-290 mmers1 / mmers2 ; naive_impl/kmer.rs:560-579 via the pinned hasher
        auto m1 = ctx.upload({std::string("AACCAAA")}).minimizers(5, 3, 5);
        CHECK(m1.first.size() == 3 && m1.first[0] == 0b010000 && m1.first[1] == 0b010100 && m1.first[2] == 0);
        CHECK(m1.second[0] == 0 && m1.second[1] == 1 && m1.second[2] == 4);
        auto m2 = ctx.upload({std::string("CACACACCAC")}).minimizers(7, 3, 3);
        CHECK(m2.second.size() == 4 && m2.second[0] == 1 && m2.second[1] == 1 && m2.second[2] == 3 && m2.second[3] == 3);
        auto mw = ctx.minimizer_word({kmer_word("ACTTGAT")}, 7, 3, 3);
        CHECK(mw.first[0] == kmer_word("ACT") && mw.second[0] == 0);
    }
    {  // seq_vector.rs:342-358 iter_kmers on the packed store
        auto b = ctx.upload({std::string("ACTTGAT")});
        b.to_packed();
        auto res = b.canonical_kmers(3, true);
        CHECK(res.fw.size() == 5 && res.fw[0] == kmer_word("act") && res.fw[4] == kmer_word("gat"));
        auto km = b.get_kmers_u64(3, {}, {0, 2, 4, 5});
        CHECK(km[0] == kmer_word("act") && km[1] == kmer_word("ttg") && km[2] == kmer_word("gat") && km[3] == KMB_SENTINEL);
    }
    {  // naive_impl/kmer.rs:325-384 test_append / test_prepend
        auto a = ctx.append_base_u8({kmer_word("att"), kmer_word("ttcga")}, "cg", 3);
        CHECK(a.first[0] == kmer_word("ttc") && a.second[0] == 0 /* A */);
        auto a5 = ctx.append_base_u8({kmer_word("ttcga")}, "g", 5);
        CHECK(a5.first[0] == kmer_word("tcgag") && a5.second[0] == 3 /* T */);
        auto p3 = ctx.prepend_base_u8({kmer_word("att")}, "c", 3);
        CHECK(p3.first[0] == kmer_word("cat") && p3.second[0] == 3);
        auto p5 = ctx.prepend_base_u8({kmer_word("ttcga")}, "g", 5);
        CHECK(p5.first[0] == kmer_word("gttcg") && p5.second[0] == 0);
    }
    {  // naive_impl/kmer.rs:530-542 test_sub_kmer
        const std::string s = "ACTTGAT";
        for (size_t i = 0; i < s.size(); ++i)
            for (size_t j = i; j < s.size(); ++j)
                CHECK(ctx.sub_kmer({kmer_word(s)}, 7, i, j - i)[0] == kmer_word(s.substr(i, j - i)));
    }
    {  // kmer.rs:46-53 get / get_prefix on [u64;2] and kmer.rs:71-91 bitmer_to_bytes
        const std::string s45 = "TAAGGATTCTAATCATAAGGATTCTAATCATAAGGATTCTAATCA";
        auto a = ctx.encode<uint64_t, 45>(Naive::ACGT, reinterpret_cast<const uint8_t*>(s45.data()), 1);
        const char* letters = "ACGT";
        for (uint32_t i = 0; i < 45; ++i) CHECK(letters[ctx.get(a, i)[0]] == s45[i]);
        CHECK(ctx.get_prefix(a, 5)[0] == (a[0][0] & 0x7ffull));  // the reference's inclusive 0..=2*len range
        CHECK(ctx.bitmer_to_bytes({kmer_word("acttgat"), kmer_word("ttttttt")}, 7) == (std::vector<std::string>{"ACTTGAT", "TTTTTTT"}));
    }
    {  // the pipelined host call = upload + canonical_kmers on the same reads
        std::string flat;
        std::vector<std::string> reads;
        for (int r = 0; r < 64; ++r) {
            std::string x(50, 'A');
            for (int i = 0; i < 50; ++i) x[i] = "ACGTN"[(r * 7 + i * i + i / 3) % (r % 5 == 0 ? 5 : 4)];
            reads.push_back(x);
            flat += x;
        }
        auto h = ctx.canonical_kmers_host(reinterpret_cast<const uint8_t*>(flat.data()), 64, 50, 21);
        auto d = ctx.upload(reads).canonical_kmers(21);
        CHECK(h.canon == d.canon && h.hash == d.hash && h.digest.n_valid == d.digest.n_valid &&
              h.digest.checksum_canon == d.digest.checksum_canon);
    }
    std::printf(failures ? "FAILED (%d)\n" : "OK\n", failures);
    return failures ? 1 : 0;
}
