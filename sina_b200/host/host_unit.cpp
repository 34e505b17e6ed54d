// Unit checks of the host-side mirror (no GPU needed). Vectors are the reference's own:
// src/unit_tests/cseq_test.cpp:49-52 (strings), :99-131 (append), :133-183 (setWidth), :185-203 (dna),
// :226-243 (case); FASTA reader/writer behaviour follows src/rw_fasta.cpp:229-315,438-528.
#include <algorithm>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>

#include "../csrc/find_layout.h"
#include "align.h"
#include "famfinder.h"
#include "reference_db.h"
#include "rw_fasta.h"

using namespace sina;
static int n_checks = 0, n_fail = 0;
#define CHECK(cond) do { n_checks++; if (!(cond)) { n_fail++; std::cerr << "FAIL " << __FILE__ << ":" << __LINE__ << ": " #cond << std::endl; } } while (0)
#define EQUAL(a, b) do { n_checks++; if (!((a) == (b))) { n_fail++; std::cerr << "FAIL " << __FILE__ << ":" << __LINE__ << ": " #a " == " #b " [" << (a) << " vs " << (b) << "]" << std::endl; } } while (0)
#define THROWS(stmt, ex) do { n_checks++; bool t_ = false; try { stmt; } catch (ex&) { t_ = true; } if (!t_) { n_fail++; std::cerr << "FAIL " << __FILE__ << ":" << __LINE__ << ": no " #ex << std::endl; } } while (0)

static const std::string rna = "AGCURYKMSWBDHVN";
static const std::string rna_aligned = "--A-G---CUR-YKM-S---WBD-HVN---";
static const std::string rna_aligned_dots = "..A-G---CUR-YKM-S---WBD-HVN...";

static std::string strip(std::string s) { s.erase(std::remove(s.begin(), s.end(), '-'), s.end()); return s; }
static std::string lower(std::string s) { for (auto& c : s) c = (char)tolower(c); return s; }
static void test_data(const cseq& c, const std::string& name, const std::string& aligned) {
    EQUAL(c.size(), strip(aligned).size());
    EQUAL(c.getWidth(), aligned.size());
    EQUAL(c.getBases(), strip(aligned));
    EQUAL(c.getAligned(true), aligned);
    EQUAL(c.getName(), name);
}

static void test_cseq() {
    { cseq c; test_data(c, "", ""); }
    { cseq c("thename", rna.c_str()); test_data(c, "thename", rna); cseq d = c; test_data(d, "thename", rna); }
    {
        cseq c;
        c.append(rna); test_data(c, "", rna);
        c.append(""); test_data(c, "", rna);
        c.append(rna); test_data(c, "", rna + rna);
        c.clearSequence(); test_data(c, "", "");
        c.append(rna_aligned); test_data(c, "", rna_aligned);
        c.append(rna); test_data(c, "", rna_aligned + rna);
        c.append(rna_aligned); test_data(c, "", rna_aligned + rna + rna_aligned);
        c.append(aligned_base(0, base_iupac::from_char('A')));  // wrong order: forced to the current width
        EQUAL(c.getAligned(true), rna_aligned + rna + rna_aligned + "A");
        EQUAL(c.getWidth(), 75u);  // the reference's known off-by-one (expected failure in its own test)
    }
    {
        cseq c;
        const std::string g20(20, '-');
        c.setWidth(20); test_data(c, "", g20);
        c.setWidth(40); test_data(c, "", g20 + g20);
        c.setWidth(20); test_data(c, "", g20);
        c.setWidth(0); test_data(c, "", "");
        c.append(rna_aligned);
        c.setWidth(rna_aligned.size() + 20); test_data(c, "", rna_aligned + g20);
        c.setWidth(rna_aligned.size()); test_data(c, "", rna_aligned);
        const char* steps[] = {"--A-G---CUR-YKM-S---WBD-HVN", "--A-G---CUR-YKM-S---WBDHVN", "--A-G---CUR-YKM-S--WBDHVN", nullptr,
                               "--A-G---CUR-YKM-SWBDHVN", "--A-G---CUR-YKMSWBDHVN", "--A-G---CURYKMSWBDHVN", "--A-G--CURYKMSWBDHVN",
                               "--A-G-CURYKMSWBDHVN", "--A-GCURYKMSWBDHVN", "--AGCURYKMSWBDHVN", "-AGCURYKMSWBDHVN", "AGCURYKMSWBDHVN"};
        for (int w = 27, i = 0; w >= 15; w--, i++) {
            if (!steps[i]) continue;  // the reference's test skips width 24
            c.setWidth(w);
            test_data(c, "", steps[i]);
        }
        cseq d("", rna_aligned.c_str());
        THROWS(d.setWidth(14), std::runtime_error);
    }
    {
        std::string r = lower(rna_aligned), d = r, D = rna_aligned;
        std::replace(d.begin(), d.end(), 'u', 't');
        std::replace(D.begin(), D.end(), 'U', 'T');
        cseq c("", r.c_str()), e("", d.c_str());
        EQUAL(c.getAligned(true, true), d); EQUAL(c.getAligned(true, false), r);
        EQUAL(e.getAligned(true, true), d); EQUAL(e.getAligned(true, false), r);
        c.upperCaseAll(); e.upperCaseAll();
        EQUAL(c.getAligned(true, true), D); EQUAL(c.getAligned(true, false), rna_aligned);
        EQUAL(e.getAligned(true, false), rna_aligned);
    }
    { cseq c("", rna_aligned.c_str()); EQUAL(c.getAligned(false), rna_aligned_dots); }
    { cseq c("", lower(rna).c_str()); EQUAL(c.getAligned(true), lower(rna)); c.upperCaseAll(); EQUAL(c.getAligned(true), rna); }
    THROWS(cseq("", "AGCX"), base_iupac::bad_character_exception);
    { cseq c("x", " A G\tC\r\nU . - N"); EQUAL(c.getAligned(true), "AGCU--N"); }
}

static void test_options() {
    po::options_description main_od("m"), adv("a");
    rw_fasta::get_options_description(main_od, adv);
    famfinder::get_options_description(main_od, adv);
    aligner::get_options_description(main_od, adv);
    po::options_description all;
    all.add(main_od).add(adv);
    {   // reference defaults (src/famfinder.cpp:155-195, src/align.cpp:232-259)
        EQUAL(famfinder::opts.fs_kmer_len, 10u); EQUAL(famfinder::opts.fs_min, 40u); EQUAL(famfinder::opts.fs_max, 40u);
        EQUAL(famfinder::opts.fs_req, 1u); EQUAL(famfinder::opts.fs_req_full, 1u); EQUAL(famfinder::opts.fs_full_len, 1400u);
        EQUAL(famfinder::opts.fs_req_gaps, 10u); EQUAL(famfinder::opts.fs_min_len, 150u);
        CHECK(famfinder::opts.fs_msc == 0.7f); CHECK(famfinder::opts.fs_msc_max == 2.f);
        CHECK(aligner::opts->match_score == 2.f); CHECK(aligner::opts->mismatch_score == -1.f);
        CHECK(aligner::opts->gap_penalty == 5.f); CHECK(aligner::opts->gap_ext_penalty == 2.f); CHECK(aligner::opts->fs_weight == 1.f);
        EQUAL((int)aligner::opts->overhang, (int)OVERHANG_ATTACH); EQUAL((int)aligner::opts->lowercase, (int)LOWERCASE_NONE);
        EQUAL((int)aligner::opts->insertion, (int)INSERTION_SHIFT);
    }
    {
        const char* argv[] = {"sina", "--db", "ref.fa", "--fs-engine", "internal", "--fs-kmer-len=8", "--fs-min", "15", "--fs-max", "20",
                              "--fs-msc", "0.5", "--fs-req", "2", "--pen-gap", "4.5", "--pen-gapext", "1.5", "--match-score", "3",
                              "--mismatch-score", "-2", "--overhang", "edge", "--lowercase", "unaligned", "--insertion", "remove",
                              "--realign", "--fs-kmer-no-fast", "--fs-kmer-mm", "1", "--fs-kmer-norel", "-t", "none"};
        po::variables_map vm;
        po::store(sizeof(argv) / sizeof(*argv), argv, all, vm);
        famfinder::validate_vm(vm, all);
        EQUAL(famfinder::opts.database, std::string("ref.fa")); EQUAL(famfinder::opts.fs_kmer_len, 8u);
        EQUAL(famfinder::opts.fs_min, 15u); EQUAL(famfinder::opts.fs_max, 20u); CHECK(famfinder::opts.fs_msc == 0.5f);
        EQUAL(famfinder::opts.fs_req, 2u); CHECK(famfinder::opts.fs_no_fast); EQUAL(famfinder::opts.fs_kmer_mm, 1u);
        CHECK(aligner::opts->gap_penalty == 4.5f); CHECK(aligner::opts->gap_ext_penalty == 1.5f);
        CHECK(aligner::opts->match_score == 3.f); CHECK(aligner::opts->mismatch_score == -2.f);
        EQUAL((int)aligner::opts->overhang, (int)OVERHANG_EDGE); EQUAL((int)aligner::opts->lowercase, (int)LOWERCASE_UNALIGNED);
        EQUAL((int)aligner::opts->insertion, (int)INSERTION_REMOVE); CHECK(aligner::opts->realign);
        EQUAL(vm.count("db"), 1u); EQUAL(vm.count("fs-min-len"), 0u);
    }
    auto rejects = [&](std::vector<const char*> extra) {
        std::vector<const char*> argv{"sina"};
        argv.insert(argv.end(), extra.begin(), extra.end());
        po::variables_map vm;
        try { po::store((int)argv.size(), argv.data(), all, vm); } catch (std::logic_error&) { return true; }
        return false;
    };
    CHECK(rejects({"--fs-engine", "pt-server"})); CHECK(rejects({"--fs-no-graph"})); CHECK(rejects({"--use-subst-matrix"}));
    CHECK(rejects({"--filter", "x"})); CHECK(rejects({"--insertion", "bogus"})); CHECK(!rejects({"--insertion", "forbid"})); CHECK(!rejects({"--insertion", "shift"})); CHECK(rejects({"--overhang", "bogus"}));
    CHECK(rejects({"--no-such-option"})); CHECK(rejects({"--fs-min"})); CHECK(rejects({"--fs-min", "abc"}));
    CHECK(rejects({"--turn", "sideways"})); CHECK(rejects({"stray"}));
    CHECK(!rejects({"--turn", "all"})); CHECK(famfinder::opts.turn_which == TURN_ALL);
    CHECK(!rejects({"--turn", "revcomp"})); CHECK(famfinder::opts.turn_which == TURN_REVCOMP);
    CHECK(!rejects({"--turn", "none"})); CHECK(famfinder::opts.turn_which == TURN_NONE);
    {   // cseq::reverse / complement (src/cseq.cpp:284-296; src/unit_tests/cseq_test.cpp has reverse on aligned data)
        cseq c("", "-AG--CUR-n");
        c.reverse();
        EQUAL(c.getAligned(true, false), std::string("n-RUC--GA-"));
        c.complement();
        EQUAL(c.getAligned(true, false), std::string("n-YAG--CU-"));
        c.reverse(); c.complement();
        EQUAL(c.getAligned(true, false), std::string("-AG--CUR-n"));
    }
    CHECK(!rejects({"--fs-min", "7"}));
    { po::variables_map vm; THROWS(famfinder::validate_vm(vm, all), std::logic_error); }  // --db is mandatory
}

static void test_fasta(const std::string& tmpdir) {
    const std::string in = tmpdir + "/host_unit_in.fasta", out = tmpdir + "/host_unit_out.fasta";
    {
        std::ofstream f(in);
        f << "junk before the first record\n>seq1 first sequence\n; key = value\nAGCU\nagcu\n>bad\nAGXU\n>seq2\r\nAC-GU.N\r\n";
    }
    rw_fasta::reader rd(in);
    tray t1, t2, t3;
    CHECK(rd(t1)); CHECK(rd(t2)); CHECK(!rd(t3));
    EQUAL(t1.seqno, 1u); EQUAL(t1.input_sequence->getName(), std::string("seq1"));
    EQUAL(t1.input_sequence->get_attr_string(fn_fullname), std::string("first sequence"));
    EQUAL(t1.input_sequence->get_attr_string("key"), std::string("value"));
    EQUAL(t1.input_sequence->getBases(), std::string("AGCUagcu"));
    EQUAL(t2.seqno, 3u);  // the skipped record consumed number 2
    EQUAL(t2.input_sequence->getName(), std::string("seq2")); EQUAL(t2.input_sequence->getAligned(true), std::string("AC-GU-N"));
    EQUAL(rd.skipped(), 1u);
    {
        rw_fasta::writer wr(out);
        t1.aligned_sequence = new cseq("seq1", "--AG-CU");
        t1.aligned_sequence->set_attr<std::string>(fn_fullname, "first sequence");
        wr(t1);
        wr(t2);  // not aligned: excluded
        EQUAL(wr.written(), 1u); EQUAL(wr.excluded(), 1u);
    }
    std::ifstream f(out);
    std::stringstream ss; ss << f.rdbuf();
    EQUAL(ss.str(), std::string(">seq1 first sequence\n--AG-CU\n"));
    t1.destroy(); t2.destroy();
    remove(in.c_str()); remove(out.c_str());

    // record_size / format_into against format() over the writer's options and awkward sequences (bases pushed past
    // the alignment width, an empty sequence, lower case, attributes)
    {
        std::vector<cseq> cs;
        cs.emplace_back("plain", "--AG-CU--ACGUACGU-------");
        cs.emplace_back("lower", "acgu--ACGU");
        cs.emplace_back("empty", "-------");
        { cseq c("past", "AC"); c.append(aligned_base(40, 1)); c.setWidth(30); cs.push_back(c); }   // last base beyond the width
        { cseq c("attrs", "--ACGU--"); c.set_attr<std::string>(fn_fullname, "a full name"); c.set_attr(fn_qual, 97); c.set_attr<std::string>("turn", "none"); cs.push_back(c); }
        const FASTA_META_TYPE metas[] = {FASTA_META_NONE, FASTA_META_HEADER, FASTA_META_COMMENT};
        for (FASTA_META_TYPE meta : metas) for (int ll : {0, 7, 60}) for (int dots = 0; dots < 2; dots++) for (int dna = 0; dna < 2; dna++) {
            rw_fasta::opts->fastameta = meta; rw_fasta::opts->line_length = ll; rw_fasta::opts->out_dots = dots; rw_fasta::opts->out_dna = dna;
            std::string rec;
            for (const cseq& c : cs) {
                const std::string want = rw_fasta::writer::format(c);
                rw_fasta::writer::format_into(c, rec);
                EQUAL(rec, want);
                EQUAL(rw_fasta::writer::record_size(c), want.size());
                const std::string seq = c.getAligned(!dots, dna);   // the sequence lines are getAligned(), wrapped
                std::string body = want.substr(want.find('\n') + 1), joined;
                if (meta == FASTA_META_COMMENT) while (!body.empty() && body[0] == ';') body = body.substr(body.find('\n') + 1);
                for (char ch : body) if (ch != '\n') joined.push_back(ch);
                EQUAL(joined, seq);
            }
        }
        rw_fasta::opts->fastameta = FASTA_META_NONE; rw_fasta::opts->line_length = 0; rw_fasta::opts->out_dots = false; rw_fasta::opts->out_dna = false;
    }
    // the block reader: records cut at any buffer boundary, parsing as a separate step
    {
        const std::string big = tmpdir + "/host_unit_big.fasta";
        std::vector<std::string> names, seqs;
        {
            std::ofstream f(big);
            for (int i = 0; i < 3000; i++) {
                std::string s2;
                for (int j = 0; j < 1000 + (i * 37) % 900; j++) s2.push_back("ACGU-"[(i * 7 + j * 13) % 5]);
                names.push_back("r" + std::to_string(i));
                seqs.push_back(s2);
                f << ">" << names.back() << " desc " << i << "\n";
                for (size_t j = 0; j < s2.size(); j += 70) f << s2.substr(j, 70) << (i % 3 == 0 ? "\r\n" : "\n");
            }
        }
        rw_fasta::reader r2(big);
        std::string rec;
        unsigned int seqno = 0, lineno = 0, n = 0;
        while (r2.next_record(rec, seqno, lineno)) {
            tray t;
            CHECK(rw_fasta::reader::parse_record(rec, seqno, lineno, big, t));
            EQUAL(seqno, n + 1);
            if (n < names.size()) {
                EQUAL(t.input_sequence->getName(), names[n]);
                EQUAL(t.input_sequence->getAligned(true), seqs[n]);
            }
            t.destroy();
            n++;
        }
        EQUAL(n, 3000u);
        remove(big.c_str());
    }
    // ".gz" on both sides (src/rw_fasta.cpp:200-202,358-360): the writer functor emits gzip members, the reader takes the
    // concatenated stream back
    {
        const std::string gz = tmpdir + "/host_unit_rt.fasta.gz";
        std::vector<std::string> rows;
        {
            rw_fasta::writer wr(gz);
            CHECK(wr.compressed() && !wr.positional());
            for (int i = 0; i < 40; i++) {
                std::string row(500, '-');
                for (int j = 0; j < 120; j++) row[(size_t)((i * 31 + j * 17) % 500)] = "ACGU"[(i + j) % 4];
                rows.push_back(row);
                tray t;
                t.input_sequence = new cseq(("z" + std::to_string(i)).c_str(), row.c_str());
                t.aligned_sequence = new cseq(*t.input_sequence);
                wr(t);
                t.destroy();
            }
            EQUAL(wr.written(), 40u);
        }
        rw_fasta::reader rd2(gz);
        for (int i = 0; i < 40; i++) {
            tray t;
            CHECK(rd2(t));
            if (t.input_sequence) { EQUAL(t.input_sequence->getName(), "z" + std::to_string(i)); EQUAL(t.input_sequence->getAligned(true), rows[(size_t)i]); }
            t.destroy();
        }
        tray tend;
        CHECK(!rd2(tend));
        remove(gz.c_str());
        // a reference database may be compressed too (SILVA ships .fasta.gz): same rows, same packed form
        {
            const std::string db = tmpdir + "/host_unit_db.fasta", dbz = db + ".gz";
            std::string text;
            for (int i = 0; i < 25; i++) {
                std::string row(70000, '-');   // lines longer than the gz line buffer
                for (int j = 0; j < 900; j++) row[(size_t)((i * 131 + j * 77) % 70000)] = "ACGUN"[(i + j) % 5];
                text += ">ref" + std::to_string(i) + " full name " + std::to_string(i) + "\n" + row + (i % 2 ? "\r\n" : "\n");
            }
            { std::ofstream f(db, std::ios::binary); f << text; }
            { const std::string m = rw_fasta::writer::gzip_member(text.data(), text.size()); std::ofstream f(dbz, std::ios::binary); f.write(m.data(), (std::streamsize)m.size()); }
            reference_db* a = reference_db::getDB(db);
            reference_db* b = reference_db::getDB(dbz);
            EQUAL(a->getSeqCount(), 25u); EQUAL(b->getSeqCount(), 25u);
            EQUAL(a->getAlignmentWidth(), b->getAlignmentWidth());
            CHECK(a->getSequenceNames() == b->getSequenceNames());
            CHECK(a->masks() == b->masks() && a->cols() == b->cols() && a->offsets() == b->offsets());
            EQUAL(b->getCseq(7).get_attr_string(fn_fullname), std::string("full name 7"));
            remove(db.c_str()); remove(dbz.c_str());
        }
        // the loader cuts the file into chunks of 64 records parsed on several threads: order and content survive, comment
        // lines are skipped wherever they stand, a bad character fails the load with its line number, an empty file too
        {
            const std::string db = tmpdir + "/host_unit_db2.fasta", bad = tmpdir + "/host_unit_db3.fasta", none = tmpdir + "/host_unit_db4.fasta";
            std::vector<std::string> rows;
            {
                std::ofstream f(db, std::ios::binary);
                f << "text before the first record\n";
                for (int i = 0; i < 1500; i++) {
                    std::string row(300 + (size_t)(i % 50), '-');
                    for (size_t j = 0; j < row.size(); j += 3 + (size_t)(i % 4)) row[j] = "ACGUacguN"[(i + (int)j) % 9];
                    rows.push_back(row);
                    f << ">d" << i << (i % 5 ? " described" : "") << (i % 2 ? "\r\n" : "\n");
                    if (i % 7 == 0) f << "; a comment\n";
                    f << row.substr(0, 100) << "\n" << (i % 11 == 0 ? "; another one\n" : "") << row.substr(100) << (i % 2 ? "\r\n" : "\n") << (i % 13 == 0 ? "\n" : "");
                }
            }
            reference_db* d = reference_db::getDB(db);
            EQUAL(d->getSeqCount(), 1500u);
            bool same = true;
            for (uint32_t i = 0; i < 1500 && same; i++) {
                const cseq& c = d->getCseq(i);
                std::string want = rows[i];
                want.resize(d->getAlignmentWidth(), '-');   // rows are padded to the database's width
                same = c.getName() == "d" + std::to_string(i) && c.getAligned(true) == want &&
                       c.get_attr_string(fn_fullname) == (i % 5 ? "described" : "");
            }
            CHECK(same);
            { std::ofstream f(bad); f << ">ok\nACGU\n>broken\nAC\nGXU\n"; }
            bool threw = false;
            try { reference_db::getDB(bad); } catch (std::runtime_error& e) { threw = std::string(e.what()).find("line 5") != std::string::npos && std::string(e.what()).find("'X'") != std::string::npos; }
            CHECK(threw);
            { std::ofstream f(none); f << "no records here\n"; }
            THROWS(reference_db::getDB(none), std::runtime_error);
            THROWS(reference_db::getDB(tmpdir + "/host_unit_does_not_exist.fasta"), std::runtime_error);
            remove(db.c_str()); remove(bad.c_str()); remove(none.c_str());
        }
        // --add-relatives (src/rw_fasta.cpp:419-433) with --meta-fmt csv: the first N relatives not written before follow the
        // sequence's record, the search result taking precedence over the alignment family
        {
            const std::string outp = tmpdir + "/host_unit_rel.fasta", csvp = tmpdir + "/host_unit_rel.csv";
            cseq r1("ref1", "AC--GU"), r2("ref2", "A-C-GU"), r3("ref3", "ACG--U");
            r2.set_attr<std::string>(fn_fullname, "second, reference");
            rw_fasta::opts->copy_relatives = 2; rw_fasta::opts->fastameta = FASTA_META_CSV;
            {
                rw_fasta::writer wr(outp);
                CHECK(!wr.positional());
                tray a, b;
                a.input_sequence = new cseq("qa", "ACGU"); a.aligned_sequence = new cseq("qa", "AC-G-U");
                a.alignment_reference = new search::result_vector{{3.f, &r1}, {2.f, &r2}, {1.f, &r3}};
                b.input_sequence = new cseq("qb", "ACGU"); b.aligned_sequence = new cseq("qb", "A-CG-U");
                b.alignment_reference = new search::result_vector{{3.f, &r1}};
                b.search_result = new search::result_vector{{0.9f, &r2}, {0.8f, &r3}, {0.7f, &r1}};
                wr(a); wr(b);
                a.destroy(); b.destroy();
            }
            rw_fasta::opts->copy_relatives = 0; rw_fasta::opts->fastameta = FASTA_META_NONE;
            std::ifstream f1(outp); std::stringstream s1; s1 << f1.rdbuf();
            EQUAL(s1.str(), std::string(">qa\nAC-G-U\n>ref1\nAC--GU\n>ref2 second, reference\nA-C-GU\n>qb\nA-CG-U\n>ref3\nACG--U\n"));
            std::ifstream f2(csvp, std::ios::binary); std::stringstream s2; s2 << f2.rdbuf();
            EQUAL(s2.str(), std::string("name\r\nqa\r\nref1\r\nref2,\"second, reference\"\r\nqb\r\nref3\r\n"));
            remove(outp.c_str()); remove(csvp.c_str());
        }
        const std::string one = rw_fasta::writer::gzip_member("", 0);   // an empty member is a valid stream
        CHECK(one.size() >= 18 && (unsigned char)one[0] == 0x1f && (unsigned char)one[1] == 0x8b);
    }
}

// Launch geometry of the k-mer search kernels (csrc/find_layout.h) over every layout an index can have: sub-tiles of
// 32..32768 references, 1..24 warps per tile (api.cu clamps warps x sub-tile to 24 x 4096 counters).
static void test_find_layout() {
    using namespace sg;
    for (uint32_t sub = 32; sub <= SUB_MAX; sub <<= 1) {
        const uint32_t tw_max = std::max<uint32_t>(1, std::min<uint32_t>(TILE_WARPS_MAX, TILE_WARPS_MAX * SUB_DEFAULT / sub));
        for (uint32_t tw = 1; tw <= tw_max; tw++) {
            const FindLayout L = find_layout(tw, sub);
            const FindVariant& V = FIND_VARIANTS[L.variant];
            const uint32_t g2 = 2u * (uint32_t)V.g, ow = tw + 1, nt = 32 * tw;
            CHECK(tw <= V.max_warps && (L.variant == 0 || tw > FIND_VARIANTS[L.variant - 1].max_warps));
            CHECK(L.kc >= g2 && L.kc % g2 == 0 && L.kc <= (uint32_t)FIND_KC);   // requests read whole groups of G lists
            CHECK(L.kc * ow <= (uint32_t)FIND_PRE * nt);                          // every staged offset has a thread
            CHECK(L.ks % 4 == 0 && L.ks >= L.kc + g2);                            // LDS.128 rows, zero columns past the chunk
            CHECK(L.scratch_words >= ow * L.ks && L.scratch_words >= SEL_BINS + TIE_CAP);
            CHECK(L.smem == (size_t)tw * sub * 2 + (size_t)L.scratch_words * 4);
            CHECK(L.smem + FIND_SMEM_FIXED <= FIND_SMEM_CTA);                      // the launch fits one CTA's shared memory
            CHECK((size_t)tw * sub * 2 % 16 == 0);                                // the scratch starts 16-byte aligned
            // the production tiles keep their two CTAs per SM
            if (sub == SUB_DEFAULT && tw <= FIND_VARIANTS[0].max_warps) CHECK(2 * (L.smem + FIND_SMEM_FIXED) <= FIND_SMEM_SM);
            if (sub == SUB_DEFAULT && tw == 13) CHECK(2 * (L.smem + FIND_SMEM_FIXED) <= FIND_SMEM_SM);   // 50 k references: one tile
        }
    }
    // tiles of an index: one tile up to 14 sub-tiles, then balanced tiles of at most 12
    for (uint32_t n_sub = 1; n_sub <= 2000; n_sub++) {
        const uint32_t tw = find_auto_tile_warps(n_sub), n_tiles = (n_sub + tw - 1) / tw;
        CHECK(tw >= 1 && tw <= 14 && (n_sub <= 14 ? tw == n_sub : tw <= 12));
        CHECK((uint64_t)n_tiles * tw >= n_sub && (uint64_t)(n_tiles - 1) * tw < n_sub);
        if (n_sub > 14) CHECK(tw >= 8);                                           // no sliver tiles
    }
    EQUAL(find_auto_tile_warps(13), 13u);     // 50 000 references
    EQUAL(find_auto_tile_warps(123), 12u);    // 500 000 references: 11 tiles
    // top-k merge: one level while max x tiles fits the sort, two levels of groups beyond, failure past that
    for (uint32_t n_tiles : {1u, 2u, 6u, 11u, 16u, 21u, 41u, 256u, 257u, 1000u})
        for (uint64_t mx : {1ull, 41ull, 410ull, 1000ull, 4100ull, 8192ull, 8193ull, 16384ull, 16385ull}) {
            uint32_t gs = 0, ng = 0;
            const bool ok = find_merge_plan(mx, n_tiles, &gs, &ng);
            if (ok) {
                CHECK(gs >= 1 && (uint64_t)gs * mx <= FIND_MAX_SORT);             // a group's keys fit one sort
                CHECK((uint64_t)ng * gs >= n_tiles && (uint64_t)(ng - 1) * gs < n_tiles);
                CHECK(ng == 1 || (uint64_t)ng * mx <= FIND_MAX_SORT);             // and so do the groups' winners
            } else {
                CHECK(mx * n_tiles > FIND_MAX_SORT);                              // never refuses what one level holds
            }
        }
    uint32_t gs = 0, ng = 0;
    CHECK(find_merge_plan(41, 11, &gs, &ng) && ng == 1);            // family window at 500 000 references
    CHECK(find_merge_plan(1000, 41, &gs, &ng) && ng == 3);          // --search candidates at 2 M references: two levels
    CHECK(!find_merge_plan(4100, 41, &gs, &ng));                    // third family window there: full ranking instead
}

int main(int argc, char** argv) {
    test_find_layout();
    test_cseq();
    test_options();
    test_fasta(argc > 1 ? argv[1] : "/tmp");
    std::cout << (n_fail ? "FAILED " : "ok ") << n_checks << " checks, " << n_fail << " failures" << std::endl;
    return n_fail ? 1 : 0;
}
