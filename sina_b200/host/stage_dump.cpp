// Test driver for the per-query stage functors: runs famfinder::operator()(tray) and aligner::operator()(tray)
// one tray at a time (the way the reference's flow graph calls them, src/sina.cpp:511,516) plus a
// kmer_search::find per query, and prints everything as text for tests/test_host_cli.py.
//   stage_dump <ref.fasta> <queries.fasta> [famfinder/aligner options...]
#include <iostream>

#include "align.h"
#include "famfinder.h"
#include "kmer_search.h"
#include "rw_fasta.h"

using namespace sina;

int main(int argc, const char** argv) {
    if (argc < 3) { std::cerr << "usage: stage_dump ref.fasta queries.fasta [options]" << std::endl; return 2; }
    try {
        po::options_description main_od, adv, all;
        rw_fasta::get_options_description(main_od, adv);
        famfinder::get_options_description(main_od, adv);
        aligner::get_options_description(main_od, adv);
        all.add(main_od).add(adv);
        std::vector<const char*> args{"stage_dump", "--db", argv[1]};
        for (int i = 3; i < argc; i++) args.push_back(argv[i]);
        po::variables_map vm;
        po::store((int)args.size(), args.data(), all, vm);
        famfinder::validate_vm(vm, all);
        famfinder ff;
        aligner al;
        famfinder ff2(ff);  // TBB copies node bodies: copies must work too
        aligner al2(al);
        kmer_search* ks = kmer_search::get_kmer_search(famfinder::opts.database, famfinder::opts.fs_kmer_len, famfinder::opts.fs_no_fast);
        std::cout << "size " << ks->size() << std::endl;
        rw_fasta::reader rd(argv[2]);
        tray t;
        while (rd(t)) {
            search::result_vector res;
            ks->find(*t.input_sequence, res, 5);
            std::cout << "query " << t.input_sequence->getName() << "\nfind";
            for (auto& r : res) std::cout << " " << r.sequence->getName() << ":" << r.score;
            std::cout << std::endl;
            t = ff2(t);
            std::cout << "family";
            if (t.alignment_reference) for (auto& r : *t.alignment_reference) std::cout << " " << r.sequence->getName() << ":" << r.score;
            else std::cout << " none";
            std::cout << std::endl;
            if (t.alignment_reference) t = al2(t);
            if (t.aligned_sequence)
                std::cout << "aligned " << t.aligned_sequence->getAligned(true) << "\nattrs " << t.aligned_sequence->get_attr_string(fn_qual)
                          << " " << t.aligned_sequence->get_attr_string(fn_head) << " " << t.aligned_sequence->get_attr_string(fn_tail) << std::endl;
            else std::cout << "aligned none" << std::endl;
            std::cout << "log " << t.log.str() << std::endl;
            t.destroy();
            t = tray();
        }
        delete ks;
    } catch (std::exception& e) {
        std::cout << "exception " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
