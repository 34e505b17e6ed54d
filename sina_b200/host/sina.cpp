// `sina` command line of the sina_b200 drop-in: the reference's option surface for the hot path
// (src/sina.cpp:224-264,379-440) and its pipeline reader -> famfinder -> aligner -> writer (src/sina.cpp:443-593),
// with the TBB flow graph replaced by batches: a reader thread cuts the input into batches, one worker thread per
// GPU runs famfinder + aligner on whole batches against its replica of the index, and the writer emits the
// batches in input order (the reference's sequencer_node, src/sina.cpp:529-538).
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <iostream>
#include <map>
#include <mutex>
#include <thread>

#include "../../include/sina_b200.h"
#include "align.h"
#include "famfinder.h"
#include "kmer_search.h"
#include "rw_fasta.h"

using namespace sina;

namespace {

struct cli_options {
    std::string in = "-", out = "-";
    unsigned int threads = 0, batch = 4096, gpus = 0, max_trays = 0;
    bool inorder = false, noalign = false, skip_align = false, show_log = false;
};
cli_options opts;

struct batch_t {
    uint64_t no = 0;
    std::vector<tray> trays;
};

template <typename T>
class bounded_queue {  // the reference bounds in-flight trays with a limiter node (src/sina.cpp:485-489)
public:
    explicit bounded_queue(size_t cap) : cap_(cap) {}
    void push(T&& v) {
        std::unique_lock<std::mutex> l(mu_);
        not_full_.wait(l, [&] { return q_.size() < cap_; });
        q_.emplace_back(std::move(v));
        not_empty_.notify_one();
    }
    bool pop(T& v) {
        std::unique_lock<std::mutex> l(mu_);
        not_empty_.wait(l, [&] { return !q_.empty() || closed_; });
        if (q_.empty()) return false;
        v = std::move(q_.front());
        q_.erase(q_.begin());
        not_full_.notify_one();
        return true;
    }
    void close() {
        std::lock_guard<std::mutex> l(mu_);
        closed_ = true;
        not_empty_.notify_all();
    }
private:
    std::mutex mu_;
    std::condition_variable not_full_, not_empty_;
    std::vector<T> q_;
    size_t cap_;
    bool closed_ = false;
};

int real_main(int argc, const char* const* argv) {
    po::options_description main_od("Options"), adv("Advanced Options");
    bool help = false, help_all = false, version = false;
    main_od.flag("help,h", &help, "show short help");
    main_od.flag("help-all,H", &help_all, "show full help (long)");
    main_od.value<std::string>("in,i", &opts.in, "-", "input file (fasta)");
    main_od.value<std::string>("out,o", &opts.out, "-", "output file (fasta)");
    main_od.unsupported("add-relatives", true, "writing relatives next to the query");
    main_od.unsupported("search,S", false, "the search/classification stage is outside the accelerated path");
    main_od.flag("prealigned,P", &opts.skip_align, "skip alignment stage");
    main_od.value<unsigned int>("threads,p", &opts.threads, 0u, "accepted for compatibility (the GPU path batches instead)");
    main_od.unsupported("num-pts", true, "PT servers");
    main_od.flag("version,V", &version, "show version");
    adv.flag("preserve-order", &opts.inorder, "maintain order of sequences (always on)");
    adv.value<unsigned int>("max-in-flight", &opts.max_trays, 0u, "max number of sequences processed at a time (0: 4 batches per GPU)");
    adv.flag("no-align", &opts.noalign, "disable alignment stage (same as prealigned)");
    adv.unsupported("intype", true, "only FASTA input");
    adv.unsupported("outtype", true, "only FASTA output");
    adv.unsupported("fields,f", true, "field selection");
    adv.value<unsigned int>("gpus", &opts.gpus, 0u, "[sina_b200] number of GPUs to shard the queries over (0: all visible)");
    adv.value<unsigned int>("batch-size", &opts.batch, 4096u, "[sina_b200] queries per device batch");
    adv.flag("show-log", &opts.show_log, "[sina_b200] print each query's log line to stderr");
    rw_fasta::get_options_description(main_od, adv);
    famfinder::get_options_description(main_od, adv);
    aligner::get_options_description(main_od, adv);
    po::options_description all("");
    all.add(main_od).add(adv);
    po::variables_map vm;
    po::store(argc, argv, all, vm);
    if (help || help_all) {
        std::cerr << "Usage:\n sina -i input -o output --db reference.fasta [--fs-engine internal] [options]\n\n" << main_od.usage();
        if (help_all) std::cerr << "\n" << adv.usage();
        return 0;
    }
    if (version) { std::cout << "SINA (sina_b200, B200-native hot path) 1.7.3-compatible" << std::endl; return 0; }
    const bool do_align = !(opts.skip_align || opts.noalign);
    if (do_align) {
        famfinder::validate_vm(vm, all);
        aligner::validate_vm(vm, all);
    }
    rw_fasta::validate_vm(vm, all);
    if (opts.batch == 0) throw std::logic_error("--batch-size must be > 0");

    unsigned int ngpu = 0;
    if (do_align) {
        const int have = sg_device_count();
        if (have < 1) throw std::runtime_error("no CUDA device: sina_b200 has no CPU path");
        ngpu = opts.gpus == 0 ? (unsigned)have : std::min<unsigned>(opts.gpus, (unsigned)have);
    }

    rw_fasta::reader reader(opts.in);
    rw_fasta::writer writer(opts.out);

    // stage instances, one pair per GPU (each builds / shares the device's replica of the index)
    std::vector<std::unique_ptr<famfinder>> ff;
    std::vector<std::unique_ptr<aligner>> al;
    for (unsigned int d = 0; d < ngpu; d++) {
        ff.emplace_back(new famfinder((int)d));
        al.emplace_back(new aligner((int)d));
    }
    std::cerr << "Aligner ready. Processing sequences" << std::endl;  // src/sina.cpp:581
    const auto before = std::chrono::steady_clock::now();

    const size_t inflight = opts.max_trays ? std::max<size_t>(1, opts.max_trays / opts.batch) : 4 * std::max(1u, ngpu);
    bounded_queue<batch_t> todo(inflight);
    std::mutex done_mu;
    std::condition_variable done_cv;
    std::map<uint64_t, batch_t> done;
    std::atomic<bool> failed(false);
    std::string failure;
    uint64_t n_batches = 0;
    bool reading_done = false;

    std::thread rd([&] {
        try {
            batch_t b;
            for (;;) {
                tray t;
                if (!reader(t)) break;
                b.trays.push_back(t);
                if (b.trays.size() == opts.batch) {
                    b.no = n_batches++;
                    todo.push(std::move(b));
                    b = batch_t();
                }
                if (failed) break;
            }
            if (!b.trays.empty()) { b.no = n_batches++; todo.push(std::move(b)); }
        } catch (std::exception& e) {
            std::lock_guard<std::mutex> l(done_mu);
            failure = e.what();
            failed = true;
        }
        todo.close();
        std::lock_guard<std::mutex> l(done_mu);
        reading_done = true;
        done_cv.notify_all();
    });

    auto work = [&](unsigned int d) {
        batch_t b;
        while (todo.pop(b)) {
            try {
                if (do_align && !failed) {
                    ff[d]->run(b.trays);
                    al[d]->run(b.trays);
                }
            } catch (std::exception& e) {
                std::lock_guard<std::mutex> l(done_mu);
                if (!failed) failure = e.what();
                failed = true;
            }
            std::lock_guard<std::mutex> l(done_mu);
            done.emplace(b.no, std::move(b));
            done_cv.notify_all();
        }
    };
    std::vector<std::thread> workers;
    for (unsigned int d = 0; d < std::max(1u, ngpu); d++) workers.emplace_back(work, d);

    uint64_t count = 0, next = 0;
    for (;;) {  // sink: batches in input order
        batch_t b;
        {
            std::unique_lock<std::mutex> l(done_mu);
            done_cv.wait(l, [&] { return done.count(next) || (reading_done && next >= n_batches); });
            if (!done.count(next)) break;
            b = std::move(done[next]);
            done.erase(next);
        }
        next++;
        for (auto& t : b.trays) {
            if (!failed) {
                if (!do_align && t.input_sequence) t.aligned_sequence = new cseq(*t.input_sequence);  // --prealigned: pass through
                writer(t);
                if (opts.show_log) std::cerr << "sequence_number: " << t.seqno << " sequence_identifier: "
                                             << t.input_sequence->getName() << " " << t.log.str() << std::endl;
            }
            count++;
            t.destroy();  // src/sina.cpp:573-579
        }
    }
    rd.join();
    for (auto& w : workers) w.join();
    if (failed) throw std::runtime_error(failure);
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - before).count();
    char buf[256];
    snprintf(buf, sizeof(buf), "Took %.3fs to align %llu sequences (%.1f sequences/s)", secs, (unsigned long long)count,
             secs > 0 ? count / secs : 0.0);  // src/sina.cpp:588-589
    std::cerr << buf << std::endl;
    if (writer.excluded()) std::cerr << writer.excluded() << " sequences were not aligned and not written" << std::endl;
    std::cerr << "SINA finished." << std::endl;
    return 0;
}

}  // namespace

int main(int argc, const char** argv) {
    try {
        return real_main(argc, argv);
    } catch (std::logic_error& e) {  // configuration errors (src/sina.cpp:429-438)
        std::cerr << "Configuration error:" << std::endl << e.what() << std::endl << "Use \"--help\" to show options" << std::endl;
        return 1;
    } catch (std::exception& e) {    // src/sina.cpp:595-607
        std::cerr << "Error during program execution: " << e.what() << std::endl;
        return 1;
    }
}
