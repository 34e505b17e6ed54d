// `sina` command line of the sina_b200 drop-in: the reference's option surface for the hot path
// (src/sina.cpp:224-264,379-440) and its pipeline reader -> famfinder -> aligner -> writer (src/sina.cpp:443-593),
// with the TBB flow graph replaced by batches: a reader thread cuts the input into batches, one worker thread per
// GPU runs famfinder + aligner on whole batches against its replica of the index, and the writer emits the
// batches in input order (the reference's sequencer_node, src/sina.cpp:529-538). The alignment_width-long output lines
// are rendered by a small pool of threads between the GPU workers and the writer (at 10^5
// sequences/s x 50 000 columns a single rendering + writing thread is the bottleneck).
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <thread>

#include "../../include/sina_b200.h"
#include "align.h"
#include "famfinder.h"
#include "search_filter.h"
#include "kmer_search.h"
#include "rw_fasta.h"

using namespace sina;

namespace {

struct cli_options {
    std::string in = "-", out = "-";
    unsigned int threads = 0, batch = 9472, gpus = 0, max_trays = 0;
    bool inorder = false, noalign = false, skip_align = false, show_log = false, do_search = false;
};
cli_options opts;

struct batch_t {
    uint64_t no = 0;
    std::vector<tray> trays;
    std::vector<std::string> raw;       // text of the input records, parsed into trays by the worker that takes the batch
    std::vector<unsigned int> raw_seqno, raw_lineno;
    std::vector<std::string> records;   // FASTA record of every tray, rendered off the writer thread
    std::vector<char> has_record;
    std::string members;                // ".gz" output: the batch's records as gzip members, compressed by the render pool
    unsigned int n_rec = 0, n_exc = 0;
    uint64_t file_offset = 0;           // where the batch's records start in the output file (positional writer)
};

// A run of consecutive trays of one batch whose records go to consecutive bytes of the output file: rendered and
// written by one thread of the output pool, record by record through a buffer that stays in that core's cache.
struct slice_t {
    std::shared_ptr<batch_t> batch;
    std::shared_ptr<std::atomic<int>> left;   // slices of the batch still to be written
    size_t begin = 0, end = 0;
    uint64_t file_offset = 0, bytes = 0;      // the slice's byte range of the output file
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
    main_od.flag("search,S", &opts.do_search, "enable search stage");
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
    adv.value<unsigned int>("batch-size", &opts.batch, 9472u, "[sina_b200] queries per device batch");
    adv.flag("show-log", &opts.show_log, "[sina_b200] print each query's log line to stderr");
    rw_fasta::get_options_description(main_od, adv);
    famfinder::get_options_description(main_od, adv);
    aligner::get_options_description(main_od, adv);
    search_filter::get_options_description(main_od, adv);
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
    if (opts.do_search) search_filter::validate_vm(vm, all);
    rw_fasta::validate_vm(vm, all);
    if (opts.batch == 0) throw std::logic_error("--batch-size must be > 0");

    unsigned int ngpu = 0;
    if (do_align || opts.do_search) {
        const int have = sg_device_count();
        if (have < 1) throw std::runtime_error("no CUDA device: sina_b200 has no CPU path");
        ngpu = opts.gpus == 0 ? (unsigned)have : std::min<unsigned>(opts.gpus, (unsigned)have);
    }

    rw_fasta::reader reader(opts.in);
    rw_fasta::writer writer(opts.out);
    const bool direct = writer.positional();   // a regular file: records are written at reserved offsets by a pool
    // ".gz": the render pool compresses whole runs of records (relatives are interleaved by the sink, one member each)
    const bool packed_members = writer.compressed() && rw_fasta::opts->copy_relatives == 0;

    // stage instances, one pair per GPU (each builds / shares the device's replica of the index)
    std::vector<std::unique_ptr<famfinder>> ff;
    std::vector<std::unique_ptr<aligner>> al;
    std::vector<std::unique_ptr<search_filter>> sf;
    for (unsigned int d = 0; d < ngpu; d++) {
        if (do_align) {
            ff.emplace_back(new famfinder((int)d));
            al.emplace_back(new aligner((int)d));
        }
        if (opts.do_search) sf.emplace_back(new search_filter((int)d));   // src/sina.cpp:519-527
    }
    std::cerr << "Aligner ready. Processing sequences" << std::endl;  // src/sina.cpp:581
    const auto before = std::chrono::steady_clock::now();

    const size_t inflight = opts.max_trays ? std::max<size_t>(1, opts.max_trays / opts.batch) : 3 * std::max(1u, ngpu) + 1;
    bounded_queue<batch_t> todo(inflight), torender(inflight);
    // Batches alive between the reader and the last byte written (the reference's limiter node, src/sina.cpp:485-489):
    // a rendered batch holds its FASTA records (4096 x 50 kB at 50 000 columns), so nothing but this bound keeps a slow
    // output device from growing the `done` map until memory runs out. The reader takes a slot per batch, the slot comes
    // back when the batch's trays are destroyed.
    const size_t alive_cap = 2 * inflight + 4;   // x 9472 trays of ~26 kB (input + aligned sequence): 0.25 GB per batch
    std::mutex alive_mu;
    std::condition_variable alive_cv;
    size_t alive = 0;
    auto alive_acquire = [&] { std::unique_lock<std::mutex> l(alive_mu); alive_cv.wait(l, [&] { return alive < alive_cap; }); alive++; };
    auto alive_release = [&] { { std::lock_guard<std::mutex> l(alive_mu); alive--; } alive_cv.notify_one(); };
    std::mutex done_mu;
    std::condition_variable done_cv;
    std::map<uint64_t, batch_t> done;
    std::atomic<bool> failed(false);
    std::string failure;
    uint64_t n_batches = 0;
    bool reading_done = false;

    // busy seconds of every pipeline role, printed with SINA_B200_TIMING=1 (where does a file-to-file run spend its time)
    std::atomic<uint64_t> us_read(0), us_parse(0), us_family(0), us_align(0), us_render(0), us_write(0);
    std::atomic<unsigned int> n_skipped(0);
    auto usec = [](std::chrono::steady_clock::time_point a) {
        return (uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - a).count();
    };
    std::thread rd([&] {
        try {
            batch_t b;
            for (;;) {
                std::string rec;
                unsigned int seqno = 0, lineno = 0;
                const auto t0 = std::chrono::steady_clock::now();
                const bool more = reader.next_record(rec, seqno, lineno);
                us_read += usec(t0);
                if (!more) break;
                b.raw.push_back(std::move(rec));
                b.raw_seqno.push_back(seqno);
                b.raw_lineno.push_back(lineno);
                if (b.raw.size() == opts.batch) {
                    b.no = n_batches++;
                    alive_acquire();
                    todo.push(std::move(b));
                    b = batch_t();
                }
                if (failed) break;
            }
            if (!b.raw.empty()) { b.no = n_batches++; alive_acquire(); todo.push(std::move(b)); }
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
                {   // the records of the batch become trays here, on the worker (rw_fasta::reader::parse_record)
                    const auto t0 = std::chrono::steady_clock::now();
                    b.trays.reserve(b.raw.size());
                    for (size_t i = 0; i < b.raw.size(); i++) {
                        tray t;
                        if (rw_fasta::reader::parse_record(b.raw[i], b.raw_seqno[i], b.raw_lineno[i], reader.filename(), t)) b.trays.push_back(t);
                        else n_skipped++;
                    }
                    b.raw.clear(); b.raw.shrink_to_fit();
                    us_parse += usec(t0);
                }
                if (do_align && !failed) {
                    auto t0 = std::chrono::steady_clock::now();
                    ff[d]->run(b.trays);
                    us_family += usec(t0);
                    t0 = std::chrono::steady_clock::now();
                    al[d]->run(b.trays);
                    us_align += usec(t0);
                }
                if (opts.do_search && !failed) {
                    if (!do_align)   // --prealigned: the input alignment is what gets searched (src/sina.cpp:505-509)
                        for (tray& t : b.trays) if (t.input_sequence && !t.aligned_sequence) t.aligned_sequence = new cseq(*t.input_sequence);
                    sf[d]->run(b.trays);
                }
            } catch (std::exception& e) {
                std::lock_guard<std::mutex> l(done_mu);
                if (!failed) failure = e.what();
                failed = true;
            }
            if (direct) {   // positional output: records are rendered by the output pool, straight into the write
                std::lock_guard<std::mutex> l(done_mu);
                done.emplace(b.no, std::move(b));
                done_cv.notify_all();
            } else {
                torender.push(std::move(b));
            }
        }
    };
    auto render = [&] {
        batch_t b;
        while (torender.pop(b)) {
            const auto t0r = std::chrono::steady_clock::now();
            try {
                b.records.resize(b.trays.size());
                b.has_record.assign(b.trays.size(), 0);
                if (!failed) {
                    for (size_t i = 0; i < b.trays.size(); i++) {
                        tray& t = b.trays[i];
                        if (!do_align && t.input_sequence && !t.aligned_sequence) t.aligned_sequence = new cseq(*t.input_sequence);  // --prealigned: pass through
                        if (t.input_sequence == nullptr) throw std::runtime_error("Received broken tray in rw_fasta writer");
                        if (t.aligned_sequence && rw_fasta::writer::passes_min_idty(*t.aligned_sequence)) { b.records[i] = rw_fasta::writer::format(*t.aligned_sequence); b.has_record[i] = 1; }
                    }
                }
                if (!failed && packed_members) {   // one gzip member per 256 records, the records freed as they are packed
                    std::string plain;
                    for (size_t i0 = 0; i0 < b.trays.size(); i0 += 256) {
                        plain.clear();
                        for (size_t i = i0; i < std::min(b.trays.size(), i0 + 256); i++) {
                            if (b.has_record[i]) { plain += b.records[i]; b.n_rec++; } else b.n_exc++;
                            std::string().swap(b.records[i]);
                        }
                        if (!plain.empty()) b.members += rw_fasta::writer::gzip_member(plain.data(), plain.size());
                    }
                }
            } catch (std::exception& e) {
                std::lock_guard<std::mutex> l(done_mu);
                if (!failed) failure = e.what();
                failed = true;
            }
            us_render += usec(t0r);
            std::lock_guard<std::mutex> l(done_mu);
            done.emplace(b.no, std::move(b));
            done_cv.notify_all();
        }
    };
    std::atomic<uint64_t> us_pwrite(0);
    bounded_queue<slice_t> towrite(16 * inflight);
    auto write_pool = [&] {
        slice_t sl;
        std::string rec;
        while (towrite.pop(sl)) {
            const auto t0 = std::chrono::steady_clock::now();
            try {
                uint64_t at = sl.file_offset;
                char* map = writer.mapped() && sl.bytes > 0 && !failed ? writer.map_range(sl.file_offset, sl.bytes) : nullptr;
                for (size_t i = sl.begin; i < sl.end; i++) {
                    tray& t = sl.batch->trays[i];
                    if (sl.batch->has_record[i] && !failed) {
                        rw_fasta::writer::format_into(*t.aligned_sequence, rec);
                        if (map) memcpy(map + (at - sl.file_offset), rec.data(), rec.size());
                        else writer.write_at(at, rec.data(), rec.size());
                        at += rec.size();
                    }
                    t.destroy();  // src/sina.cpp:573-579
                }
                if (map) writer.unmap_range(map, sl.file_offset, sl.bytes);
            } catch (std::exception& e) {
                std::lock_guard<std::mutex> l(done_mu);
                if (!failed) failure = e.what();
                failed = true;
            }
            us_pwrite += usec(t0);
            if (sl.left->fetch_sub(1) == 1) alive_release();
        }
    };
    std::vector<std::thread> workers, renderers, writers;
    const unsigned int hw = std::max(1u, std::thread::hardware_concurrency());
    // render + pwrite, one record at a time. A tmpfs / page-cache file takes about 2.5 GB/s (50 k records of 50 kB per
    // second) however many threads write to it (B200 box: 4 threads 2.2 GB/s, 8 threads 2.5 GB/s, 12 threads and twice the
    // workers 1.6 GB/s), so the pool stays small and the rest of the cores go to the per-GPU workers
    unsigned int n_write = direct ? std::max(4u, std::min(8u, hw / 3)) : 0u;
    if (direct && getenv("SINA_B200_WRITERS")) n_write = std::max(1, atoi(getenv("SINA_B200_WRITERS")));
    for (unsigned int r = 0; r < n_write; r++) writers.emplace_back(write_pool);
    // several host threads per GPU: the library serialises the device calls of one index, so while one thread's batch
    // is on the GPU the others pack queries / build the result sequences of theirs (B200 box, 1 GPU, 160k queries:
    // 3 threads 44.8k seq/s, 6 threads 49.1k); never more threads than the cores left beside the output pool
    unsigned int wpg = 1u;
    {
        const unsigned int spare = hw > n_write + 2 ? hw - n_write - 2 : 2;
        if (do_align || opts.do_search) wpg = std::max(2u, std::min(6u, spare / std::max(1u, ngpu)));
        else wpg = std::max(1u, std::min(4u, spare));   // --prealigned without --search: the workers only parse
    }
    if (getenv("SINA_B200_WORKERS")) wpg = std::max(1, atoi(getenv("SINA_B200_WORKERS")));
    for (unsigned int d = 0; d < std::max(1u, ngpu); d++)
        for (unsigned int k = 0; k < wpg; k++) workers.emplace_back(work, d);
    const unsigned int n_render = direct ? 0u : std::max(2u, std::min(16u, hw / 2));
    for (unsigned int r = 0; r < n_render; r++) renderers.emplace_back(render);

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
        const auto t0w = std::chrono::steady_clock::now();
        if (direct) {
            // byte ranges are handed out here, in input order (the size of a record follows from its header and the
            // alignment width); the output pool renders and writes them, slice by slice, and frees the trays
            auto sb = std::make_shared<batch_t>(std::move(b));
            batch_t& B = *sb;
            const size_t n = B.trays.size(), per = 256;
            B.has_record.assign(n, 0);
            std::vector<uint64_t> size(n, 0);
            uint64_t total = 0;
            unsigned int nrec = 0, nexc = 0;
            for (size_t i = 0; i < n; i++) {
                tray& t = B.trays[i];
                if (!do_align && t.input_sequence && !t.aligned_sequence) t.aligned_sequence = new cseq(*t.input_sequence);  // --prealigned: pass through
                if (t.input_sequence == nullptr) { if (!failed) failure = "Received broken tray in rw_fasta writer"; failed = true; }
                if (t.aligned_sequence && !failed && rw_fasta::writer::passes_min_idty(*t.aligned_sequence)) { size[i] = rw_fasta::writer::record_size(*t.aligned_sequence); B.has_record[i] = 1; total += size[i]; nrec++; } else nexc++;
                if (opts.show_log && t.input_sequence) std::cerr << "sequence_number: " << t.seqno << " sequence_identifier: "
                                                                 << t.input_sequence->getName() << " " << t.log.str() << std::endl;
            }
            count += n;
            uint64_t at = writer.reserve(total, nrec, nexc);
            auto left = std::make_shared<std::atomic<int>>((int)((n + per - 1) / per));
            for (size_t i0 = 0; i0 < n; i0 += per) {
                slice_t sl;
                sl.batch = sb; sl.left = left; sl.begin = i0; sl.end = std::min(n, i0 + per); sl.file_offset = at;
                for (size_t i = sl.begin; i < sl.end; i++) at += size[i];
                sl.bytes = at - sl.file_offset;
                towrite.push(std::move(sl));
            }
            if (n == 0) alive_release();
            us_write += usec(t0w);
            continue;
        }
        if (packed_members && !failed) writer.write_members(b.members, b.n_rec, b.n_exc);
        for (size_t i = 0; i < b.trays.size(); i++) {
            tray& t = b.trays[i];
            if (!failed) {
                if (!packed_members) writer.write_formatted(b.has_record[i] ? &b.records[i] : nullptr);
                if (b.has_record[i]) writer.write_csv(*t.aligned_sequence);   // --meta-fmt csv
                if (b.has_record[i]) writer.write_relatives(t);                // --add-relatives
                if (opts.show_log) std::cerr << "sequence_number: " << t.seqno << " sequence_identifier: "
                                             << t.input_sequence->getName() << " " << t.log.str() << std::endl;
            }
            count++;
            t.destroy();  // src/sina.cpp:573-579
        }
        us_write += usec(t0w);
        alive_release();
    }
    towrite.close();
    for (auto& w : writers) w.join();
    rd.join();
    for (auto& w : workers) w.join();
    torender.close();
    for (auto& r : renderers) r.join();
    if (failed) throw std::runtime_error(failure);
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - before).count();
    char buf[256];
    snprintf(buf, sizeof(buf), "Took %.3fs to align %llu sequences (%.1f sequences/s)", secs, (unsigned long long)count,
             secs > 0 ? count / secs : 0.0);  // src/sina.cpp:588-589
    std::cerr << buf << std::endl;
    if (getenv("SINA_B200_TIMING")) {
        snprintf(buf, sizeof(buf), "busy seconds: read %.3f | parse %.3f + family %.3f + align %.3f over %u worker threads | render %.3f over %u threads | sink %.3f | render+pwrite %.3f over %u threads",
                 us_read / 1e6, us_parse / 1e6, us_family / 1e6, us_align / 1e6, wpg * std::max(1u, ngpu), us_render / 1e6, n_render, us_write / 1e6, us_pwrite / 1e6, n_write);
        std::cerr << buf << std::endl;
    }
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
