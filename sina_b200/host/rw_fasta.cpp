#include "rw_fasta.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>
#include <cerrno>
#include <cstdlib>
#include <cstring>

#include <algorithm>
#include <fstream>
#include <iostream>
#include <unordered_set>
#include <vector>

namespace sina {

rw_fasta::options* rw_fasta::opts = nullptr;

// file names ending in ".gz" are gzip streams, on both sides (src/rw_fasta.cpp:200-202,358-360)
static bool is_gz(const std::string& name) { return name.size() > 3 && name.compare(name.size() - 3, 3, ".gz") == 0; }

void rw_fasta::get_options_description(po::options_description& main, po::options_description& adv) {
    if (!opts) opts = new options();
    main.custom("meta-fmt", "none", "meta data in (*none*|header|comment)", [](const std::string& v) {
        if (v == "none") opts->fastameta = FASTA_META_NONE;
        else if (v == "header") opts->fastameta = FASTA_META_HEADER;
        else if (v == "comment") opts->fastameta = FASTA_META_COMMENT;
        else if (v == "csv") opts->fastameta = FASTA_META_CSV;
        else throw std::logic_error("Illegal value for meta-fmt");
    });
    main.value<unsigned int>("add-relatives", &opts->copy_relatives, 0u, "add the ARG nearest relatives for each sequence to output");
    po::options_description od("FASTA I/O");
    od.value<int>("line-length", &opts->line_length, 0, "wrap output sequence (unlimited)");
    od.value<float>("min-idty", &opts->min_idty, 0.f, "only write sequences with align_idty_slv > X, implies calc-idty");
    od.flag("fasta-write-dna", &opts->out_dna, "Write DNA sequences (default: RNA)");
    od.flag("fasta-write-dots", &opts->out_dots, "Use dots instead of dashes to distinguish unknown sequence data from indels");
    od.value<long>("fasta-idx", &opts->fasta_idx, 0L, "process only sequences beginning in block <arg>");
    od.value<long>("fasta-block", &opts->fasta_block, 0L, "length of blocks");
    adv.add(od);
}
void rw_fasta::validate_vm(po::variables_map&, po::options_description&) {}

// ------------------------------------------------------------------------------------------------ reader
// The input is read in blocks and cut into records (from a '>' at the start of a line to the next); turning a record
// into a cseq is a separate, thread-safe step (parse_record), so that the command line can leave it to its worker
// threads: one thread parsing 1.5 kB records reaches about 130 k sequences/s, eight GPUs take five times that.
struct rw_fasta::reader::priv_data {
    int fd = -1;                 // file or stdin (0): read() into buf, no stream layer in between
    bool own_fd = false;
    gzFile gz = nullptr;         // ".gz" input
    std::string filename;
    std::vector<char> buf;       // [pos, size) = bytes not yet handed out
    size_t pos = 0, size = 0;
    uint64_t base = 0;           // file offset of buf[0]
    bool eof = false;
    long block = 0, block_idx = 0;   // --fasta-block / --fasta-idx as they apply to this file
    unsigned int seqno = 0, lineno = 1, skipped = 0;
    ~priv_data() { if (gz) gzclose(gz); if (own_fd && fd >= 0) ::close(fd); }
    bool refill() {   // drop the consumed part, read another block; false when nothing was added
        if (eof) return false;
        const size_t block = 4u << 20;
        if (pos > 0) { memmove(buf.data(), buf.data() + pos, size - pos); base += pos; size -= pos; pos = 0; }
        if (buf.size() < size + block) buf.resize(std::max(buf.size() * 2, size + block));   // a record longer than the buffer
        size_t got = 0;
        while (got < block) {   // a pipe hands out less than asked for: keep the blocks large
            long n;
            if (gz) {
                n = gzread(gz, buf.data() + size + got, (unsigned)(block - got));
                if (n < 0) throw std::runtime_error("Error reading compressed file " + filename);
            } else {
                n = ::read(fd, buf.data() + size + got, block - got);
                if (n < 0) { if (errno == EINTR) continue; throw std::runtime_error("Error reading file " + filename + ": " + strerror(errno)); }
            }
            if (n == 0) { eof = true; break; }
            got += (size_t)n;
        }
        size += got;
        return got > 0;
    }
};

rw_fasta::reader::reader(const std::string& infile, bool whole_file) : data(new priv_data) {
    if (!opts) opts = new options();
    data->filename = infile;
    if (!whole_file) { data->block = opts->fasta_block; data->block_idx = opts->fasta_idx; }
    if (infile == "-") data->fd = 0;
    else if (is_gz(infile)) {
        data->gz = gzopen(infile.c_str(), "rb");
        if (!data->gz) throw std::runtime_error("Unable to open file " + infile + " for reading.");
        gzbuffer(data->gz, 1u << 20);
        // (the reference seeks its filter chain here, which a gzip stream cannot do)
        if (data->block > 0) throw std::logic_error("Cannot use --fasta-idx on compressed input");
    } else {
        data->fd = ::open(infile.c_str(), O_RDONLY);
        if (data->fd < 0) throw std::runtime_error("Unable to open file " + infile + " for reading.");
        data->own_fd = true;
    }
    // --fasta-block / --fasta-idx (src/rw_fasta.cpp:209-216,237-242): start at byte block * idx, at the next title line
    if (data->block > 0 && !data->gz) {
        if (infile == "-") throw std::logic_error("Cannot use --fasta-idx when input is piped");
        if (::lseek(data->fd, (off_t)(data->block * data->block_idx), SEEK_SET) < 0)
            throw std::runtime_error("Unable to seek in file " + infile);
        data->base = (uint64_t)(data->block * data->block_idx);
    }
}
rw_fasta::reader::~reader() = default;
unsigned int rw_fasta::reader::skipped() const { return data->skipped; }
const std::string& rw_fasta::reader::filename() const { return data->filename; }
void rw_fasta::reader::count_skipped() { data->skipped++; }

bool rw_fasta::reader::next_record(std::string& record, unsigned int& seqno, unsigned int& lineno, bool append) {
    priv_data& d = *data;
    // block-wise input: stop once the previous sequence ended past the block (the reference tests tellg() here)
    if (d.block > 0 && d.base + d.pos > (uint64_t)(d.block * (d.block_idx + 1))) return false;
    // skip to the next title line
    for (;;) {
        if (d.pos >= d.size && !d.refill()) return false;
        if (d.buf[d.pos] == '>') break;
        const void* nl = memchr(d.buf.data() + d.pos, '\n', d.size - d.pos);
        if (nl) { d.pos = (size_t)((const char*)nl - d.buf.data()) + 1; d.lineno++; }
        else d.pos = d.size;
    }
    // the record ends behind the newline that a '>' follows, or with the input: one memchr per line, which also
    // counts the lines (the scan position is kept relative to the record's start: a refill moves that to 0)
    size_t scan = 1, end = 0;
    unsigned int lines = 0;
    for (;;) {
        const char* rec = d.buf.data() + d.pos;
        const size_t avail = d.size - d.pos;
        const void* nl = scan < avail ? memchr(rec + scan, '\n', avail - scan) : nullptr;
        if (nl) {
            const size_t at = (size_t)((const char*)nl - rec);
            if (at + 1 < avail) {
                lines++;
                scan = at + 1;
                if (rec[at + 1] == '>') { end = at + 1; break; }
                continue;
            }
            if (d.eof) { lines++; end = avail; break; }
            scan = at;                 // the byte behind this newline is not here yet: look at it again after the refill
        } else {
            if (d.eof) { end = avail; break; }
            scan = avail;
        }
        d.refill();
    }
    if (append) record.append(d.buf.data() + d.pos, end);
    else record.assign(d.buf.data() + d.pos, end);
    d.pos += end;
    seqno = ++d.seqno;
    lineno = d.lineno;
    d.lineno += lines;
    return true;
}

bool rw_fasta::reader::parse_record(const std::string& record, unsigned int seqno, unsigned int lineno, const std::string& filename, tray& t) {
    auto next_line = [&](size_t& at, std::string& line) {
        if (at >= record.size()) return false;
        size_t e = record.find('\n', at);
        if (e == std::string::npos) e = record.size();
        line.assign(record, at, e - at);
        if (!line.empty() && line.back() == '\r') line.pop_back();
        at = e + 1;
        return true;
    };
    size_t at = 0;
    std::string line;
    next_line(at, line);
    t.seqno = seqno;
    t.input_sequence = new cseq();
    cseq& c = *t.input_sequence;
    size_t blank = line.find_first_of(" \t");
    if (blank == 0 || blank == std::string::npos) blank = line.size();
    c.setName(line.substr(1, blank - 1));
    if (blank < line.size()) c.set_attr<std::string>(fn_fullname, line.substr(blank + 1));
    while (at < record.size() && record[at] == ';' && next_line(at, line)) {  // comment lines may carry key=value attributes
        lineno++;
        const size_t eq = line.find('=');
        if (eq != std::string::npos) {
            auto trim = [](std::string s) {
                const size_t a = s.find_first_not_of(" \t\r"), b = s.find_last_not_of(" \t\r");
                return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
            };
            c.set_attr<std::string>(trim(line.substr(1, eq - 1)), trim(line.substr(eq + 1)));
        }
    }
    try {
        if (at < record.size()) c.append(record.data() + at, record.size() - at);
        return true;
    } catch (base_iupac::bad_character_exception& e) {  // src/rw_fasta.cpp:294-304: skip the sequence, keep going
        size_t bad = record.find((char)e.character, at);
        for (size_t i = 0; i < std::min(bad, record.size()); i++) lineno += record[i] == '\n';
        std::cerr << "Skipping sequence " << seqno << " (>" << c.getName() << ") at " << filename << ":" << lineno
                  << " (contains character '" << (char)e.character << "')" << std::endl;
        delete t.input_sequence;
        t.input_sequence = nullptr;
        return false;
    }
}

bool rw_fasta::reader::operator()(tray& t) {
    std::string record;
    unsigned int seqno = 0, lineno = 0;
    while (next_record(record, seqno, lineno)) {
        if (parse_record(record, seqno, lineno, data->filename, t)) return true;
        data->skipped++;
    }
    return false;
}

// ------------------------------------------------------------------------------------------------ writer
struct rw_fasta::writer::priv_data {
    int fd = -1;                 // regular file: positional writes
    uint64_t offset = 0;         // next free byte of the file
    std::ostream* out = nullptr; // stdout
    bool gz = false;             // ".gz" output: every put() becomes a gzip member (members concatenate to one stream)
    bool map = false;            // SINA_B200_MMAP_OUT=1 on a regular, uncompressed file
    std::ofstream out_csv;       // --meta-fmt csv
    std::unordered_set<std::string> relatives_written;   // --add-relatives
    bool csv_started = false;
    unsigned int count = 0, excluded = 0;
    void write(const cseq& c);
    void put(const char* p, size_t n);
    void put_raw(const char* p, size_t n);
    ~priv_data();
};

static void pwrite_all(int fd, const char* p, size_t n, uint64_t off) {
    while (n > 0) {
        const ssize_t w = ::pwrite(fd, p, n, (off_t)off);
        if (w < 0) {
            if (errno == EINTR) continue;
            throw std::runtime_error(std::string("write failed: ") + strerror(errno));
        }
        p += w; n -= (size_t)w; off += (uint64_t)w;
    }
}

rw_fasta::writer::priv_data::~priv_data() {
    if (fd < 0) return;
    if (gz && offset == 0) {   // nothing was written: still a valid (empty) gzip stream
        try { const std::string m = rw_fasta::writer::gzip_member("", 0); pwrite_all(fd, m.data(), m.size(), 0); } catch (...) {}
    }
    ::close(fd);
}
void rw_fasta::writer::priv_data::put_raw(const char* p, size_t n) {
    if (fd >= 0) { pwrite_all(fd, p, n, offset); offset += n; }
    else out->write(p, (std::streamsize)n);
}
void rw_fasta::writer::priv_data::put(const char* p, size_t n) {
    if (!gz) { put_raw(p, n); return; }
    const std::string m = rw_fasta::writer::gzip_member(p, n);
    put_raw(m.data(), m.size());
}

// One complete gzip member holding p[0..n). A gzip file is any number of members back to back, so the command line
// compresses the records of a batch on its render threads and the sink only appends the members in input order (the
// reference pushes a single-threaded gzip_compressor in front of its file). Level: zlib's fastest unless
// SINA_B200_GZIP_LEVEL says otherwise -- 50 kB alignment rows compress 12:1 at 180 MB/s per thread at level 1, 18:1 at
// 37 MB/s at the default level 6.
std::string rw_fasta::writer::gzip_member(const char* p, size_t n) {
    static const int level = [] { const char* e = getenv("SINA_B200_GZIP_LEVEL"); const int l = e ? atoi(e) : 1; return l < 0 || l > 9 ? 1 : l; }();
    z_stream z;
    memset(&z, 0, sizeof(z));
    if (deflateInit2(&z, level, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) throw std::runtime_error("deflateInit2 failed");
    std::string out;
    out.resize(deflateBound(&z, (uLong)n) + 32);
    size_t done = 0, produced = 0;
    int rc = Z_OK;
    do {   // avail_in is 32 bit: feed the input in pieces below 1 GiB
        const size_t piece = std::min<size_t>(n - done, (size_t)1 << 30);
        z.next_in = reinterpret_cast<Bytef*>(const_cast<char*>(p + done));
        z.avail_in = (uInt)piece;
        done += piece;
        do {
            if (produced == out.size()) out.resize(out.size() * 2);
            z.next_out = reinterpret_cast<Bytef*>(&out[produced]);
            z.avail_out = (uInt)std::min<size_t>(out.size() - produced, (size_t)1 << 30);
            const size_t before = z.avail_out;
            rc = deflate(&z, done == n ? Z_FINISH : Z_NO_FLUSH);
            produced += before - z.avail_out;
            if (rc == Z_STREAM_ERROR) { deflateEnd(&z); throw std::runtime_error("deflate failed"); }
        } while (z.avail_out == 0 || (done == n && rc != Z_STREAM_END));
    } while (done < n);
    deflateEnd(&z);
    out.resize(produced);
    return out;
}

rw_fasta::writer::writer(const std::string& outfile) : data(new priv_data) {
    if (!opts) opts = new options();
    if (outfile == "-") data->out = &std::cout;
    else {
        data->fd = ::open(outfile.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
        if (data->fd < 0) throw std::runtime_error("Unable to open file " + outfile + " for writing.");
        data->gz = is_gz(outfile);
        struct stat st;
        const char* e = getenv("SINA_B200_MMAP_OUT");
        data->map = e && atoi(e) > 0 && !data->gz && fstat(data->fd, &st) == 0 && S_ISREG(st.st_mode);
        if (data->map) {   // the mappings need read access to the file
            ::close(data->fd);
            data->fd = ::open(outfile.c_str(), O_RDWR | O_CREAT | O_TRUNC, 0644);
            if (data->fd < 0) throw std::runtime_error("Unable to open file " + outfile + " for writing.");
        }
    }
    if (opts->fastameta == FASTA_META_CSV) {   // the output name with its extension replaced (src/rw_fasta.cpp:362-372)
        std::string csv = outfile;
        const size_t slash = csv.find_last_of('/'), dot = csv.find_last_of('.');
        if (dot != std::string::npos && (slash == std::string::npos || dot > slash)) csv.erase(dot);
        csv += ".csv";
        data->out_csv.open(csv, std::ios::binary);
        if (!data->out_csv) throw std::runtime_error("Unable to open file \"" + csv + "\" for writing.");
    }
}

// src/rw_fasta.cpp:379-392
static std::string escape_string(const std::string& in) {
    if (in.find_first_of("\",\r\n") == std::string::npos) return in;
    std::string o = "\"";
    for (const char ch : in) { if (ch == '"') o += "\"\""; else o.push_back(ch); }
    o += "\"";
    return o;
}
void rw_fasta::writer::write_csv(const cseq& c) {
    if (opts->fastameta != FASTA_META_CSV) return;
    std::string o;
    if (!data->csv_started) {   // column names from the first record written
        data->csv_started = true;
        o = "name";
        for (const auto& ap : c.get_attrs()) if (ap.first != fn_family) { o += ","; o += escape_string(ap.first); }
        o += "\r\n";
    }
    o += c.getName();
    for (const auto& ap : c.get_attrs()) if (ap.first != fn_family) { o += ","; o += escape_string(ap.second); }
    o += "\r\n";
    data->out_csv.write(o.data(), (std::streamsize)o.size());
}
bool rw_fasta::writer::positional() const {   // compressed sizes are not known up front; the csv lines go out in order
    return data->fd >= 0 && !data->gz && opts->fastameta != FASTA_META_CSV && opts->copy_relatives == 0;   // ... and so do the relatives
}
bool rw_fasta::writer::compressed() const { return data->gz; }
void rw_fasta::writer::write_members(const std::string& members, unsigned int n_records, unsigned int n_excluded) {
    data->put_raw(members.data(), members.size());
    data->count += n_records;
    data->excluded += n_excluded;
}
bool rw_fasta::writer::mapped() const { return data->map; }
char* rw_fasta::writer::map_range(uint64_t offset, size_t n) const {
    const uint64_t page = (uint64_t)sysconf(_SC_PAGESIZE), lo = offset / page * page;
    void* p = mmap(nullptr, (size_t)(offset - lo) + n, PROT_READ | PROT_WRITE, MAP_SHARED, data->fd, (off_t)lo);
    if (p == MAP_FAILED) throw std::runtime_error(std::string("mmap of the output file failed: ") + strerror(errno));
    return static_cast<char*>(p) + (offset - lo);
}
void rw_fasta::writer::unmap_range(char* p, uint64_t offset, size_t n) const {
    const uint64_t page = (uint64_t)sysconf(_SC_PAGESIZE), lo = offset / page * page;
    munmap(p - (offset - lo), (size_t)(offset - lo) + n);
}
uint64_t rw_fasta::writer::reserve(uint64_t nbytes, unsigned int n_records, unsigned int n_excluded) {
    const uint64_t at = data->offset;
    if (data->map && nbytes > 0 && ftruncate(data->fd, (off_t)(at + nbytes)) != 0)
        throw std::runtime_error(std::string("growing the output file failed: ") + strerror(errno));
    data->offset += nbytes;
    data->count += n_records;
    data->excluded += n_excluded;
    return at;
}
void rw_fasta::writer::write_at(uint64_t offset, const char* p, size_t n) const { pwrite_all(data->fd, p, n, offset); }
rw_fasta::writer::~writer() = default;
unsigned int rw_fasta::writer::written() const { return data->count; }
unsigned int rw_fasta::writer::excluded() const { return data->excluded; }

// header line (+ comment lines) of a record, src/rw_fasta.cpp:438-500
static void append_header(const cseq& c, std::string& o) {
    o += ">";
    o += c.getName();
    const std::string fname = c.get_attr_string(fn_fullname);
    if (!fname.empty()) { o += " "; o += fname; }
    if (rw_fasta::opts->fastameta == FASTA_META_HEADER) {
        for (const auto& ap : c.get_attrs()) {
            if (ap.first == fn_family || ap.first == fn_fullname || ap.second.empty()) continue;
            o += " ["; o += ap.first; o += "="; o += ap.second; o += "]";
        }
        o += "\n";
    } else if (rw_fasta::opts->fastameta == FASTA_META_COMMENT) {
        o += "\n";
        for (const auto& ap : c.get_attrs()) {
            if (ap.first == fn_family || ap.first == fn_fullname) continue;
            o += "; "; o += ap.first; o += "="; o += ap.second; o += "\n";
        }
    } else {
        o += "\n";
    }
}

// length of cseq::getAligned(): the alignment width, or one past the last base where the gap placement pushed bases
// beyond it (src/cseq.cpp:135-174)
static size_t aligned_length(const cseq& c) {
    const auto& b = c.getAlignedBases();
    size_t n = c.getWidth();
    if (!b.empty()) n = std::max<size_t>(n, (size_t)b.back().getPosition() + 1);
    return n;
}

size_t rw_fasta::writer::record_size(const cseq& c) {
    if (!opts) opts = new options();
    std::string h;
    append_header(c, h);
    const size_t n = aligned_length(c);
    return h.size() + n + (opts->line_length > 0 ? (n + opts->line_length - 1) / opts->line_length : 1);
}

void rw_fasta::writer::format_into(const cseq& c, std::string& o) {  // src/rw_fasta.cpp:438-528
    if (!opts) opts = new options();
    o.clear();
    append_header(c, o);
    const auto& bases = c.getAlignedBases();
    bool sorted = true;
    for (size_t i = 1; i < bases.size() && sorted; i++) sorted = bases[i].getPosition() > bases[i - 1].getPosition();
    if (opts->line_length > 0 || !sorted) {   // wrapped lines (or an unplaced sequence): render, then cut
        const std::string seq = c.getAligned(!opts->out_dots, opts->out_dna);
        if (opts->line_length > 0) {
            for (size_t i = 0; i < seq.size(); i += opts->line_length) { o.append(seq, i, opts->line_length); o += "\n"; }
        } else {
            o += seq;
            o += "\n";
        }
        return;
    }
    // one line: fill with gap characters and drop the bases in place (what getAligned produces, without the second copy)
    const size_t h = o.size(), n = aligned_length(c);
    o.resize(h + n + 1);
    char* p = &o[h];
    memset(p, '-', n);
    if (opts->out_dots) {   // dots before the first and after the last base (cseq::getAligned(nodots = false))
        const size_t first = bases.empty() ? n : bases.front().getPosition();
        memset(p, '.', first);
        const size_t last = bases.empty() ? 0 : (size_t)bases.back().getPosition() + 1;
        if (!bases.empty() && last < n) memset(p + last, '.', n - last);
    }
    if (opts->out_dna) for (const auto& b : bases) p[b.getPosition()] = base_iupac::iupac_dna(b.getBase());
    else for (const auto& b : bases) p[b.getPosition()] = base_iupac::iupac_rna(b.getBase());
    p[n] = '\n';
}

bool rw_fasta::writer::passes_min_idty(const cseq& c) {
    if (!opts) opts = new options();
    if (!(opts->min_idty > 0)) return true;
    return !(opts->min_idty > c.get_attr<float>(fn_idty, 0.f));
}

std::string rw_fasta::writer::format(const cseq& c) {
    std::string o;
    format_into(c, o);
    return o;
}

void rw_fasta::writer::priv_data::write(const cseq& c) {
    const std::string rec = format(c);
    put(rec.data(), rec.size());
    count++;
}

void rw_fasta::writer::write_formatted(const std::string* record) {
    if (record == nullptr) { ++data->excluded; return; }  // src/rw_fasta.cpp:399-404
    data->put(record->data(), record->size());
    data->count++;
}

tray rw_fasta::writer::operator()(tray t) {
    if (t.input_sequence == nullptr) throw std::runtime_error("Received broken tray in rw_fasta writer");
    if (t.aligned_sequence == nullptr) {  // src/rw_fasta.cpp:399-404
        ++data->excluded;
        return t;
    }
    if (!passes_min_idty(*t.aligned_sequence)) {  // src/rw_fasta.cpp:405-414
        ++data->excluded;
        return t;
    }
    write_csv(*t.aligned_sequence);
    data->write(*t.aligned_sequence);
    write_relatives(t);
    return t;
}

void rw_fasta::writer::write_relatives(const tray& t) {
    if (opts->copy_relatives == 0) return;
    const search::result_vector* relatives = t.search_result != nullptr ? t.search_result : t.alignment_reference;
    if (relatives == nullptr) return;
    int i = (int)opts->copy_relatives;
    for (const auto& item : *relatives) {
        if (data->relatives_written.insert(item.sequence->getName()).second) {
            write_csv(*item.sequence);
            data->write(*item.sequence);
        }
        if (--i == 0) break;
    }
}

}  // namespace sina
