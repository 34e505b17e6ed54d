// FASTA source and sink of the pipeline (src/rw_fasta.h:46-105, src/rw_fasta.cpp:229-315,394-528).
#ifndef SINA_B200_HOST_RW_FASTA_H
#define SINA_B200_HOST_RW_FASTA_H
#include <istream>
#include <memory>
#include <ostream>
#include <string>

#include "options.h"
#include "tray.h"

namespace sina {

enum FASTA_META_TYPE { FASTA_META_NONE = 0, FASTA_META_HEADER = 1, FASTA_META_COMMENT = 2, FASTA_META_CSV = 3 };

class rw_fasta {
public:
    struct options {
        FASTA_META_TYPE fastameta = FASTA_META_NONE;
        int line_length = 0;
        float min_idty = 0.f;              // --min-idty: only sequences with align_ident_slv above it are written
        long fasta_block = 0, fasta_idx = 0;   // --fasta-block B --fasta-idx i: the records starting in bytes (B*i, B*(i+1)] of the input
        bool out_dots = false, out_dna = false;
        unsigned int copy_relatives = 0;   // --add-relatives N: the N nearest relatives of every sequence are written too
    };
    static options* opts;
    static void get_options_description(po::options_description& main, po::options_description& adv);
    static void validate_vm(po::variables_map& vm, po::options_description& desc);

    class reader {
    public:
        // "-" = stdin; throws std::runtime_error if unreadable. whole_file: ignore --fasta-block / --fasta-idx (they select
        // a part of the QUERY file; a reference database read through this class is read whole)
        explicit reader(const std::string& infile, bool whole_file = false);
        ~reader();
        bool operator()(tray& t);                    // false at end of input; bad sequences are skipped with a message
        // the two halves of operator(): the text of the next record (one thread), and its parsing (any thread). parse_record
        // returns false for a record with an illegal character (message printed, t.input_sequence left null)
        // (append: the record's text is appended to `record` instead of replacing it)
        bool next_record(std::string& record, unsigned int& seqno, unsigned int& lineno, bool append = false);
        static bool parse_record(const std::string& record, unsigned int seqno, unsigned int lineno, const std::string& filename, tray& t);
        const std::string& filename() const;
        void count_skipped();
        unsigned int skipped() const;
    private:
        struct priv_data;
        std::shared_ptr<priv_data> data;
    };
    class writer {
    public:
        explicit writer(const std::string& outfile);  // "-" = stdout
        ~writer();
        tray operator()(tray t);
        // the record operator() would write for this sequence (header + sequence lines), so that the rendering of the
        // alignment_width-long lines can run on several threads; write_formatted() then emits it (null: sequence excluded)
        static std::string format(const cseq& c);
        // the same into a caller-owned buffer (reused from record to record), and the size it will have
        static void format_into(const cseq& c, std::string& record);
        static size_t record_size(const cseq& c);
        // --min-idty (src/rw_fasta.cpp:405-414): false for a sequence whose align_ident_slv is below the threshold
        static bool passes_min_idty(const cseq& c);
        void write_formatted(const std::string* record);
        // ".gz" output (src/rw_fasta.cpp:358-360): records go out as gzip members. gzip_member() compresses any number of
        // formatted records into one member (any thread), write_members() appends finished members in output order.
        bool compressed() const;
        static std::string gzip_member(const char* p, size_t n);
        void write_members(const std::string& members, unsigned int n_records, unsigned int n_excluded);
        // --meta-fmt csv (src/rw_fasta.cpp:362-372,484-515): the attributes of every written record as a line of
        // <outfile>.csv, the column names from the first record; to be called in output order for the records written
        void write_csv(const cseq& c);
        // --add-relatives (src/rw_fasta.cpp:419-433): the first N of the tray's search result (or, without one, of its
        // alignment family) that have not been written yet, behind the sequence's own record; in output order
        void write_relatives(const tray& t);
        // positional output (regular files): the caller reserves byte ranges in record order and any thread fills them
        // with pwrite, so that writing 50 kB records is not bound to one thread. Not available on stdout.
        bool positional() const;
        uint64_t reserve(uint64_t nbytes, unsigned int n_records, unsigned int n_excluded);
        void write_at(uint64_t offset, const char* p, size_t n) const;
        // SINA_B200_MMAP_OUT=1: reserve() grows the file and the pool fills mappings of the reserved ranges instead of
        // calling pwrite (buffered writes to one file serialise on its inode lock, page faults of a mapping do not).
        // map_range() returns the address of file byte `offset`; unmap_range() takes the same arguments back.
        bool mapped() const;
        char* map_range(uint64_t offset, size_t n) const;
        void unmap_range(char* p, uint64_t offset, size_t n) const;
        unsigned int written() const;
        unsigned int excluded() const;
    private:
        struct priv_data;
        std::shared_ptr<priv_data> data;
    };
};

}  // namespace sina
#endif
