// Per-query carrier passed between the pipeline stages (src/tray.h:41-57): raw owning pointers freed by
// destroy() at the sink (src/tray.cpp:77-86).
#ifndef SINA_B200_HOST_TRAY_H
#define SINA_B200_HOST_TRAY_H
#include <sstream>

#include "search.h"

namespace sina {

class alignment_stats;  // positional-variability weights need ARB SAI data: always absent here (width 0 => simple scheme)

class tray {
public:
    unsigned int seqno{0};
    cseq* input_sequence{nullptr};
    cseq* aligned_sequence{nullptr};
    search::result_vector* alignment_reference{nullptr};
    search::result_vector* search_result{nullptr};
    std::stringstream log;
    alignment_stats* astats{nullptr};

    tray() = default;
    tray(const tray& o)
        : seqno(o.seqno), input_sequence(o.input_sequence), aligned_sequence(o.aligned_sequence),
          alignment_reference(o.alignment_reference), search_result(o.search_result), astats(o.astats) {
        log.str(o.log.str());
    }
    tray& operator=(const tray& o) {
        seqno = o.seqno; input_sequence = o.input_sequence; aligned_sequence = o.aligned_sequence;
        alignment_reference = o.alignment_reference; search_result = o.search_result; astats = o.astats;
        log.str(o.log.str());
        return *this;
    }
    void destroy() {
        delete input_sequence; delete aligned_sequence; delete alignment_reference; delete search_result;
        input_sequence = aligned_sequence = nullptr;
        alignment_reference = search_result = nullptr;
    }
};

}  // namespace sina
#endif
