// kmat_reader.cpp -- FASTA/FASTQ ingest for the read_label host, and the per-read tally classification.
//
// Restates the single-producer parser inside read_label main() (read_label.cpp:1651-1713) as a line state
// machine over large file chunks, and the header substitution done at pop time (:1728-1732).  The reference
// refills its queue in rounds of 2*n_threads reads; a round always ends right after a read was pushed, and the
// header variables it resets per round are always reassigned before the next push, so one continuous state
// machine yields the same (header, read) sequence.
#include <cerrno>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <fcntl.h>
#include <unistd.h>

#include "kmat_internal.h"

struct kmat_read_batch {
    std::string bases, hdrs;
    std::vector<uint64_t> offs, hdr_offs;
    uint64_t first_ordinal = 1;
    uint32_t n = 0;
};

struct kmat_reader {
    int fd = -1;
    bool own_fd = false, fastq = false;
    bool file_eof = false;      // read(2) returned 0
    bool in_finished = false;   // the reference's flag: a getline failed
    std::vector<char> buf;
    size_t pos = 0, end = 0;
    std::string read_buff, hdr_buff, last_hdr_buff;
    uint64_t n_emitted = 0;     // read_count_in == read_count_out ordinal
};

static const size_t kChunk = 8u << 20;

extern "C" int kmat_reader_open(const char *path, int fastq, kmat_reader **out) {
    if (!path || !out) { kmat_set_error("kmat_reader_open: bad argument"); return KMAT_ERR_ARG; }
    kmat_reader *r = new kmat_reader();
    r->fastq = fastq != 0;
    if (strcmp(path, "-") == 0) r->fd = 0;
    else {
        r->fd = open(path, O_RDONLY);
        r->own_fd = true;
        if (r->fd < 0) { kmat_set_error("Did not open for reading: %s (%s)", path, strerror(errno)); delete r; return KMAT_ERR_IO; }
    }
    r->buf.resize(kChunk);
    *out = r;
    return KMAT_OK;
}
extern "C" void kmat_reader_close(kmat_reader *r) {
    if (!r) return;
    if (r->own_fd && r->fd >= 0) close(r->fd);
    delete r;
}
extern "C" kmat_read_batch *kmat_read_batch_new(void) { return new kmat_read_batch(); }
extern "C" void kmat_read_batch_free(kmat_read_batch *b) { delete b; }

// std::getline over the chunk buffer: false when no byte is left.  The line excludes the '\n'; a final line
// without '\n' is still returned.
static bool next_line(kmat_reader *r, const char **line, size_t *len) {
    for (;;) {
        if (r->pos < r->end) {
            const char *p = r->buf.data() + r->pos;
            const char *nl = (const char *)memchr(p, '\n', r->end - r->pos);
            if (nl) { *line = p; *len = (size_t)(nl - p); r->pos += *len + 1; return true; }
            if (r->file_eof) { *line = p; *len = r->end - r->pos; r->pos = r->end; return true; }
        } else if (r->file_eof) return false;
        // need more bytes: move the tail to the front, grow when a single line fills the buffer
        if (r->pos > 0) { memmove(r->buf.data(), r->buf.data() + r->pos, r->end - r->pos); r->end -= r->pos; r->pos = 0; }
        if (r->end == r->buf.size()) r->buf.resize(r->buf.size() * 2);
        ssize_t got;
        do { got = read(r->fd, r->buf.data() + r->end, r->buf.size() - r->end); } while (got < 0 && errno == EINTR);
        if (got <= 0) r->file_eof = true; else r->end += (size_t)got;
    }
}

static void emit(kmat_reader *r, kmat_read_batch *b, const std::string &hdr) {
    r->n_emitted++;
    b->bases.append(r->read_buff);
    b->offs.push_back(b->bases.size());
    if (hdr.empty() || hdr[0] == '\0') {                 // :1728-1732
        char tmp[48];
        snprintf(tmp, sizeof tmp, "unknown_hdr:%llu", (unsigned long long)r->n_emitted);
        b->hdrs.append(tmp);
    } else b->hdrs.append(hdr);
    b->hdr_offs.push_back(b->hdrs.size());
    b->n++;
    r->read_buff.clear();
}

extern "C" int64_t kmat_reader_next(kmat_reader *r, uint32_t max_reads, uint64_t max_bases, kmat_read_batch *b) {
    if (!r || !b) { kmat_set_error("kmat_reader_next: bad argument"); return KMAT_ERR_ARG; }
    if (max_reads == 0) max_reads = 1;
    b->bases.clear(); b->hdrs.clear(); b->offs.assign(1, 0); b->hdr_offs.assign(1, 0); b->n = 0;
    b->first_ordinal = r->n_emitted + 1;
    while (!r->in_finished && b->n < max_reads && b->bases.size() < max_bases) {
        const char *line = ""; size_t len = 0;
        if (!next_line(r, &line, &len)) { r->in_finished = true; line = ""; len = 0; }      // :1663-1669
        char c0 = len ? line[0] : '\0';
        if (c0 == '>' || (r->fastq && c0 == '@')) {                                           // :1672-1677
            r->last_hdr_buff.swap(r->hdr_buff);
            r->hdr_buff.assign(line + 1, len - 1);
        }
        if (c0 != '>' && len > 1 && !r->fastq) { r->read_buff.append(line, len); len = 0; c0 = '\0'; }              // :1679-1682
        if (r->fastq && c0 != '@' && c0 != '+' && c0 != '-') { r->read_buff.append(line, len); len = 0; c0 = '\0'; } // :1684-1687
        if (((c0 == '>' || r->in_finished) || (r->fastq && (c0 == '+' || c0 == '-'))) && !r->read_buff.empty()) {    // :1688-1707
            emit(r, b, r->in_finished ? r->hdr_buff : r->last_hdr_buff);
            if (r->fastq) { const char *q; size_t ql; next_line(r, &q, &ql); }                // the quality line is skipped
        }
    }
    return (int64_t)b->n;
}

extern "C" int kmat_read_batch_view(const kmat_read_batch *b, const char **bases, const uint64_t **offs, const char **hdrs,
                                    const uint64_t **hdr_offs, uint32_t *n_reads, uint64_t *first_ordinal) {
    if (!b) return KMAT_ERR_ARG;
    if (bases) *bases = b->bases.data();
    if (offs) *offs = b->offs.data();
    if (hdrs) *hdrs = b->hdrs.data();
    if (hdr_offs) *hdr_offs = b->hdr_offs.data();
    if (n_reads) *n_reads = b->n;
    if (first_ordinal) *first_ordinal = b->first_ordinal;
    return KMAT_OK;
}

// proc_line's bookkeeping (read_label.cpp:1217-1277)
extern "C" int kmat_tally_class(const kmat_read_result *r, float min_score, int32_t min_kmer) {
    if (!r) return -1;
    switch (r->status) {
        case KMAT_ST_SHORT_LEN: case KMAT_ST_SHORT_VALID: return 1;
        case KMAT_ST_NODBHITS: case KMAT_ST_SILENT: return 2;
        case KMAT_ST_PHIX: case KMAT_ST_LABELED:
            if (r->match == KMAT_NOMATCH) return r->valid_kmers < min_kmer ? 1 : 2;     // :1241-1253
            if (r->score >= min_score && r->valid_kmers >= min_kmer) return 0;            // :1254-1261
            if (r->score < min_score) return 3;                                           // :1262-1268
            return -1;
        default: return -1;
    }
}
