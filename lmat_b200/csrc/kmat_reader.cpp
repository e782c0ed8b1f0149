// kmat_reader.cpp -- FASTA/FASTQ ingest for the read_label host, and the per-read tally classification.
//
// Restates the single-producer parser inside read_label main() (read_label.cpp:1651-1713) as a line state
// machine over large file chunks, and the header substitution done at pop time (:1728-1732).  The reference
// refills its queue in rounds of 2*n_threads reads; a round always ends right after a read was pushed, and the
// header variables it resets per round are always reassigned before the next push, so one continuous state
// machine yields the same (header, read) sequence.
#include <algorithm>
#include <atomic>
#include <new>
#include <cerrno>
#include <condition_variable>
#include <map>
#include <mutex>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "kmat_internal.h"

// Page-locked blocks for the bases of batches that asked for them (kmat_read_batch_new_pinned): allocated once, then handed
// from batch to batch -- cudaMallocHost costs milliseconds per block, and a parser thread takes one per 8 MB segment.  The pool
// lives for the whole process (never destroyed: the CUDA runtime may be gone by the time static destructors run).
struct KmPinPool {
    std::mutex m;
    std::vector<std::pair<char *, size_t>> free_;
    bool failed = false;                         // pinning does not work here (no GPU): plain memory from now on
    char *get(size_t want, size_t *cap) {
        {
            std::lock_guard<std::mutex> l(m);
            if (failed) return nullptr;
            size_t best = free_.size();
            for (size_t i = 0; i < free_.size(); i++)
                if (free_[i].second >= want && (best == free_.size() || free_[i].second < free_[best].second)) best = i;
            if (best < free_.size()) { char *p = free_[best].first; *cap = free_[best].second; free_.erase(free_.begin() + (long)best); return p; }
        }
        const size_t sz = (want + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
        char *p = (char *)kmat_host_alloc(sz);
        if (!p) { std::lock_guard<std::mutex> l(m); failed = true; return nullptr; }
        *cap = sz;
        return p;
    }
    void put(char *p, size_t cap) {
        {
            std::lock_guard<std::mutex> l(m);
            if (free_.size() < 96) { free_.emplace_back(p, cap); return; }
        }
        kmat_host_free(p);
    }
};
static KmPinPool &km_pin_pool() { static KmPinPool *pool = new KmPinPool(); return *pool; }

// The bases of a batch: a growable byte buffer (the std::string subset the parser uses) whose storage is a pool block when
// the batch wants page-locked memory
class KmBytes {
    char *p_ = nullptr; size_t n_ = 0, cap_ = 0; bool from_pool_ = false;
    void release() { if (p_) { if (from_pool_) km_pin_pool().put(p_, cap_); else free(p_); } p_ = nullptr; cap_ = 0; from_pool_ = false; }
  public:
    bool pinned_mode = false;
    KmBytes() = default;
    KmBytes(const KmBytes &) = delete;
    KmBytes &operator=(const KmBytes &) = delete;
    ~KmBytes() { release(); }
    size_t size() const { return n_; }
    const char *data() const { return p_ ? p_ : ""; }
    bool is_pinned() const { return from_pool_; }
    void clear() { n_ = 0; }
    void reserve(size_t want) {
        if (want <= cap_) return;
        char *q = nullptr; size_t qc = 0; bool qp = false;
        if (pinned_mode) { q = km_pin_pool().get(want, &qc); qp = q != nullptr; }
        if (!q) { qc = want; q = (char *)malloc(qc ? qc : 1); if (!q) throw std::bad_alloc(); }
        if (n_) memcpy(q, p_, n_);
        const size_t keep = n_;
        release();
        p_ = q; cap_ = qc; from_pool_ = qp; n_ = keep;
    }
    void append(const char *s, size_t len) {
        if (!len) return;
        if (n_ + len > cap_) reserve(std::max(n_ + len, cap_ + cap_ / 2 + 4096));
        memcpy(p_ + n_, s, len);
        n_ += len;
    }
    void assign(const KmBytes &src, size_t pos, size_t len) { n_ = 0; if (len > cap_) reserve(len); if (len) memcpy(p_, src.p_ + pos, len); n_ = len; }
    void swap(KmBytes &o) { std::swap(p_, o.p_); std::swap(n_, o.n_); std::swap(cap_, o.cap_); std::swap(from_pool_, o.from_pool_); }
};

struct kmat_read_batch {
    KmBytes bases;
    std::string hdrs;
    std::vector<uint64_t> offs, hdr_offs;
    std::vector<uint32_t> unknown;      // reads whose header is "unknown_hdr:<ordinal>" (needs the global ordinal)
    std::string tail_hdr;               // parallel FASTQ: the header line of the segment's last record (the next segment's first read carries it)
    bool open_end = false;              // parallel FASTQ: the segment did not end between two records (malformed input)
    uint64_t first_ordinal = 1;
    uint32_t n = 0;
    kmat_read_batch() = default;
    kmat_read_batch(const kmat_read_batch &) = delete;
    kmat_read_batch &operator=(const kmat_read_batch &) = delete;
    void clear() { bases.clear(); hdrs.clear(); offs.assign(1, 0); hdr_offs.assign(1, 0); unknown.clear(); tail_hdr.clear(); n = 0; }
};

// The line state machine of read_label main() (:1651-1713), independent of where the lines come from.
struct KmParseState {
    bool fastq = false;
    bool in_finished = false;   // the reference's flag: a getline failed
    std::string hdr_buff, last_hdr_buff;      // the pending read itself accumulates at the tail of the batch's `bases`
    uint64_t n_emitted = 0;     // read_count_in == read_count_out ordinal
    bool open_end = false;      // FASTQ: the lines ran out with a read pending or inside a quality-line skip (see km_reader_worker)
};

struct kmat_reader {
    int fd = -1;
    bool own_fd = false;
    bool file_eof = false;      // read(2) returned 0
    int io_error = 0;           // errno of a failed read(2): reported by kmat_reader_next instead of a silent truncation
    std::vector<char> buf;
    size_t pos = 0, end = 0;
    KmParseState st;
    // ---- parallel mode (FASTA, regular file): the file is mapped and cut at header lines into segments that a pool
    //      of threads parses independently; kmat_reader_next hands the parsed segments out in file order
    bool mt = false;
    const char *map = nullptr; size_t map_len = 0;
    std::vector<size_t> seg;                 // segment boundaries, seg.size() - 1 segments
    std::atomic<size_t> next_seg{0};
    size_t next_out = 0, window = 4;
    std::mutex m;
    std::condition_variable cv_done, cv_space;
    std::map<size_t, kmat_read_batch *> done;
    std::vector<std::thread> workers;
    bool stop = false;
    uint64_t ordinal = 0;
    bool mt_fastq = false;
    bool fallback = false;                   // parallel FASTQ met a segment with an open end: sequential from fb on
    const char *fb_p = nullptr;
    std::string carry_hdr;                   // parallel FASTQ: header of the last record handed out so far
    std::atomic<bool> pin_segments{false};   // a caller batch wants page-locked bases: the parser threads fill pool blocks (no copy on hand-over)
    kmat_read_batch *cur = nullptr;          // parallel mode: the parsed segment being handed out in max_reads / max_bases slices
    uint32_t cur_i = 0;
};

static const size_t kChunk = 8u << 20;

// std::getline over the chunk buffer: false when no byte is left.  The line excludes the '\n'; a final line
// without '\n' is still returned.
static bool next_line_fd(kmat_reader *r, const char **line, size_t *len) {
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
        if (got < 0) r->io_error = errno;
        if (got <= 0) r->file_eof = true; else r->end += (size_t)got;
    }
}
struct KmMemLines {             // the same over a memory range
    const char *p, *e;
    bool next(const char **line, size_t *len) {
        if (p >= e) return false;
        const char *nl = (const char *)memchr(p, '\n', (size_t)(e - p));
        *line = p;
        if (nl) { *len = (size_t)(nl - p); p = nl + 1; } else { *len = (size_t)(e - p); p = e; }
        return true;
    }
};

static void emit(KmParseState &st, kmat_read_batch *b, const std::string &hdr) {
    st.n_emitted++;
    b->offs.push_back(b->bases.size());
    if (hdr.empty() || hdr[0] == '\0') {                 // :1728-1732
        char tmp[48];
        snprintf(tmp, sizeof tmp, "unknown_hdr:%llu", (unsigned long long)st.n_emitted);
        b->hdrs.append(tmp);
        b->unknown.push_back(b->n);
    } else b->hdrs.append(hdr);
    b->hdr_offs.push_back(b->hdrs.size());
    b->n++;
}

// Runs the state machine until the batch is full or the lines run out.  read_buff of the reference = the bytes of
// `bases` past the last emitted read; it is empty whenever the loop is left (a batch only fills up right after an
// emit), so batches never split a read.
template <typename NextLine>
static void parse_lines(KmParseState &st, NextLine &&next_line, uint32_t max_reads, uint64_t max_bases, kmat_read_batch *b) {
    while (!st.in_finished && b->n < max_reads && b->bases.size() < max_bases) {
        const char *line = ""; size_t len = 0;
        if (!next_line(&line, &len)) { st.in_finished = true; line = ""; len = 0; }          // :1663-1669
        char c0 = len ? line[0] : '\0';
        if (c0 == '>' || (st.fastq && c0 == '@')) {                                           // :1672-1677
            st.last_hdr_buff.swap(st.hdr_buff);
            st.hdr_buff.assign(line + 1, len - 1);
        }
        if (c0 != '>' && len > 1 && !st.fastq) { b->bases.append(line, len); len = 0; c0 = '\0'; }              // :1679-1682
        if (st.fastq && c0 != '@' && c0 != '+' && c0 != '-') { b->bases.append(line, len); len = 0; c0 = '\0'; } // :1684-1687
        if (((c0 == '>' || st.in_finished) || (st.fastq && (c0 == '+' || c0 == '-'))) && b->bases.size() > b->offs.back()) {    // :1688-1707
            emit(st, b, st.in_finished ? st.hdr_buff : st.last_hdr_buff);
            if (st.fastq && st.in_finished) st.open_end = true;
            if (st.fastq) { const char *q; size_t ql; if (!next_line(&q, &ql)) st.open_end = true; }     // the quality line is skipped
        }
    }
}

// ---- parallel mode ----------------------------------------------------------------------------------------------
// FASTQ (-q): the state machine emits a read at its '+' line, paired with the header that was current BEFORE this record's
// '@' line (:1689-1692, the reference's quirk), then skips the quality line.  A file cut right before a record's '@' line
// parses independently on both sides except for that pairing: the first read of the right part must carry the header of
// the left part's last record -- the worker returns it (tail_hdr) and kmat_reader_next patches it in, in file order.
// Telling a record's '@' line from a quality line that happens to start with '@' needs context: a cut is only made at an
// '@' line that is followed by >= 1 plain lines, a '+' / '-' line, one more line (the quality) and then an '@' line or the
// end of the file.  A quality line starting with '@' is followed by the next header, so it never qualifies.
// That rule is enough for well-formed files.  For anything else the cut is checked after the fact: the state machine is in
// its start state at a cut exactly when the left part ended neither with a read still pending (pushed by the end-of-input
// branch) nor inside the skip of a quality line; a segment that ends otherwise is flagged (open_end) and kmat_reader_next
// then parses the rest of the file sequentially from that segment's first byte, where the state is known.
//
// A FASTA file cut right before a header line parses independently on both sides: at a '>' line the reference
// pushes the pending read with the header that preceded it, which is also what the end-of-input branch (:1663-1669)
// does for the last read of the left part, and the right part starts from the same fresh state as the file does.
// Only "unknown_hdr:<n>" (empty header, :1728-1732) needs the global read ordinal; it is patched when the segment is
// handed out in order.
static void km_reader_worker(kmat_reader *r) {
    for (;;) {
        size_t s;
        {
            std::unique_lock<std::mutex> l(r->m);
            r->cv_space.wait(l, [&] { return r->stop || r->next_seg.load() < r->next_out + r->window; });
            if (r->stop) return;
            s = r->next_seg.fetch_add(1);
        }
        if (s + 1 >= r->seg.size()) return;
        kmat_read_batch *b = new kmat_read_batch();
        b->clear();
        b->bases.pinned_mode = r->pin_segments.load();
        b->bases.reserve(r->seg[s + 1] - r->seg[s]);
        KmParseState st;
        st.fastq = r->mt_fastq;
        KmMemLines src{r->map + r->seg[s], r->map + r->seg[s + 1]};
        while (!st.in_finished) parse_lines(st, [&](const char **ln, size_t *n) { return src.next(ln, n); }, 0xFFFFFFFFu, ~0ull, b);
        b->tail_hdr = st.hdr_buff;
        b->open_end = st.open_end;
        {
            std::lock_guard<std::mutex> l(r->m);
            r->done[s] = b;
        }
        r->cv_done.notify_all();
    }
}

// Is the line starting at `p` the '@' line of a FASTQ record (see above)?
static bool km_fastq_record_start(const char *map, size_t len, size_t p) {
    if (p >= len || map[p] != '@') return false;
    auto next_line = [&](size_t at) -> size_t { const char *nl = (const char *)memchr(map + at, '\n', len - at); return nl ? (size_t)(nl - map) + 1 : len; };
    size_t q = next_line(p);
    int plain = 0;
    for (;;) {
        if (q >= len) return false;
        const char c = map[q];
        if (c == '@') return false;
        if (c == '+' || c == '-') break;
        plain++;
        q = next_line(q);
    }
    if (!plain) return false;
    q = next_line(q);                        // the quality line
    if (q >= len) return false;              // no quality line: leave the tail to one segment
    q = next_line(q);
    return q >= len || map[q] == '@';
}

static bool km_reader_start_mt(kmat_reader *r, int threads) {
    struct stat sb;
    if (fstat(r->fd, &sb) != 0 || !S_ISREG(sb.st_mode) || sb.st_size <= 0) return false;
    void *m = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, r->fd, 0);
    if (m == MAP_FAILED) return false;
    madvise(m, (size_t)sb.st_size, MADV_SEQUENTIAL);
    r->map = (const char *)m; r->map_len = (size_t)sb.st_size;
    size_t seg_bytes = 8u << 20;
    if (const char *e = getenv("KMAT_READER_SEG_BYTES")) { const long v = atol(e); if (v > 0) seg_bytes = (size_t)v; }
    r->seg.push_back(0);
    for (size_t at = seg_bytes; at < r->map_len; at += seg_bytes) {
        // first header line at or after `at`
        size_t p = at;
        const char *hit = nullptr;
        while (p < r->map_len) {
            const char *nl = (const char *)memchr(r->map + p, '\n', r->map_len - p);
            if (!nl) break;
            if (r->mt_fastq ? km_fastq_record_start(r->map, r->map_len, (size_t)(nl - r->map) + 1) : ((size_t)(nl - r->map) + 1 < r->map_len && nl[1] == '>')) { hit = nl + 1; break; }
            p = (size_t)(nl - r->map) + 1;
        }
        if (!hit) break;
        const size_t cut = (size_t)(hit - r->map);
        if (cut > r->seg.back()) r->seg.push_back(cut);
        if (cut > at) at = cut - (cut % seg_bytes);          // a very long record: continue after it
    }
    r->seg.push_back(r->map_len);
    r->window = (size_t)threads * 2 + 2;
    r->mt = true;
    for (int t = 0; t < threads; t++) r->workers.emplace_back(km_reader_worker, r);
    return true;
}

extern "C" int kmat_reader_open_mt(const char *path, int fastq, int threads, kmat_reader **out) {
    if (!path || !out) { kmat_set_error("kmat_reader_open: bad argument"); return KMAT_ERR_ARG; }
    kmat_reader *r = new kmat_reader();
    r->st.fastq = fastq != 0;
    if (strcmp(path, "-") == 0) r->fd = 0;
    else {
        r->fd = open(path, O_RDONLY);
        r->own_fd = true;
        if (r->fd < 0) { kmat_set_error("Did not open for reading: %s (%s)", path, strerror(errno)); delete r; return KMAT_ERR_IO; }
    }
    // stdin cannot be mapped
    r->mt_fastq = fastq != 0;
    if (!(threads > 1 && r->own_fd && km_reader_start_mt(r, threads))) r->buf.resize(kChunk);
    *out = r;
    return KMAT_OK;
}
extern "C" int kmat_reader_open(const char *path, int fastq, kmat_reader **out) { return kmat_reader_open_mt(path, fastq, 1, out); }
extern "C" void kmat_reader_close(kmat_reader *r) {
    if (!r) return;
    if (r->mt) {
        { std::lock_guard<std::mutex> l(r->m); r->stop = true; }
        r->cv_space.notify_all();
        for (auto &t : r->workers) t.join();
        for (auto &kv : r->done) delete kv.second;
        delete r->cur;
        munmap((void *)r->map, r->map_len);
    }
    if (r->own_fd && r->fd >= 0) close(r->fd);
    delete r;
}
extern "C" kmat_read_batch *kmat_read_batch_new(void) { return new kmat_read_batch(); }
extern "C" kmat_read_batch *kmat_read_batch_new_pinned(void) { kmat_read_batch *b = new kmat_read_batch(); b->bases.pinned_mode = true; return b; }
extern "C" void kmat_read_batch_free(kmat_read_batch *b) { delete b; }

extern "C" int64_t kmat_reader_next(kmat_reader *r, uint32_t max_reads, uint64_t max_bases, kmat_read_batch *b) {
    if (!r || !b) { kmat_set_error("kmat_reader_next: bad argument"); return KMAT_ERR_ARG; }
    if (max_reads == 0) max_reads = 1;
    b->clear();
    if (b->bases.pinned_mode) r->pin_segments.store(true);
    // reads [cur_i, ...) of the current parsed segment, cut to max_reads / max_bases (at least one read)
    auto slice = [&]() -> int64_t {
        kmat_read_batch *d = r->cur;
        const uint32_t i0 = r->cur_i;
        uint32_t i1 = i0;
        while (i1 < d->n && i1 - i0 < max_reads && (i1 == i0 || d->offs[i1 + 1] - d->offs[i0] <= max_bases)) i1++;
        const bool last = i1 >= d->n;
        if (i0 == 0 && last) {                                   // the whole segment fits: no copy
            b->bases.swap(d->bases); std::swap(b->hdrs, d->hdrs); std::swap(b->offs, d->offs); std::swap(b->hdr_offs, d->hdr_offs);
            std::swap(b->unknown, d->unknown); std::swap(b->tail_hdr, d->tail_hdr);
            b->open_end = d->open_end; b->first_ordinal = d->first_ordinal; b->n = d->n;
        } else {
            b->bases.assign(d->bases, d->offs[i0], (size_t)(d->offs[i1] - d->offs[i0]));
            b->hdrs.assign(d->hdrs, d->hdr_offs[i0], d->hdr_offs[i1] - d->hdr_offs[i0]);
            for (uint32_t i = i0; i < i1; i++) { b->offs.push_back(d->offs[i + 1] - d->offs[i0]); b->hdr_offs.push_back(d->hdr_offs[i + 1] - d->hdr_offs[i0]); }
            b->n = i1 - i0;
            b->first_ordinal = d->first_ordinal + i0;
        }
        r->cur_i = i1;
        if (last) { delete r->cur; r->cur = nullptr; r->cur_i = 0; }
        return (int64_t)b->n;
    };
    if (r->mt && r->cur) return slice();
    if (r->mt && !r->fallback) {
        // segments come out in file order; empty ones are skipped
        for (;;) {
            if (r->next_out + 1 >= r->seg.size()) return 0;
            kmat_read_batch *d = nullptr;
            {
                std::unique_lock<std::mutex> l(r->m);
                r->cv_done.wait(l, [&] { return r->done.count(r->next_out) != 0; });
                d = r->done[r->next_out];
                r->done.erase(r->next_out);
                r->next_out++;
                if (d->open_end && r->next_out + 1 < r->seg.size()) {     // not the last segment: its end is not the end of the input
                    r->stop = true;
                    r->fallback = true;
                }
            }
            r->cv_space.notify_all();
            if (r->fallback) {
                delete d;
                r->fb_p = r->map + r->seg[r->next_out - 1];
                r->st = KmParseState();
                r->st.fastq = true; r->st.hdr_buff = r->carry_hdr; r->st.n_emitted = r->ordinal;
                break;
            }
            d->first_ordinal = r->ordinal + 1;
            if (!d->unknown.empty()) {                         // rebuild the headers with the global ordinals
                std::string h; std::vector<uint64_t> ho(1, 0);
                size_t u = 0;
                for (uint32_t i = 0; i < d->n; i++) {
                    if (u < d->unknown.size() && d->unknown[u] == i) {
                        u++;
                        if (r->mt_fastq && i == 0 && !r->carry_hdr.empty() && r->carry_hdr[0] != '\0') h.append(r->carry_hdr);   // the previous segment's last header
                        else {
                            char tmp[48];
                            snprintf(tmp, sizeof tmp, "unknown_hdr:%llu", (unsigned long long)(r->ordinal + i + 1));
                            h.append(tmp);
                        }
                    } else h.append(d->hdrs, d->hdr_offs[i], d->hdr_offs[i + 1] - d->hdr_offs[i]);
                    ho.push_back(h.size());
                }
                d->hdrs.swap(h); d->hdr_offs.swap(ho);
            }
            r->ordinal += d->n;
            if (r->mt_fastq) r->carry_hdr = d->tail_hdr;
            if (!d->n) { delete d; continue; }
            r->cur = d; r->cur_i = 0;
            return slice();
        }
    }
    b->first_ordinal = r->st.n_emitted + 1;
    if (r->fallback) {
        KmMemLines src{r->fb_p, r->map + r->map_len};
        parse_lines(r->st, [&](const char **ln, size_t *n) { return src.next(ln, n); }, max_reads, max_bases, b);
        r->fb_p = src.p;
        return (int64_t)b->n;
    }
    parse_lines(r->st, [&](const char **ln, size_t *n) { return next_line_fd(r, ln, n); }, max_reads, max_bases, b);
    if (r->io_error) { kmat_set_error("read error on the input: %s", strerror(r->io_error)); return KMAT_ERR_IO; }
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
