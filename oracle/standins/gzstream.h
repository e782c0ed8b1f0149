/* TEST INFRASTRUCTURE -- stand-in for the un-vendored gzstream library
 * (reference CMakeLists.txt:272-285 downloads it).  The reference only uses
 * `igzstream ifs(path); ifs.getline(buf, n)` (read_label.cpp:570-585,
 * gene_label.cpp:658-681).  zlib's gzread handles plain and gzip files. */
#ifndef KMAT_ORACLE_GZSTREAM_H
#define KMAT_ORACLE_GZSTREAM_H
#include <istream>
#include <streambuf>
#include <zlib.h>
class kmat_gzbuf : public std::streambuf {
    gzFile f_ = nullptr;
    char buf_[1 << 16];
public:
    bool open(const char *name) { f_ = gzopen(name, "rb"); return f_ != nullptr; }
    ~kmat_gzbuf() override { if (f_) gzclose(f_); }
    int underflow() override {
        if (!f_) return traits_type::eof();
        int n = gzread(f_, buf_, sizeof buf_);
        if (n <= 0) return traits_type::eof();
        setg(buf_, buf_, buf_ + n);
        return traits_type::to_int_type(buf_[0]);
    }
};
class igzstream : public std::istream {
    kmat_gzbuf b_;
public:
    explicit igzstream(const char *name) : std::istream(&b_) { if (!b_.open(name)) setstate(std::ios::failbit); }
};
#endif
