/* TEST INFRASTRUCTURE -- stand-in for perm-je's <jemalloc/pallocator.h>.
 *
 * The reference (LivGen/LMAT) depends on PERM / perm-je, a persistent jemalloc
 * that is NOT vendored under /root/reference (CMakeLists.txt:223-234 clones
 * github.com/khyox/perm-je at build time, no tag pinned).  perm-je contributes
 * no arithmetic to the read_label path: it only allocates the SortedDb arrays
 * and persists them.  This header supplies the handful of symbols the
 * reference calls (perm, mopen, mclose, mflush, JEMALLOC_P(malloc), PERM_NEW,
 * PERM_DELETE, PERM_NS::allocator) so that the UNMODIFIED reference sources
 * compile into oracle/_ref/.  Call sites: read_label.cpp:1481-1482,
 * make_db_table.cpp:330-343,429, SortedDb.hpp:164-166, TaxTable.hpp:70-74,353.
 *
 * On-disk image written by this stand-in ("KMPERM01", documented in
 * DESIGN.md; NOT the real perm-je heap format, which is unpinned):
 *   page 0 (4096 B): char magic[8]="KMPERM01"; u64 map_addr; u64 file_size;
 *                    u64 brk; u64 nreg; then nreg x { u64 nbytes; bytes
 *                    padded to 8 }  -- the perm()-registered globals
 *   page 1..      : bump-allocated heap, 64-byte aligned blocks; pointers
 *                    stored inside are absolute for a mapping at map_addr.
 */
#ifndef KMAT_ORACLE_PALLOCATOR_H
#define KMAT_ORACLE_PALLOCATOR_H
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace PERM_NS { template <class T> using allocator = std::allocator<T>; }

namespace kmperm {
static const uint64_t kMapAddr = 0x100000000000ULL;
static const size_t kHeader = 4096;
static const int kMaxReg = 16;
struct State {
    char *base = nullptr;
    size_t size = 0, brk = 0;
    int fd = -1, nreg = 0;
    bool writing = false;
    void *reg_ptr[kMaxReg];
    size_t reg_len[kMaxReg];
};
inline State &st() { static State s; return s; }
inline void *bump(size_t n) {
    State &s = st();
    size_t at = (s.brk + 63) & ~size_t(63);
    if (!s.base || at + n > s.size) {
        fprintf(stderr, "kmperm stand-in: heap exhausted (want %zu at %zu of %zu)\n", n, at, s.size);
        abort();
    }
    s.brk = at + n;
    return s.base + at;
}
inline void globals_io(bool save) {
    State &s = st();
    char *h = s.base;
    uint64_t *w = reinterpret_cast<uint64_t *>(h);
    if (save) {
        memcpy(h, "KMPERM01", 8);
        w[1] = kMapAddr; w[2] = s.size; w[3] = s.brk; w[4] = (uint64_t)s.nreg;
    }
    size_t off = 40;
    for (int i = 0; i < s.nreg; i++) {
        uint64_t n = s.reg_len[i];
        if (save) { memcpy(h + off, &n, 8); memcpy(h + off + 8, s.reg_ptr[i], n); }
        else memcpy(s.reg_ptr[i], h + off + 8, n);
        off += 8 + ((n + 7) & ~uint64_t(7));
    }
}
}  // namespace kmperm

inline int perm(void *p, size_t n) {
    kmperm::State &s = kmperm::st();
    if (s.nreg >= kmperm::kMaxReg) return -1;
    s.reg_ptr[s.nreg] = p; s.reg_len[s.nreg] = n; s.nreg++;
    return 0;
}
inline int mopen(const char *fn, const char *mode, size_t size) {
    kmperm::State &s = kmperm::st();
    s.writing = (mode[0] == 'w');
    s.fd = open(fn, s.writing ? (O_RDWR | O_CREAT | O_TRUNC) : O_RDONLY, 0644);
    if (s.fd < 0) return -1;
    if (s.writing) { if (ftruncate(s.fd, (off_t)size) != 0) return -1; }
    else { struct stat sb; if (fstat(s.fd, &sb) != 0) return -1; size = (size_t)sb.st_size; }
    void *m = mmap((void *)kmperm::kMapAddr, size, PROT_READ | PROT_WRITE,
                   (s.writing ? MAP_SHARED : MAP_PRIVATE) | MAP_FIXED_NOREPLACE, s.fd, 0);
    if (m != (void *)kmperm::kMapAddr) return -1;
    s.base = (char *)m; s.size = size; s.brk = kmperm::kHeader;
    if (!s.writing) {
        if (memcmp(s.base, "KMPERM01", 8) != 0) return -1;
        kmperm::globals_io(false);
    }
    return 0;
}
inline int mflush(void) { return 0; }
inline int mclose(void) {
    kmperm::State &s = kmperm::st();
    if (!s.base) return -1;
    if (s.writing) { kmperm::globals_io(true); msync(s.base, s.size, MS_SYNC); }
    munmap(s.base, s.size); close(s.fd);
    s.base = nullptr; s.fd = -1;
    return 0;
}
#define JEMALLOC_P(name) kmperm::bump
#define PERM_NEW(T) new (kmperm::bump(sizeof(T))) T
#define PERM_DELETE(p, T) ((void)0)
#endif
