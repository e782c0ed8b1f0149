/* TEST INFRASTRUCTURE.  LD_PRELOAD shim for the UNMODIFIED reference rand_read_label: its only source of
 * irreproducibility is `std::srand(unsigned(std::time(0)))` (src/rand_read_label.cpp:412).  With this library
 * preloaded, time() returns $KMAT_FIXED_TIME, so a `-t 1` run draws a known glibc rand() sequence and its .rand_lst
 * becomes a golden vector (tests/golden/make_nullgen_golden.py).  Nothing of the reference is modified or copied. */
#include <stdlib.h>
#include <time.h>
time_t time(time_t *t) {
    const char *s = getenv("KMAT_FIXED_TIME");
    time_t v = s ? (time_t)strtoll(s, NULL, 10) : (time_t)0;
    if (t) *t = v;
    return v;
}
