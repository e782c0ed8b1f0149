"""CPU: pin the two pieces of 'library behaviour' the oracle (and the CUDA path) restate:
glibc's logf and libstdc++'s std::sort / priority_queue tie order."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle_py as op

LOGF_CHECK = r"""
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
float kmo_logf(float);
int main(int argc, char **argv) {
    uint64_t step = argc > 1 ? strtoull(argv[1], 0, 10) : 1; long bad = 0;
    for (uint64_t u = 1; u < 0x7f800000ull; u += step) {
        float x; uint32_t w = (uint32_t)u; memcpy(&x, &w, 4);
        float a = logf(x), b = kmo_logf(x);
        if (memcmp(&a, &b, 4)) bad++;
    }
    printf("%ld\n", bad); return 0;
}
"""


def _build_logf_check(tmp_path):
    src = tmp_path / "logf_check.c"
    src.write_text("#include <stdlib.h>\n" + LOGF_CHECK)
    exe = tmp_path / "logf_check"
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", str(src), "-o", str(exe), "-L" + op.HERE,
                           "-lkmat_oracle", "-Wl,-rpath," + op.HERE, "-lm"])
    return exe


def test_logf_port_matches_host_libm_sampled(tmp_path):
    """Every 257th positive float (8.3 M values); the exhaustive sweep is test_logf_port_exhaustive."""
    op.build_lib()
    exe = _build_logf_check(tmp_path)
    assert subprocess.check_output([str(exe), "257"]).strip() == b"0"


@pytest.mark.slow
@pytest.mark.skipif(not os.environ.get("KMAT_SLOW"), reason="set KMAT_SLOW=1 for the exhaustive 2^31 sweep")
def test_logf_port_exhaustive(tmp_path):
    op.build_lib()
    exe = _build_logf_check(tmp_path)
    assert subprocess.check_output([str(exe), "1"]).strip() == b"0"


SORT_CHECK = r"""
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <queue>
#include <vector>
struct P { uint32_t tid; float score; };
struct ByMod { bool operator()(const P& a, const P& b) const { return (a.tid % 7) > (b.tid % 7); } };
struct Tol { bool operator()(const P& a, const P& b) const {
    if (std::fabs(a.score - b.score) < 0.001) return (a.tid % 5) < (b.tid % 5); return a.score < b.score; } };
struct MyPair { unsigned first; uint32_t second; bool operator<(const MyPair& o) const { return first < o.first; } };
int main() {
    uint32_t n, mode;
    while (scanf("%u %u", &mode, &n) == 2) {
        std::vector<P> v(n);
        for (auto& p : v) if (scanf("%u %f", &p.tid, &p.score) != 2) return 1;
        if (mode == 0) std::sort(v.begin(), v.end(), ByMod());
        else if (mode == 1) std::sort(v.begin(), v.end(), Tol());
        else {
            std::priority_queue<MyPair> q;
            for (auto& p : v) q.push(MyPair{p.tid % 7, p.tid});
            for (auto& p : v) { p.tid = q.top().second; q.pop(); }
        }
        for (auto& p : v) printf("%u ", p.tid);
        printf("\n");
    }
}
"""


def test_std_sort_and_heap_emulation_match_libstdcxx(tmp_path):
    L = op.lib()
    src = tmp_path / "sort_check.cpp"
    src.write_text(SORT_CHECK)
    exe = tmp_path / "sort_check"
    subprocess.check_call(["g++", "-O2", str(src), "-o", str(exe)])
    LESS = C.CFUNCTYPE(C.c_int, C.POINTER(op.Pair), C.POINTER(op.Pair), C.c_void_p)
    by_mod = LESS(lambda a, b, _: int((a[0].tid % 7) > (b[0].tid % 7)))

    def tol(a, b, _):
        if abs(float(np.float32(a[0].score) - np.float32(b[0].score))) < 0.001:
            return int((a[0].tid % 5) < (b[0].tid % 5))
        return int(a[0].score < b[0].score)
    tol = LESS(tol)
    L.kmo_std_sort.argtypes = [C.POINTER(op.Pair), C.c_size_t, LESS, C.c_void_p]
    L.kmo_heap_push.argtypes = [C.POINTER(op.Pair), C.POINTER(C.c_size_t), op.Pair]
    L.kmo_heap_pop.argtypes = [C.POINTER(op.Pair), C.POINTER(C.c_size_t)]
    L.kmo_heap_pop.restype = op.Pair
    rng = np.random.default_rng(11)
    cases, text = [], []
    for n in list(range(0, 40)) + [64, 100, 257, 1000, 5000]:
        for mode in (0, 1, 2):
            tids = rng.permutation(100000)[:n].astype(np.uint32)
            scores = (rng.integers(0, 12, size=n) * 0.0007).astype(np.float32)
            cases.append((mode, tids, scores))
            text.append(f"{mode} {n} " + " ".join(f"{t} {s:.9g}" for t, s in zip(tids, scores)))
    # adversarial for introsort's depth limit: organ-pipe and sorted inputs under the tie-heavy comparator
    for n in (3000,):
        t = np.arange(n, dtype=np.uint32)
        for arr in (t, t[::-1].copy(), np.concatenate([t[::2], t[1::2][::-1]])):
            cases.append((0, arr, np.zeros(n, np.float32)))
            text.append(f"0 {n} " + " ".join(f"{x} 0" for x in arr))
    out = subprocess.run([str(exe)], input="\n".join(text).encode(), stdout=subprocess.PIPE, check=True).stdout.decode().split("\n")
    for (mode, tids, scores), line in zip(cases, out):
        n = len(tids)
        arr = (op.Pair * max(n, 1))()
        for i in range(n):
            arr[i].tid, arr[i].score = int(tids[i]), float(scores[i])
        if mode in (0, 1):
            L.kmo_std_sort(arr, n, by_mod if mode == 0 else tol, None)
            got = [arr[i].tid for i in range(n)]
        else:
            heap = (op.Pair * max(n, 1))()
            hn = C.c_size_t(0)
            for i in range(n):
                p = op.Pair(int(tids[i]), 0.0)
                p.score = np.frombuffer(np.uint32(int(tids[i]) % 7).tobytes(), dtype=np.float32)[0]
                L.kmo_heap_push(heap, C.byref(hn), p)
            got = [L.kmo_heap_pop(heap, C.byref(hn)).tid for _ in range(n)]
        want = [int(x) for x in line.split()]
        assert got == want, (mode, n)
