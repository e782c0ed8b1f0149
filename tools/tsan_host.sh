#!/bin/bash
# The threaded host code (parallel FASTA/FASTQ reader incl. early close and the sequential fallback, parallel table build)
# under ThreadSanitizer: tools/tsan_host_driver.cpp + the three .cpp sources, no CUDA needed.  Usage: tools/tsan_host.sh
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
W=/tmp/kmat_tsan; mkdir -p $W
g++ -O1 -g -fsanitize=thread -std=c++17 -I$ROOT/include $ROOT/tools/tsan_host_driver.cpp $ROOT/lmat_b200/csrc/kmat_host.cpp \
    $ROOT/lmat_b200/csrc/kmat_reader.cpp $ROOT/lmat_b200/csrc/kmat_build.cpp -o $W/drv -lz -lpthread
python - "$W" "$ROOT" <<'PY'
import gzip, random, sys
W, ROOT = sys.argv[1], sys.argv[2]
rng = random.Random(3)
with open(W + "/r.fa", "w") as f:
    for i in range(60000):
        f.write(">r%d\n" % i if i % 97 else ">\n")
        s = "".join(rng.choice("ACGT") for _ in range(rng.choice([30, 150, 151])))
        f.write(s[:80] + "\n" + s[80:] + "\n")
for name, broken in (("r.fq", -1), ("bad.fq", 20000)):
    with open(W + "/" + name, "w") as f:
        for i in range(60000):
            s = "".join(rng.choice("ACGT") for _ in range(50))
            f.write("@q%d\n%s\n" % (i, s) if i == broken else "@q%d\n%s\n+\n%s\n" % (i, s, "@" + "I" * 49))
for i in range(4):
    open(W + "/th.%d.bin" % i, "wb").write(gzip.open(ROOT + "/tests/golden/dbbuild/th.%d.bin.gz" % i).read())
PY
cd $W
export KMAT_READER_SEG_BYTES=65536 TSAN_OPTIONS=halt_on_error=1
./drv r.fa r.fq $ROOT/tests/golden/dbbuild/map16.txt th.0.bin th.1.bin th.2.bin th.3.bin
./drv r.fa bad.fq
echo "tsan: clean"
