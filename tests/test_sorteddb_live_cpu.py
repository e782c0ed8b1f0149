"""CPU, build container only (needs the reference headers under /root/reference): kmat_table_from_sorteddb driven from a
LIVE reference SortedDb object -- the documented conversion path for databases that only the reference's own allocator can
map.  tests/sorteddb_live.cpp is compiled against the reference's headers and oracle/_ref/libmetag.a,
opens a DB that the unmodified make_db_table built, converts it and checks every k-mer, count and stored id against the
reference's own begin_/next on the same object; the saved .kmat image must equal the golden dump of that DB."""
import gzip
import json
import os
import shutil
import subprocess

import numpy as np
import pytest

from lmat_b200 import api, build
from oracle import refchain as rc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden", "dbbuild")
REF = "/root/reference"


@pytest.mark.parametrize("variant,kw", [("plain", dict()), ("prune2", dict(prune=2, numrank=os.path.join(G, "numrank.txt")))])
def test_from_sorteddb_in_a_reference_process(variant, kw, tmp_path):
    if not os.path.isdir(os.path.join(REF, "src", "kmerdb")) or not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libmetag.a")) or not rc.have_ref("make_db_table"):
        pytest.skip("reference sources / oracle/_ref not present (GPU box)")
    lib, _ = build.build_all()
    exe = str(tmp_path / "sorteddb_live")
    ora = os.path.join(ROOT, "oracle")
    gomp = os.path.dirname(subprocess.run(["g++", "-print-file-name=libgomp.so"], stdout=subprocess.PIPE, text=True).stdout.strip())
    cmd = ["g++", "-std=gnu++17", "-O2", "-w", "-fopenmp", "-DIDX_CONFIG=2027", "-DTID_SIZE=16", "-DDBTID_T=uint16_t", "-DUSE_SORTED_DB=1", "-DWITH_PJMALLOC=1",
           f"-I{ora}/_ref/gen", f"-I{ora}/standins", f"-I{REF}/include", f"-I{REF}/src/kmerdb", f"-I{REF}/src", f"-I{ROOT}/include",
           "-c", os.path.join(ROOT, "tests", "sorteddb_live.cpp"), "-o", exe + ".o"]
    subprocess.run(cmd, check=True)
    subprocess.run(["g++", exe + ".o", "-o", exe, f"{ora}/_ref/libmetag.a", f"-L{os.path.dirname(lib)}", "-lkmat", f"-Wl,-rpath,{os.path.dirname(lib)}",
                    "-L/usr/lib/gcc/x86_64-linux-gnu/13", f"-L{gomp}", "-lgomp", "-lpthread", "-lz"], check=True)
    ths = []
    for i in range(4):
        p = str(tmp_path / f"th.{i}.bin")
        with gzip.open(os.path.join(G, f"th.{i}.bin.gz"), "rb") as f, open(p, "wb") as o:
            shutil.copyfileobj(f, o)
        ths.append(p)
    db = rc.make_db_table(ths, str(tmp_path / "ref.db"), 20, 2, str(tmp_path), map16=os.path.join(G, "map16.txt"), **kw)
    out = str(tmp_path / "live.kmat")
    p = subprocess.run([exe, db, out], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    st = json.loads(p.stdout.strip().splitlines()[-1])
    want = np.load(os.path.join(G, f"{variant}.npz"))
    assert st["kmers"] == len(want["kmers"]) and st["lists"] > 0 and st["absent_checked"] > 100000 and st["k"] == 20
    kmers, offs, ids = api.Table.open(out).arrays()
    assert np.array_equal(kmers, want["kmers"]) and np.array_equal(offs, want["offs"]) and np.array_equal(ids, want["ids"])
