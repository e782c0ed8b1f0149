import gzip
import hashlib
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running CPU check (exhaustive sweeps)")


def _sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


class GoldenScenario:
    """Seeded text inputs regenerated into a tmp dir (checked against manifest.json) + the committed
    reference table dump and reference read_label outputs."""

    def __init__(self, name, workdir):
        import scenarios as S
        self.name = name
        self.manifest = json.load(open(os.path.join(GOLDEN, "manifest.json")))[name]
        self.inp = S.build_inputs(name, workdir)
        self.paths = self.inp["paths"]
        self.workdir = workdir
        for k, want in self.manifest["inputs"].items():
            if k.startswith("null."):
                got = hashlib.sha256(gzip.open(os.path.join(workdir, k)).read()).hexdigest()
            else:
                got = _sha(self.paths[k])
            assert got == want, f"seeded input {k} of scenario {name} drifted from the golden manifest"
        t = np.load(os.path.join(GOLDEN, f"{name}.table.npz"))
        self.kmers, self.offs, self.ids = t["kmers"], t["offs"], t["ids"].astype(np.uint32)
        self.kmer_len, self.tid_bytes = int(t["kmer_len"]), int(t["tid_bytes"])

    def golden_out(self, tag):
        return gzip.open(os.path.join(GOLDEN, f"{self.name}.{tag}.out.gz")).read().decode("latin-1")

    def golden_file(self, suffix):
        return open(os.path.join(GOLDEN, f"{self.name}.{suffix}")).read()


@pytest.fixture(scope="session")
def golden_small(tmp_path_factory):
    return GoldenScenario("small", str(tmp_path_factory.mktemp("small")))


@pytest.fixture(scope="session")
def golden_lists(tmp_path_factory):
    return GoldenScenario("lists", str(tmp_path_factory.mktemp("lists")))


def oracle_for(g, opts_name):
    """Oracle instance configured like scenarios.OPTION_SETS[opts_name] over golden scenario g."""
    import scenarios as S
    from oracle import oracle_py as op
    o = S.OPTION_SETS[opts_name]
    sd = op.SortedDbArrays(g.kmers, g.offs, g.ids, g.kmer_len, g.tid_bytes)
    orc = op.Oracle(cdb=sd.cdb(), keep=sd)
    orc.set_opts(min_kmer=o["min_kmer"], hbias=o["hbias"], sdiff=o["sdiff"], min_score=o["min_score"],
                 prn_all=int(o["prn_all"]), permissive=int(bool(o.get("permissive"))),
                 phix_screen=0 if o.get("phix_off") else 1, min_fnd_kmer=o.get("min_fnd", 1),
                 max_count=o.get("prune", 65535), prn_read=0 if o.get("hide_read") else 1)
    P = g.paths
    orc.load_files(tree=P["tree"], depth=P["depth"], rank=P["rank"], map16=P["map16"],
                   numrank=P["numrank"] if o.get("prune") else None,
                   plasmids=P["plasmids"] if o.get("plasmids") else None,
                   null_lst=P["null_lst"] if o["null"] else None, lmat_dir=g.workdir)
    return orc
