"""The read_label host binary (lmat_b200/csrc/read_label_main.cpp): option handling on CPU, and on the GPU the
whole drop-in path -- DB file + FASTA/FASTQ in, <ofbase><t>.out / .fastsummary / .nomatchsum out -- against
the outputs of the unmodified reference committed under tests/golden/."""
import os
import re
import subprocess

import pytest

import scenarios as S
from lmat_b200 import api, build


@pytest.fixture(scope="session")
def cli():
    lib, exe = build.build_all()
    assert exe and os.path.exists(exe)
    return exe


def run_cli(exe, args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([exe] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=e, timeout=600)


def test_cli_version_and_usage(cli):
    p = run_cli(cli, ["-V"])
    assert p.returncode == 0 and "LMAT version" in p.stdout
    p = run_cli(cli, ["-H"])
    assert p.returncode == 0 and "usage" in p.stdout


def test_cli_missing_required_options(cli):
    """read_label.cpp:1443-1453: every missing required option is named on stderr, exit status -1."""
    p = run_cli(cli, ["-d", "x.db"])
    assert p.returncode == 255
    for name in ("depth_file", "ofbase", "n_threads", "query_fn"):
        assert f"ERROR! Missing {name}" in p.stderr


def test_cli_bad_db(cli, tmp_path):
    p = run_cli(cli, ["-d", str(tmp_path / "nope.db"), "-i", "r.fa", "-o", str(tmp_path / "o"), "-t", "1", "-e", "d"])
    assert p.returncode == 255 and "unable to open kmer db" in p.stderr


def test_cli_fails_loudly_without_gpu(cli, golden_small, tmp_path):
    """No CPU fallback: with no device the binary must exit non-zero, not produce labels."""
    if api.device_count() > 0:
        pytest.skip("a GPU is present")
    g = golden_small
    db = str(tmp_path / "small.kmat")
    api.Table.from_arrays(g.kmers, g.offs, g.ids, g.kmer_len, g.tid_bytes).save(db)
    P = g.paths
    p = run_cli(cli, ["-d", db, "-i", P["reads"], "-o", str(tmp_path / "o"), "-t", "1", "-e", P["depth"], "-c", P["tree"]])
    assert p.returncode != 0 and "No CUDA device" in p.stderr
    assert not os.path.exists(str(tmp_path / "o0.out"))


def ref_args(g, o, db, reads, ofbase, threads, fastq=False):
    """The flag set of bin/run_rl.sh:243 as oracle/refchain.read_label passes it to the reference."""
    P = g.paths
    a = ["-f", P["map16"]]
    if o.get("prune"):
        a += ["-g", str(o["prune"]), "-m", P["numrank"]]
    a += ["-u", P["names"], "-w", P["rank"], "-x", str(o["min_score"]), "-j", str(o["min_kmer"]), "-l", str(o["hbias"]), "-b", str(o["sdiff"])]
    if o["null"]:
        a += ["-n", P["null_lst"]]
    a += ["-e", P["depth"]]
    if o["prn_all"]:
        a += ["-p"]
    if o.get("plasmids"):
        a += ["-r", P["plasmids"]]
    a += ["-t", str(threads), "-i", reads, "-d", db, "-c", P["tree"], "-o", ofbase]
    if fastq:
        a += ["-q"]
    if o.get("phix_off"):
        a += ["-h"]
    if o.get("hide_read"):
        a += ["-a"]
    if o.get("min_fnd"):
        a += ["-z", str(o["min_fnd"])]
    return a


@pytest.fixture(scope="module")
def flat_dbs(golden_small, golden_lists, tmp_path_factory):
    d = tmp_path_factory.mktemp("flatdb")
    out = {}
    for g in (golden_small, golden_lists):
        out[g.name] = str(d / f"{g.name}.kmat")
        api.Table.from_arrays(g.kmers, g.offs, g.ids, g.kmer_len, g.tid_bytes).save(out[g.name])
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("scen", ["golden_small", "golden_lists"])
def test_cli_reproduces_reference_run(scen, request, cli, flat_dbs, tmp_path):
    """-t 1: .out, .fastsummary and .nomatchsum byte-identical to the reference's (run_rl.sh option set)."""
    g = request.getfixturevalue(scen)
    ofb = str(tmp_path / "rl_")
    p = run_cli(cli, ref_args(g, S.OPTION_SETS["run_rl"], flat_dbs[g.name], g.paths["reads"], ofb, 1), env={"LMAT_DIR": g.workdir, "KMAT_BATCH_READS": "100", "KMAT_CHUNK_READS": "37"})   # 3 pipelined chunks per batch
    assert p.returncode == 0, p.stderr
    assert "Total query time" in p.stdout and "Total reads loaded" in p.stdout
    assert open(ofb + "0.out", encoding="latin-1").read() == g.golden_out("run_rl")
    assert open(ofb + ".0.30.fastsummary").read() == g.golden_file("run_rl.fastsummary")
    assert open(ofb + ".0.30.nomatchsum").read() == g.golden_file("run_rl.nomatchsum")


@pytest.mark.gpu
@pytest.mark.parametrize("scen", ["golden_small", "golden_lists"])
@pytest.mark.parametrize("host_format", [False, True])
def test_cli_verbose_prints_negative_candidates(scen, host_format, request, cli, flat_dbs, tmp_path):
    """-y: the debug traces on stdout are not produced, but its effect on the .out file is -- under -p the candidates with a
    negative score are printed too (read_label.cpp:901).  Golden: the unmodified reference run with -p -y
    (tests/golden/make_golden_verbose.py); 10 / 39 lines differ from the plain -p output.  Both formatters (device, host)."""
    g = request.getfixturevalue(scen)
    ofb = str(tmp_path / "rl_")
    env = {"LMAT_DIR": g.workdir}
    if host_format:
        env["KMAT_HOST_FORMAT"] = "1"
    p = run_cli(cli, ref_args(g, S.OPTION_SETS["run_rl"], flat_dbs[g.name], g.paths["reads"], ofb, 1) + ["-y"], env=env)
    assert p.returncode == 0, p.stderr
    got = open(ofb + "0.out", encoding="latin-1").read()
    assert got == g.golden_out("run_rl_verbose")
    assert got != g.golden_out("run_rl")


@pytest.mark.gpu
@pytest.mark.parametrize("opts", ["defaults", "tight", "prune3", "nophix_hide", "quirk", "plasmid", "nonull"])
def test_cli_option_sets(opts, golden_lists, cli, flat_dbs, tmp_path):
    g = golden_lists
    ofb = str(tmp_path / "rl_")
    p = run_cli(cli, ref_args(g, S.OPTION_SETS[opts], flat_dbs[g.name], g.paths["reads"], ofb, 1), env={"LMAT_DIR": g.workdir})
    assert p.returncode == 0, p.stderr
    assert open(ofb + "0.out", encoding="latin-1").read() == g.golden_out(opts)


@pytest.mark.gpu
@pytest.mark.parametrize("tag,key,fastq", [("wrapped", "reads_wrapped", False), ("fastq", "reads_fq", True)])
def test_cli_input_formats(tag, key, fastq, golden_small, cli, flat_dbs, tmp_path):
    g = golden_small
    ofb = str(tmp_path / "rl_")
    p = run_cli(cli, ref_args(g, S.OPTION_SETS["run_rl"], flat_dbs[g.name], g.paths[key], ofb, 1, fastq=fastq), env={"LMAT_DIR": g.workdir})
    assert p.returncode == 0, p.stderr
    assert open(ofb + "0.out", encoding="latin-1").read() == g.golden_out(tag)


@pytest.mark.gpu
def test_cli_multi_file_output_is_same_multiset(golden_small, cli, flat_dbs, tmp_path):
    """-t 3 with small batches: the records of the three .out files are the reference's records (as a multiset),
    read from stdin ('-i -')."""
    g = golden_small
    ofb = str(tmp_path / "rl_")
    args = ref_args(g, S.OPTION_SETS["nonull"], flat_dbs[g.name], "-", ofb, 3)
    e = dict(os.environ, LMAT_DIR=g.workdir, KMAT_BATCH_READS="17")
    p = subprocess.run([cli] + args, stdin=open(g.paths["reads"], "rb"), stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=e, timeout=600)
    assert p.returncode == 0, p.stderr
    got = []
    for t in range(3):
        got += open(f"{ofb}{t}.out", encoding="latin-1").read().split("\n")
    want = g.golden_out("nonull").split("\n")

    def records(lines):
        # the silent-NoMatch read writes no newline, so the record after it shares its line
        out = []
        for ln in lines:
            m = re.match(r"^(period25\t[A-Za-z]+\t)(.*)$", ln)
            out += [m.group(1), m.group(2)] if m else [ln]
        return sorted(x for x in out if x)
    assert records(got) == records(want)


@pytest.mark.gpu
def test_cli_sharded_table_mode(golden_lists, cli, flat_dbs, tmp_path):
    """KMAT_TABLE_MODE=sharded: each worker holds one shard and reads the others' buckets through the peer table (here
    three workers on the same device); the records are the reference's, as a multiset over the workers' batches."""
    g = golden_lists
    ofb = str(tmp_path / "rl_")
    args = ref_args(g, S.OPTION_SETS["run_rl"], flat_dbs[g.name], g.paths["reads"], ofb, 1)
    p = run_cli(cli, args, env={"LMAT_DIR": g.workdir, "KMAT_TABLE_MODE": "sharded", "KMAT_DEVICES": "0,0,0", "KMAT_BATCH_READS": "50"})
    assert p.returncode == 0, p.stderr
    assert "Table sharded over 3 GPUs" in p.stdout
    assert open(ofb + "0.out", encoding="latin-1").read() == g.golden_out("run_rl")      # -t 1: one writer, input order
    assert open(ofb + ".0.30.fastsummary").read() == g.golden_file("run_rl.fastsummary")


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["replicated", "sharded", "exchange"])
def test_cli_two_distinct_devices(mode, golden_lists, cli, flat_dbs, tmp_path):
    """Two different GPUs driven by one process (one worker thread and one CUDA context per device): replicated table, direct
    peer reads, and the NCCL exchange (KMAT_TABLE_MODE=exchange: kmat_shard_label_batch, the workers in lockstep).  -t 1 and
    input order, so the outputs are byte-identical to the reference's."""
    if api.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    g = golden_lists
    ofb = str(tmp_path / "rl_")
    args = ref_args(g, S.OPTION_SETS["run_rl"], flat_dbs[g.name], g.paths["reads"], ofb, 1)
    env = {"LMAT_DIR": g.workdir, "KMAT_DEVICES": "0,1", "KMAT_BATCH_READS": "50"}
    if mode != "replicated":
        env["KMAT_TABLE_MODE"] = mode
    p = run_cli(cli, args, env=env)
    assert p.returncode == 0, p.stderr
    if mode == "exchange":
        assert "exchanged over NCCL" in p.stdout
    assert open(ofb + "0.out", encoding="latin-1").read() == g.golden_out("run_rl")
    assert open(ofb + ".0.30.fastsummary").read() == g.golden_file("run_rl.fastsummary")
