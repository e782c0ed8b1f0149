#!/usr/bin/env python
"""bench.py -- reads/s and k-mer lookups/s of the read_label hot path on synthetic 150 bp reads, k = 20.

Workload (BASELINE.json configs[1]): a synthetic table built from G random genomes of
L bp (defaults 2,000 x 500 kbp ~ 1e9 distinct 20-mers, 10 % of each genome shared with a sibling) through
the C ABI (kmat_db_build_device), replicated per GPU; R = 10 M Illumina-like 150 bp reads per GPU.

A step = one pass of the hot path (encode -> probe -> candidate sets -> scoring/LCA) over the R reads.
  value : whole-job reads/s with the reads already resident in HBM (kmat_label_batch_device), CUDA events on
          the launching stream, K steps after W warm-ups, barrier + synchronize on both sides, max over ranks
  e2e   : the same metric through kmat_label_batch with HOST buffers (H2D of the reads and D2H of the results
          inside the timed region)
  roofline : the encode+probe kernel (the table-bound one): algorithmic table bytes of SURVEY.md 8(d)
          (1 sector for a prefix miss, 2 for any other lookup, + ceil((2+2n)/32) per fetched list) / its
          CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs and the random-gather rate measured here
  cpu_baseline : the UNMODIFIED reference read_label (oracle/_ref, all host threads) on a bounded sample of the
          same workload (a sub-table of the first genomes, reads drawn from them); rank 0, N = 1 only
`--impl reference` prints the reference arm's line instead (same metric/config, CPU only).
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "reads_per_s"
UNIT = "reads/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="kmat", choices=["kmat", "reference"])
    ap.add_argument("--workload", default="C2", choices=["C2", "C3", "C4"],
                    help="BASELINE.json configs[1] (default: 2000 genomes, ~0.96 G k-mers, table replicated), configs[2] (C3: marker-scale, 3550 genomes = "
                         "1.7 G k-mers, run-time pruning -g 10, denser first level so that one GPU holds it) or configs[3] (C4: 5200 genomes = 2.5 G k-mers, a "
                         "~185 GB table at the default density: sharded over the ranks, query k-mers exchanged over NCCL)")
    ap.add_argument("--genomes", type=int, default=None)
    ap.add_argument("--genome-len", type=int, default=500000)
    ap.add_argument("--reads", type=int, default=10_000_000)
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--cpu-genomes", type=int, default=256,
                    help="genomes in the CPU-baseline sub-table (256 -> 1.2e8 k-mers, a 2 GB reference DB: far outside the host's L3)")
    ap.add_argument("--no-parity", action="store_true", help="skip the untimed bench-scale parity leg (GPU binary vs the unmodified reference)")
    ap.add_argument("--cpu-reads", type=int, default=200_000, help="reads in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--table-mode", default="replicated", choices=["replicated", "sharded", "direct"],
                    help="replicated: every GPU holds the whole table (configs[1]); sharded: table partitioned by k-mer hash over "
                         "the ranks, query k-mers exchanged with an NCCL all-to-all (configs[3] layout on the configs[1] table)")
    ap.add_argument("--round-reads", type=int, default=1 << 20, help="reads per exchange round in sharded mode")
    ap.add_argument("--exchange-slots", type=int, default=2, choices=[1, 2],
                    help="sharded mode: 2 = two contexts per rank, the encode of round i+1 and the finish of round i overlap the exchanges")
    ap.add_argument("--exchange", default="nccl", choices=["nccl", "torch"],
                    help="sharded mode: nccl = the exchange inside libkmat (kmat_shard_label_device: ncclSend/ncclRecv groups driven from C++); "
                         "torch = the Python driver over torch.distributed all_to_all_single (two-slot pipeline)")
    ap.add_argument("--pipeline", type=int, default=1, help="sub-batches per pass (kmat_ctx_set_pipeline); -1 automatic, 1 serial")
    a = ap.parse_args()
    if a.genomes is None:
        a.genomes = {"C2": 2000, "C3": 3550, "C4": 5200}[a.workload]
    a.prune = 10 if a.workload == "C3" else 0
    if a.workload == "C3":
        os.environ.setdefault("KMAT_LINE_DENSITY", "4")          # 2^29 lines = 69 GB instead of 137 GB: the replicated table + its build temporaries fit one GPU
    if a.workload == "C4" and a.table_mode == "replicated" and not os.environ.get("KMAT_LINE_DENSITY"):
        a.table_mode = "sharded"                                  # with an explicit density (8: a 69 GB first level) one GPU can hold C4 as well
    return a


def workload_name(a):
    if a.workload != "C2":
        return (f"{a.workload}: {a.genomes} random genomes x {a.genome_len} bp (10% sibling-shared, 2% mutated), k=20, 16-bit ids"
                f"{', lists pruned at run time to <= 10 taxids (-g 10 -m)' if a.prune else ''}; {a.reads} x {a.read_len} bp reads/GPU (90% genomic, 10% novel, 0.1%-2% subst, 0.05% N)")
    return (f"C2-replicated: {a.genomes} random genomes x {a.genome_len} bp (10% sibling-shared, 2% mutated), k=20, "
            f"16-bit ids; {a.reads} x {a.read_len} bp reads/GPU (90% genomic, 10% novel, 0.1%-2% subst, 0.05% N)")


class ClockSampler:
    """nvidia-smi sampled DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([x.strip() for x in ln.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, streaming copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# -------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the unmodified reference read_label on a bounded sample
# -------------------------------------------------------------------------------------------------
def _sample_cache_dir(a):
    """The CPU sample (reference DB, reads, taxonomy files) is built once per box and shared by the two arms the driver
    runs back to back (`--impl reference`, then the kmat arm): /dev/shm/kmat_cpu_sample_<key>."""
    key = f"g{a.genomes}_l{a.genome_len}_cg{a.cpu_genomes}_cr{a.cpu_reads}_rl{a.read_len}"
    base = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    return os.path.join(base, "kmat_cpu_sample_" + key)


def build_cpu_sample(a, workdir, device):
    """Sub-table of the first `cpu_genomes` genomes written as a tax_histo file and built into a reference DB by
    oracle/_ref/make_db_table; reads drawn from those genomes with the workload's read model.  Cached (see above)."""
    import numpy as np
    import torch
    from lmat_b200 import fixtures as fx
    from lmat_b200 import synth
    from oracle import refchain as rc
    cdir = _sample_cache_dir(a)
    meta_p = os.path.join(cdir, "meta.json")
    if os.path.exists(meta_p):
        meta = json.load(open(meta_p))
        if all(os.path.exists(p) for p in [meta["db"] or meta_p, meta["fasta"]] + list(meta["paths"].values())):
            meta["cached"] = True
            return meta
    shutil.rmtree(cdir, ignore_errors=True)
    os.makedirs(cdir, exist_ok=True)
    t0 = time.time()
    G = min(a.cpu_genomes, a.genomes)
    tax, m16, anc_tid, anc_sid = synth.make_taxonomy_c2(20240, a.genomes)
    paths = fx.write_taxonomy_files(tax, cdir)
    null_lst = synth.write_null_models_for(tax, cdir)
    dev = device
    codes = synth.make_genomes_gpu(20240, tax, a.genomes, a.genome_len, dev)[:G].contiguous() if dev != "cpu" else \
        synth.make_genomes_gpu(20240, tax, G, a.genome_len, dev)
    tbl = synth.build_table_gpu(codes, anc_sid[:G])
    sid2tid = np.zeros(65536, dtype=np.uint32)
    for t, s in m16.items():
        sid2tid[s] = t
    kmers, offs, tids = synth.table_logical(tbl, sid2tid)
    n_kmers, n_tids = int(kmers.numel()), int(tids.numel())
    th = os.path.join(cdir, "cpu_sample.th.bin")
    synth.write_tax_histo_torch(th, 20, kmers, offs, tids)
    del kmers, offs, tids, tbl
    size_gib = int(2 + (n_kmers * 8 + n_tids * 8) / (1 << 30) * 1.3) + 1
    db = None
    kind = "port"
    if rc.have_ref("make_db_table") and rc.have_ref("read_label"):
        db = rc.make_db_table([th], os.path.join(cdir, "cpu_sample.db"), 20, size_gib, cdir, map16=paths["map16"])
        kind = "reference"
        os.unlink(th)
    reads = synth.make_reads_gpu(20241, codes, a.cpu_reads, a.read_len).cpu().numpy()
    fa = os.path.join(cdir, "cpu_sample.fa")
    hdr = np.char.add(np.char.add(">r", np.arange(reads.shape[0]).astype(str)), "\n").astype("S")
    with open(fa, "wb") as f:
        for i in range(reads.shape[0]):
            f.write(hdr[i])
            f.write(reads[i].tobytes())
            f.write(b"\n")
    meta = dict(db=db, kind=kind, paths=paths, null_lst=null_lst, fasta=fa, n_reads=int(reads.shape[0]), n_kmers=n_kmers, genomes=G,
                th=None if db else th, dir=cdir, build_s=round(time.time() - t0, 1), cached=False)
    with open(meta_p + ".tmp", "w") as f:
        json.dump(meta, f)
    os.replace(meta_p + ".tmp", meta_p)
    return meta


def run_cpu_sample(sample, workdir, threads):
    """One timed run of the reference on the sample.  Returns (reads/s from its own 'Total query time', wall s)."""
    from oracle import refchain as rc
    if sample["kind"] == "reference":
        P = sample["paths"]
        t0 = time.time()
        out, qt = rc.read_label(sample["db"], sample["fasta"], os.path.join(workdir, "cpu_rl_"), P["depth"], P["tree"], threads=threads,
                                map16=P["map16"], rank=P["rank"], names=P["names"], null_lst=sample["null_lst"], lmat_dir=sample["dir"],
                                min_kmer=30, hbias=0, sdiff=1.0, prn_all=True)
        wall = time.time() - t0
        return sample["n_reads"] / qt, wall
    # port (only where the reference binaries are missing): the plain-C oracle restatement, single thread
    import numpy as np
    from lmat_b200 import api
    from oracle import oracle_py as op
    t = api.Table.build([sample["th"]], kmer_len=20, map16=sample["paths"]["map16"])
    kmers, offs, ids = t.arrays()
    sd = op.SortedDbArrays(kmers, offs, ids)
    orc = op.Oracle(cdb=sd.cdb(), keep=sd)
    P = sample["paths"]
    orc.set_opts(min_kmer=30, hbias=0.0, sdiff=1.0, prn_all=1)
    orc.load_files(tree=P["tree"], depth=P["depth"], rank=P["rank"], map16=P["map16"], null_lst=sample["null_lst"], lmat_dir=sample["dir"])
    hdrs, seqs = op.read_fasta_like_reference(sample["fasta"])
    seqs = seqs[:20000]
    t0 = time.time()
    orc.label(seqs)
    dt = time.time() - t0
    return len(seqs) / dt, dt


_REC_SPLIT = None


def out_records(ofbase, n_files):
    """The per-read records of <ofbase><t>.out as a sorted list.  A silent NoMatch prints `hdr\tread\t` with no newline
    (read_label.cpp:727-733, 1248-1253), so the next record of the same thread runs on in the same line: which records run
    together depends on the thread partition, not on the labels -- lines are cut in front of every `r<N>\t<bases>\t`."""
    import re
    global _REC_SPLIT
    if _REC_SPLIT is None:
        _REC_SPLIT = re.compile(r"(?<=\t)(?=r[0-9]+\t[A-Za-z]+\t)")
    recs = []
    for t in range(n_files):
        p = f"{ofbase}{t}.out"
        if os.path.exists(p):
            for ln in open(p, encoding="latin-1").read().split("\n"):
                if ln:
                    recs.extend(x for x in _REC_SPLIT.split(ln) if x)
    recs.sort()
    return recs


def parity_at_scale(sample, workdir, threads):
    """GPU labels vs the UNMODIFIED reference at bench scale: the drop-in binary (lmat_b200/bin/read_label -> libkmat) and the
    reference read_label run on the same DB file, the same reads file and the same flags (bin/run_rl.sh:243); their per-read
    records are compared as sorted multisets (the reference's own -t 1 vs -t N runs agree only in that sense)."""
    from lmat_b200 import build
    if sample["kind"] != "reference":
        return {"reads": 0, "lines_equal": None, "note": "reference binaries not available"}
    exe = build.BIN
    P = sample["paths"]
    ref_base = os.path.join(workdir, "cpu_rl_")
    if not os.path.exists(ref_base + "0.out"):
        run_cpu_sample(sample, workdir, threads)
    ofb = os.path.join(workdir, "gpu_rl_")
    n_out = 4
    cmd = [exe, "-f", P["map16"], "-u", P["names"], "-w", P["rank"], "-x", "0", "-j", "30", "-l", "0", "-b", "1.0", "-n", sample["null_lst"],
           "-e", P["depth"], "-p", "-t", str(n_out), "-i", sample["fasta"], "-d", sample["db"], "-c", P["tree"], "-o", ofb]
    env = dict(os.environ)
    env["LMAT_DIR"] = sample["dir"]
    env.setdefault("KMAT_DEVICES", os.environ.get("LOCAL_RANK", "0"))
    t0 = time.time()
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
    if p.returncode != 0:
        return {"reads": sample["n_reads"], "lines_equal": False, "note": ("read_label (kmat) rc=%d: " % p.returncode + p.stderr[-300:])}
    ref = out_records(ref_base, threads)
    got = out_records(ofb, n_out)
    n_diff = 0
    first = None
    if ref != got:
        import collections
        cr, cg = collections.Counter(ref), collections.Counter(got)
        d = (cr - cg) + (cg - cr)
        n_diff = sum(d.values())
        first = next(iter(d))[:200]
    return {"reads": sample["n_reads"], "records_ref": len(ref), "records_kmat": len(got), "lines_equal": ref == got, "n_diff": n_diff,
            "first_diff": first, "db_kmers": sample["n_kmers"], "wall_s": round(time.time() - t0, 1),
            "how": "lmat_b200/bin/read_label vs oracle/_ref/read_label: same DB file, reads file and flags; .out records as sorted multisets"}


def sharded_bench(a, ctx, db, reads, world, rank, local, dev, n_kmers, n_lists, setup_s, launches0, tstream, ctx2=None):
    """DB-sharded arm: every rank holds 1/world of the table and its own reads; per round of --round-reads reads the
    first-occurrence k-mers go to their owner ranks (all-to-all), hit words and list records come back (all-to-all)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from lmat_b200 import api, sharded
    n, L = reads.shape
    stream = tstream.cuda_stream
    ex = sharded.DistExchange(dev) if world > 1 else sharded.LocalExchange(sharded.LocalGroup(1), 0, sync=torch.cuda.synchronize)
    if ctx2 is None:
        lab = sharded.ShardedLabeler(sharded.CudaPhases(ctx, dev, world, stream), ex, round_reads=a.round_reads)
    else:
        lab = sharded.ShardedLabeler(sharded.CudaPhases(ctx, dev, world, torch_stream=torch.cuda.Stream()), ex, round_reads=a.round_reads,
                                     phases2=sharded.CudaPhases(ctx2, dev, world, torch_stream=torch.cuda.Stream()))
    comm = None
    if a.exchange == "nccl":
        # the exchange lives in libkmat: one NCCL communicator per rank; the 128-byte unique id travels over torch.distributed once
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(api.Comm.unique_id()), dtype=torch.uint8))
        if world > 1:
            dist.broadcast(uid, 0)
        comm = api.Comm(local, rank, world, uid.cpu().numpy().tobytes())
        h_offs_all = (np.arange(n + 1, dtype=np.uint64) * np.uint64(L))
        d_offs_all = torch.as_tensor(h_offs_all.astype(np.int64), device=dev)
    rr = min(a.round_reads, n)
    offs_full = (torch.arange(rr + 1, device=dev, dtype=torch.int64) * L).contiguous()       # chunk-local offsets, every round
    d_out = torch.empty(n * api.RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    rounds = [(r0, min(n, r0 + rr)) for r0 in range(0, n, rr)]

    def round_args(r0, r1):
        return (reads.data_ptr() + r0 * L, offs_full.data_ptr(), r1 - r0, (r1 - r0) * L, L, d_out.data_ptr() + r0 * api.RESULT_DTYPE.itemsize)

    comm_stats = [0, 0, 0, 0]

    def step():
        if comm is not None:
            st_ = comm.label_device(ctx, reads.data_ptr(), h_offs_all, d_offs_all.data_ptr(), n, d_out.data_ptr(), a.round_reads, stream)
            for i_ in range(4):
                comm_stats[i_] += st_[i_]
        else:
            lab.run(rounds, round_args)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ctx.set_stats(False)
    if ctx2 is not None:
        ctx2.set_stats(False)
    for _ in range(a.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    lab.lookups = lab.served = lab.payload_words = lab.rounds = 0
    comm_stats[:] = [0, 0, 0, 0]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(tstream)
    for _ in range(a.steps):
        step()
    e1.record(tstream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop()
    phase_ms = {}
    if comm is not None:
        lab.lookups, lab.served, lab.payload_words, lab.rounds = comm_stats
    else:
        # one more, untimed step with a CUDA event after every phase: where a round's time goes
        lab.timing = {}
        keep = (lab.lookups, lab.served, lab.payload_words, lab.rounds)
        ph2_keep, lab.ph2 = lab.ph2, None              # the phase split is taken on the serial schedule
        step()
        lab.ph2 = ph2_keep
        torch.cuda.synchronize()
        phase_ms = {k: round(v, 2) for k, v in lab.timing.items()}
        lab.timing = None
        lab.lookups, lab.served, lab.payload_words, lab.rounds = keep
    res = d_out.cpu().numpy().view(api.RESULT_DTYPE)
    labels_checksum = int(((res["status"].astype(np.int64) * 1000003 + res["tid"].astype(np.int64) * 7919 + res["score"].view(np.int32).astype(np.int64)) & 0xFFFFFFFF).sum())
    errs = int((res["status"] == 6).sum())
    labeled = int((res["status"] == 5).sum())
    if errs:
        bad = np.nonzero(res["status"] == 6)[0]
        print(f"[rank {rank}] {errs} reads in error: err codes {np.unique(res['err'][bad], return_counts=True)}, first reads {bad[:8].tolist()}, "
              f"valid_kmers {res['valid_kmers'][bad[:8]].tolist()}", file=sys.stderr, flush=True)
    t = torch.tensor([ms_total, 0.0], device=dev, dtype=torch.float64)
    tot = torch.tensor([lab.lookups, lab.payload_words, errs, labeled], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_step = float(t[0].item()) / a.steps
    lookups_step = int(tot[0].item()) / a.steps
    if rank == 0:
        hbm_peak, peak_src = measured_peaks()
        value = world * n / (ms_step * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64/f32", "data": "synthetic",
            "config": {"workload": workload_name(a).replace("C2-replicated", "C2 table, DB-sharded"), "db_kmers": int(n_kmers), "db_lists": n_lists,
                       "db_bytes_per_gpu": int(db.bytes), "db_kmers_this_shard": int(db.size), "reads_per_gpu": n, "read_len": L, "k": 20,
                       "options": "run_rl.sh:243 (-j 30 -l 0 -b 1 -p, null models on)", "round_reads": a.round_reads,
                       "l2_policy": "inputs larger than L2; no flush needed",
                       "parallelism": f"table sharded x{world} by kmat_shard_of (k-mer hash), reads stay home; per round: all-to-all of query k-mers "
                                      f"(8 B), all-to-all of hit words (4 B) + list records back ({'ncclSend/ncclRecv groups inside libkmat (kmat_shard_label_device)' if comm is not None else 'NCCL, torch.distributed all_to_all_single'})"},
            "kmer_lookups_per_s": lookups_step / (ms_step * 1e-3), "lookups_per_read": lookups_step / (world * n),
            "exchange_bytes_per_step": int(lookups_step * 12 + int(tot[1].item()) / a.steps * 4), "reads_error": int(tot[2].item()),
            "reads_labeled": int(tot[3].item()), "phase_ms_rank0": phase_ms, "exchange": a.exchange, "exchange_slots": (2 if ctx2 is not None else 1) if comm is None else 1, "labels_checksum_rank0": labels_checksum,
            "roofline": {"bound": "hbm", "kernel": "km_shard_probe_kernel", "achieved": None, "peak": hbm_peak, "unit": "GB/s", "frac": None, "traffic": None,
                         "peak_source": peak_src, "note": "per-kernel split not measured in sharded mode; see the replicated line"},
            "e2e": None, "gpu_launches": int(api.lib().kmat_launch_count() - launches0), "clocks": clocks, "setup_s": setup_s,
        }
        print(json.dumps(line), flush=True)


def traffic_from_profile(a):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the encode+probe kernel, from the committed ncu
    pass of this same workload (profiles/traffic.json, written by tools/ncu_traffic.py); None for other workloads."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if t["reads"] == a.reads and t["genomes"] == a.genomes and t["read_len"] == a.read_len and t["genome_len"] == a.genome_len:
            return t["dram_bytes_per_launch"]
    except Exception:
        pass
    return None


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    workdir = tempfile.mkdtemp(prefix="kmat_ref_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        dev = "cuda:0" if torch.cuda.is_available() else "cpu"
        if dev == "cpu":
            a.cpu_genomes = min(a.cpu_genomes, 8)
        sample = build_cpu_sample(a, workdir, dev)
        threads = os.cpu_count() or 1
        vals, walls = [], []
        for i in range(a.warmup + a.steps):
            v, w = run_cpu_sample(sample, workdir, threads)
            if i >= a.warmup:
                vals.append(v)
                walls.append(w)
        value = sum(vals) / len(vals)
        cores = threads if sample["kind"] == "reference" else 1
        desc = (f"{sample['n_reads']} x {a.read_len} bp reads vs a {sample['n_kmers']}-k-mer sub-table (first {sample['genomes']} of "
                f"{a.genomes} genomes) per step; reference read_label -t {cores}, its own 'Total query time'")
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": 1e3 * sample["n_reads"] / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u64/f32", "data": "synthetic", "config": {"workload": workload_name(a), "sample": desc},
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": sample["kind"], "sample": desc},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "kmer_lookups_per_s": value * (a.read_len - 19), "wall_s_per_step": sum(walls) / len(walls), "gpu_launches": 0}
        print(json.dumps(line), flush=True)
    finally:
        shutil.rmtree(workdir, ignore_errors=True)


# -------------------------------------------------------------------------------------------------
def main():
    a = parse_args()
    if a.impl == "reference":
        reference_arm(a)
        return
    import numpy as np
    import torch
    import torch.distributed as dist
    from lmat_b200 import api
    from lmat_b200 import fixtures as fx
    from lmat_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- libkmat has no CPU fallback")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    launches0 = api.lib().kmat_launch_count()

    workdir = tempfile.mkdtemp(prefix=f"kmat_bench_{rank}_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        t_setup = time.time()
        tax, m16, anc_tid, anc_sid = synth.make_taxonomy_c2(20240, a.genomes)
        paths = fx.write_taxonomy_files(tax, workdir)
        null_lst = synth.write_null_models_for(tax, workdir)
        codes = synth.make_genomes_gpu(20240, tax, a.genomes, a.genome_len, dev)
        tbl = synth.build_table_gpu(codes, anc_sid)
        n_kmers = tbl.n
        sharded_mode = a.table_mode == "sharded"
        direct_mode = a.table_mode == "direct"        # table sharded, probes go to the owner GPU's memory over NVLink (no rounds)
        split = sharded_mode or direct_mode
        del codes
        torch.cuda.empty_cache()
        db = synth.upload_table(tbl, local, shard_index=rank if split else 0, shard_count=world if split else 1)
        n_lists = tbl.n_lists
        del tbl
        torch.cuda.empty_cache()
        inputs = api.Inputs(tree=paths["tree"], depth=paths["depth"], rank=paths["rank"], map16=paths["map16"], null_lst=null_lst, lmat_dir=workdir,
                            numrank=paths["numrank"] if a.prune else None)
        # options of bin/run_rl.sh:243: -j 30 -l 0 -b 1.0 -x 0 -p, null models on (C3 adds -g 10 -m numeric_ranks)
        kw_opts = dict(min_kmer=30, hbias=0.0, sdiff=1.0, min_score=0.0, want_lineage=0)
        if a.prune:
            kw_opts["max_count"] = a.prune
        ctx = api.Ctx(db, inputs, api.default_opts(**kw_opts))
        if direct_mode and world > 1:
            from lmat_b200 import sharded
            sharded.attach_peers(ctx, device=dev)
        codes = synth.make_genomes_gpu(20240, tax, a.genomes, a.genome_len, dev)          # again: they were freed to make room for the table build
        reads = synth.make_reads_gpu(20241 + rank, codes, a.reads, a.read_len)           # weak scaling: every rank its own R reads
        del codes
        torch.cuda.empty_cache()
        n, L = reads.shape
        d_offs = (torch.arange(n + 1, device=dev, dtype=torch.int64) * L).contiguous()
        total = n * L
        setup_s = time.time() - t_setup

        # a non-default torch stream: its handle is passed to the library, so the kernels and the CUDA events
        # below are on the same stream (handle 0 would mean "the ctx's own stream" to the C ABI)
        tstream = torch.cuda.Stream()
        torch.cuda.synchronize()
        torch.cuda.set_stream(tstream)
        stream = tstream.cuda_stream

        def step():
            ctx.label_device(reads.data_ptr(), d_offs.data_ptr(), n, total, L, None, stream)

        if sharded_mode:
            ctx2 = api.Ctx(db, inputs, api.default_opts(**kw_opts)) if (a.exchange_slots == 2 and a.exchange == "torch") else None
            sharded_bench(a, ctx, db, reads, world, rank, local, dev, n_kmers, n_lists, setup_s, launches0, tstream, ctx2)
            return

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        # statistics counters (instrumentation) are off in every timed region and collected by one extra step below
        ctx.set_stats(False)
        ctx.set_pipeline(a.pipeline)
        for _ in range(a.warmup):
            step()
        barrier()
        sampler = ClockSampler(local)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        probe_ms, cand_ms, score_ms = [], [], []
        barrier()
        e0.record()
        for _ in range(a.steps):
            step()
        e1.record()
        barrier()
        ms_total = e0.elapsed_time(e1)
        clocks = sampler.stop()
        # per-kernel durations (CUDA events on the same stream): extra untimed steps with the kernels run one after the
        # other (the timed steps above overlap them on two streams, see kmat_ctx_set_pipeline)
        ctx.set_pipeline(1)
        for _ in range(3):
            step()
            p, c_, s = ctx.kernel_ms()
            probe_ms.append(p)
            cand_ms.append(c_)
            score_ms.append(s)
        torch.cuda.synchronize()
        ctx.set_pipeline(a.pipeline)
        ctx.set_stats(True)
        step()
        torch.cuda.synchronize()
        # order-independent checksum of the labels of rank 0's reads (status, taxid, score bits): the replicated, the
        # exchange and the direct arm label the same seeded reads against the same table, so their checksums must agree
        labels_checksum = None
        try:
            from lmat_b200 import sharded as _sh
            optr, _, _ = ctx.device_results()
            rv = _sh._wrap(optr, n * api.RESULT_DTYPE.itemsize, torch.uint8, dev).view(torch.int32).view(n, api.RESULT_DTYPE.itemsize // 4).to(torch.int64)
            labels_checksum = int(((rv[:, 0] * 1000003 + rv[:, 6] * 7919 + rv[:, 7]) & 0xFFFFFFFF).sum().item())
            if os.environ.get("KMAT_BENCH_DUMP") and rank == 0:        # debugging aid: the raw results of rank 0's reads
                np.save(os.environ["KMAT_BENCH_DUMP"], rv.to(torch.int32).cpu().numpy())
        except Exception as ex:
            labels_checksum = repr(ex)[:100]
        st = ctx.stats()
        ctx.set_stats(False)
        ctx.sync()          # raises if a pass overflowed its candidate buffer (reads left in error): the line would be invalid
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step = float(t.item()) / a.steps
        value = world * n / (ms_step * 1e-3)
        lookups_per_read = st.lookups / n

        # ---- e2e through the host-buffer API: kmat_label_batch_packed (the compact interface: 2-bit packed reads in, 32-byte
        #      results + the candidate list out), every host buffer page-locked, copies inside the timed region.  The ASCII
        #      interface (kmat_label_batch: 150 B in, 64 B + lists out per read) is timed next to it as e2e_ascii.
        e2e = e2e_ascii = None
        if not a.no_e2e:
            import ctypes as C
            L_ = api.lib()
            h_reads = torch.empty((n, L), dtype=torch.uint8, pin_memory=True)
            h_reads.copy_(reads)
            t_offs = torch.empty(n + 1, dtype=torch.int64, pin_memory=True)
            t_offs.copy_(torch.arange(n + 1, dtype=torch.int64) * L)
            h_offs = t_offs.numpy().view(np.uint64)
            n_c = C.c_uint64()
            steps_e = max(1, min(a.steps, 3))

            def timed(fn):
                fn()
                barrier()
                t0 = time.perf_counter()
                for _ in range(steps_e):
                    fn()
                barrier()
                dt = (time.perf_counter() - t0) / steps_e
                tt = torch.tensor([dt], device=dev, dtype=torch.float64)
                if world > 1:
                    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                return float(tt.item())

            # compact interface; the reads are packed once, outside the timed region (the reader's job in the file pipeline)
            n_words = int(L_.kmat_pack_words(total))
            t_codes = torch.empty(n_words, dtype=torch.int32, pin_memory=True)
            t_inv = torch.empty(max(1, total // 64), dtype=torch.int64, pin_memory=True)
            n_inv = C.c_uint64()
            rc = L_.kmat_pack_reads(h_reads.data_ptr(), total, min(16, os.cpu_count() or 1), t_codes.data_ptr(), t_inv.data_ptr(), t_inv.numel(), C.byref(n_inv))
            if rc < 0:
                raise api.KmatError(rc, L_.kmat_last_error().decode())
            t_res32 = torch.empty(n * api.RESULT32_DTYPE.itemsize, dtype=torch.uint8, pin_memory=True)
            pairs_per_read = 24 + L // 40                      # ~10 candidates per 150-base read, ~100 per 10 kbp read
            t_list = torch.empty(max(1, pairs_per_read * n) * api.PAIR_DTYPE.itemsize, dtype=torch.uint8, pin_memory=True)
            list_cap = t_list.numel() // api.PAIR_DTYPE.itemsize
            n_w = C.c_uint64()

            def packed_rl_step():
                # the pair lists in run-length form (a taxid word, its score once per run of equal scores): fewer bytes on the way out
                rc = L_.kmat_label_batch_packed_rl(ctx.h, t_codes.data_ptr(), t_inv.data_ptr(), n_inv.value, h_offs.ctypes.data, n, t_res32.data_ptr(),
                                                   t_list.data_ptr(), 2 * list_cap, C.byref(n_w))
                if rc < 0:
                    raise api.KmatError(rc, L_.kmat_last_error().decode())

            def packed_step():
                rc = L_.kmat_label_batch_packed(ctx.h, t_codes.data_ptr(), t_inv.data_ptr(), n_inv.value, h_offs.ctypes.data, n, t_res32.data_ptr(),
                                                t_list.data_ptr(), list_cap, C.byref(n_c))
                if rc < 0:
                    raise api.KmatError(rc, L_.kmat_last_error().decode())
            dt = timed(packed_rl_step)
            e2e = {"value": world * n / dt, "unit": UNIT, "h2d_bytes_per_step": int(4 * n_words + 8 * n_inv.value + 8 * (n + 1)),
                   "d2h_bytes_per_step": int(n * api.RESULT32_DTYPE.itemsize + n_w.value * 4),
                   "api": "kmat_label_batch_packed_rl: 2-bit packed reads + invalid-base positions + offsets in (packed by kmat_pack_reads before the timed region), "
                          "32-byte results + the rank_label pairs in run-length form out (kmat_list_decode rebuilds them); pinned host buffers"}
            r32 = t_res32.numpy().view(api.RESULT32_DTYPE)
            e2e_checksum = int((((r32["flags"] & 7).astype(np.int64) * 1000003 + r32["tid"].astype(np.int64) * 7919 + r32["score"].view(np.int32).astype(np.int64)) & 0xFFFFFFFF).sum())
            # the plain compact interface (8-byte pairs) next to it; tests/test_compact_io.py checks that the words decode to the same pairs
            dt_plain = timed(packed_step)
            e2e["plain_pairs"] = {"value": world * n / dt_plain, "d2h_bytes_per_step": int(n * api.RESULT32_DTYPE.itemsize + n_c.value * 8),
                                  "api": "kmat_label_batch_packed (8-byte pairs out)"}
            del t_codes, t_inv, t_res32, t_list
            # ASCII interface
            t_res = torch.empty(n * api.RESULT_DTYPE.itemsize, dtype=torch.uint8, pin_memory=True)
            t_cands = torch.empty(max(1, pairs_per_read * n) * api.PAIR_DTYPE.itemsize, dtype=torch.uint8, pin_memory=True)
            res = t_res.numpy().view(api.RESULT_DTYPE)
            cands = t_cands.numpy().view(api.PAIR_DTYPE)

            def ascii_step():
                rc = L_.kmat_label_batch(ctx.h, h_reads.data_ptr(), h_offs.ctypes.data, n, res.ctypes.data, cands.ctypes.data, len(cands),
                                         C.byref(n_c), None, 0, None)
                if rc < 0:
                    raise api.KmatError(rc, L_.kmat_last_error().decode())
            dt = timed(ascii_step)
            e2e_ascii = {"value": world * n / dt, "unit": UNIT, "h2d_bytes_per_step": int(total + 8 * (n + 1)),
                         "d2h_bytes_per_step": int(n * api.RESULT_DTYPE.itemsize + n_c.value * 8), "api": "kmat_label_batch: ASCII reads in, 64-byte results + pairs out"}
            ascii_checksum = int(((res["status"].astype(np.int64) * 1000003 + res["tid"].astype(np.int64) * 7919 + res["score"].view(np.int32).astype(np.int64)) & 0xFFFFFFFF).sum())
            e2e["labels_checksum"] = e2e_checksum
            e2e_ascii["labels_checksum"] = ascii_checksum
            del h_reads, t_offs, t_res, t_cands

        if rank == 0:
            hbm_peak, peak_src = measured_peaks()
            gather_gps, gather_gbps = api.gather_bench(local, 16 << 30, 8, 1 << 29, 10)
            pm = sorted(probe_ms)[len(probe_ms) // 2]
            sm_ = sorted(score_ms)[len(score_ms) // 2]
            cm_ = sorted(cand_ms)[len(cand_ms) // 2]
            achieved = st.algorithmic_bytes / (pm * 1e-3) / 1e9
            errs = st.reads_error
            line = {
                "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64/f32", "data": "synthetic",
                "config": {"workload": workload_name(a).replace("C2-replicated", "C2 table, DB-sharded (direct peer reads)") if direct_mode else workload_name(a), "db_kmers": int(n_kmers), "db_lists": n_lists, "db_bytes": int(db.bytes),
                           "reads_per_gpu": n, "read_len": L, "k": 20, "options": "run_rl.sh:243 (-j 30 -l 0 -b 1 -p, null models on)",
                           "l2_policy": "inputs larger than L2 (reads 1.5 GB, table >> 126 MB); no flush needed",
                           "parallelism": (f"table sharded x{world} by kmat_shard_of (k-mer hash), {int(db.bytes)} B per GPU; every bucket gather goes to the owner "
                                           f"GPU's memory (CUDA IPC peer mapping, NVLink reads); reads stay home, no exchange rounds, no collective")
                           if direct_mode else f"read-sharded x{world}, table replicated, no data-path collective"},
                "kmer_lookups_per_s": value * lookups_per_read, "lookups_per_read": lookups_per_read,
                "hit_rate": st.hits / max(1, st.lookups), "extra_buckets_per_lookup": st.probe_extra_buckets / max(1, st.lookups), "reads_error": int(errs), "labels_checksum_rank0": labels_checksum,
                "kernels_ms": {"encode_probe": pm, "candidates": cm_, "score": sm_, "how": "CUDA events around each kernel of one serial pass"},
                "pipeline_sub_batches": a.pipeline,
                "roofline": {"bound": "hbm", "kernel": "km_encode_probe_fast_kernel<5>" if L <= 160 else ("km_encode_probe_fast_kernel<8>" if L <= 256 else "km_encode_probe_kernel"), "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                             "frac": achieved / hbm_peak, "traffic": None if direct_mode else traffic_from_profile(a), "peak_source": peak_src,
                             "request_rate_peak_G": gather_gps / 1e9, "request_rate_frac": (st.lookups / (pm * 1e-3)) / gather_gps,
                             "request_rate_how": "probe-kernel lookups/s over the measured ceiling of independent random requests (uniform random 8-byte loads, one per "
                                                 "32-byte sector, 16 GiB, best of 10: kmat_gather_bench); a first-level lookup of the two-level table shares its request "
                                                 "with the neighbouring k-mers of the read (~0.41 L2 requests per lookup, profiles/), so this can exceed 1",
                             "algorithmic_bytes_per_lookup": st.algorithmic_bytes / max(1, st.lookups)},
                "e2e": e2e, "e2e_ascii": e2e_ascii, "gpu_launches": int(api.lib().kmat_launch_count() - launches0), "clocks": clocks, "setup_s": setup_s,
            }
            if world == 1 and not a.no_cpu_baseline:
                try:
                    sample = build_cpu_sample(a, workdir, dev)
                    threads = os.cpu_count() or 1
                    v, w = run_cpu_sample(sample, workdir, threads)
                    cores = threads if sample["kind"] == "reference" else 1
                    line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": sample["kind"],
                                            "sample": f"{sample['n_reads']} reads vs a {sample['n_kmers']}-k-mer sub-table (first {sample['genomes']} genomes); "
                                                      f"read_label -t {cores} 'Total query time'; wall {w:.1f} s"}
                    if not a.no_parity:
                        line["parity_at_scale"] = parity_at_scale(sample, workdir, threads)
                except Exception as ex:          # the baseline must never take the GPU number down with it
                    line.setdefault("cpu_baseline", {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(ex)[:200]})
                    if not a.no_parity:
                        line.setdefault("parity_at_scale", {"reads": 0, "lines_equal": None, "note": repr(ex)[:200]})
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.barrier()          # direct mode: peers read this rank's table until they are done
    finally:
        shutil.rmtree(workdir, ignore_errors=True)
        if world > 1:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
