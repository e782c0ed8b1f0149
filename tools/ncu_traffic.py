#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --csv` log of bench.py:
DRAM bytes of ONE launch of the encode+probe kernel at the bench workload (roofline.traffic).
usage: ncu_traffic.py <ncu.csv> <out.json> <reads> <genomes> <read_len> <genome_len>"""
import csv
import json
import sys

log, out = sys.argv[1], sys.argv[2]
reads, genomes, read_len, genome_len = map(int, sys.argv[3:7])
rows = [r for r in csv.reader(l for l in open(log) if l.startswith('"'))]
hdr = rows[0]
ik, im, iu, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
per = {}
for r in rows[1:]:
    if "km_encode_probe" not in r[ik]:
        continue
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(r[iu], 1)
    per.setdefault((r[0], r[ik]), {})[r[im]] = float(r[iv].replace(",", "")) * mult
launches = [v for v in per.values() if "dram__bytes_read.sum" in v and "dram__bytes_write.sum" in v]
tot = sorted(v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"] for v in launches)
res = {"kernel": list(per)[0][1], "launches_seen": len(launches), "dram_bytes_per_launch": tot[len(tot) // 2],
       "dram_read_bytes": sorted(v["dram__bytes_read.sum"] for v in launches)[len(launches) // 2],
       "l2_read_requests": sorted(v.get("lts__t_requests_srcunit_tex_op_read.sum", 0) for v in launches)[len(launches) // 2],
       "reads": reads, "genomes": genomes, "read_len": read_len, "genome_len": genome_len,
       "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, median over the captured launches"}
json.dump(res, open(out, "w"), indent=1)
print(res)
