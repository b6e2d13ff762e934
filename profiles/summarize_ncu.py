#!/usr/bin/env python
"""Turn the two ncu outputs of the recipe (gpurun_out/r2_ncu_tcr.ncu-rep from `--set full`, gpurun_out/r2_launches_tcr.csv
from `--metrics gpu__time_duration.sum`) into the committed summaries: profiles/r2_ncu_full_esn_recur_tcr.txt,
profiles/r2_launches_tcr.csv and the esn_recur_tcr entry of profiles/ncu_traffic.json (keyed by a hash of the kernel
sources, which bench.py checks before it reports roofline.traffic).  Runs here, without a GPU."""
import csv
import hashlib
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REP = os.path.join(ROOT, "gpurun_out", "r2_ncu_tcr.ncu-rep")
LAUNCHES = os.path.join(ROOT, "gpurun_out", "r2_launches_tcr.csv")
CMD = "python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-dropin --fit-pilots 0 --shared-pilots 0"
KEYS = ['Kernel Name', 'gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__grid_size', 'launch__block_size',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed']


def ncu(page):
    return subprocess.run(["ncu", "-i", REP, "--page", page, "--csv"], capture_output=True, text=True).stdout


rows = list(csv.reader(ncu("raw").splitlines()))
h, u, v = rows[0], rows[1], rows[-1]
d = dict(zip(h, v))
lines = [f"{k}: {d[k]} {u[h.index(k)]}" for k in KEYS if k in d]
cyc, rd, wr = float(d['sm__cycles_elapsed.max']), float(d['dram__bytes_read.sum']), float(d['dram__bytes_write.sum'])
unit = u[h.index('dram__bytes_read.sum')]
mult = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[unit]
rows = list(csv.reader(ncu("source").splitlines()))
hh, data = rows[1], rows[2:]
ix = {k: i for i, k in enumerate(hh)}
stalls = [k for k in hh if k.startswith('stall_') and 'Not Issued' not in k]
tot, S, top = {k: 0 for k in stalls}, 0, []
for r in data:
    try:
        n = int(r[ix['# Samples']])
    except (ValueError, IndexError):
        continue
    S += n
    for k in stalls:
        try:
            tot[k] += int(r[ix[k]])
        except ValueError:
            pass
    top.append((n, r[ix['Source']].strip()[:80]))
top.sort(key=lambda x: -x[0])
out = f"""ncu --set full --clock-control none --import-source on -k regex:esn_recur_tcr -s 4 -c 1
of: {CMD}
(one launch: 9472 frames x 522 steps, cfg3; B200, driver 580.159; final kernel of round 2; written by profiles/summarize_ncu.py)

""" + "\n".join(lines) + f"""

cycles per time step: {cyc / 1e6:.2f} M / 522 = {cyc / 522 / 1e3:.1f} K (in-kernel stamps: profiles/r2_tcr_timeline.txt, 24.4 K)
DRAM per launch: {rd:.1f} + {wr:.1f} {unit} = inputs in (316 MB) + outputs out (the transient rows are not written); the weight
image (1.2 MB) and the readout tables stay in L2
the CUDA-core readout sweep (512 FFMA + 48 LDS.128 + 31 SHFL per thread and step) shares the shared-memory pipe with the
MMA operand reads and the bulk-copy writes of the weight ring: bulk copies issued during the sweep land 1.5-4 K cycles
after the request instead of ~1 K (per-item trace in the timeline file)

warp stall samples ({S} total): """ + ", ".join(f"{k[6:]} {100 * v / S:.1f}%" for k, v in sorted(tot.items(), key=lambda x: -x[1])[:8]) + """
top sampled instructions:
""" + "\n".join(f"  {n:7d}  {s}" for n, s in top[:12]) + "\n"
open(os.path.join(ROOT, "profiles", "r2_ncu_full_esn_recur_tcr.txt"), "w").write(out)


def kh(files):
    x = hashlib.sha256()
    for f in files:
        x.update(open(os.path.join(ROOT, "esn-ofdm-mimo_b200", "csrc", f), "rb").read())
    return x.hexdigest()[:16]


tj = os.path.join(ROOT, "profiles", "ncu_traffic.json")
ent = [e for e in json.load(open(tj)) if e.get("kernel") != "esn_recur_tcr"]
ent.append({"kernel": "esn_recur_tcr", "frames_per_launch": 9472, "dram_bytes_read": int(rd * mult), "dram_bytes_write": int(wr * mult),
            "source_hash": kh(["recurrence_tcr.cu", "tc_common.cuh"]),
            "source": "profiles/r2_ncu_full_esn_recur_tcr.txt (ncu --set full, one launch of bench.py --steps 3 --warmup 3)"})
json.dump(ent, open(tj, "w"), indent=1)
rows = [r for r in csv.reader(open(LAUNCHES)) if len(r) > 10 and r[0].isdigit()]
o = [f"# ncu --metrics gpu__time_duration.sum --clock-control none of: {CMD}",
     "# (cold-cache, serialised launches: compare shares, not absolutes).  id, kernel, grid, block, time_us"]
tot = {}
for r in rows:
    short = re.sub(r"\(.*", "", r[4].replace("void ", "").replace("<unnamed>::", ""))[:80]
    t = float(r[-1].replace(",", "")) / 1e3
    o.append(f"{r[0]},{short},{r[8]},{r[7]},{t:.1f}")
    tot[short] = tot.get(short, 0) + t
open(os.path.join(ROOT, "profiles", "r2_launches_tcr.csv"), "w").write("\n".join(o) + "\n")
print("\n".join(lines[:8]))
for k, v in sorted(tot.items(), key=lambda x: -x[1])[:4]:
    print(f"{v / 1e3:9.2f} ms  {k}")
