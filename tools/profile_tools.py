"""Evidence helpers for profiles/ (run here, on the CPU box, on what gpurun brought back).

  python tools/profile_tools.py traffic gpurun_out/X.ncu-rep profiles/r2_traffic.json [--algo-bytes B]
      DRAM bytes and duration of the captured launch (ncu raw page) stamped with the hash of the kernel sources the
      capture was taken from; bench.py reports them as roofline.traffic / frac_dram only while that hash still matches.
  python tools/profile_tools.py sass profiles/r2_sass_summary.txt
      per-kernel counts of the SASS mnemonics that characterise the build (UBLKCP = TMA bulk copy, SYNCS = mbarrier,
      LDG.E.128 / STG.E.128 = 128-bit global accesses, DFMA/DMUL/DADD = FP64 pipe, no tensor-core instruction on this
      FP64 sparse path).
  python tools/profile_tools.py summary gpurun_out/X.ncu-rep profiles/r2_ncu_full_X.txt
      the headline metrics of a --set full capture as text.
"""
import csv
import hashlib
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "parelagmc_b200", "csrc")
KERNEL_SOURCES = ["program.cuh", "pmc_b200.cu", "rng.cuh", "detmath.h", "detmath_tables.h"]


def kernel_source_hash() -> str:
    h = hashlib.sha256()
    for f in KERNEL_SOURCES:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, units))


def num(x):
    return float(str(x).replace(",", ""))


def scale(v, unit):
    m = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0,
         "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}
    return num(v) * m.get(unit, 1.0)


def cmd_traffic(rep, dst, algo_bytes=None):
    rows, units = raw_metrics(rep)
    r = rows[0]
    rd = scale(r["dram__bytes_read.sum"], units["dram__bytes_read.sum"])
    wr = scale(r["dram__bytes_write.sum"], units["dram__bytes_write.sum"])
    dur = scale(r["gpu__time_duration.sum"], units["gpu__time_duration.sum"])
    d = {"kernel": r.get("Kernel Name", ""), "dram_bytes_read": rd, "dram_bytes_write": wr, "duration_s_under_ncu": dur,
         "source": os.path.basename(rep), "kernel_source_hash": kernel_source_hash(),
         "hash_of": KERNEL_SOURCES}
    if algo_bytes:
        d["algo_bytes"] = float(algo_bytes)
    json.dump(d, open(dst, "w"), indent=1)
    print(json.dumps(d))


def cmd_sass(dst):
    lib = os.path.join(ROOT, "parelagmc_b200", "lib", "libpmc_b200.so")
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    kern, counts = None, {}
    pats = {"UBLKCP": r"\bUBLKCP", "SYNCS": r"\bSYNCS", "LDG.E.128": r"\bLDG\.E\.128", "STG.E.128": r"\bSTG\.E\.128",
            "LDG (all)": r"\bLDG\b", "STG (all)": r"\bSTG\b", "LDS": r"\bLDS", "DFMA": r"\bDFMA", "DMUL": r"\bDMUL", "DADD": r"\bDADD",
            "MUFU.RCP64H": r"MUFU\.RCP64H", "BAR.SYNC": r"\bBAR\.SYNC", "UCGABAR": r"UCGABAR", "SHFL": r"\bSHFL",
            "HMMA/UTC*MMA (tensor)": r"\b(HMMA|DMMA|UTC\w*MMA)"}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            kern = m.group(1)
            counts[kern] = {k: 0 for k in pats}
            counts[kern]["instructions"] = 0
            continue
        if kern and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            counts[kern]["instructions"] += 1
            for k, p in pats.items():
                if re.search(p, line):
                    counts[kern][k] += 1
    with open(dst, "w") as f:
        f.write(f"cuobjdump -sass parelagmc_b200/lib/libpmc_b200.so (sm_100a), kernel source hash {kernel_source_hash()}\n")
        f.write("mnemonic counts per kernel (static instruction counts)\n\n")
        for k, c in counts.items():
            name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip() or k
            f.write(name[:150] + "\n   " + "  ".join(f"{a}={b}" for a, b in c.items() if b) + "\n")
    print("wrote", dst)


def cmd_summary(rep, dst):
    rows, units = raw_metrics(rep)
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
            "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
            "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
    with open(dst, "w") as f:
        f.write(f"ncu --set full --clock-control none: {os.path.basename(rep)}; kernel source hash {kernel_source_hash()}\n")
        for r in rows:
            f.write("\n" + r.get("Kernel Name", "")[:120] + "\n")
            for w in want:
                if w in r:
                    f.write(f"  {w:75s} {r[w]:>18s} {units.get(w, '')}\n")
    print("wrote", dst)


if __name__ == "__main__":
    c = sys.argv[1]
    if c == "traffic":
        ab = sys.argv[sys.argv.index("--algo-bytes") + 1] if "--algo-bytes" in sys.argv else None
        cmd_traffic(sys.argv[2], sys.argv[3], ab)
    elif c == "sass":
        cmd_sass(sys.argv[2])
    elif c == "summary":
        cmd_summary(sys.argv[2], sys.argv[3])
    elif c == "hash":
        print(kernel_source_hash())
