"""Condense ncu output into the small text tables kept under profiles/.

    ncu -i X.ncu-rep --page raw --csv > raw.csv ; python profiles/summarize_ncu.py full raw.csv > profiles/X.txt
    python profiles/summarize_ncu.py launches launches.csv [first_id last_id] > profiles/launches_X.txt

`full`  : one line per profiled kernel of an `ncu --set full` capture: duration, DRAM bytes read+written (the
          `traffic` figure of bench.py's roofline object), DRAM / L2 / SM throughput, tensor-pipe activity,
          occupancy, registers.
`launches`: the `--metrics gpu__time_duration.sum` launch list: per-launch time and share ('*' = kernels of
          libprism_b200.so).  Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.
"""
import csv
import re
import sys

OURS = ("gemm_kernel", "conv_", "tree_", "upd_", "store_", "qh_loss", "ens_loss", "ids_select", "greedy_select", "cos_basis", "pack_grads",
        "adam_clip", "loss_combine", "global_", "linear_", "conv3x3", "tc_gemm", "philox", "uniform_", "ln_", "layernorm",
        "iqn_", "peer_", "pb_", "nstep", "per_step", "tail_", "embed_")

SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-3, "us": 1.0, "ms": 1e3,
         "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6,
         "byte/s": 1.0, "Kbyte/s": 1e3, "Mbyte/s": 1e6, "Gbyte/s": 1e9, "Tbyte/s": 1e12}


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    return name


def is_ours(name):
    """Every kernel of libprism_b200.so lives in an anonymous namespace at file scope: ncu prints `<unnamed>::name`
    (framework kernels carry their own namespace first: at::, at::<unnamed>::, cutlass::, cublasLt::, ...)."""
    n = short(name)
    return n.startswith("<unnamed>::") or n.startswith("(anonymous namespace)::") or \
        any(n.startswith(p) for p in OURS)


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return float("nan")


def full(path):
    rows = list(csv.reader(open(path)))
    hdr, units, rows = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def get(row, key, scale=True):
        i = col.get(key)
        if i is None:
            return float("nan")
        v = num(row[i])
        return v * SCALE.get(units[i], 1.0) if scale else v

    print("%-4s %-58s %-14s %-10s %9s %11s %11s %7s %7s %7s %8s %6s %5s" % (
        "id", "kernel", "grid", "block", "time_us", "dram_rd_B", "dram_wr_B", "dram%", "L2%", "SM%", "tensor%", "occ%",
        "regs"))
    for r in rows:
        print("%-4s %-58s %-14s %-10s %9.2f %11.0f %11.0f %7.1f %7.1f %7.1f %8.2f %6.1f %5.0f" % (
            r[col["ID"]], short(r[col["Kernel Name"]])[:58], r[col["Grid Size"]].replace(" ", ""),
            r[col["Block Size"]].replace(" ", ""), get(r, "gpu__time_duration.sum"), get(r, "dram__bytes_read.sum"),
            get(r, "dram__bytes_write.sum"), get(r, "dram__throughput.avg.pct_of_peak_sustained_elapsed", False),
            get(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed", False),
            get(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed", False),
            get(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", False),
            get(r, "sm__warps_active.avg.pct_of_peak_sustained_active", False),
            get(r, "launch__registers_per_thread", False)))


def launches(path, lo=None, hi=None):
    lines = [ln for ln in open(path) if ln.startswith('"')]
    rows = list(csv.reader(lines))
    hdr, rows = rows[0], rows[1:]
    col = {h: i for i, h in enumerate(hdr)}
    rows = [r for r in rows if r[col["Metric Name"]] == "gpu__time_duration.sum"]
    if lo is not None:
        rows = [r for r in rows if lo <= int(r[col["ID"]]) <= hi]
    t = [num(r[col["Metric Value"]]) * SCALE.get(r[col["Metric Unit"]], 1.0) for r in rows]
    tot = sum(t)
    ours = 0.0
    for r, us in zip(rows, t):
        mine = is_ours(r[col["Kernel Name"]])
        ours += us if mine else 0.0
        print("%5s %9.2f us %5.1f%%  %-14s %-12s %s %s" % (
            r[col["ID"]], us, 100.0 * us / tot, r[col["Grid Size"]].replace(" ", ""),
            r[col["Block Size"]].replace(" ", ""), "*" if mine else " ", short(r[col["Kernel Name"]])[:100]))
    print("launches %d   sum %.1f us   ours(*) %.1f us = %.1f%% of the listed launches" % (
        len(rows), tot, ours, 100.0 * ours / max(tot, 1e-9)))


if __name__ == "__main__":
    if sys.argv[1] == "full":
        full(sys.argv[2])
    else:
        launches(sys.argv[2], *(int(a) for a in sys.argv[3:5]))
