#!/usr/bin/env python
"""Turn an .ncu-rep (read here, no GPU needed) into the small evidence files kept under profiles/:
   python tools/ncu_summary.py gpurun_out/r2_prof_k22.ncu-rep profiles/r2_fused_kernel_2_2
writes <out>_ncu_raw.csv (selected raw metrics) and <out>_phases.txt (time / instruction share and top stalls of
every region between block barriers / mbarrier waits, from the source page; needs -lineinfo / --import-source)."""
import collections
import csv
import io
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "launch__", "smsp__inst_executed.sum", "sm__issue_active", "sm__inst_executed_pipe_",
        "sm__pipe_", "smsp__average_warps_issue_stalled", "smsp__average_warp_latency", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared",
        "sm__warps_active", "sm__throughput", "smsp__sass_inst_executed_op_shared", "lts__t_bytes.sum", "sm__cycles_elapsed.max")


def page(rep, which):
    out = subprocess.run(["ncu", "-i", rep, "--page", which, "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                         text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, out = sys.argv[1], sys.argv[2]
    rows = page(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    with open(out + "_ncu_raw.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit", "value"])
        for h, u, v in zip(hdr, units, vals):
            if h in ("Kernel Name", "Device Name", "Block Size", "Grid Size") or (
                    any(h.startswith(k) for k in KEEP) and "per_second" not in h and ".max." not in h and ".min." not in h):
                w.writerow([h, u, v])
    src = page(rep, "source")
    if len(src) < 3:
        return
    h2 = src[1]
    ix = {h: i for i, h in enumerate(h2)}
    data = src[2:]
    stalls = [h for h in h2 if h.startswith("stall_")]
    tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
    toti = sum(int(r[ix["Instructions Executed"]] or 0) for r in data)
    lines = ["kernel: %s" % src[0][1],
             "regions between BAR.SYNC / mbarrier try_wait (SYNCS...TRYWAIT) / EXIT (SASS instruction index range):"]
    start = 0
    for i, r in enumerate(data):
        if ("BAR.SYNC" in r[ix["Source"]] or "TRYWAIT" in r[ix["Source"]] or "EXIT" in r[ix["Source"]]
                or i == len(data) - 1):
            a, b = start, i
            start = i + 1
            s = sum(int(x[ix["# Samples"]] or 0) for x in data[a:b + 1])
            if b - a < 3 or s < tot * 0.005:
                continue
            n = sum(int(x[ix["Instructions Executed"]] or 0) for x in data[a:b + 1])
            st = collections.Counter()
            for x in data[a:b + 1]:
                for hh in stalls:
                    if x[ix[hh]]:
                        st[hh] += int(x[ix[hh]])
            top = ", ".join("%s %.0f%%" % (k[6:], 100.0 * v / max(s, 1)) for k, v in st.most_common(5))
            lines.append("  %5d-%5d (%4d instr)  time %5.1f%%  executed instr %5.1f%%  | %s" % (a, b, b - a + 1,
                                                                                               100.0 * s / tot, 100.0 * n / toti, top))
    with open(out + "_phases.txt", "w") as f:
        f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
