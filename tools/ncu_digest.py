#!/usr/bin/env python
"""Digest of an `ncu --set full --import-source on` report for profiles/: per captured kernel the headline metrics
(duration, issue slots, pipe utilisation, stall reasons per issued instruction, DRAM bytes) and the source lines that
executed the most instructions / collected the most stall samples.

    python tools/ncu_digest.py gpurun_out/prof_x.ncu-rep [top_n] > profiles/rNN_ncu_x_digest.md
"""
import csv
import io
import subprocess
import sys


def ncu(args):
    return subprocess.run(["ncu"] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 14
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, rows = raw[0], raw[1], raw[2:]
    col = {k: i for i, k in enumerate(hdr)}
    want = [("gpu__time_duration.sum", "duration"), ("smsp__inst_executed.sum", "warp instructions"),
            ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
            ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
            ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %"),
            ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
            ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory wavefronts %"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
            ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
            ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
            ("launch__registers_per_thread", "registers / thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]
    stalls = [k for k in hdr if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")]
    print("# %s\n" % rep.split("/")[-1])
    for r in rows:
        print("## %s\n" % r[col["Kernel Name"]][:110])
        print("| metric | value |\n|---|---|")
        for key, name in want:
            if key in col:
                print("| %s | %s %s |" % (name, r[col[key]], units[col[key]]))
        st = sorted(((float(r[col[k]]), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for k in stalls
                     if r[col[k]] not in ("", "n/a")), reverse=True)
        print("| stall reasons (warps stalled per issued instruction) | %s |" % ", ".join("%s %.2f" % (n, v) for v, n in st[:6]))
        print()
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    cur, hd, kid, per = None, None, -1, {}
    for row in src:
        if len(row) >= 2 and row[0] == "Function Name":
            kid += 1
            continue
        if len(row) >= 2 and row[0] in ("File Path", "File Name"):
            cur = row[1].split("/")[-1]
            continue
        if row and row[0] == "Line No":
            hd = row
            continue
        if hd and len(row) == len(hd) and row[0].isdigit():
            d = {}
            for k, v in zip(hd, row):
                d.setdefault(k, v)
            per.setdefault(kid, []).append((cur, int(row[0]), row[1].strip(), int(d["Instructions Executed"]), int(d["# Samples"])))
    for kid in sorted(per):
        rs = per[kid]
        ti, ts = sum(x[3] for x in rs) or 1, sum(x[4] for x in rs) or 1
        if ti < 1000:
            continue
        print("### source lines, captured kernel %d (share of executed warp instructions / of stall samples)\n" % kid)
        print("| file:line | instr % | samples % | source |\n|---|---|---|---|")
        seen = set()
        for x in sorted(rs, key=lambda x: -x[3])[:top] + sorted(rs, key=lambda x: -x[4])[:top // 2]:
            if (x[0], x[1]) in seen:
                continue
            seen.add((x[0], x[1]))
            print("| %s:%d | %.1f | %.1f | `%s` |" % (x[0], x[1], 100.0 * x[3] / ti, 100.0 * x[4] / ts, x[2][:110].replace("|", "\\|")))
        print()


if __name__ == "__main__":
    main()
