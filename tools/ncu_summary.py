#!/usr/bin/env python3
"""Summarise an .ncu-rep (run here, no GPU needed): headline metrics, stall reasons, executed
instructions per barrier-delimited segment and the most-sampled SASS lines of every kernel.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--top 12]"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_warps", "launch__grid_size", "launch__block_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sectors.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]


def ncu(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    top_n = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 12
    raw = ncu(rep, "raw")
    hdr, units = raw[0], raw[1]
    for r in raw[2:]:
        print("=" * 100)
        print(r[hdr.index("Kernel Name")], " grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
        for w in WANT:
            if w in hdr:
                print("  %-62s %s %s" % (w, r[hdr.index(w)], units[hdr.index(w)]))
        st = [(hdr[i], r[i]) for i in range(len(hdr))
              if "smsp__average_warps_issue_stalled" in hdr[i] and hdr[i].endswith("_per_issue_active.ratio")]
        st = sorted(st, key=lambda x: -float(x[1].replace(",", "") or 0))[:7]
        print("  stalls/issue: " + ", ".join("%s %.2f" % (a.replace("smsp__average_warps_issue_stalled_", "")
                                                        .replace("_per_issue_active.ratio", ""), float(b)) for a, b in st))
    src = ncu(rep, "source")
    # the source page concatenates kernels: split on the "Kernel Name" rows
    blocks, cur = [], None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    for b in blocks:
        rows = b["rows"]
        h = rows[0]
        iS, iE, iSm, iT = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples"), \
            h.index("Thread Instructions Executed")
        body = [r for r in rows[1:] if len(r) > iE and r[iE].isdigit()]
        tot = sum(int(r[iE]) for r in body) or 1
        tots = sum(int(r[iSm]) for r in body) or 1
        print("-" * 100)
        print(b["name"], " executed warp-instr:", tot, " samples:", tots)
        # stall reasons of the samples taken in each segment (a warp waiting at a barrier is sampled at
        # the instruction AFTER it, i.e. it shows up as "barrier" at the top of the next segment)
        stall_cols = [(i, c.replace("stall_", "")) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
        seg = acc = accs = acct = n = 0
        why = [0] * len(stall_cols)
        for idx, r in enumerate(body):
            acc += int(r[iE]); accs += int(r[iSm]); acct += int(r[iT]); n += 1
            for q, (ci, _) in enumerate(stall_cols):
                why[q] += int(r[ci] or 0)
            s = r[iS]
            if "BAR." in s or "EXIT" in s or "RET." in s:
                if acc * 200 > tot or accs * 200 > tots:
                    top = sorted(zip(why, [c for _, c in stall_cols]), reverse=True)[:4]
                    wsum = sum(why) or 1
                    print("  seg %2d ..%5d %-34s sass %4d  exec %5.1f%%  samples %5.1f%%  lanes %4.1f  | %s" % (
                        seg, idx, s.strip()[:34], n, 100 * acc / tot, 100 * accs / tots, acct / max(acc, 1),
                        ", ".join("%s %.0f%%" % (c, 100 * v / wsum) for v, c in top)))
                seg += 1; acc = accs = acct = n = 0
                why = [0] * len(stall_cols)
        for r in sorted(body, key=lambda r: -int(r[iSm]))[:top_n]:
            print("    %6s samples  %9s exec   %s" % (r[iSm], r[iE], r[iS].strip()[:80]))


if __name__ == "__main__":
    main()
