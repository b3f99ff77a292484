#!/usr/bin/env python
"""Markdown digest of `ncu --set full` reports (run where ncu is installed; reads .ncu-rep, prints tables):
    python scripts/ncu_summary.py gpurun_out/a.ncu-rep [b.ncu-rep ...] > profiles/rNN_ncu_summary.md
Per kernel launch: duration, DRAM bytes, lanes per instruction, issue / warps active, pipe utilisation, L1 / L2 hit rates and the
top stall reasons; for reports captured with --import-source on, the instruction mix per 25-instruction block of the first kernel."""
import csv
import io
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
        ("smsp__inst_executed.sum", "warp instructions"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / instruction"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "registers / thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"), ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %")]


def export(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    for rep in sys.argv[1:]:
        rows = export(rep, "raw")
        if len(rows) < 3:
            print("## %s: empty\n" % rep)
            continue
        h, u = rows[0], rows[1]
        print("## %s\n" % rep.split("/")[-1])
        for d in rows[2:]:
            name = d[h.index("Kernel Name")]
            print("### `%s`\n\n| metric | value | unit |\n|---|---|---|" % name[:90])
            for key, label in WANT:
                if key in h:
                    i = h.index(key)
                    print("| %s (`%s`) | %s | %s |" % (label, key, d[i], u[i]))
            st = sorted(((float(d[i]), n) for i, n in enumerate(h) if n.startswith("smsp__average_warps_issue_stalled") and n.endswith("per_issue_active.ratio")), reverse=True)
            print("\nstalls (warps per issue): " + ", ".join("%s %.2f" % (n.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v) for v, n in st[:7]) + "\n")
        src = export(rep, "source", ("--print-source", "sass"))
        hdr_i = next((i for i, r in enumerate(src) if r and r[0] == "Address"), None)
        if hdr_i is None:
            continue
        hdr, data = src[hdr_i], [r for r in src[hdr_i + 1:] if len(r) > 8 and r[0].startswith("0x")]
        kernels = [r for r in src[:hdr_i] if r and r[0] == "Kernel Name"]
        try:
            ie, it, isamp, isrc = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
        except ValueError:
            continue
        # only the first kernel of the report (the source page lists them one after another; addresses restart)
        first = []
        last = -1
        for r in data:
            a = int(r[0], 16)
            if a < last:
                break
            last = a
            first.append(r)
        tot = sum(int(r[ie]) for r in first) or 1
        tots = sum(int(r[isamp]) for r in first) or 1
        print("#### instruction mix per 25-instruction block of `%s` (source page)\n\n| SASS index | %% of warp instructions | lanes | %% of stall samples | opcodes |\n|---|---|---|---|---|"
              % (kernels[0][1][:60] if kernels else "first kernel"))
        for k in range(0, len(first), 25):
            blk = first[k:k + 25]
            e = sum(int(r[ie]) for r in blk)
            t = sum(int(r[it]) for r in blk)
            sm = sum(int(r[isamp]) for r in blk)
            if e * 200 < tot:
                continue
            ops = " ".join(sorted({(r[isrc].split()[1] if r[isrc].strip().startswith("@") else r[isrc].split()[0]).split(".")[0] for r in blk if r[isrc].strip()}))
            print("| %d | %.1f | %.1f | %.1f | %s |" % (k, 100.0 * e / tot, t / max(e, 1), 100.0 * sm / tots, ops[:80]))
        print()


if __name__ == "__main__":
    main()
