"""Text summary of one kernel of an ncu --set full report: headline counters, warp stall samples, hottest instructions
(by stall samples and by executions).  usage: python tools/ncu_summary.py report.ncu-rep kernel_regex [launch_index]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size" , "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.max"]


def blocks_of(text, first):
    out, cur = [], []
    for line in text.splitlines():
        if line.startswith(first):
            if cur:
                out.append(cur)
            cur = []
        cur.append(line)
    if cur:
        out.append(cur)
    return out


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "-k", "regex:" + rx], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    r = data[which]
    print("Kernel Name".ljust(82), r[hdr.index("Kernel Name")][:90])
    for k in KEYS:
        if k in hdr:
            print(k.ljust(82), r[hdr.index(k)], units[hdr.index(k)])
    stalls = [(h, float(r[i])) for i, h in enumerate(hdr) if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued") and r[i]]
    tot = sum(v for _, v in stalls) or 1.0
    print("\nwarp stall samples (all warps):")
    for h, v in sorted(stalls, key=lambda kv: -kv[1])[:10]:
        print(f"  {h.replace('smsp__pcsamp_warps_issue_stalled_', 'stall_'):28s} {int(v):9d} {100 * v / tot:6.1f}%")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + rx], capture_output=True, text=True).stdout
    blk = blocks_of(src, '"Kernel Name"')[which]
    srows = list(csv.reader(io.StringIO("\n".join(blk[1:]))))
    sh = srows[0]
    isrc, isamp, iex = sh.index("Source"), sh.index("# Samples"), sh.index("Instructions Executed")
    ins = []
    for x in srows[1:]:
        try:
            ins.append((int(x[isamp]), int(x[iex]), x[isrc].strip()))
        except Exception:
            pass
    ts, te = sum(i[0] for i in ins) or 1, sum(i[1] for i in ins) or 1
    print(f"\n{len(ins)} SASS instructions, {te} warp-level executions, {ts} stall samples")
    print("hottest instructions by samples (samples, executions, SASS):")
    for s, e, t in sorted(ins, reverse=True)[:14]:
        print(f"  {s:7d} {e:10d}  {t[:90]}")
    print("most executed instructions (executions, samples, SASS):")
    for s, e, t in sorted(ins, key=lambda i: -i[1])[:10]:
        print(f"  {e:10d} {s:7d}  {t[:90]}")


if __name__ == "__main__":
    main()
