"""Top stall-sample SASS lines of one kernel from an ncu report's source page.
usage: python tools/ncu_hot.py report.ncu-rep kernel_regex [top] [launch_index]"""
import csv
import io
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + rx],
                         capture_output=True, text=True).stdout
    # several launches are concatenated; keep the first block unless told otherwise
    which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    blocks, cur = [], []
    for line in out.splitlines():
        if line.startswith('"Kernel Name"'):
            if cur:
                blocks.append(cur)
            cur = []
        cur.append(line)
    if cur:
        blocks.append(cur)
    blk = blocks[which]
    rows = list(csv.reader(io.StringIO("\n".join(blk[1:]))))
    hdr = rows[0]
    ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    data = []
    for k, r in enumerate(rows[1:]):
        try:
            data.append((int(r[isamp]), k, r[ia][-5:], int(r[iex]), r[isrc].strip()))
        except Exception:
            pass
    tot = sum(d[0] for d in data)
    print(f"{blk[0][:120]}\n total samples {tot}, instructions {len(data)}")
    for s, k, a, ex, src in sorted(data, reverse=True)[:top]:
        print(f"{a} #{k:5d} {s:7d} {100.0*s/max(tot,1):5.1f}%  exec={ex:9d}  {src[:100]}")


if __name__ == "__main__":
    main()
