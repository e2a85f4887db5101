"""Group the stall samples of an ncu source page (csv) by execution-count region."""
import collections
import csv
import io
import sys


def main(path, top=16):
    txt = open(path).read().splitlines()
    idx = [i for i, l in enumerate(txt) if l.startswith('"Kernel Name"')]
    blk = txt[idx[0] + 1: idx[1] if len(idx) > 1 else None]
    rows = list(csv.reader(io.StringIO("\n".join(blk))))
    hdr = rows[0]
    H = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    data = rows[1:]
    reg = collections.Counter(); ninstr = collections.Counter()
    regstall = collections.defaultdict(collections.Counter)
    tot = 0
    for r in data:
        e = int(r[H['Instructions Executed']]); s = int(r[H['# Samples']])
        reg[e] += s; ninstr[e] += 1; tot += s
        for st in stalls:
            regstall[e][st] += int(r[H[st]])
    print("total samples", tot)
    for k, v in reg.most_common(top):
        print(f"exec={k:9d} ninstr={ninstr[k]:4d} samples={v:6d} ({100*v/tot:4.1f}%)  {regstall[k].most_common(4)}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 16)
