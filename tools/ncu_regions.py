"""share of executed warp instructions / stall samples per source-line range of one file, from an .ncu-rep
usage: python tools/ncu_regions.py rep file.cuh name:lo-hi [name:lo-hi ...]   (lines outside: by file name)"""
import collections
import csv
import io
import subprocess
import sys

rep, fname, specs = sys.argv[1], sys.argv[2], sys.argv[3:]
regions = [(s.split(":")[0], *map(int, s.split(":")[1].split("-"))) for s in specs]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
inst, samp, cur, hdr = collections.Counter(), collections.Counter(), None, None
for r in csv.reader(io.StringIO(src)):
    if not r:
        continue
    if r[0] in ("File Path", "File Name"):
        cur = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit():
        try:
            i, s = int(r[hdr.index("Instructions Executed")] or 0), int(r[hdr.index("# Samples")] or 0)
        except ValueError:
            continue
        key = cur
        if cur == fname:
            for name, lo, hi in regions:
                if lo <= int(r[0]) <= hi:
                    key = name
                    break
        inst[key] += i
        samp[key] += s
ti, ts = sum(inst.values()) or 1, sum(samp.values()) or 1
for k, v in inst.most_common():
    print("%-24s inst %5.1f%%  samples %5.1f%%" % (k, 100 * v / ti, 100 * samp[k] / ts))
