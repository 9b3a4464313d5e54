"""Executed-instruction mix of one launch from the SASS page: python tools/ncu_mix.py report.ncu-rep <kernel regex> <launch-skip>"""
import csv, io, subprocess, sys, collections
rep, rx, idx = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx,
                      "--launch-skip", idx, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ia = hdr.index('Instructions Executed'); isrc = hdr.index('Source'); isamp = hdr.index('# Samples')
mix = collections.Counter(); samp = collections.Counter(); tot = 0
for r in rows[2:]:
    try:
        n = int(r[ia]); s = int(r[isamp])
    except (ValueError, IndexError):
        continue
    src = r[isrc].strip()
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.split(".")[0]
    mix[op] += n; samp[op] += s; tot += n
print("total warp instructions", tot)
ts = sum(samp.values())
for op, n in mix.most_common(28):
    print("%-10s %6.2f%% exec  %6.2f%% samples" % (op, 100.0 * n / tot, 100.0 * samp[op] / max(ts, 1)))
