"""Top SASS instructions by stall samples for one launch: python tools/ncu_sass.py report.ncu-rep <launch-index> [top N]"""
import csv, io, subprocess, sys
rep, idx = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:stage_kernel",
                      "--launch-skip", str(idx), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ia = hdr.index('Instructions Executed'); isrc = hdr.index('Source'); isamp = hdr.index('# Samples')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data = []
for n, r in enumerate(rows[2:]):
    try: data.append((int(r[isamp]), int(r[ia]), n, r[isrc].strip(), {hdr[i][6:]: int(r[i]) for i in stall_cols if r[i] not in ('', '0')}))
    except (ValueError, IndexError): pass
tot = sum(d[0] for d in data)
print('samples', tot)
for s, n, ln, src, st in sorted(data, reverse=True)[:top]:
    main = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f'{100*s/tot:5.2f}% #{ln:6d} exec {n:9d} {src[:60]:60s} {main}')
