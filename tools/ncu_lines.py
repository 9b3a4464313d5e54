"""Per-source-line executed instructions / stall samples of one kernel launch in an ncu report.
usage: python tools/ncu_lines.py report.ncu-rep <launch-index> [top N]"""
import csv, io, subprocess, sys, re, collections
rep, idx = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + (sys.argv[4] if len(sys.argv) > 4 else "stage_kernel"),
                      "--launch-skip", str(idx), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur = None; out = []; ops = collections.Counter(); opsamp = collections.Counter()
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': hdr = r; continue
    if hdr is None or len(r) < 10: continue
    ie = hdr.index('Instructions Executed'); isamp = hdr.index('# Samples')
    if r[0] not in ('', 'Line No'):
        try: out.append((int(r[ie]), int(r[isamp]), cur, int(r[0]), r[1].strip()[:105]))
        except ValueError: pass
    elif r[0] == '' and r[2] not in ('', '...'):
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[3])
        try:
            n = int(r[ie]); s = int(r[isamp])
            op = m.group(2) if m else '?'
            ops[op] += n; opsamp[op] += s
        except ValueError: pass
tot = sum(o[0] for o in out); tsamp = sum(o[1] for o in out)
print('total warp-instructions', tot, 'samples', tsamp)
out.sort(reverse=True)
for n, s, f, l, src in out[:top]: print(f'{100*n/tot:5.1f}% smp{100*s/max(tsamp,1):5.1f}% {f}:{l} {src}')
print('--- opcodes')
for op, n in ops.most_common(24): print(f'{op:10s} {100*n/tot:5.1f}%  samples {100*opsamp[op]/max(tsamp,1):5.1f}%')
