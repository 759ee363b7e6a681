"""Print the hot SASS regions of one kernel from an .ncu-rep captured with --import-source on (development aid)."""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
detail = len(sys.argv) > 3
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern, "--launch-count", "1"] if False else
                     ["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = [r for r in csv.reader(raw.splitlines())]
hdr = rows[1]
ia, ie, it, isamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Avg. Threads Executed'), hdr.index('# Samples')
body = [r for r in rows[2:] if len(r) > isamp and r[ie].isdigit()]
# keep the first launch only (the listing repeats per launch)
first = body[0][ia]
for k in range(1, len(body)):
    if body[k][ia] == first and k > 8:
        body = body[:k]; break
tot = sum(int(r[ie]) for r in body); mx = max(int(r[ie]) for r in body); tots = max(1, sum(int(r[isamp]) for r in body))
print('total warp-instr', tot, 'sass', len(body), 'max count', mx)
i = 0
while i < len(body):
    j = i
    while j + 1 < len(body) and body[j + 1][ie] == body[i][ie]: j += 1
    e = int(body[i][ie]); n = j - i + 1
    samp = sum(int(body[k][isamp]) for k in range(i, j + 1)) * 100 / tots
    ops = {}
    for k in range(i, j + 1):
        t = body[k][ia].split()
        op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]; ops[op] = ops.get(op, 0) + 1
    top = ' '.join(f"{k}:{v}" for k, v in sorted(ops.items(), key=lambda x: -x[1])[:8])
    if e * n / tot > 0.004:
        print(f"[{i:3d}-{j:3d}] n={n:3d} x{e / mx:5.2f} share={e * n * 100 / tot:5.1f}% samp={samp:5.1f}% thr={float(body[i][it]):4.1f} {top}")
        if detail:
            for k in range(i, j + 1): print("        ", body[k][ia].strip()[:100])
    i = j + 1
