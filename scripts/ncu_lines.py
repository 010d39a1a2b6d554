"""Joins an `ncu --page source --csv` dump (SASS rows with stall samples) with `nvdisasm -g` line info of the same
kernel and prints where the samples / executed instructions fall, per source line and per device function.
usage: ncu_lines.py <ncu_source.csv> <nvdisasm -g -c of the kernel's section> [top]"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
isamp = hdr.index('# Samples'); iex = hdr.index('Instructions Executed'); isrc = hdr.index('Source')
cols = {k: hdr.index(k) for k in ('stall_long_sb', 'stall_wait', 'stall_no_inst', 'stall_short_sb', 'stall_branch_resolving', 'stall_math', 'stall_not_selected', 'stall_selected', 'stall_dispatch', 'stall_mio', 'stall_lg')}
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
ins = []; cur = ("?", 0); func = "kernel"
for ln in open(sys.argv[2]):
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'^(\$?[\w$.]+):', ln.strip())
    if m and not ln.strip().startswith('.L_'):
        func = m.group(1)[-50:]; continue
    if re.match(r'^\s+/\*[0-9a-f]{4,6}\*/', ln):
        ins.append((cur, func, ln.strip()))
print("sass rows", len(data), "disasm instrs", len(ins))
n = min(len(data), len(ins))
byline = collections.defaultdict(lambda: [0, 0]); byfunc = collections.defaultdict(lambda: [0, 0])
stall = collections.defaultdict(lambda: collections.Counter())
ts = te = 0
for r, (cur, func, txt) in zip(data[:n], ins[:n]):
    s = int(r[isamp]); e = int(r[iex]); ts += s; te += e
    byline[cur][0] += s; byline[cur][1] += e
    byfunc[func][0] += s; byfunc[func][1] += e
    for k, c in cols.items():
        stall[cur][k] += int(r[c] or 0)
print("total samples", ts, "instr", te)
print("--- by function")
for f, (s, e) in sorted(byfunc.items(), key=lambda kv: -kv[1][0])[:30]:
    print("%6.2f%% samples %6.2f%% instr  %s" % (100.0 * s / ts, 100.0 * e / te, f))
print("--- by line")
for l, (s, e) in sorted(byline.items(), key=lambda kv: -kv[1][0])[:top]:
    st = stall[l]; tot = sum(st.values()) or 1
    print("%5.2f%% samp %5.2f%% instr  %s:%d   %s" % (100.0 * s / ts, 100.0 * e / te, l[0], l[1], " ".join("%s=%.0f%%" % (k.replace('stall_', ''), 100.0 * v / tot) for k, v in st.most_common(3))))
