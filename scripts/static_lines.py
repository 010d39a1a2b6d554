"""Static SASS instruction count per source line / function of one kernel section of an `nvdisasm -g -c` dump.
usage: static_lines.py <dis> <section substring> [source file for function names]"""
import re, sys, collections, bisect
cur = None; cnt = collections.Counter(); on = False
for ln in open(sys.argv[1]):
    if ln.startswith('.text.'):
        on = sys.argv[2] in ln; continue
    if not on: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'^\s+/\*[0-9a-f]{4,6}\*/', ln): cnt[cur] += 1
srcname = sys.argv[3] if len(sys.argv) > 3 else 'trajectory_generator_b200/csrc/tg_sqp.h'
base = srcname.split('/')[-1]
src = open(srcname).read().split('\n')
funcs = []
for i, l in enumerate(src, 1):
    m = re.match(r'^(TG_QFN|TG_FN|TG_HD|template|static).*?\b(tg_\w+)\(', l)
    if m: funcs.append((i, m.group(2)))
starts = [f[0] for f in funcs]
agg = collections.Counter()
for (f, l), c in cnt.items():
    if f == base:
        k = bisect.bisect_right(starts, l) - 1
        agg[funcs[k][1] if k >= 0 else 'pre'] += c
    else: agg[f] += c
for k, v in agg.most_common(): print(v, k)
print(sum(cnt.values()), "total")
for (f, l), c in sorted(cnt.items(), key=lambda kv: -kv[1])[:30]:
    print(c, f, l, src[l - 1].strip()[:100] if f == base else '')
