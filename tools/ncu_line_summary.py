#!/usr/bin/env python
"""Join an `ncu --page source --csv` SASS dump with nvdisasm line info: stall samples per CUDA source line.
usage: ncu_line_summary.py <src.csv> <object.o> <kernel-substring> [top]"""
import csv, subprocess, sys, re, collections
csv_path, obj, kname = sys.argv[1:4]; top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
dis = subprocess.run(['nvdisasm', '-g', '-c', obj], capture_output=True, text=True).stdout.splitlines()
lines = []; cur = None; infunc = False
for l in dis:
    if l.startswith('.text.') or re.match(r'\s*\.section\s+\.text\.', l):
        infunc = kname in l
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2)))
    if infunc and re.match(r'\s+/\*[0-9a-f]{4,}\*/\s', l):
        lines.append(cur)
rows = list(csv.reader(open(csv_path))); hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
print('sass instr: disasm', len(lines), 'ncu', len(data))
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = collections.Counter(); reasons = collections.defaultdict(collections.Counter); execd = collections.Counter()
for ln, r in zip(lines, data):
    s = float(r[ix['# Samples']] or 0); agg[ln] += s; execd[ln] += float(r[ix['Instructions Executed']] or 0)
    for h in stall_cols: reasons[ln][h[6:]] += float(r[ix[h]] or 0)
tot = sum(agg.values())
src_cache = {}
def src(ln):
    if ln is None: return ''
    f, n = ln
    import glob
    if f not in src_cache:
        c = glob.glob('/root/repo/torch_de_b200/csrc/' + f) + glob.glob('/usr/local/cuda/include/**/' + f, recursive=True)
        src_cache[f] = open(c[0]).read().splitlines() if c else []
    L = src_cache[f]
    return L[n - 1].strip()[:90] if 0 < n <= len(L) else ''
for ln, v in agg.most_common(top):
    rs = [(k, int(x)) for k, x in reasons[ln].most_common(3) if x]
    print(f'{100*v/tot:5.1f}% {str(ln):28s} ex={int(execd[ln]):11d} {rs}  | {src(ln)}')
