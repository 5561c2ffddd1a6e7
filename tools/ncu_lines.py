"""Attribute an ncu source-page export (SASS rows with execution counts and stall samples) to CUDA source lines.

    ncu -i rep.ncu-rep --page source --csv > src.csv
    cuobjdump -xelf all lib.so ; nvdisasm --print-line-info x.cubin > dis.txt
    python tools/ncu_lines.py src.csv dis.txt <kernel-name-substring> [top]

The SASS rows of ncu and nvdisasm come in the same order, so they are joined by position."""
import csv, re, sys, collections

src_csv, dis_txt, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(src_csv)))
hi = [i for i, r in enumerate(rows) if 'Source' in r and any('Sampl' in x for x in r)][0]
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
sass = []
for r in rows[hi + 1:]:
    try:
        sass.append((int(r[idx['Instructions Executed']]), int(r[idx['Thread Instructions Executed']]),
                     int(r[idx['Warp Stall Sampling (All Samples)']]), r[idx['Source']].strip()))
    except Exception:
        pass
# nvdisasm: find the .text section of the kernel; lines "//## File "...", line N" precede instructions
lines = open(dis_txt).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('\t.section\t.text.') and kname in l)
cur = None; stack = []
insts = []
for l in lines[start + 1:]:
    if l.startswith('\t.section') or l.startswith('//-----'):
        if insts: break
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        f = m.group(1).split('/')[-1]
        cur = f"{f}:{m.group(2)}"
        inl = re.search(r'inlined at "([^"]+)", line (\d+)', l)
        if inl: cur += f" <- {inl.group(1).split('/')[-1]}:{inl.group(2)}"
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        insts.append((cur, m.group(2).strip()))
print(f"{len(sass)} ncu rows, {len(insts)} nvdisasm instructions")
n = min(len(sass), len(insts))
agg = collections.defaultdict(lambda: [0, 0, 0])
for (e, t, s, _), (loc, _) in zip(sass[:n], insts[:n]):
    a = agg[loc]; a[0] += e; a[1] += t; a[2] += s
tot_e = sum(a[0] for a in agg.values()); tot_s = sum(a[2] for a in agg.values())
print(f"total warp instructions {tot_e}, stall samples {tot_s}")
def srcline(loc):
    try:
        f, ln = loc.split(' <- ')[0].split(':')
        import glob
        p = [x for x in glob.glob('**/' + f, recursive=True)][0]
        return open(p).read().split('\n')[int(ln) - 1].strip()[:90]
    except Exception:
        return ''
for loc, (e, t, s) in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{100*e/tot_e:5.1f}% inst {100*s/max(tot_s,1):5.1f}% stall  lanes {t/max(e,1):5.1f}  {loc}   | {srcline(loc)}")
