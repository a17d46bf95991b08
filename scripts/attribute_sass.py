"""Attributes executed SASS instructions of one kernel (ncu --page source --csv) to the C++ functions they were inlined
from, using nvdisasm -g line info.  usage: attribute_sass.py <nvdisasm.txt> <ncu_source.csv> <kernel-name-substring>"""
import collections, csv, re, sys, pathlib
lines = open(sys.argv[1]).read().split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and sys.argv[3] in l][0]
end = [i for i, l in enumerate(lines) if ".section" in l and i > start][0]
cur, seq = None, []
for l in lines[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        seq.append(cur)
rows = list(csv.reader(open(sys.argv[2])))
hdr, data = rows[1], rows[2:]
ie, isamp, it = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
assert len(seq) == len(data), (len(seq), len(data))
csrc = pathlib.Path(__file__).resolve().parent.parent / "balance_robot_b200" / "csrc"
src = {p.name: p.read_text().split("\n") for p in csrc.iterdir() if p.suffix in (".cu", ".cuh", ".inc")}
sig = re.compile(r"^(?!#|//|\s|\}|template|struct|enum|extern \"C\" \{)(?:[\w:<>\*&\s]+?)\b(\w+)\s*\(")
def func_of(fn, ln):
    if fn not in src:
        return fn
    if fn.endswith(".inc"):
        return fn
    L = src[fn]
    for i in range(min(ln, len(L)) - 1, -1, -1):
        m = sig.match(L[i])
        if m and m.group(1) not in ("if", "for", "while", "switch", "return", "sizeof"):
            return m.group(1)
    return fn
fa, fs, ft = collections.Counter(), collections.Counter(), collections.Counter()
for key, r in zip(seq, data):
    f = func_of(*key) if key else "?"
    fa[f] += int(r[ie]); fs[f] += int(r[isamp]); ft[f] += int(r[it])
tot, ts = sum(fa.values()), sum(fs.values())
print(f"{'function':34s} {'inst %':>7s} {'samples %':>9s} {'lanes':>6s}")
for f, c in fa.most_common(24):
    print(f"{f:34s} {100*c/tot:7.2f} {100*fs[f]/ts:9.2f} {ft[f]/max(1,c):6.1f}")
