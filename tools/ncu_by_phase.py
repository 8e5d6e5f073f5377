#!/usr/bin/env python
"""Attribute an ncu SASS-level profile to the kernel's phases (dev tool).

  cuobjdump -xelf all ramdsir_b200/libramdsir.so ; nvdisasm -gi -c ram_fused.sm_100a.cubin > fused_gi.sass
  ncu -i prof.ncu-rep --page source --csv > src.csv
  python tools/ncu_by_phase.py fused_gi.sass <mangled-kernel-substring> src.csv [top_file.cu [first-last line]]

Groups executed warp-instructions and stall samples by the OUTERMOST source line (the call site in
the kernel body), i.e. by phase of the fused pipeline.
"""
import collections, csv, re, sys

sass_file, kernel_pat, ncu_csv = sys.argv[1:4]
top_file = sys.argv[4] if len(sys.argv) > 4 else "ram_fused.cu"
lo_line, hi_line = (int(v) for v in sys.argv[5].split("-")) if len(sys.argv) > 5 else (0, 1 << 30)   # e.g. a task body's lines
in_k, chain, seq = False, [], []
for ln in open(sass_file):
    if ".section" in ln and ".text." in ln:
        in_k = kernel_pat in ln
        continue
    if not in_k:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        chain.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        seq.append((int(m.group(1), 16), m.group(2).strip(), chain))
        chain = []
# instructions without their own marker inherit the previous one
last = []
for i, (a, t, c) in enumerate(seq):
    if c:
        last = c
    seq[i] = (a, t, last)
rows = list(csv.reader(open(ncu_csv)))
h = rows[1]
ix = {n: i for i, n in enumerate(h)}
data = rows[2:]
base = int(data[0][0], 16)
stallcols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
agg = collections.defaultdict(lambda: [0, 0, collections.Counter(), collections.Counter()])
tot_i = tot_s = 0
for r, (addr, txt, ch) in zip(data, seq):
    assert int(r[0], 16) - base == addr
    ie, sm = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
    outer = [c for c in ch if c[0] == top_file and lo_line <= c[1] <= hi_line]
    key = outer[-1][1] if outer else -1
    a = agg[key]
    a[0] += ie
    a[1] += sm
    for c in stallcols:
        v = int(r[ix[c]])
        if v:
            a[2][c.replace("stall_", "")] += v
    a[3][txt.split()[0].split(".")[0] if not txt.startswith("@") else txt.split()[1].split(".")[0]] += ie
    tot_i += ie
    tot_s += sm
print("total warp-inst %d, samples %d" % (tot_i, tot_s))
for key, (ie, sm, st, ops) in sorted(agg.items()):
    print("%s:%-4d inst %5.1f%% (%9d)  samples %5.1f%%  stalls[%s]  ops[%s]" % (
        top_file, key, 100 * ie / tot_i, ie, 100 * sm / max(tot_s, 1),
        ", ".join("%s:%d" % kv for kv in st.most_common(4)), ", ".join("%s:%.0f%%" % (k, 100 * v / max(ie, 1)) for k, v in ops.most_common(6))))
