#!/usr/bin/env python
"""Per-source-line hot spots of one kernel from an ncu report: joins `ncu --page source --csv` (SASS rows with instruction
counts and stall samples) with the line table of `nvdisasm -g` of the same cubin (instruction order is the same).
  python tools/ncu_lines.py <report.ncu-rep> <lib.so> <kernel-name-substring, mangled> [top]"""
import collections, csv, re, subprocess, sys, tempfile, os
rep, lib, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
body = [r for r in rows[2:] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
sass = None
for f in os.listdir(d):
    if f.endswith(".cubin") and "sm_100a" in f:
        txt = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(d, f)], capture_output=True, text=True).stdout
        if kern in txt:
            sass = txt
            break
assert sass, "kernel not found"
lines = sass.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l)
insts = []
cur = None
in_chain = False
cur_file = ""
for l in lines[start + 1:]:
    if l.startswith("\t.section") and ".text." in l:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        # -gi prints the inlining chain innermost first: attribute to the innermost frame that is not a toolkit header
        here = (os.path.basename(m.group(1)), int(m.group(2)))
        if not in_chain:
            cur, in_chain = here, True
        if "/cuda/" in cur_file and "/cuda/" not in m.group(1):
            cur = here
        if not in_chain or cur == here:
            cur_file = m.group(1)
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        insts.append((cur, l.strip()))
        in_chain = False
print(f"{len(insts)} SASS instructions in the cubin, {len(body)} rows in the report")
n = min(len(insts), len(body))
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0, 0])
def num(r, name):
    try:
        return int(float(r[col[name]] or 0))
    except (KeyError, ValueError):
        return 0
tot_i = tot_s = 0
for (src, _), r in zip(insts[:n], body[:n]):
    ie = int(r[col["Instructions Executed"]] or 0); te = int(r[col["Thread Instructions Executed"]] or 0); ss = int(r[col["# Samples"]] or 0)
    a = agg[src]; a[0] += ie; a[1] += te; a[2] += ss
    a[3] += num(r, "L2 Theoretical Sectors Global"); a[4] += num(r, "L1 Wavefronts Shared"); a[5] += num(r, "L2 Theoretical Sectors Local")
    tot_i += ie; tot_s += ss
src_lines = {}
for (fn, ln) in agg:
    if fn and fn not in src_lines:
        for root in (".", "physimglobalpose_b200/csrc"):
            pth = os.path.join(root, fn)
            if os.path.exists(pth):
                src_lines[fn] = open(pth).read().splitlines()
print(f"total warp instructions {tot_i:.3e}, samples {tot_s}")
mem = os.environ.get("NCU_LINES_MEM")     # NCU_LINES_MEM=1: order by memory traffic (global sectors + shared wavefronts + local sectors) instead
tot_g = sum(a[3] for a in agg.values()); tot_w = sum(a[4] for a in agg.values()); tot_l = sum(a[5] for a in agg.values())
if mem:
    print(f"global sectors {tot_g:.3e}, shared wavefronts {tot_w:.3e}, local sectors {tot_l:.3e}")
    for (src, a) in sorted(agg.items(), key=lambda kv: -(kv[1][3] + kv[1][4] + kv[1][5]))[:top]:
        fn, ln = src if src else ("?", 0)
        text = src_lines.get(fn, [""] * (ln + 1))[ln - 1].strip()[:100] if fn in src_lines and 0 < ln <= len(src_lines[fn]) else ""
        print(f"{fn}:{ln:4d} gsect {a[3] / 1e6:7.2f}M  shwave {a[4] / 1e6:7.2f}M  lsect {a[5] / 1e6:6.2f}M  inst {a[0] / tot_i * 100:4.1f}%  stall {a[2] / max(tot_s, 1) * 100:4.1f}%  | {text}")
    sys.exit(0)
for (src, (ie, te, ss, _g, _w, _l)) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    fn, ln = src if src else ("?", 0)
    text = src_lines.get(fn, [""] * (ln + 1))[ln - 1].strip()[:110] if fn in src_lines and 0 < ln <= len(src_lines[fn]) else ""
    print(f"{fn}:{ln:4d} inst {ie / tot_i * 100:5.1f}%  lanes {te / max(ie, 1):4.1f}  stall {ss / max(tot_s, 1) * 100:5.1f}%  | {text}")
