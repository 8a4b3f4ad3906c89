"""Local-memory loads / stores (spills, stack) of one kernel by source line.  python tools/sass_spills.py <object.o> <kernel name fragment> [rows]"""
import re, subprocess, sys, tempfile, os, glob
obj, frag = sys.argv[1], sys.argv[2]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, check=True, capture_output=True)
sass = subprocess.run(["nvdisasm", "-g", "-c", glob.glob(d + "/*.cubin")[0]], capture_output=True, text=True).stdout.split("\n")
heads = [i for i, l in enumerate(sass) if l.startswith(".text.")]
start = [i for i in heads if frag in sass[i]][0]
end = min([i for i in heads if i > start] + [len(sass)])
line, out, files = None, {}, {}
for l in sass[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (m.group(1), int(m.group(2)))
    if re.search(r"\b(LDL|STL)", l):
        out.setdefault(line, []).append(re.sub(r"/\*[0-9a-f]+\*/", "", l).strip()[:70])
for k in sorted(out, key=lambda k: (k[0], k[1])):
    try:
        files.setdefault(k[0], open(k[0]).read().split("\n"))
        src = files[k[0]][k[1] - 1].strip()[:110]
    except Exception:
        src = ""
    print(os.path.basename(k[0]), k[1], len(out[k]), "|", src)
    for x in out[k][:int(sys.argv[3]) if len(sys.argv) > 3 else 4]:
        print("       ", x)
