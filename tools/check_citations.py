"""tools/check_citations.py -- every `file.c:line[-line]` citation of the reference in include/, libvips_b200/csrc/, oracle/,
DESIGN.md and INTEGRATION.md must name a file that exists under /root/reference/libvips (or the tree root) and lines that
exist in it.  CPU, needs /root/reference."""
import glob
import os
import re
import sys

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
index = {}
for dp, dn, fn in os.walk(REF):
    if "/.git" in dp:
        continue
    for f in fn:
        index.setdefault(f, []).append(os.path.join(dp, f))


def lines_of(path, cache={}):
    if path not in cache:
        with open(path, "rb") as fh:
            cache[path] = fh.read().count(b"\n") + 1
    return cache[path]


bad = 0
files = glob.glob(ROOT + "/include/*.h") + glob.glob(ROOT + "/libvips_b200/csrc/*.c*") + glob.glob(ROOT + "/libvips_b200/csrc/*.h") + \
    glob.glob(ROOT + "/oracle/*.cpp") + glob.glob(ROOT + "/oracle/ref_shim/*.c*") + [ROOT + "/DESIGN.md", ROOT + "/INTEGRATION.md"]
n = 0
for f in files:
    text = open(f, errors="replace").read()
    last = None
    for m in re.finditer(r"(?:([\w./+-]+\.(?:cpp|c|h|py|md|sh|build)):|(?<=[ ,(]):)(\d+)(?:-(\d+))?", text):
        name = m.group(1) or last
        if m.group(1):
            last = m.group(1)
        if not name:
            continue
        base = os.path.basename(name)
        if base not in index or base in ("vb200.h",):
            continue  # not a reference file (our own sources, or third-party files named in passing)
        cands = [p for p in index[base] if p.endswith("/" + name.lstrip("./"))] or index[base]
        hi = int(m.group(3) or m.group(2))
        n += 1
        if not any(lines_of(p) >= hi for p in cands):
            bad += 1
            print("%s: %s:%s beyond the end of %s (%d lines)" % (os.path.relpath(f, ROOT), name, m.group(0).split(":")[-1], cands[0][len(REF) + 1:], lines_of(cands[0])))
print("%d citations checked, %d bad" % (n, bad))
sys.exit(1 if bad else 0)
