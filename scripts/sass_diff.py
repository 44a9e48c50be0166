"""Which kernels changed? Compiles cpvs_b200/csrc/*.cu of a git revision and of the working tree for sm_100a and compares
the SASS of every kernel (branch targets and parameter offsets normalised). Used to show that a change behind a switch left
the default path's machine code alone when there is no GPU at hand to re-run the parity suite.

    python scripts/sass_diff.py <git-rev> [file ...]        # files default to svo merge emit pyramid lookup synthgen
"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--fmad=false", "-cubin"]


def kernels(cubin):
    out = subprocess.check_output(["cuobjdump", "-sass", cubin], text=True)
    res, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.check_output(["c++filt", m.group(1)], text=True).strip()
            name = re.sub(r"\(anonymous namespace\)::", "", name)
            cur = re.sub(r"^void ", "", re.sub(r"\(.*", "", name))
            res[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?);", line)
        if m and cur:
            res[cur].append(re.sub(r"c\[0x0\]\[0x[0-9a-f]+\]", "c[P]", re.sub(r"0x[0-9a-f]{5,}", "ADDR", m.group(1).strip())))
    return res


def main():
    rev = sys.argv[1]
    files = sys.argv[2:] or ["svo", "merge", "emit", "pyramid", "lookup", "synthgen"]
    with tempfile.TemporaryDirectory() as tmp:
        tar = subprocess.Popen(["git", "-C", ROOT, "archive", rev, "cpvs_b200/csrc", "cpvs_b200/synth", "include"], stdout=subprocess.PIPE)
        subprocess.check_call(["tar", "-x", "-C", tmp], stdin=tar.stdout)
        changed = 0
        for f in files:
            cub = {}
            for tag, base in (("old", tmp), ("new", ROOT)):
                cub[tag] = os.path.join(tmp, "%s_%s.cubin" % (tag, f))
                subprocess.check_call([NVCC] + FLAGS + [os.path.join(base, "cpvs_b200", "csrc", f + ".cu"), "-o", cub[tag]])
            old, new = kernels(cub["old"]), kernels(cub["new"])
            for name in sorted(set(old) | set(new)):
                # a kernel that became a template keeps its default behaviour in the <false> instantiation
                twin = next((t for t in (name, name + "<false>", name + "<false, false>") if t in new), name)
                if name not in old:
                    print("%-9s %-50s new kernel (%d instructions)" % (f, name[:50], len(new[name])))
                elif twin not in new:
                    print("%-9s %-50s REMOVED" % (f, name[:50]))
                    changed += 1
                else:
                    same = old[name] == new[twin]
                    changed += 0 if same else 1
                    print("%-9s %-50s %5d -> %5d  %s" % (f, name[:50], len(old[name]), len(new[twin]), "identical" if same else "DIFFERS"))
        print("%d kernel(s) of %s changed" % (changed, rev))


if __name__ == "__main__":
    main()
