"""Which kernels of libgfmd_b200 still have the machine code of an earlier revision?  CPU only (nvcc
cross-compiles): builds user-gfmd_b200/csrc/gfmd_b200.cu of <rev> and of the working tree for
sm_100a, dumps the SASS of both and compares it kernel by kernel (names are matched up to template
and parameter suffixes that were added since).  Used to show that code written without GPU access
left the GPU-measured kernels byte for byte as they were:

    python tools/sass_identity.py <git-rev> > profiles/<round>_sass_identity.txt        (~6 min)
"""
import os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def sass_of(srcdir, out):
    obj = os.path.join(out, "k.o")
    subprocess.check_call([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-w",
                           "-c", os.path.join(srcdir, "user-gfmd_b200", "csrc", "gfmd_b200.cu"), "-o", obj])
    txt = subprocess.check_output(["cuobjdump", "-sass", obj], text=True)
    d, cur = {}, None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            d[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            d[cur].append(re.sub(r"/\*[0-9a-f]+\*/", "", line).strip())
    return d


def demangle(names):
    out = subprocess.check_output(["cu++filt"] + names, text=True).splitlines()
    return dict(zip(names, out))


def base(dem):
    """kernel name with its template arguments, without the parameter list (the first '(' outside
    the template brackets; template arguments are printed as "(int)3")"""
    depth = 0
    for i, c in enumerate(dem):
        if c == "<":
            depth += 1
        elif c == ">":
            depth -= 1
        elif c == "(" and depth == 0:
            return dem[:i].replace("void ", "")
    return dem.replace("void ", "")


def main():
    rev = sys.argv[1]
    with tempfile.TemporaryDirectory() as tmp:
        old = os.path.join(tmp, "old")
        os.makedirs(old)
        subprocess.check_call("git -C %s archive %s user-gfmd_b200/csrc include | tar -x -C %s" % (ROOT, rev, old), shell=True)
        a = sass_of(old, old)
        new = os.path.join(tmp, "new")
        os.makedirs(new)
        b = sass_of(ROOT, new)
    da, db = demangle(list(a)), demangle(list(b))
    newnames = {base(db[k]): k for k in b}
    print("kernels at %s: %d, now: %d" % (rev, len(a), len(b)))
    same = changed = gone = 0
    for k in sorted(a, key=lambda k: da[k]):
        name = base(da[k])
        # a template argument list that grew: old "<a, b>" is a prefix of new "<a, b, default...>"
        cand = [n for n in newnames if n == name or (name.endswith(">") and n.startswith(name[:-1] + ","))]
        match = [n for n in cand if b[newnames[n]] == a[k]]
        if match:
            same += 1
            print("identical  %s%s" % (name, "" if match[0] == name else "   (now %s)" % match[0]))
        elif cand:
            changed += 1
            print("CHANGED    %s  (%d -> %d instructions)" % (name, len(a[k]), len(b[newnames[cand[0]]])))
        else:
            gone += 1
            print("GONE       %s" % name)
    oldbases = {base(da[k]) for k in a}
    added = [n for n in newnames if not any(n == o or (o.endswith(">") and n.startswith(o[:-1] + ",")) for o in oldbases)]
    for n in sorted(added):
        print("new        %s" % n)
    print("identical %d, changed %d, gone %d, new %d" % (same, changed, gone, len(added)))


if __name__ == "__main__":
    main()
