"""TEST INFRASTRUCTURE ONLY -- mechanical source rewrite for the CUDA-on-CPU emulation build
(tests/emu/include/cuda_runtime.h).  The kernel and host sources of user-gfmd_b200/csrc are
taken as they are; only the three constructs g++ cannot parse are rewritten:

  kernel<<<grid, block, smem, stream>>>(args)   ->  emu::launch(grid, block, smem, [&]() { kernel(args); })
  extern __shared__ T name[];                   ->  T *name = reinterpret_cast<T *>(emu::dyn_smem());
  asm volatile("...ptx..." ...);                ->  (removed: prefetch hints only)
"""
import re
import sys


def _match_back_angle(s, end):
    """s[end-1] == '>': index of the matching '<'."""
    depth = 0
    i = end - 1
    while i >= 0:
        if s[i] == '>':
            depth += 1
        elif s[i] == '<':
            depth -= 1
            if depth == 0:
                return i
        i -= 1
    raise ValueError("unbalanced template brackets before <<<")


def _match_paren(s, start):
    """s[start] == '(': index of the matching ')'."""
    depth = 0
    for i in range(start, len(s)):
        if s[i] == '(':
            depth += 1
        elif s[i] == ')':
            depth -= 1
            if depth == 0:
                return i
    raise ValueError("unbalanced parentheses after >>>")


def _split_top(s):
    """Split on commas that are not nested in (), <> or []."""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == ',' and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    out.append(cur)
    return out


def rewrite_launches(s):
    while True:
        k = s.find("<<<")
        if k < 0:
            return s
        # kernel name (with optional template arguments) ends at k
        j = k
        while j > 0 and s[j - 1].isspace():
            j -= 1
        name_end = j
        if s[j - 1] == '>':
            j = _match_back_angle(s, j)
        while j > 0 and (s[j - 1].isalnum() or s[j - 1] in "_:"):
            j -= 1
        name = s[j:name_end]
        e = s.find(">>>", k)
        cfg = _split_top(s[k + 3:e])      # continuation backslashes inside a macro stay where they are
        while len(cfg) < 3:
            cfg.append("0")
        p0 = e + 3
        while s[p0].isspace() or s[p0] == '\\':
            p0 += 1
        assert s[p0] == '(', "launch of %s without argument list" % name
        p1 = _match_paren(s, p0)
        args = s[p0:p1 + 1]
        repl = "emu::launch(%s, %s, %s, [&]() { %s%s; })" % (cfg[0].strip(" \t"), cfg[1].strip(" \t"),
                                                         cfg[2].strip(" \t"), name, args)
        s = s[:j] + repl + s[p1 + 1:]


def rewrite(s):
    s = rewrite_launches(s)
    s = re.sub(r"extern\s+__shared__\s+(\w+)\s+(\w+)\[\];",
               r"\1 *\2 = reinterpret_cast<\1 *>(emu::dyn_smem());", s)
    s = re.sub(r"asm\s+volatile\s*\(.*\);", "/* asm removed by tests/emu/preprocess.py */;", s)   # one-line asm only
    return s


if __name__ == "__main__":
    src, dst = sys.argv[1], sys.argv[2]
    with open(src) as f:
        text = f.read()
    with open(dst, "w") as f:
        f.write(rewrite(text))
