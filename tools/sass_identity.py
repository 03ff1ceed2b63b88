"""Are the kernels of two object files instruction-for-instruction the same?  Used when an experimental template variant
is added next to a shipped kernel: the shipped instantiations must not change (addresses and opcodes compared; encodings,
the anonymous-namespace hash and trailing defaulted template arguments ignored).

usage: python tools/sass_identity.py old.o new.o        (exit status 1 if any kernel of old.o changed or disappeared)
"""
import re, subprocess, sys


def kernels(path):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    out, cur = {}, None
    for l in txt.splitlines():
        m = re.search(r"Function : (\S+)", l)
        if m:
            cur = re.sub(r"_GLOBAL__N__[0-9a-f]+_\d+_\w+?_cu_[0-9a-f]+", "NS", m.group(1))
            out[cur] = []
            continue
        if cur is not None and re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", l):
            out[cur].append(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", l).rstrip())
    return out


if __name__ == "__main__":
    a, b = kernels(sys.argv[1]), kernels(sys.argv[2])
    bad = 0
    for k, body in sorted(a.items()):
        cands = [x for x in b if x == k or re.sub(r"(ELb[01])+(EEEv)", r"\2", x) == re.sub(r"(ELb[01])+(EEEv)", r"\2", k)]
        same = [x for x in cands if b[x] == body]
        print(("IDENTICAL " if same else "DIFFERENT ") + k[-70:] + ("" if same else f"   candidates: {len(cands)}"))
        bad += not same
    for k in sorted(set(b) - set(a)):
        print("NEW       " + k[-70:])
    sys.exit(1 if bad else 0)
