"""Compare two builds of libfreerl_b200.so kernel by kernel at the SASS level (no GPU needed).

    git archive <commit> freerl_b200/csrc include | tar -x -C /tmp/base
    (cd /tmp/base && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -o base.so freerl_b200/csrc/capi.cu)
    python tools/sass_compare.py /tmp/base/base.so freerl_b200/libfreerl_b200.so [old_symbol=new_symbol ...]

Prints, for every kernel of the first library, whether the instruction stream (opcode + operands, addresses and encodings stripped)
is identical in the second one, else the opcode-count differences.  Used to prove that a change which must not touch a measured
kernel (e.g. a new compile-time variant next to it) really left it alone, when there is no GPU time left to re-measure."""
import collections
import re
import subprocess
import sys


def kernels(so):
    text = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, errors="ignore").stdout
    d, name, buf = {}, None, []
    for line in text.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                d[name] = buf
            name, buf = m.group(1), []
        elif name:
            mm = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?)\s*/\*", line)
            if mm:
                buf.append(mm.group(1))
    if name:
        d[name] = buf
    return d


def hist(x):
    return collections.Counter((i.split()[1] if i.startswith("@") else i.split()[0]) for i in x)


def main(old, new, *renames):
    ren = dict(r.split("=", 1) for r in renames)
    a, b = kernels(old), kernels(new)
    same = 0
    for k in a:
        kb = ren.get(k, k)
        if kb not in b:
            print("MISSING", k)
        elif a[k] == b[kb]:
            same += 1
        else:
            ha, hb = hist(a[k]), hist(b[kb])
            print("DIFF", k, len(a[k]), len(b[kb]), {op: (ha[op], hb[op]) for op in set(ha) | set(hb) if ha[op] != hb[op]})
    print("identical: %d of %d kernels" % (same, len(a)))
    for k in b:
        if k not in a and k not in ren.values():
            print("NEW", k, len(b[k]))


if __name__ == "__main__":
    main(*sys.argv[1:])
