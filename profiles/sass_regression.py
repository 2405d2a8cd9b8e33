"""Instruction-level comparison of two builds of libmeb200.so: which kernels are bit-identical,
which differ only in register allocation / scheduling (same opcode histogram) and which really
changed.  Used when an experiment is added next to a measured kernel as an extra template
instantiation: the measured instantiations must come out IDENT.

    python profiles/sass_regression.py old/libmeb200.so minkowskiengine_b200/csrc/libmeb200.so
"""
import collections
import re
import subprocess
import sys


def kernels(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    cur, d = None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            d[cur] = []
            continue
        m = re.search(r"/\*[0-9a-f]+\*/\s+(.*?);", line)
        if m and cur:
            d[cur].append(m.group(1).strip())
    return d


def canonical(name):
    """Drop trailing defaulted template arguments added for experiments (..., 0>)."""
    name = re.sub(r"(k_conv_rsI.*?)Li0EEEv", r"\1EEv", name)
    return re.sub(r"(k_wgrad_tcI.*?)Li0EEEv", r"\1EEv", name)


def opcodes(instrs):
    return collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", i).split()[0] for i in instrs)


def main():
    old, new = kernels(sys.argv[1]), kernels(sys.argv[2])
    new_c = {canonical(n): v for n, v in new.items()}
    ident = 0
    for name, instrs in sorted(old.items()):
        key = canonical(name)
        if key not in new_c:
            print("REMOVED ", name)
        elif new_c[key] == instrs:
            ident += 1
        else:
            kind = "SCHEDULE" if opcodes(new_c[key]) == opcodes(instrs) else "CHANGED "
            print(kind, name, len(instrs), "->", len(new_c[key]))
    old_c = {canonical(n) for n in old}
    for name in sorted(new):
        if canonical(name) not in old_c:
            print("ADDED   ", name, len(new[name]))
    print(f"{ident} kernels identical")


if __name__ == "__main__":
    main()
