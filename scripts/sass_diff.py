"""Development: are two builds of libalphagpu.so the same GPU code?  Compares every kernel's SASS instruction by instruction, and again
with register names masked (ptxas permutes registers from run to run on identical input).

    python scripts/sass_diff.py old/libalphagpu.so alphagpu_b200/libalphagpu.so

Used before committing source changes that are meant to leave the default library untouched (compile-time development variants)."""
import re
import subprocess
import sys


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    d, cur = {}, None
    for line in out.splitlines():
        if "Function :" in line:
            cur = line.split("Function :")[1].strip()
            d[cur] = []
        elif cur and re.match(r"^\s+/\*[0-9a-f]{4,6}\*/", line):
            d[cur].append(re.sub(r"\s+", " ", re.sub(r"/\*.*?\*/", "", line)).strip())
    return d


def main():
    a, b = kernels(sys.argv[1]), kernels(sys.argv[2])
    only = set(a) ^ set(b)
    mask = lambda ls: [re.sub(r"\bU?[RP]\d+\b", "r", x) for x in ls]
    exact = [k for k in a if k in b and a[k] == b[k]]
    modulo = [k for k in a if k in b and a[k] != b[k] and mask(a[k]) == mask(b[k])]
    differ = [k for k in a if k in b and mask(a[k]) != mask(b[k])]
    print(f"{len(a)} / {len(b)} kernels; identical {len(exact)}, identical modulo register names {len(modulo)}, different {len(differ)}, "
          f"in one library only {len(only)}")
    for k in differ + sorted(only):
        print("  ", k[:160])
    sys.exit(1 if differ or only else 0)


if __name__ == "__main__":
    main()
