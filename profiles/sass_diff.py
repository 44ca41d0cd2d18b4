"""Structural SASS comparison of two builds of a library: proves that adding gated kernel variants left the validated kernels alone.

    cuobjdump -sass old/libshipyard_gemm.so > a.txt ; cuobjdump -sass new/libshipyard_gemm.so > b.txt
    python profiles/sass_diff.py a.txt b.txt

Per function (matched by mangled name; template-parameter suffixes added with default values can be mapped with --map OLD=NEW regexes)
the instruction streams are compared after normalising register numbers (ptxas renumbers uniform registers between otherwise
identical compilations).  Exit code 1 if any function present in both dumps differs structurally.
"""
import re
import sys


def funcs(path):
    out = {}
    for f in re.split(r"\n\s*Function : ", open(path).read())[1:]:
        name, _, body = f.partition("\n")
        out[name.strip()] = [re.sub(r"/\* 0x[0-9a-f]+ \*/", "", ln).rstrip() for ln in body.split("\n") if "/*" in ln]
    return out


def norm(line):
    return re.sub(r"\b(UR|R|UP|P)\d+\b", r"\1#", line)


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--map")]
    maps = [a.split("=", 1) for a in sys.argv[1:] if a.startswith("--map=") for a in [a[6:]]]
    a, b = funcs(args[0]), funcs(args[1])
    bad = same = missing = 0
    for name, body in a.items():
        new = name
        for pat, rep in maps:
            new = re.sub(pat, rep, new)
        other = b.get(new)
        if other is None:
            missing += 1
            continue
        d = sum(1 for x, y in zip(body, other) if norm(x) != norm(y)) + abs(len(body) - len(other))
        if d:
            bad += 1
            print(f"DIFF {d:5d} lines  {new[:140]}")
        else:
            same += 1
    print(f"{same} identical (modulo register numbering), {bad} different, {missing} not found in the second dump, "
          f"{len(b) - same - bad} only in the second dump")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
