"""Static opcode mix of the loops of one kernel in build/blsgpu.cubin:  python tools/loopmix.py <mangled-substring> [cubin]"""
import collections, re, subprocess, sys
cubin = sys.argv[2] if len(sys.argv) > 2 else "build/blsgpu.cubin"
txt = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
for part in re.split(r"\n\s*Function : ", txt)[1:]:
    name = part.split("\n")[0]
    if sys.argv[1] not in name:
        continue
    ins = [(int(m.group(1), 16), m.group(2).strip()) for m in re.finditer(r"/\*([0-9a-f]{4})\*/\s+(.*?);", part)]
    print(name[:90], len(ins), "instructions")
    loops = []
    for a, s in ins:
        m = re.search(r"BRA\s+0x([0-9a-f]+)", s)
        if m and int(m.group(1), 16) < a:
            loops.append((int(m.group(1), 16), a))
    for lo, hi in sorted(loops):
        c = collections.Counter()
        for a, s in ins:
            if lo <= a <= hi:
                t = s.split()
                op = t[1] if t[0].startswith("@") else t[0]
                c[".".join(op.split(".")[:2]) if op.startswith("IMAD") else op.split(".")[0]] += 1
        n = sum(c.values())
        if n > 60:
            print(f"  loop {lo:#x}..{hi:#x} n={n}", dict(c.most_common(10)))
