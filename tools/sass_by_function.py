"""Per device function of one kernel of the library: SASS instruction mix (local / shared / generic loads and stores).
   python tools/sass_by_function.py [lib.so] [kernel substring]     (symbol ranges from `cuobjdump -elf`, instructions from `cuobjdump -sass`)"""
import re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "daccord_b200/_build/libdaccord_b200.so"
kern = sys.argv[2] if len(sys.argv) > 2 else "dcus_window_kernel"
ns = "dcus" if "dcus" in kern else "dcu"
elf = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
funcs = []
for line in elf.splitlines():
    m = re.match(r"\s*0x[0-9a-f]+\s+(0x[0-9a-f]+)\s+(0x[0-9a-f]+)\s+0x2\s+0\s+0x[0-9a-f]+\s+\$.*%s.*\$_ZN\d+%s\d+([A-Za-z_0-9]+?)E" % (kern, ns), line)
    if m:
        funcs.append((int(m.group(1), 16), int(m.group(2), 16), re.sub(r"I.*", "", m.group(3))))
funcs.sort()
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
part = [p for p in re.split(r"\n\s*Function : ", sass) if p.split("\n")[0].find(kern) >= 0][0]
agg = {}
for m in re.finditer(r"/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", part):
    off = int(m.group(1), 16); op = m.group(2)
    name = "kernel_main"
    for o, s, n in funcs:
        if o <= off < o + s:
            name = n
    a = agg.setdefault(name, {})
    a[op] = a.get(op, 0) + 1
cols = ["LDL", "STL", "LDS", "STS", "ATOMS", "LD", "ST", "LDG", "STG", "ATOM", "ATOMG", "UBLKCP", "SHFL", "BAR"]
print("%-22s %6s " % ("function", "sass") + " ".join("%6s" % c for c in cols))
tot = {}
for n, a in sorted(agg.items(), key=lambda x: -sum(x[1].values())):
    print("%-22s %6d " % (n, sum(a.values())) + " ".join("%6d" % a.get(c, 0) for c in cols))
    for c, v in a.items():
        tot[c] = tot.get(c, 0) + v
print("%-22s %6d " % ("total", sum(tot.values())) + " ".join("%6d" % tot.get(c, 0) for c in cols))
