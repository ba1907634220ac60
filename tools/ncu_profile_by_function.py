"""Per device function of a window kernel, from an exported source page of an ncu capture (`ncu -i rep --page source --csv > x.csv`):
samples, share, executed warp instructions, lanes per instruction and the top stall reasons.
   python tools/ncu_profile_by_function.py source.csv [lib.so] [windows]     (symbol ranges from `cuobjdump -elf` of the library that was profiled)"""
import csv, re, subprocess, sys
src = sys.argv[1]; lib = sys.argv[2] if len(sys.argv) > 2 else "daccord_b200/_build/libdaccord_b200.so"
nwin = float(sys.argv[3]) if len(sys.argv) > 3 else 0
rows = list(csv.reader(open(src)))
kern = "dcus_window_kernel" if "dcus_window_kernel" in rows[0][1] else "dcu_window_kernel"
ns = "dcus" if "dcus" in kern else "dcu"
# lib: the library that was profiled, or a text file with its `cuobjdump -elf | grep '$_ZN'` lines saved next to the capture (the symbol ranges must be those of the profiled code)
elf = open(lib).read() if lib.endswith(".txt") else subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
funcs = []
for line in elf.splitlines():
    m = re.match(r"\s*0x[0-9a-f]+\s+(0x[0-9a-f]+)\s+(0x[0-9a-f]+)\s+0x2\s+0\s+0x[0-9a-f]+\s+\$.*\d+%s(?:I\w+?E)?E.*\$_ZN\d+%s\d+([A-Za-z_0-9]+?)E" % (kern, ns), line)
    if m:
        funcs.append((int(m.group(1), 16), int(m.group(2), 16), re.sub(r"I[A-Z].*", "", m.group(3))))
funcs.sort()
hdr = rows[1]
ia = hdr.index("Address"); isamp = hdr.index("# Samples"); iex = hdr.index("Instructions Executed"); ith = hdr.index("Thread Instructions Executed")
stalls = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = int(rows[2][ia], 16)
agg = {}; tot = [0, 0, 0]; stot = {}
for r in rows[2:]:
    off = int(r[ia], 16) - base
    name = "kernel_main"
    for o, s, n in funcs:
        if o <= off < o + s:
            name = n
    a = agg.setdefault(name, [0, 0, 0, 0, {}])
    a[0] += int(r[isamp]); a[1] += int(r[iex]); a[2] += int(r[ith]); a[3] += 1
    tot[0] += int(r[isamp]); tot[1] += int(r[iex]); tot[2] += int(r[ith])
    for i, h in stalls:
        v = int(r[i] or 0)
        if v:
            a[4][h] = a[4].get(h, 0) + v; stot[h] = stot.get(h, 0) + v
print("%s: %d samples, %d warp instructions, %.1f lanes/inst%s" % (kern, tot[0], tot[1], tot[2] / max(tot[1], 1), (", %.0f warp instructions / window" % (tot[1] / nwin)) if nwin else ""))
print("stalls: " + ", ".join("%s %.1f%%" % (h[6:], 100 * v / tot[0]) for h, v in sorted(stot.items(), key=lambda x: -x[1])[:8]))
print("%-22s %8s %6s %12s %6s %6s %6s  %s" % ("function", "samples", "%", "inst_exec", "%", "lanes", "sass", "top stalls"))
for n, a in sorted(agg.items(), key=lambda x: -x[1][0]):
    top = ", ".join("%s %.0f%%" % (h[6:], 100 * v / max(a[0], 1)) for h, v in sorted(a[4].items(), key=lambda x: -x[1])[:3])
    print("%-22s %8d %6.1f %12d %6.1f %6.1f %6d  %s" % (n, a[0], 100 * a[0] / tot[0], a[1], 100 * a[1] / tot[1], a[2] / max(a[1], 1), a[3], top))
