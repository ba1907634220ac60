"""Per-function aggregation of an ncu --set full --import-source on report: samples, executed instructions and active lanes per device
function of window_core.cuh, by symbol range from `cuobjdump -elf`.   python tools/ncu_by_function.py report.ncu-rep daccord_b200/_build/libdaccord_b200.so"""
import csv,sys,subprocess,re
rep=sys.argv[1]; lib=sys.argv[2]
out=subprocess.run(['cuobjdump','-elf',lib],capture_output=True,text=True).stdout
funcs=[]
for line in out.splitlines():
    m=re.match(r'\s*0x[0-9a-f]+\s+(0x[0-9a-f]+)\s+(0x[0-9a-f]+)\s+0x2\s+0\s+0x[0-9a-f]+\s+\$.*\$_ZN3dcu\d+([a-z_]+)E',line)
    if m: funcs.append((int(m.group(1),16),int(m.group(2),16),m.group(3)))
funcs.sort()
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hdr=rows[1]; ia=hdr.index('Address'); isamp=hdr.index('# Samples'); iex=hdr.index('Instructions Executed'); ith=hdr.index('Thread Instructions Executed')
base=int(rows[2][ia],16)
agg={}
tot=[0,0,0]
for r in rows[2:]:
    off=int(r[ia],16)-base
    name='kernel_main'
    for (o,s,n) in funcs:
        if o<=off<o+s: name=n
    a=agg.setdefault(name,[0,0,0,0])
    a[0]+=int(r[isamp]); a[1]+=int(r[iex]); a[2]+=int(r[ith]); a[3]+=1
    tot[0]+=int(r[isamp]); tot[1]+=int(r[iex]); tot[2]+=int(r[ith])
print("%-20s %8s %6s %12s %6s %8s %6s"%("function","samples","%","inst_exec","%","lanes/inst","sass"))
for n,a in sorted(agg.items(),key=lambda x:-x[1][0]):
    print("%-20s %8d %6.1f %12d %6.1f %8.1f %6d"%(n,a[0],100*a[0]/tot[0],a[1],100*a[1]/tot[1],a[2]/max(a[1],1),a[3]))
print("total", tot)
