"""Stall reasons per device function from the source page of an ncu report (same symbol-range mapping as ncu_by_function.py).
   python tools/ncu_stalls_by_function.py report.ncu-rep daccord_b200/_build/libdaccord_b200.so"""
import csv,sys,subprocess,re
rep=sys.argv[1]; lib=sys.argv[2]
out=subprocess.run(['cuobjdump','-elf',lib],capture_output=True,text=True).stdout
funcs=[]
for line in out.splitlines():
    m=re.match(r'\s*0x[0-9a-f]+\s+(0x[0-9a-f]+)\s+(0x[0-9a-f]+)\s+0x2\s+0\s+0x[0-9a-f]+\s+\$.*\$_ZN3dcu\d+([a-z_0-9]+)E',line)
    if m: funcs.append((int(m.group(1),16),int(m.group(2),16),m.group(3)))
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hdr=rows[1]; ia=hdr.index('Address')
cols={k:hdr.index(k) for k in ['# Samples','stall_long_sb','stall_no_inst','stall_wait','stall_short_sb','stall_branch_resolving','stall_selected','stall_lg','stall_mio','stall_not_selected','stall_math','Instructions Executed']}
base=int(rows[2][ia],16)
agg={}
for r in rows[2:]:
    off=int(r[ia],16)-base
    name='kernel_main'
    for (o,s,n) in funcs:
        if o<=off<o+s: name=n
    a=agg.setdefault(name,{k:0 for k in cols})
    for k,i in cols.items(): a[k]+=int(r[i] or 0)
tot=sum(a['# Samples'] for a in agg.values())
print("%-18s %6s | %6s %6s %6s %6s %6s %6s | %8s"%("function","smp%","longSB","noInst","wait","shortSB","branch","select","cyc/inst"))
T={k:0 for k in cols}
for n,a in sorted(agg.items(),key=lambda x:-x[1]['# Samples'])[:22]:
    s=a['# Samples']
    print("%-18s %6.1f | %6.1f %6.1f %6.1f %6.1f %6.1f %6.1f | %8.1f"%(n,100*s/tot,100*a['stall_long_sb']/s,100*a['stall_no_inst']/s,100*a['stall_wait']/s,100*a['stall_short_sb']/s,100*a['stall_branch_resolving']/s,100*a['stall_selected']/s, 0))
for a in agg.values():
    for k in cols: T[k]+=a[k]
print("TOTAL long_sb %.1f%% no_inst %.1f%% wait %.1f%% short_sb %.1f%% branch %.1f%% selected %.1f%% lg %.1f%% mio %.1f%% notsel %.1f%% math %.1f%%"%tuple(100*T[k]/T['# Samples'] for k in ['stall_long_sb','stall_no_inst','stall_wait','stall_short_sb','stall_branch_resolving','stall_selected','stall_lg','stall_mio','stall_not_selected','stall_math']))
