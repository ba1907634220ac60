"""Samples per source line (nvdisasm -gi line table of the built library) from the source page of an ncu report.
   python tools/ncu_by_line.py report.ncu-rep [function substring]"""
import re,sys,csv,subprocess
rep=sys.argv[1]; fn=sys.argv[2] if len(sys.argv)>2 else None
# map: offset in kernel text section -> source line, from nvdisasm -gi
cur=None; off2line={}
insec=False
for l in open('/tmp/dis.txt'):
    m=re.search(r'//## File "([^"]+)", line (\d+)',l)
    if m: cur=(m.group(1).split('/')[-1],int(m.group(2))); continue
    if l.startswith('.section') or '.text.' in l[:40]:
        insec = '.text._ZN43' in l
    m=re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+[A-Z@!]',l)
    if m and insec: off2line[int(m.group(1),16)]=cur
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hdr=rows[1]; ia=hdr.index('Address'); isamp=hdr.index('# Samples'); iex=hdr.index('Instructions Executed')
base=int(rows[2][ia],16)
agg={}
for r in rows[2:]:
    off=int(r[ia],16)-base
    k=off2line.get(off)
    a=agg.setdefault(k,[0,0]); a[0]+=int(r[isamp]); a[1]+=int(r[iex])
tot=sum(a[1] for a in agg.values())
lines=open('/root/repo/daccord_b200/csrc/window_core.cuh').read().split('\n')
items=sorted(agg.items(),key=lambda x:-x[1][1])
for k,a in items[:int(sys.argv[3]) if len(sys.argv)>3 else 40]:
    txt = lines[k[1]-1].strip()[:110] if k and k[0]=='window_core.cuh' else str(k)
    print("%6.2f%% %6.2f%%smp  L%-5s %s"%(100*a[1]/tot, 100*a[0]/sum(x[0] for x in agg.values()), k[1] if k else '?', txt))
