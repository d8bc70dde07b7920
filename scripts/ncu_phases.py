import re, sys, collections
lines_file, src_file = sys.argv[1], sys.argv[2]
src=open(src_file).read().split('\n')
marks=[(i+1,l.strip()) for i,l in enumerate(src) if l.strip().startswith('// ---- ')]
first_kernel_line = next(i+1 for i,l in enumerate(src) if '__global__' in l)
agg=collections.OrderedDict()
for line in open(lines_file).read().split('\n')[1:]:
    m=re.match(r"(\S+):\s*(\d+) inst\s+([\d.]+)%\s+samples\s+([\d.]+)%",line)
    if not m: continue
    f,ln,inst,samp=m.group(1),int(m.group(2)),float(m.group(3)),float(m.group(4))
    if f!=src_file.split('/')[-1]: key='other:'+f
    elif ln<first_kernel_line: key='helpers (cov/sqrt/exp/col_off)'
    else:
        key='prologue'
        for mk,name in marks:
            if ln>=mk: key=name[:60]
    a=agg.setdefault(key,[0,0]); a[0]+=inst; a[1]+=samp
for k,v in agg.items(): print(f"{k:62s} inst {v[0]:5.1f}%  samples {v[1]:5.1f}%")
