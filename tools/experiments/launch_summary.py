import csv,collections,sys
for f in sys.argv[1:]:
    rows=[r for r in csv.reader(open(f)) if len(r)>5 and r[0].isdigit()]
    d=collections.OrderedDict()
    for r in rows:
        k=r[4][:48]; d.setdefault(k,[]).append(float(r[-1]))
    print(f, ' '.join(f"{sorted(v)[len(v)//2]/1e3:.0f}" for k,v in d.items() if 'conv_tc' in k))
