import csv, collections, sys
path = sys.argv[1]
lines=[l for l in open(path) if not l.startswith('==')]
rd=csv.DictReader(lines)
agg=collections.defaultdict(lambda:[0,0.0])
for r in rd:
    if r.get('Metric Name')!='gpu__time_duration.sum': continue
    k=r['Kernel Name'].split('(')[0].replace('void ','').replace('cpb::','')
    v=float(r['Metric Value'].replace(',','')); u=r['Metric Unit']
    if u in ('nsecond','ns'): v/=1000
    elif u in ('msecond','ms'): v*=1000
    agg[k][0]+=1; agg[k][1]+=v
tot=sum(v[1] for v in agg.values())
print("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|")
for k,v in sorted(agg.items(), key=lambda x:-x[1][1])[:28]:
    print(f"| `{k[:60]}` | {v[0]} | {v[1]:.1f} | {v[1]/v[0]:.1f} | {100*v[1]/tot:.1f}% |")
