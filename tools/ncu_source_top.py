"""Top stalled instructions of one kernel from `ncu --page source --csv` output.
Usage: python tools/ncu_source_top.py <source.csv> <kernel index> [min percent]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]); minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.8
starts = [i for i, r in enumerate(rows) if len(r) > 3 and r[0] == 'Address' and r[1] == 'Source']
starts.append(len(rows))
h = rows[starts[which]]; seg = rows[starts[which] + 1:starts[which + 1]]
if which > 0: print("kernel header:", rows[starts[which] - 1][:1])
iS = h.index('Warp Stall Sampling (All Samples)'); iSrc = h.index('Source'); iE = h.index('Instructions Executed')
st = [i for i, x in enumerate(h) if x.startswith('stall_') and 'Not Issued' not in x]
tot = sum(int(r[iS]) for r in seg if len(r) > iS and r[iS].isdigit())
print("total samples", tot, "instructions", len(seg), "executed warp-instr", sum(int(r[iE]) for r in seg if len(r) > iE and r[iE].isdigit()))
for k, r in enumerate(seg):
    if len(r) <= iS or not r[iS].isdigit(): continue
    s = int(r[iS])
    if s > tot * minpct / 100:
        top = sorted([(int(r[i]), h[i][6:]) for i in st if r[i].isdigit() and int(r[i]) > 0], reverse=True)[:2]
        print(k, "%4.1f%%" % (100 * s / tot), r[iE], r[iSrc][:64], top)
