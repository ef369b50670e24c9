"""Condense an ncu --metrics gpu__time_duration.sum --csv launch list: kernel name, grid, time (us), for the LAST call
(the kernels after the last k_init).  usage: python profiles/launch_list.py file.csv"""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
last = max(i for i, r in enumerate(rows) if 'k_init' in r[4])
tot = 0.0
for r in rows[last:]:
    name = r[4].split('(')[0].replace('void ', '').replace('nvnl::', '')[:48]
    us = float(r[-1]) / 1000.0
    tot += us
    print('%-50s grid %-14s %9.1f us' % (name, r[8], us))
print('sum of kernels %.1f us' % tot)
