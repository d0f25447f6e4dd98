import csv
rows=list(csv.reader(open("gpurun_out/s15_qp.csv")))
h=next(i for i,r in enumerate(rows) if r and r[0]=="ID")
ki=rows[h].index("Kernel Name"); mi=rows[h].index("Metric Name"); vi=rows[h].index("Metric Value")
for r in rows[h+1:]:
    if len(r)>vi: print(r[0], r[ki].split("(")[0], r[mi], r[vi])
