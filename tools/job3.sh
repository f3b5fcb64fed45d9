mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_eig_one.csv python tools/eig_one.py > gpurun_out/ncu_eig_one.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02_launches_eig_one.csv")) if len(r)>5 and r[0].isdigit()]
# columns: ID, Process ID, Process Name, Host Name, Kernel Name, ..., Metric Name, Metric Unit, Metric Value
for r in rows[-24:]:
    print(r[4][:60], r[-2], r[-1])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 260 --csv --log-file gpurun_out/r02_launches_c3_v3.csv python bench.py --steps 4 --warmup 3 --no-cpu --no-sampler > gpurun_out/ncu_bench.log 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r02_launches_c3_v3.csv")) if len(r)>5 and r[0].isdigit()]
names=[r[4] for r in rows]
# find one step: from one k_dgemm<1,1> merge... print a contiguous window of ~40 launches in the middle
mid=len(rows)//2
# locate first 'k_psi' after mid and print 30 launches before its preceding merge through next psi
idx=[i for i,n in enumerate(names) if 'k_psi' in n]
i0=[i for i in idx if i>=mid][0]
i1=[i for i in idx if i>i0][0]
tot=0
for r in rows[i0-1:i1-1]:
    print(r[4][:70], r[-1]); tot+=float(r[-1].replace(',',''))
print('step total ns', tot, 'launches', i1-i0)
PY
