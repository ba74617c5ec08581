# usage: r2_ncu_list.sh <tag> [bench args] -- per-launch durations of one short bench run (cold-cache, serialised)
TAG=$1; shift
GCM_B200_CONCURRENT=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e "$@" > gpurun_out/ncu_list_$TAG.log 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_$TAG.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); 
agg=collections.OrderedDict()
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    k=r[ki][:70]
    a=agg.setdefault(k,[0,0.0,1e30,0]); a[0]+=1; a[1]+=v; a[2]=min(a[2],v); a[3]=max(a[3],v)
for k,a in agg.items(): print('%-72s n=%4d avg=%10.1f us min=%10.1f max=%10.1f' % (k,a[0],a[1]/a[0]/1e3,a[2]/1e3,a[3]/1e3))
PY
