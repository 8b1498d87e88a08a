run() {
env "$@" ncu --metrics gpu__time_duration.sum --clock-control none -c 16 --csv --log-file gpurun_out/l_x.csv python bench.py --no-cpu --no-e2e --steps 1 --warmup 3 > /dev/null 2>&1
python - "$*" <<PY
import csv,sys
rows=list(csv.reader(open("gpurun_out/l_x.csv")))
hi=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
out=[]
for r in rows[hi+9:hi+14]:
    out.append(r[4].split("(")[0].split("::")[-1][:12]+"="+r[-1])
print(sys.argv[1], ":", " ".join(out))
PY
}
run NTC_CHUNK_WAVES=0
run NTC_CHUNK_WAVES=0 NTC_NO_STAGE=1
run NTC_CHUNK_WAVES=0 NTC_PL_DEBUG=32
run NTC_CHUNK_WAVES=0 NTC_PL_DEBUG=64
