run() {
env "$@" python bench.py --no-cpu --no-e2e --steps 5 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('$*', j['value'], j['ms_per_step'], j['roofline']['kernel_ms'])"
}
run NTC_CHUNK_WAVES=0
run NTC_CHUNK_WAVES=1
run NTC_CHUNK_WAVES=2
run NTC_CHUNK_WAVES=3
