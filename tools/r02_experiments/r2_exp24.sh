#!/bin/bash
# hit kernel: prefetch of the next unit's mask rows to L2 (NTC_MASK_PREFETCH), A/B on config 2 and k=64/s=11
mkdir -p gpurun_out
for p in 0 1 0 1; do
  NTC_MASK_PREFETCH=$p timeout 300 python bench.py --steps 200 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_e24_p$p.json 2> gpurun_out/r2_e24_p$p.err
  python - <<PY
import json; d=json.loads(open("gpurun_out/r2_e24_p$p.json").read().strip().splitlines()[-1])
print("prefetch=$p value %.4e ms/step %.4f scan %.4f hit %.4f apply %.4f" % (d["value"], d["ms_per_step"], d.get("scan_ms",0), d.get("hit_ms",0), d.get("apply_ms",0)))
PY
done
for p in 0 1; do
  NTC_MASK_PREFETCH=$p timeout 300 python bench.py --workload k64s11 --steps 100 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_e24_k64_p$p.json 2> gpurun_out/r2_e24_k64_p$p.err
  python - <<PY
import json; d=json.loads(open("gpurun_out/r2_e24_k64_p$p.json").read().strip().splitlines()[-1])
print("k64s11 prefetch=$p value %.4e ms/step %.4f scan %.4f hit %.4f apply %.4f" % (d["value"], d["ms_per_step"], d.get("scan_ms",0), d.get("hit_ms",0), d.get("apply_ms",0)))
PY
done
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "at_size or golden or fused" > gpurun_out/r2_e24_pytest.log 2>&1; tail -2 gpurun_out/r2_e24_pytest.log
