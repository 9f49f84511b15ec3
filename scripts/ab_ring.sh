#!/bin/bash
# A/B of the two K3 kernels on the bench line (run under gpurun): CRB_BPR_RING=0 register-staged, =1 bulk-copy ring
for ring in 0 1; do
  for opt in Adam SGD; do
    CRB_BPR_RING=$ring timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --eval-users 0 --optimizer $opt 2>/dev/null | grep "^{" > gpurun_out/r2_ring${ring}_${opt}.json
  done
done
python - <<'PY'
import json
for ring in (0, 1):
    for opt in ("Adam", "SGD"):
        try:
            d = json.loads(open("gpurun_out/r2_ring%d_%s.json" % (ring, opt)).read())
            r = d["roofline"]
            print("ring", ring, opt, "value %.4e" % d["value"], "ms/step %.3f" % d["ms_per_step"], "K3 %.3f" % r["kernel_ms"], "frac %.3f" % r["frac"], "step_frac %.3f" % r["step_frac"], r.get("other_kernels_ms"))
        except Exception as e:
            print("ring", ring, opt, "ERR", e)
PY
