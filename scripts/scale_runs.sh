#!/bin/bash
# multi-GPU measurement batch (run under gpurun --gpus 8): $1 = list of "N:extra args" items separated by ';'
run() { # N tag args...
  n=$1; tag=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $n "$@" 2>gpurun_out/r2_${tag}.err | grep "^{" > gpurun_out/r2_${tag}.json
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.load(open("gpurun_out/r2_%s.json" % tag))
    r = d["roofline"]
    print(tag, "N", d["n_gpus"], "value %.4e" % d["value"], "ms/step %.3f" % d["ms_per_step"], "step kernel %.3f" % r["kernel_ms"], r.get("other_kernels_ms"), r.get("phase_ms_per_step"),
          "e2e %.3e" % d["e2e"]["value"], "nvlink", {k: v for k, v in (r.get("nvlink") or {}).items() if k.startswith("measured")}, "eval", (d.get("eval") or {}).get("value"), (d.get("eval") or {}).get("full_sweep"))
except Exception as e:
    print(tag, "ERR", e, open("gpurun_out/r2_%s.err" % tag).read()[-1500:])
PY
}
case "$1" in
  eight)
    run 8 n8_full --steps 20 --warmup 5
    run 8 n8_d64 --steps 20 --warmup 5 --dim 64 --no-cpu-baseline --no-eval-full --phases
    run 8 n8_zipf --steps 20 --warmup 5 --item-popularity zipf --no-cpu-baseline --eval-users 0 --phases
    ;;
  eight_full)
    run 8 n8_full --steps 20 --warmup 5
    ;;
  four)
    run 4 n4 --steps 20 --warmup 5 --no-cpu-baseline --no-eval-full --phases
    ;;
  two)
    run 2 n2 --steps 20 --warmup 5 --no-cpu-baseline --no-eval-full --phases
    ;;
esac
