#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
t0=$(date +%s)
timeout 900 python bench.py > gpurun_out/c12_cfg3.json 2> gpurun_out/c12_cfg3.err; echo "cfg3 full: $(( $(date +%s) - t0 )) s"; cut -c1-300 gpurun_out/c12_cfg3.json; tail -3 gpurun_out/c12_cfg3.err
t0=$(date +%s)
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c12_ref.json 2> gpurun_out/c12_ref.err; echo "reference arm: $(( $(date +%s) - t0 )) s"; cut -c1-400 gpurun_out/c12_ref.json
timeout 600 python bench.py --config 2 --no-cpu-baseline > gpurun_out/c12_cfg2.json 2> gpurun_out/c12_cfg2.err; tail -2 gpurun_out/c12_cfg2.err
timeout 900 python bench.py --config 4 --no-cpu-baseline > gpurun_out/c12_cfg4.json 2> gpurun_out/c12_cfg4.err; tail -2 gpurun_out/c12_cfg4.err
timeout 600 python bench.py --config 5 --arch medium > gpurun_out/c12_cfg5_medium.json 2> gpurun_out/c12_cfg5_medium.err; tail -2 gpurun_out/c12_cfg5_medium.err
timeout 900 python bench.py --config 5 --arch high > gpurun_out/c12_cfg5_high.json 2> gpurun_out/c12_cfg5_high.err; tail -2 gpurun_out/c12_cfg5_high.err
python - <<'PY'
import json
for n in ("cfg3", "cfg2", "cfg4", "cfg5_medium", "cfg5_high"):
    try:
        d = json.loads(open(f"gpurun_out/c12_{n}.json").read().strip().splitlines()[-1])
        print(n, f"{d['value'] / 1e6:.1f} M samples/s, {d['ms_per_step']:.3f} ms/step, e2e {d['e2e']['value'] / 1e6:.1f} M, stage_ms {d.get('stage_ms')}, batch1 {d.get('batch1', {}).get('latency_ms')}, streaming {d.get('streaming')}, roofline {d['roofline']['kernel']} bound {d['roofline']['bound']} frac {d['roofline']['frac']:.3f}, cpu {d.get('cpu_baseline')}")
        if 'sweep' in d:
            for s in d['sweep']: print("   ", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in s.items()})
        if n == "cfg3":
            for k, f in d['roofline']['families'].items(): print("    ", k, {kk: (round(vv, 3) if isinstance(vv, float) else vv) for kk, vv in f.items()})
    except Exception as e:
        print(n, "FAILED", e)
PY
