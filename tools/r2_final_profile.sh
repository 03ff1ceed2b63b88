#!/bin/bash
# Final evidence run of a round: full GPU suite, the bench line, the ncu launch list of the bench command, DRAM traffic of
# the dominant conv family on THIS build (-> gpurun_out/r02_traffic.json, keyed by the library hash) and one --set full capture.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
if [ "${SUITE:-1}" = "1" ]; then
  timeout -k 10 1500 python -m pytest tests -m gpu -q -n 4 --timeout 900 > gpurun_out/f_gpu_suite.log 2>&1; tail -3 gpurun_out/f_gpu_suite.log | cut -c1-200
else
  timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "tail_kernel or batch32" > gpurun_out/f_gpu_suite.log 2>&1; tail -3 gpurun_out/f_gpu_suite.log | cut -c1-200
fi
timeout 900 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; cut -c1-220 gpurun_out/f_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/f_launches_bench.csv python bench.py --steps 2 --warmup 1 --quick > gpurun_out/f_ncu_bench.log 2>&1; wc -l gpurun_out/f_launches_bench.csv
# launches per step (the fused MRF stage is the last one; before it: up3, 6 ResBlock convs, up2, 6 ResBlock convs, up1, conv_pre)
NL=$(PIPER_B200_GRAPH=0 timeout 300 python tools/ncu_step.py 1 | sed -n 's/.*launches so far \([0-9]*\).*/\1/p' | tail -1)
echo "launches per step: $NL"
# DRAM traffic of the layer-wise generator ResBlock launches (the 13 launches NL-15 .. NL-3 of a step, the second upsampler in the middle)
PIPER_B200_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --launch-skip $((NL + NL - 15)) --launch-count 13 --csv --log-file gpurun_out/f_traffic.csv python tools/ncu_step.py 2 > gpurun_out/f_ncu_traffic.log 2>&1
python - <<'PY'
import csv, hashlib, json
rows = list(csv.reader(open("gpurun_out/f_traffic.csv")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]; idc = h.index("ID"); kn = h.index("Kernel Name"); mn = h.index("Metric Name"); mu = h.index("Metric Unit"); mv = h.index("Metric Value")
per = {}
for r in rows[hdr + 1:]:
    if len(r) <= mv: continue
    v = float(r[mv].replace(",", ""))
    unit = r[mu].lower()
    if "byte" in unit:
        v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
    per.setdefault(r[idc], {"kernel": r[kn][:70]})[r[mn]] = v
launches = [per[k] for k in sorted(per, key=int)]
rb = [l for i, l in enumerate(launches) if i != 6]            # index 6 of the window is the second upsample conv
tot = sum(l.get("dram__bytes_read.sum", 0) + l.get("dram__bytes_write.sum", 0) for l in rb)
sha = hashlib.sha256(open("piper_b200/libpiper_b200.so", "rb").read()).hexdigest()[:16]
out = {"lib_sha16": sha, "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum, 12 layer-wise generator ResBlock launches of one step (tools/r2_final_profile.sh)",
       "families": {"dec.rb.mma": {"dram_bytes_per_launch": tot / max(len(rb), 1), "launches": len(rb)}}, "launch_list": launches}
json.dump(out, open("gpurun_out/r02_traffic.json", "w"), indent=1)
print("traffic per launch", tot / max(len(rb), 1) / 1e6, "MB over", len(rb), "launches; lib", sha)
PY
PIPER_B200_GRAPH=0 timeout 600 ncu --set full --import-source on --clock-control none --launch-skip $((NL + NL - 3)) --launch-count 1 -f -o gpurun_out/f_full_rb2k7 python tools/ncu_step.py 2 > gpurun_out/f_ncu_full.log 2>&1
ncu -i gpurun_out/f_full_rb2k7.ncu-rep --page raw --csv > gpurun_out/f_full_rb2k7.raw.csv 2>/dev/null
ncu -i gpurun_out/f_full_rb2k7.ncu-rep --page details --csv > gpurun_out/f_full_rb2k7.details.csv 2>/dev/null
ls -la gpurun_out/f_full_rb2k7.ncu-rep | awk '{print $5}'
if [ $(stat -c %s gpurun_out/f_full_rb2k7.ncu-rep) -gt 30000000 ]; then rm gpurun_out/f_full_rb2k7.ncu-rep; fi
