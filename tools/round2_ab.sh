#!/bin/bash
# First GPU call of round 2: does each opt-in variant written at the end of round 1 keep parity, and what does it buy?
#   gpurun --timeout 1500 -- 'bash tools/round2_ab.sh'
# Everything lands in gpurun_out/ab_*.  Variants (all off by default):
#   PIPER_B200_UNI=1     uniform-issue TMA warps in conv_mma_persist_kernel (expected: the big one)
#   PIPER_B200_SMALL=1   double-buffered plan for small one-tile-per-CTA launches (batch-1 latency)
#   PIPER_B200_MMA=31    fused MRF stage kernel for the 32-channel generator stage (mrf_fused.cu, never run before)
#   PIPER_B200_LN2=1     LayerNorm with batched global loads (encoder.cu layernorm_kernel2)
#   PIPER_B200_POST2=1   conv_post with batched staging loads (vocoder_tail.cu conv_post_kernel2)
#   PIPER_B200_ATT2=1    attention with one 16-byte query broadcast per 4 FMAs (encoder.cu rel_attention_kernel2)
#   PIPER_B200_ATT3=1    attention on the tensor cores (att_mma.cu: fp16x3, two passes over 64-key blocks)
#   PIPER_B200_V2=1      second-generation conv kernel: uniform TMA issue + stacked [W_hi;W_lo] weights (conv_mma2.cu)
#   PIPER_B200_V2_PREC=f16  ... with FP16 hi/lo operands for every family (22 bits at K = 16: half the MMAs of tf32x3)
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== experimental parity tests" > gpurun_out/ab_tests.log
PIPER_B200_EXPERIMENTAL=1 timeout -k 10 900 python -m pytest tests/test_gpu_experimental.py -m gpu -q --timeout 600 >> gpurun_out/ab_tests.log 2>&1
echo "rc=$?" >> gpurun_out/ab_tests.log
tail -25 gpurun_out/ab_tests.log
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout -k 10 300 python bench.py --steps 10 --warmup 3 --quick > gpurun_out/ab_bench_$name.json 2> gpurun_out/ab_bench_$name.err
  python - "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/ab_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(f"{sys.argv[1]:12s} {d['value'] / 1e6:8.1f} M samples/s  {d['ms_per_step']:7.3f} ms  stages {d.get('stage_ms')}  batch1 {d.get('batch1', {}).get('latency_ms')}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run base PIPER_B200_NOP=1
run graph PIPER_B200_GRAPH=1
run uni PIPER_B200_UNI=1
run uni_fused PIPER_B200_UNI=1 PIPER_B200_MMA=31
run ln2_post2_att2 PIPER_B200_LN2=1 PIPER_B200_POST2=1 PIPER_B200_ATT2=1
run att3 PIPER_B200_ATT3=1
run v2 PIPER_B200_V2=1
run v2_f16 PIPER_B200_V2=1 PIPER_B200_V2_PREC=f16
run v2_tm PIPER_B200_V2=1 PIPER_B200_V2_TM=1
run v2_f16_tm_graph PIPER_B200_V2=2 PIPER_B200_V2_PREC=f16 PIPER_B200_V2_TM=1 PIPER_B200_GRAPH=1
run everything PIPER_B200_V2=2 PIPER_B200_V2_PREC=f16 PIPER_B200_MMA=31 PIPER_B200_LN2=1 PIPER_B200_POST2=1 PIPER_B200_ATT3=1   # v2 also for launches with < 148 tiles (batch-1 latency)
PIPER_B200_UNI=1 timeout -k 10 200 python tools/layer_report.py > gpurun_out/ab_layer_report_uni.txt 2>&1
tail -9 gpurun_out/ab_layer_report_uni.txt
PIPER_B200_V2=1 PIPER_B200_V2_PREC=f16 timeout -k 10 200 python tools/layer_report.py > gpurun_out/ab_layer_report_v2_f16.txt 2>&1
tail -9 gpurun_out/ab_layer_report_v2_f16.txt
