#!/bin/bash
# value / e2e against the number of extractor handles (streams with a batch in flight), with and without launch priority
set -u
mkdir -p gpurun_out
run() {
  name=$1; shift
  hh=$1; shift
  env "$@" timeout 600 python bench.py --handles $hh --no-cpu-baseline --e2e-repeats 1 --latency-frames 5 > gpurun_out/r02_h_bench_$name.json 2> gpurun_out/r02_h_bench_$name.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r02_h_bench_$name.json"))
print("$name: value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "serial", round(d["roofline"]["stage_ms_per_batch"]["serial_step_total"], 3), "chain", round(d["chained_flow"]["e2e_frames_per_s"]))
PY
}
run h1 1 ORB_PRIO=0
run h2 2 ORB_PRIO=0
run h3 3 ORB_PRIO=0
run h6 6 ORB_PRIO=0
run h8 8 ORB_PRIO=0
run h8p 8 ORB_PRIO=1
run h2p 2 ORB_PRIO=1
