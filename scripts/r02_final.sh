#!/bin/bash
# end-of-round check: the whole GPU tier, both bench arms, the ncu launch list of the final kernels
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r02_pytest_final.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
timeout 600 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_final.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02_bench_final.json")); r = json.load(open("gpurun_out/r02_bench_ref.json"))
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ref", round(r["value"]), "cores", r["cpu_baseline"]["cores"], "ratio e2e", round(d["e2e"]["value"] / r["value"], 1))
print({k: round(x, 3) for k, x in d["roofline"]["stage_ms_per_batch"].items()})
print("lba", d["lba"].get("ms_per_solve_e2e"), "liba", d["inertial_ba"].get("ms_per_solve_e2e"), d["inertial_ba"].get("ms_per_solve_batch8_e2e"))
print("chain", d["chained_flow"].get("e2e_frames_per_s"), "latency", d["latency_b1"]["headline_step"]["graph_replay_device_resident"]["p50_ms"], d["latency_b1"]["headline_step"]["e2e_host_buffers"]["p50_ms"])
PY
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -c 2500 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-repeats 1 --latency-frames 5 > gpurun_out/r02_launches_final.log 2>&1
echo "launches rc=$?"; wc -l gpurun_out/r02_launches_final.csv
# the side kernels (LocalBundleAdjustment, LocalInertialBA, brute-force 2-NN, isInFrustum, the chained-flow glue, the keyframe packer)
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -k regex:"k_lba|k_liba|k_hamming|k_in_frustum|k_chain|k_pack|k_bow|k_pose_edges" -c 400 --csv --log-file gpurun_out/r02_launches_side.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-repeats 1 --latency-frames 2 > gpurun_out/r02_launches_side.log 2>&1
echo "side rc=$?"; wc -l gpurun_out/r02_launches_side.csv
