#!/bin/bash
# copies the evidence of the last tools/gpu_round2_final.sh run (gpurun_out/r2z_*) into profiles/ and prints the numbers DESIGN.md quotes
set -u
T=gpurun_out/r2z
cp ${T}_bench_notebook.json profiles/r2_bench_notebook.json
cp ${T}_bench_test_suite.json profiles/r2_bench_test_suite.json
cp ${T}_bench_reference_arm.json profiles/r2_bench_reference_arm.json
cp ${T}_launches.csv profiles/r2_launches.csv
cp ${T}_configs.json profiles/r2_configs.json
ncu -i ${T}_warp.ncu-rep --page details > profiles/r2_warp_final_ncu_details.txt 2>/dev/null
python tools/ncu_warp_stalls.py ${T}_warp.ncu-rep > profiles/r2_warp_final_stalls_by_window.txt 2>/dev/null
ncu -i ${T}_tiny.ncu-rep --page details > profiles/r2_tiny_final_ncu_details.txt 2>/dev/null
python tools/ncu_lines_exec.py ${T}_tiny.ncu-rep "qpc_tiny_tick_kernelILi320ELi128ELi2" 40 > profiles/r2_tiny_final_instructions_by_line.txt 2>&1
python - <<PY
import csv, json, subprocess
out = subprocess.run(["ncu", "-i", "${T}_warp.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); h = rows[0]; r = rows[2]
rd = float(r[h.index("dram__bytes_read.sum")]); wr = float(r[h.index("dram__bytes_write.sum")])
u = rows[1][h.index("dram__bytes_read.sum")]
scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[u]
uw = rows[1][h.index("dram__bytes_write.sum")]
rd, wr = rd * scale, wr * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[uw]
d = json.load(open("profiles/admm_warp_traffic.json"))
d.update(dram_bytes_read=int(rd), dram_bytes_write=int(wr), dram_bytes_per_launch=int(rd + wr), bytes_per_solve=int((rd + wr) / 16384),
         source="ncu --set full (profiles/r2_warp_final_ncu_details.txt), one launch, 16384 solves, notebook settings, final build of round 2",
         note="assembled QP read once from HBM: G 10.2 KB, the contact blocks of P_bb, diagonal of P_aa, vectors; compulsory traffic of the tick is 1,632 B/solve")
json.dump(d, open("profiles/admm_warp_traffic.json", "w"), indent=1)
print("traffic", d["dram_bytes_per_launch"] / 1e6, "MB per launch,", d["bytes_per_solve"], "B per solve")
for n in ("notebook", "test_suite"):
    b = json.loads(open("profiles/r2_bench_%s.json" % n).read().strip().splitlines()[-1])
    print(n, "value %.2f M  %.3f ms  e2e %.2f M  stages %s  frac %.3f  flops/solve %.0f  iters %.1f/%s nfac %.2f  seq %.3f ms %.2f M (%.0f its)  lat %s  cpu %.2f k  c4 %.2f M %.2f ms" % (
        b["value"] / 1e6, b["ms_per_step"], b["e2e"]["value"] / 1e6, {k: round(v, 3) for k, v in b["stage_ms"].items()}, b["roofline"]["frac"],
        b["roofline"]["flops_per_solve"], b["roofline"]["iters_mean"], b.get("iters_max"), b["roofline"]["factorizations_mean"],
        b["sequential_ticks"]["ms_per_tick"], b["sequential_ticks"]["solves_per_s"] / 1e6, b["sequential_ticks"]["iters_mean_last_tick"],
        {k: round(v) for k, v in b["latency_us_by_batch"].items()}, b["cpu_baseline"]["value"] / 1e3,
        b["config4_strong_split"]["value"] / 1e6, b["config4_strong_split"]["ms_per_step"]))
c = json.load(open("profiles/r2_configs.json"))
c1 = c["config1_atlas_single_instance"]
print("config1 cpu cold %.0f warm %.0f us, gpu %.0f us" % (c1["cpu_oracle_1_thread_cold_us_per_solve"], c1["cpu_oracle_1_thread_warm_repeat_us_per_solve"], c1["gpu_batch_of_one_latency_us"]))
print("config2 %.2f ms %.0f M" % (c["config2_acrobot_point_task"]["ms_per_tick"], c["config2_acrobot_point_task"]["solves_per_s"] / 1e6))
c4 = c["config4_atlas_contact_masks"]
print("config4 %.2f ms %.2f M iters %.1f max %d acc %.4f" % (c4["ms_per_tick"], c4["solves_per_s"] / 1e6, c4["iters_mean"], c4["iters_max"], c4["accepted_frac"]))
for r in c["config5_dense_qp_sweep"]:
    print("config5", r["n"], r["m"], r["batch"], "%.0f solves/s" % r["solves_per_s"], "%.3f" % r["frac_of_dfma_peak"])
PY
grep -E "Duration|Issue Slots Busy|Registers Per|Achieved Occ" profiles/r2_warp_final_ncu_details.txt profiles/r2_tiny_final_ncu_details.txt
