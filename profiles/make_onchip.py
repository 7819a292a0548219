"""profiles/r02_onchip.json and profiles/roofline_traffic.json from the committed ncu summaries of the benched
kernels (profiles/r02_ncu_k_refine_hd_{ss,cs3}.txt, written by profiles/summarize.py kernel <report>).
bench.py copies the entry of the dominant kernel into the `roofline.onchip` / `roofline.traffic` of its JSON line."""
import json, os, re

HERE = os.path.dirname(os.path.abspath(__file__))
PICK = {
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1_data_pipe_pct": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "fp64_pipe_pct": "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "xu_pipe_pct": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "alu_pipe_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "live_lanes_per_instruction": "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1_hit_rate_pct": "l1tex__t_sector_hit_rate.pct",
    "smem_wavefronts": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smem_bank_conflict_wavefronts": "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "registers_per_thread": "launch__registers_per_thread",
    "launch_ms_under_ncu": "gpu__time_duration.sum",
    "dram_read_mb": "dram__bytes_read.sum",
    "dram_write_mb": "dram__bytes_write.sum",
}


def parse(path):
    vals = {}
    for line in open(path):
        parts = line.split()
        if len(parts) >= 2 and not line.startswith("#"):
            try:
                vals[parts[0]] = float(parts[1].replace(",", ""))
            except ValueError:
                pass
    return vals


onchip, traffic = {}, {}
for key, name in (("ss", "r02_ncu_k_refine_hd_ss.txt"), ("cs", "r02_ncu_k_refine_hd_cs3.txt")):
    path = os.path.join(HERE, name)
    if not os.path.exists(path):
        continue
    v = parse(path)
    entry = {k: v.get(m) for k, m in PICK.items()}
    entry["source"] = f"profiles/{name} (ncu --set full --clock-control none, one k_refine launch of the HD step)"
    onchip["k_refine_" + key] = entry
    traffic[key] = int(round((v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"]) * 1e6))
json.dump(onchip, open(os.path.join(HERE, "r02_onchip.json"), "w"), indent=1)
traffic["note"] = "dram__bytes_read.sum + dram__bytes_write.sum of ONE k_refine launch at HD (ss: single-scale, cs: S=3), from profiles/r02_ncu_k_refine_hd_*.txt"
json.dump(traffic, open(os.path.join(HERE, "roofline_traffic.json"), "w"), indent=1)
print(json.dumps(onchip, indent=1)[:600], traffic)
