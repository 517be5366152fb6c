"""Per-kernel averages of one `ncu --set full` capture -> profiles/<tag>_ncu_traffic.json (read by bench.py for
`roofline.traffic`) and a markdown table (stdout).  Runs in the authoring container (ncu -i needs no GPU).

    python tools/ncu_traffic.py gpurun_out/r2_final.ncu-rep r2 "bench.py --steps 4 --warmup 3 --eager (cfg2, 1xB200)"
"""
import collections
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_summary import load  # noqa: E402

KEY = {"project_kernel": "project", "tile_scan_kernel": "tile_scan", "sh_color16_kernel": "sh_color",
       "emit_kernel": "emit", "sort_pack_kernel": "sort_pack", "blend_fwd2_kernel": "blend_fwd",
       "blend_fwd_kernel": "blend_fwd", "blend_bwd2_kernel": "blend_bwd", "blend_bwd_kernel": "blend_bwd",
       "preprocess_bwd16_kernel": "preprocess_bwd", "photometric_l1_fwd_vec_kernel": "photometric_fwd",
       "photometric_l1_bwd_vec_kernel": "photometric_bwd", "color_fill_kernel": "color_fill"}
M = {"gpu__time_duration.sum": "duration_us", "dram__bytes_read.sum": "dram_bytes_read",
     "dram__bytes_write.sum": "dram_bytes_write", "sm__inst_executed.avg.per_cycle_active": "ncu_ipc",
     "sm__instruction_throughput.avg.pct_of_peak_sustained_active": "ncu_issue_slot_pct",
     "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "ncu_dram_pct",
     "smsp__thread_inst_executed_per_inst_executed.ratio": "ncu_active_threads_per_inst",
     "sm__warps_active.avg.pct_of_peak_sustained_active": "ncu_occupancy_pct",
     "launch__registers_per_thread": "registers", "smsp__inst_executed.sum": "warp_instructions",
     "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
     "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
     "lts__t_sector_hit_rate.pct": "l2_hit_pct"}
SCALE = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3}


def main():
    path, tag, what = sys.argv[1], sys.argv[2], sys.argv[3]
    hdr, units, data = load(path)
    idx = {h: i for i, h in enumerate(hdr)}
    acc = collections.defaultdict(lambda: collections.defaultdict(list))
    for r in data:
        name = r[idx["Kernel Name"]].split("(")[0].replace("gg::", "").replace("void ", "").split("<")[0]
        key = KEY.get(name)
        if key is None:
            continue
        for m, out in M.items():
            if m not in idx or r[idx[m]] in ("", "n/a"):
                continue
            v = float(r[idx[m]].replace(",", ""))
            u = units[idx[m]]
            if out.startswith("dram_bytes") or out == "duration_us":
                v *= SCALE.get(u, 1.0)
            acc[key][out].append(v)
    kernels = {}
    for key, ms in acc.items():
        d = {k: sum(v) / len(v) for k, v in ms.items()}
        d["launches_averaged"] = len(next(iter(ms.values())))
        d["traffic"] = d.get("dram_bytes_read", 0.0) + d.get("dram_bytes_write", 0.0)
        kernels[key] = {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d.items()}
    out = {"source": f"ncu --set full --clock-control none, {what}; see profiles/{tag}_ncu_summary.md", "kernels": kernels}
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", f"{tag}_ncu_traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    cols = ["duration_us", "traffic", "ncu_dram_pct", "ncu_ipc", "ncu_issue_slot_pct", "ncu_active_threads_per_inst",
            "ncu_occupancy_pct", "registers", "warp_instructions", "smem_bank_conflicts", "smem_wavefronts", "l2_hit_pct",
            "launches_averaged"]
    print("| kernel | " + " | ".join(cols) + " |")
    print("|---|" + "---|" * len(cols))
    for key, d in sorted(kernels.items(), key=lambda kv: -kv[1].get("duration_us", 0)):
        print(f"| {key} | " + " | ".join(str(d.get(c, "")) for c in cols) + " |")


if __name__ == "__main__":
    main()
