"""Summarise an `ncu --set full` capture for profiles/:

    ncu -i X.ncu-rep --page raw --csv > X.csv
    python tools/ncu_summary.py X.csv profiles/NAME.md [--traffic B N heads]

Writes a markdown table of the metrics the roofline discussion uses (one section per kernel launch) and, with --traffic,
refreshes profiles/traffic.json: DRAM bytes per launch per attention kernel, keyed to the sha of the kernel sources so that
bench.py only reports the figure while the kernels are unchanged."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

METRICS = [
    "gpu__time_duration.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__inst_executed.sum",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg.per_second",
]

SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(src, errors="replace")))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units = rows[hi], rows[hi + 1]
    ix = {h: i for i, h in enumerate(hdr)}
    out = [f"# ncu --set full summary of `{os.path.basename(src)}`", ""]
    traffic = {}
    for r in rows[hi + 2:]:
        if len(r) < len(hdr):
            continue
        name = r[ix["Kernel Name"]]
        short = name.replace("void ", "").replace("<unnamed>::", "").split("(")[0].split("<")[0].split("::")[-1].strip()
        out += [f"## {short}", "", f"`{name[:160]}`", "", "| metric | value | unit |", "|---|---|---|"]
        for m in METRICS:
            if m in ix:
                out.append(f"| {m} | {r[ix[m]]} | {units[ix[m]]} |")
        out.append("")
        try:
            rd = float(r[ix["dram__bytes_read.sum"]].replace(",", "")) * SCALE.get(units[ix["dram__bytes_read.sum"]], 1.0)
            wr = float(r[ix["dram__bytes_write.sum"]].replace(",", "")) * SCALE.get(units[ix["dram__bytes_write.sum"]], 1.0)
            traffic[short] = rd + wr
        except Exception:
            pass
    open(dst, "w").write("\n".join(out))
    print("wrote", dst)
    if "--traffic" in sys.argv:
        import bench
        k = sys.argv.index("--traffic")
        B, N, heads = (int(v) for v in sys.argv[k + 1:k + 4])
        names = {"attn_fwd_tc_kernel": "attn_fwd", "attn_fwd_kernel": "attn_fwd", "attn_bwd_fused_kernel": "attn_bwd_fused", "attn_bwd_dkv_kernel": "attn_bwd_dkv",
                 "attn_bwd_dq_kernel": "attn_bwd_dq", "attn_fwd_v2_kernel": "attn_fwd", "attn_bwd_v3_kernel": "attn_bwd_fused"}
        kern = {names[s]: {"dram_bytes_per_launch": v, "B": B, "N": N, "heads": heads} for s, v in traffic.items() if s in names}
        path = os.path.join(ROOT, "profiles", "traffic.json")
        json.dump({"src_sha": bench.kernel_source_hash(), "from": os.path.basename(src), "kernels": kern}, open(path, "w"), indent=1)
        print("wrote", path, kern)


if __name__ == "__main__":
    main()
