"""Summarise an `ncu --set full --import-source on` report into a small JSON for profiles/ (read here, no GPU needed).

    python tools/ncu_summary.py gpurun_out/r02_c3_pass_f32.ncu-rep profiles/r02_pass_f32_c3_ncu_summary.json [--top 25]

Per kernel in the report: duration, DRAM bytes, pipe / issue utilisation, occupancy, the warp-stall breakdown summed over
the SASS lines, and the `top` SASS lines by stall samples (the "hot spots").
"""
import csv
import io
import json
import subprocess
import sys

RAW_KEYS = {
    "gpu__time_duration.sum": "duration_ns",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__registers_per_thread": "registers",
    "launch__shared_mem_per_block_dynamic": "smem_dynamic_bytes",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__inst_executed.avg.per_cycle_elapsed": "ipc_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slots_busy_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active": "pipe_fmaheavy_pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "pipe_fp64_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "pipe_xu_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active": "pipe_tensor_dmma_pct",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active": "pipe_tensor_dmma_pct",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active": "pipe_tensor_inst_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "pipe_tensor_cycles_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "smsp__inst_executed.sum": "warp_instructions",
}


def _run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True, check=True).stdout


def _num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return v


def summarise(rep, top):
    raw = list(csv.reader(io.StringIO(_run(["-i", rep, "--page", "raw", "--csv"]))))
    head, units = raw[0], raw[1]
    kernels = []
    for row in raw[2:]:
        k = {"kernel": row[head.index("Kernel Name")]}
        for i, name in enumerate(head):
            if name in RAW_KEYS and row[i] != "":
                k[RAW_KEYS[name]] = _num(row[i])
                if units[i] and RAW_KEYS[name] in ("duration_ns", "dram_read_bytes", "dram_write_bytes"):
                    k[RAW_KEYS[name] + "_unit"] = units[i]
            if "pipe_tensor" in name and name.endswith("pct_of_peak_sustained_active") and row[i] != "":
                k.setdefault("tensor_pipe_metrics", {})[name] = _num(row[i])
        kernels.append(k)
    # SASS page: one table per kernel, separated by a "Kernel Name" line
    src = _run(["-i", rep, "--page", "source", "--csv", "--print-source", "sass"])
    blocks, cur = [], None
    for row in csv.reader(io.StringIO(src)):
        if row and row[0] == "Kernel Name":
            cur = {"name": row[1], "head": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and row and row[0] == "Address":
            cur["head"] = row
        elif cur is not None and cur["head"] is not None and len(row) == len(cur["head"]):
            cur["rows"].append(row)
    if len(blocks) == 2 * len(kernels):          # ncu prints every kernel's table twice for multi-launch reports
        blocks = blocks[::2]
    for k, b in zip(kernels, blocks):
        h = b["head"]
        i_src, i_all, i_ex = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
        stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
        total = sum(int(r[i_all] or 0) for r in b["rows"]) or 1
        stalls = {h[i]: sum(int(r[i] or 0) for r in b["rows"]) for i in stall_cols}
        k["stall_samples_total"] = total
        k["stall_breakdown_pct"] = {n: round(100.0 * v / total, 2) for n, v in sorted(stalls.items(), key=lambda t: -t[1]) if v > 0}
        ranked = sorted(b["rows"], key=lambda r: -int(r[i_all] or 0))[:top]
        k["sass_hot_spots"] = [{"pct": round(100.0 * int(r[i_all] or 0) / total, 2), "executed": int(r[i_ex] or 0),
                                "sass": " ".join(r[i_src].split())} for r in ranked]
        mn = {}
        for r in b["rows"]:
            op = r[i_src].split()
            op = next((t for t in op if not t.startswith("@")), "?").split(".")[0]
            mn[op] = mn.get(op, 0) + int(r[i_ex] or 0)
        tot_ex = sum(mn.values()) or 1
        k["sass_mnemonic_mix_pct"] = {n: round(100.0 * v / tot_ex, 2) for n, v in sorted(mn.items(), key=lambda t: -t[1])[:16]}
    return kernels


def main():
    rep, out = sys.argv[1], sys.argv[2]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    res = {"report": rep.split("/")[-1], "command": "ncu --set full --import-source on --clock-control none (one launch per kernel, "
           "warm caches after 3 warm-up iterations)", "kernels": summarise(rep, top)}
    with open(out, "w") as f:
        json.dump(res, f, indent=1)
    for k in res["kernels"]:
        print(k["kernel"][:70], k.get("duration_ns"), "stalls:", list(k["stall_breakdown_pct"].items())[:4])


if __name__ == "__main__":
    main()
