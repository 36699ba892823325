"""Summarise an `ncu --set full` capture into profiles/: a CSV of the metrics the notes quote and a JSON record
(dram bytes per member-step, git head of the captured build) that bench.py reads for `roofline.traffic`.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r02_ncu_k_run_c5 --member-steps 1184 \
        --cmd "python tools/short_run.py 296 4" [--kernel-index 0]
"""
import argparse
import csv
import hashlib
import json
import os
import subprocess
import sys


def csrc_sha256():
    """Hash of the CUDA sources the library is built from: a capture record is valid for a build with the same hash
    (bench.py refuses a record whose hash differs, whatever else was committed since)."""
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pnmol-experiments_b200", "csrc")
    h = hashlib.sha256()
    for name in sorted(os.listdir(root)):
        if name.endswith((".cu", ".cuh", ".h")):
            h.update(name.encode())
            h.update(open(os.path.join(root, name), "rb").read())
    return h.hexdigest()

KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "launch__", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct", "sm__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "sm__inst_executed_pipe_",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct", "sm__warps_active.avg.per_cycle_active",
        "smsp__average_warps_issue_stalled_", "smsp__issue_active.avg.pct", "sass__inst_executed_local", "sm__icc_request_hit_rate",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("out_prefix")
    ap.add_argument("--member-steps", type=int, required=True, help="member-steps (or steps) the captured launch processed")
    ap.add_argument("--cmd", default="")
    ap.add_argument("--kernel-index", type=int, default=0)
    ap.add_argument("--note", default="")
    args = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", args.rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2 + args.kernel_index]
    kname = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""
    head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True,
                          cwd=os.path.dirname(os.path.abspath(__file__))).stdout.strip()
    picked = [(h, u, v) for h, u, v in zip(hdr, units, vals)
              if any(h.startswith(k) or k in h for k in KEEP) and "not_issued" not in h and v != ""]
    with open(args.out_prefix + "_summary.csv", "w") as f:
        f.write("metric,unit,value\n")
        f.write(f"# ncu --set full --clock-control none --import-source on; {args.cmd}; kernel {kname}; {args.member_steps} member-steps "
                f"per launch; git head {head}; csrc sha256 {csrc_sha256()[:16]}; {args.note},,\n")
        for h, u, v in picked:
            f.write(f"{h},{u},{v}\n")
    def get(name):
        i = hdr.index(name)
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[units[i]]
        return float(vals[i]) * scale
    rec = {"git_head": head, "csrc_sha256": csrc_sha256(), "kernel": kname, "command": args.cmd, "member_steps_per_launch": args.member_steps,
           "dram_bytes_read": get("dram__bytes_read.sum"), "dram_bytes_write": get("dram__bytes_write.sum"),
           "dram_bytes_per_member_step": (get("dram__bytes_read.sum") + get("dram__bytes_write.sum")) / args.member_steps,
           "duration_ms_under_ncu": float(vals[hdr.index("gpu__time_duration.sum")]), "note": args.note}
    json.dump(rec, open(args.out_prefix + ".json", "w"), indent=1)
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
