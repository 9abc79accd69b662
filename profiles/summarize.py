"""Turn an .ncu-rep (read here with `ncu -i`, no GPU needed) into the text summaries committed under profiles/.
    python profiles/summarize.py gpurun_out/r01_loss.ncu-rep profiles/r01_loss"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum [", "dram__bytes_write.sum [", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread [",
        "launch__occupancy_limit", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "smsp__inst_executed.sum [", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__average_warps_issue_stalled",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sass__inst_executed_local_loads", "sm__maximum_warps_per_active_cycle_pct", "lts__t_sector_hit_rate.pct"]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


def num(x):
    try:
        return int(float(x))
    except Exception:
        return 0


def main(rep, out):
    rows = list(csv.reader(ncu(rep, "--page", "raw", "--csv").splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    lines = [f"# {rep}: kernel {vals[hdr.index('Kernel Name')]}", "# (ncu --set full --clock-control none; cold-cache, serialised: compare shares, not absolutes)"]
    for h, u, v in zip(hdr, units, vals):
        key = f"{h} [{u}]"
        if any(w in key for w in WANT):
            lines.append(f"{key} = {v}")
    rows = list(csv.reader(ncu(rep, "--page", "source", "--csv", "--print-source", "cuda,sass").splitlines()))
    hdr = rows[2]
    i_s, i_i = hdr.index("# Samples"), hdr.index("Instructions Executed")
    cur, src = None, []
    for r in rows:
        if r and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif len(r) > i_i and r[0].isdigit():
            src.append((cur, int(r[0]), r[1].strip()[:110], num(r[i_s]), num(r[i_i])))
    ts, ti = sum(x[3] for x in src) or 1, sum(x[4] for x in src) or 1
    lines.append(f"\n# source lines by executed instructions (total {ti}) / by stall samples (total {ts})")
    for x in sorted(src, key=lambda x: -x[4])[:20]:
        lines.append(f"inst {100 * x[4] / ti:5.1f}%  samp {100 * x[3] / ts:5.1f}%  {x[0]}:{x[1]}  {x[2]}")
    lines.append("")
    for x in sorted(src, key=lambda x: -x[3])[:20]:
        lines.append(f"samp {100 * x[3] / ts:5.1f}%  inst {100 * x[4] / ti:5.1f}%  {x[0]}:{x[1]}  {x[2]}")
    open(out + "_summary.txt", "w").write("\n".join(lines) + "\n")
    print(out + "_summary.txt")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
