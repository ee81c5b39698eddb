"""Markdown summary of the kernels in an .ncu-rep (run where ncu is installed):
    python profiles/ncu_extract.py gpurun_out/x.ncu-rep [more.ncu-rep ...] > profiles/rNN_ncu_summary.md
Prints, per captured launch: duration, DRAM bytes, warp instructions, occupancy, issue-slot utilisation,
the stall reasons above 0.3 warps per issue, cache hit rates and the launch shape."""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / instruction"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor-pipe instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / cycle"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % (achieved occupancy)"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers), blocks"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (shared memory), blocks"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("launch__shared_mem_per_block_static", "static smem / block"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
]


def rows_of(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def main():
    for path in sys.argv[1:]:
        hdr, units, rows = rows_of(path)
        print(f"## `{path.split('/')[-1]}`\n")
        for r in rows:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            print(f"### `{d['Kernel Name']}`  (launch id {d.get('ID', '?')})\n")
            print("| metric | value |\n|---|---|")
            for k, label in KEYS:
                if k in d and d[k] != "":
                    print(f"| {label} (`{k}`) | {d[k]} {u[k]} |")
            stalls = []
            for k in hdr:
                if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
                    try:
                        v = float(d[k].replace(",", ""))
                    except ValueError:
                        continue
                    if v >= 0.3:
                        stalls.append((v, k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            stalls.sort(reverse=True)
            print("| stalled warps per issue (>= 0.3) | " + ", ".join(f"{n} {v:.2f}" for v, n in stalls) + " |")
            print()


if __name__ == "__main__":
    main()
