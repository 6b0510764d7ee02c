"""Summarises a `ncu --set full` report of the tree kernels: one line per kernel launch with duration, DRAM bytes and
rate against the measured copy peak, cache hit rates, occupancy and issue utilisation.
Usage: python tools/ncu_tree_summary.py report.ncu-rep [more reports ...]   (runs `ncu -i ... --page raw --csv` itself)"""
import csv
import io
import subprocess
import sys

HBM_GBS = 6539.2  # MEASURED_PEAKS.json


def main():
    for rep in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        idx = {h: i for i, h in enumerate(rows[0])}

        def f(r, name):
            return float(r[idx[name]].replace(",", "") or 0)

        print(rep)
        for r in rows[2:]:
            name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
            us = f(r, "gpu__time_duration.sum")
            rd, wr = f(r, "dram__bytes_read.sum"), f(r, "dram__bytes_write.sum")
            gbs = (rd + wr) * 1e6 / (us * 1e-6) / 1e9
            print("  %-44s %6.1f us  dram read %7.2f MB write %6.2f MB -> %6.0f GB/s (%4.1f %% of %d)  L2 hit %4.1f %%  L1 hit %4.1f %%  "
                  "warps active %4.1f %%  issue active %4.1f %%  regs %d  grid %d x %d"
                  % (name, us, rd, wr, gbs, 100 * gbs / HBM_GBS, HBM_GBS, f(r, "lts__t_sector_hit_rate.pct"),
                     f(r, "l1tex__t_sector_hit_rate.pct"), f(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                     f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"), int(f(r, "launch__registers_per_thread")),
                     int(f(r, "launch__grid_size")), int(f(r, "launch__block_size"))))


if __name__ == "__main__":
    main()
