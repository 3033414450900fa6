"""Turn a .ncu-rep into the compact text summary committed under profiles/ (run here, no GPU needed)."""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__cycles_elapsed.avg",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]


def main(rep, note=""):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print("# ncu --set full --clock-control none summary of %s" % rep.split("/")[-1])
    if note:
        print("# " + note)
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print("\nkernel: %s" % name[:110])
        for k in KEYS:
            for i, h in enumerate(hdr):
                if h == k or h.endswith("." + k):
                    print("  %-82s %14s %s" % (h[-82:], r[i], units[i]))
                    break


if __name__ == "__main__":
    main(sys.argv[1], " ".join(sys.argv[2:]))
