"""Key metrics of an `ncu -i X.ncu-rep --page raw --csv` dump, one block per captured launch (markdown)."""
import csv, sys

WANT = [
    ("duration", "gpu__time_duration.sum"),
    ("DRAM read", "dram__bytes_read.sum"),
    ("DRAM write", "dram__bytes_write.sum"),
    ("DRAM throughput %", "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L2 throughput %", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L1/TEX throughput %", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor pipe active %", "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active"),
    ("tensor pipe (any) %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("fp64 pipe %", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    ("DMMA sub-pipe %", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active"),
    ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("shared-memory wavefronts %", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
    ("registers / thread", "launch__registers_per_thread"),
    ("dynamic shared memory / CTA", "launch__shared_mem_per_block_dynamic"),
    ("achieved occupancy %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("warp instructions", "smsp__inst_executed.sum"),
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("### `%s`" % r[idx["Kernel Name"]][:110])
        for label, key in WANT:
            if key in idx and r[idx[key]] != "":
                print("* %s: %s %s" % (label, r[idx[key]], units[idx[key]]))
        stalls = []
        for h in hdr:
            if "issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(r[idx[h]].replace(",", "")), h.split("issue_stalled_")[1].split("_per")[0]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print("* top stalls (warps per issue): " + ", ".join("%s %.2f" % (n, v) for v, n in stalls[:4]))
        print()


if __name__ == "__main__":
    main(sys.argv[1])
