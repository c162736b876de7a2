"""Summarise an .ncu-rep (read on the CPU box with `ncu -i`) into a small text table for profiles/."""
import csv, subprocess, sys

KEYS = [("gpu__time_duration.sum", "time_us"), ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct_of_ncu_peak"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
        ("smsp__inst_executed.sum", "warp_insts"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct")]

def main(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {k: hdr.index(k) for k, _ in KEYS if k in hdr}
    kn = hdr.index("Kernel Name")
    print(f"# {rep}")
    for r in rows[2:]:
        print(f"kernel: {r[kn][:90]}")
        for k, short in KEYS:
            if k in ix:
                print(f"    {short:24s} {r[ix[k]]:>18s} {units[ix[k]]}")

if __name__ == "__main__":
    main(sys.argv[1])
