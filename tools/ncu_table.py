"""One row per launch out of ncu reports (`ncu --set full` captures brought back in gpurun_out/): the counters the
DESIGN / VERDICT discussion uses.

    python tools/ncu_table.py gpurun_out/r02_prof_estep.ncu-rep gpurun_out/r02_prof_stream.ncu-rep > profiles/r02_estep_and_stream_ncu.txt
"""
import csv, io, re, subprocess, sys

COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
UNIT = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def table(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[hdr], rows[hdr + 1]
    idx = {n: i for i, n in enumerate(names)}
    for r in rows[hdr + 2:]:
        if len(r) != len(names):
            continue
        name = re.sub(r"^void ", "", r[idx["Kernel Name"]]).replace("(anonymous namespace)::", "").replace("vmm::", "")
        vals = []
        for c in COLS:
            if c not in idx:
                vals.append("n/a")
                continue
            try:
                v = float(r[idx[c]].replace(",", ""))
            except ValueError:
                vals.append("n/a")
                continue
            v *= UNIT.get(units[idx[c]], 1.0)
            vals.append(f"{v:.3f}" if abs(v) < 1e6 else f"{v:.0f}")
        yield name[:64], vals


if __name__ == "__main__":
    print("# ncu --set full --clock-control none; one row per launch; time in us, dram bytes in MB")
    print("# columns: kernel | " + " | ".join(COLS))
    for rep in sys.argv[1:]:
        print(f"# ---- {rep}")
        for name, vals in table(rep):
            print(f"{name:64s} " + " ".join(f"{v:>12s}" for v in vals))
