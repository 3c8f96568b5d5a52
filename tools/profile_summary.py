"""Post-process ncu CSV logs into the summaries committed under profiles/.

    python tools/profile_summary.py launches gpurun_out/launches.csv <steps_in_log> > profiles/rNN_launch_summary.txt
    python tools/profile_summary.py bench gpurun_out/launches.csv > profiles/rNN_launch_summary.txt   # log of `bench.py --value-only`:
                                                    # steps are delimited by their (single) loss_forward_kernel launch
    python tools/profile_summary.py traffic gpurun_out/gemm_traffic.csv <steps_in_log> <precision> <batch>   # -> profiles/gemm_traffic.json
"""
import collections, csv, json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rows(path):
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    r = list(csv.reader(lines))
    h = r[0]
    return h, [dict(zip(h, x)) for x in r[1:] if len(x) == len(h)]


def per_launch(path):
    """-> list of (kernel name, {metric: value}) in launch order (ncu --csv long format: one row per metric)."""
    h, rs = rows(path)
    out, index = [], {}
    for d in rs:
        i = d["ID"]
        if i not in index:
            index[i] = len(out)
            out.append((d["Kernel Name"], {}))
        v = d["Metric Value"].replace(",", "")
        try:
            val = float(v)
        except ValueError:
            continue
        unit = d.get("Metric Unit", "")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        out[index[i]][1][d["Metric Name"]] = val * scale
    return out


def short(name):
    name = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "").replace("<unnamed>::", ""))
    return re.sub(r"^void ", "", name)


if sys.argv[1] in ("launches", "bench"):
    L = per_launch(sys.argv[2])
    if sys.argv[1] == "bench":
        marks = [i for i, (k, _) in enumerate(L) if "loss_forward_kernel" in k]
        steps = len(marks)
        n = marks[-1] - marks[-2]                   # launches from one step's loss kernel to the next one's = one step
        last = L[len(L) - n:]                       # the last step of the log (after warm-up)
    else:
        steps = int(sys.argv[3])
        n = len(L) // steps
        last = L[-n:]                               # the last step of the log (after warm-up)
    tot = sum(m.get("gpu__time_duration.sum", 0.0) for _, m in last)
    agg = collections.OrderedDict()
    for k, m in last:
        a = agg.setdefault(short(k), [0, 0.0])
        a[0] += 1
        a[1] += m.get("gpu__time_duration.sum", 0.0)
    fam = sum(v[1] for k, v in agg.items() if "gemm_tcgen05" in k)
    print(f"# last of {steps} steps in the log: {n} launches, total kernel time {tot:.1f} us (cold-cache, serialised: compare SHARES)")
    print(f"# gemm_tcgen05_kernel family share of kernel time: {100 * fam / tot:.1f}%")
    for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{us:12.1f} us {100 * us / tot:6.2f}%  x{c:5d}  {k[:110]}")
else:
    L = per_launch(sys.argv[2])
    steps, prec, batch = int(sys.argv[3]), sys.argv[4], int(sys.argv[5])
    n = len(L) // steps
    last = L[-n:]
    rd = sum(m.get("dram__bytes_read.sum", 0.0) for _, m in last)
    wr = sum(m.get("dram__bytes_write.sum", 0.0) for _, m in last)
    us = sum(m.get("gpu__time_duration.sum", 0.0) for _, m in last)
    sys.path.insert(0, ROOT)
    import bench
    out = dict(precision=prec, batch=batch, source_hash=bench.kernel_source_hash(), launches_per_step=n, dram_bytes_per_step=rd + wr, dram_read_bytes=rd,
               dram_write_bytes=wr, kernel_us_under_ncu=us,
               source="ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none "
                      "-k regex:gemm_tcgen05 python tools/gemm_breakdown.py %d %s (last step of the log)" % (batch, prec))
    with open(os.path.join(ROOT, "profiles", "gemm_traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))
