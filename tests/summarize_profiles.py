"""Regenerates profiles/r02_dram_bytes.json and prints the markdown launch table from the ncu launch list (profiles/r02_launches_dram.csv),
and, given .ncu-rep files, a table of the headline metrics per kernel:  python tests/summarize_profiles.py [rep ...]"""
import collections, csv, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = ("profiles/r02_launches_dram.csv: DT_NCU_REGION=1 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
       "dram__bytes_write.sum --clock-control none python bench.py --steps 1 --warmup 3 (one timed step, batch 8); sum over the conv_tcgen05_kernel launches")


def launch_table():
    rows = [r for r in csv.reader(open(os.path.join(ROOT, "profiles", "r02_launches_dram.csv"))) if len(r) > 10]
    hdr = rows[0]
    iK, iM, iV, iU, iID = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
    L = collections.OrderedDict()
    for r in rows[1:]:
        d = L.setdefault(r[iID], {"k": r[iK]})
        v = float(r[iV].replace(",", ""))
        if r[iM].startswith("dram"):
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[r[iU]]
        else:
            v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}.get(r[iU], 1)
        d[r[iM]] = v
    tot = sum(d["gpu__time_duration.sum"] for d in L.values())
    conv = [d for d in L.values() if "conv_tcgen05" in d["k"]]
    ct = sum(d["gpu__time_duration.sum"] for d in conv)
    rd = sum(d["dram__bytes_read.sum"] for d in conv); wr = sum(d["dram__bytes_write.sum"] for d in conv)
    json.dump({"conv_family_bytes_per_step": rd + wr, "conv_family_read_bytes": rd, "conv_family_write_bytes": wr, "conv_family_ms_under_ncu": ct / 1e3,
               "launches": len(conv), "step_launches": len(L), "step_ms_under_ncu": tot / 1e3, "conv_share_of_step_under_ncu": ct / tot, "source": SRC},
              open(os.path.join(ROOT, "profiles", "r02_dram_bytes.json"), "w"), indent=1)
    agg = collections.OrderedDict()
    for d in L.values():
        k = d["k"].split("(")[0].replace("void ", "").replace("dt::", "")
        a = agg.setdefault(k, [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += d["gpu__time_duration.sum"]; a[2] += d["dram__bytes_read.sum"]; a[3] += d["dram__bytes_write.sum"]
    print("%d launches, %.1f us under ncu; conv family %d launches, %.1f us = %.1f %%; DRAM %.2f GB read + %.2f GB written" %
          (len(L), tot, len(conv), ct, ct / tot * 100, rd / 1e9, wr / 1e9))
    print("| kernel | launches | time (us) | share | DRAM read (MB) | DRAM write (MB) | DRAM GB/s |\n|---|---|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("| `%s` | %d | %.1f | %.1f %% | %.1f | %.1f | %.0f |" % (k, a[0], a[1], a[1] / tot * 100, a[2] / 1e6, a[3] / 1e6, (a[2] + a[3]) / a[1] / 1e3))


def rep_table(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    want = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "us"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
            ("l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "L2->SM TB/s"), ("dram__bytes_read.sum", "DRAM rd MB"), ("dram__bytes_write.sum", "DRAM wr MB"),
            ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("launch__registers_per_thread", "regs"),
            ("smsp__inst_executed.sum", "warp insts"), ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem conflicts")]
    idx = [(hdr.index(k), n) for k, n in want if k in hdr]
    print("| " + " | ".join(n for _, n in idx) + " |\n|" + "---|" * len(idx))
    for r in rows[2:]:
        print("| " + " | ".join(r[i].replace("void ", "").replace("(ConvParams)", "")[:60] for i, _ in idx) + " |")


if __name__ == "__main__":
    launch_table()
    for p in sys.argv[1:]:
        print("\n" + p)
        rep_table(p)
