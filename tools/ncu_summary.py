"""Condense .ncu-rep captures (ncu --set full) into a small JSON summary for profiles/."""
import csv, io, json, subprocess, sys

WANT = [
    ("duration", "gpu__time_duration.sum"),
    ("dram_read", "dram__bytes_read.sum"),
    ("dram_write", "dram__bytes_write.sum"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor_pipe_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("tensor_realtime_pct", "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
    ("l2_to_sm_read", "l1tex__m_xbar2l1tex_read_bytes.sum"),
    ("l2_to_sm_pct", "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed"),
    ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
    ("sm_throughput_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("registers", "launch__registers_per_thread"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("smem_dyn", "launch__shared_mem_per_block_dynamic"),
    ("sm_clock", "sm__cycles_elapsed.avg.per_second"),
]


def summarize(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    idx = {h: i for i, h in enumerate(hdr)}
    res = {"kernel": vals[idx["Kernel Name"]] if "Kernel Name" in idx else "?"}
    for key, name in WANT:
        if name in idx:
            res[key] = "%s %s" % (vals[idx[name]], units[idx[name]])
    return res


if __name__ == "__main__":
    allres = {}
    for p in sys.argv[1:]:
        allres[p.split("/")[-1]] = summarize(p)
    print(json.dumps(allres, indent=1))
