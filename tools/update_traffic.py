#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` capture of k_score_units on a bench workload: DRAM bytes per launch (what
bench.py reports as roofline.traffic) and the SASS thread-instructions the kernel executes per anchor pair.
    python tools/update_traffic.py gpurun_out/r5l_score.ncu-rep ont 3229434240"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, wl, pairs = sys.argv[1], sys.argv[2], int(sys.argv[3])
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, u, v = rows[0], rows[1], rows[2]

    def val(k):
        x = float(v[h.index(k)].replace(",", ""))
        unit = u[h.index(k)].lower()
        return x * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
    name = v[h.index("Kernel Name")]
    assert "k_score_units<512, 1>" in name or "k_score_units" in name, name
    dram = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
    tinst = val("smsp__thread_inst_executed.sum") if "smsp__thread_inst_executed.sum" in h else val("smsp__inst_executed.sum") * val("smsp__thread_inst_executed_per_inst_executed.ratio")
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        tj = json.load(open(path))
    except Exception:
        tj = {}
    tj[wl] = {"kernel": name.split("(")[0], "dram_bytes_per_launch": int(dram), "thread_instr_per_pair": round(tinst / pairs, 2),
              "warp_instr": int(val("smsp__inst_executed.sum")), "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
              "ms_under_ncu": val("gpu__time_duration.sum") / 1e6 if u[h.index("gpu__time_duration.sum")].lower() in ("ns", "nsecond") else val("gpu__time_duration.sum"),
              "source": os.path.basename(rep)}
    json.dump(tj, open(path, "w"), indent=1)
    print(json.dumps(tj[wl]))


if __name__ == "__main__":
    main()
