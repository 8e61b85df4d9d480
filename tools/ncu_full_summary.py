"""Turn an `ncu --set full` report into the small JSON `bench.py` reads for `roofline.traffic`
(profiles/step_kernel_ncu.json) or prints a markdown digest for any other kernel.

    ncu --set full --clock-control none --import-source on -k regex:step_kernel -c 4 -o gpurun_out/prof_step python tools/kbench.py --profile 1
    python tools/ncu_full_summary.py gpurun_out/prof_step.ncu-rep --json profiles/step_kernel_ncu.json \
        --kernel-regex 'step_kernel<0, __nv_bfloat16, __nv_bfloat16, 0' --algorithmic-bytes 50331648 --shape '(12,4096,64)'

Reads the report with `ncu -i <rep> --page raw --csv` (ncu must be on PATH; no GPU needed)."""
import argparse
import csv
import io
import json
import re
import subprocess

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum",
        "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_bytes.sum",
        "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"]
UNIT_BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def load(rep):
    if rep.endswith(".csv"):                      # a raw page exported on the GPU box: ncu -i X.ncu-rep --page raw --csv > X.csv
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out[out.index('"ID"'):])))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, units))


def stalls(d):
    st = {}
    for k, v in d.items():
        m = re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active\.ratio", k)
        if m and v not in ("", "n/a"):
            st[m.group(1)] = float(v.replace(",", ""))
    tot = sum(st.values()) or 1.0
    return {k: round(100 * v / tot, 1) for k, v in sorted(st.items()) if 100 * v / tot >= 0.5}


def num(s):
    return float(s.replace(",", ""))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--kernel-regex", default=".")
    ap.add_argument("--json", default=None)
    ap.add_argument("--algorithmic-bytes", type=int, default=0)
    ap.add_argument("--shape", default="")
    ap.add_argument("--kernel-label", default=None)
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    launches, units = load(a.report)
    sel = [d for d in launches if re.search(a.kernel_regex, d["Kernel Name"])]
    if not sel:
        raise SystemExit(f"no launch matches {a.kernel_regex!r}; kernels: {sorted({d['Kernel Name'] for d in launches})}")
    per = []
    for d in sel:
        e = {k: [num(d[k]), units[k]] for k in KEEP if k in d and d[k] not in ("", "n/a")}
        e["stall_pct"] = stalls(d)
        e["kernel"] = d["Kernel Name"]
        per.append(e)
    rd = sum(p["dram__bytes_read.sum"][0] * UNIT_BYTES[p["dram__bytes_read.sum"][1]] for p in per) / len(per)
    wr = sum(p["dram__bytes_write.sum"][0] * UNIT_BYTES[p["dram__bytes_write.sum"][1]] for p in per) / len(per)
    doc = {"kernel": a.kernel_label or sel[0]["Kernel Name"], "shape": a.shape, "launches_captured": len(per), "dram_bytes_read": int(rd),
           "dram_bytes_write": int(wr), "algorithmic_bytes": a.algorithmic_bytes, "note": a.note, "per_launch": per}
    if a.json:
        with open(a.json, "w") as f:
            json.dump(doc, f, indent=1)
    print(f"{len(per)} launch(es) of {sel[0]['Kernel Name']}")
    print(f"dram read {rd / 1e6:.2f} MB, write {wr / 1e6:.3f} MB per launch; duration "
          f"{', '.join(str(p['gpu__time_duration.sum'][0]) + ' ' + p['gpu__time_duration.sum'][1] for p in per)}")
    print("stalls:", per[0]["stall_pct"])


if __name__ == "__main__":
    main()
