"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (prints markdown)."""
import collections
import csv
import io
import re
import sys


def short_name(n):
    if "step_kernel" in n:
        m = re.search(r"step_kernel<(?:\(int\))?(\d), ([^,]+), ([^,]+), (?:\(int\))?(\d), (?:\(int\))?(\d), (?:\(bool\))?(\d), (?:\(bool\))?(\d), (?:\(bool\))?(\d), (?:\(bool\))?(\d)>", n)
        if m:
            fam = {"0": "flow", "1": "dance", "2": "dpm"}[m.group(1)]
            src = {"0": "SDE(noise)", "1": "train(given)", "2": "ODE"}[m.group(4)]
            return f"mg::step_kernel<{fam},{'bf16' if 'bfloat' in m.group(2) else 'f32'},{src},rnd={m.group(6)},mean={m.group(9)}>"
    if "logprob_bwd_kernel" in n:
        return "mg::logprob_bwd_kernel"
    if "grpo_loss_kernel" in n:
        return "mg::grpo_loss_kernel"
    if "group_adv_kernel" in n:
        return "mg::group_adv_kernel"
    n = re.sub(r"\(.*", "", n)
    n = re.sub(r"<.*", "<...>", n)
    return n[:80]


def main(path):
    text = open(path).read()
    text = text[text.index('"ID"'):]
    rows = list(csv.DictReader(io.StringIO(text)))
    agg = collections.OrderedDict()
    for r in rows:
        d = agg.setdefault(short_name(r["Kernel Name"]), [0, 0.0, r["Grid Size"], r["Block Size"]])
        d[0] += 1
        d[1] += float(r["Metric Value"].replace(",", ""))
    unit = rows[0]["Metric Unit"]
    scale = {"ns": 1e-3, "us": 1.0, "nsecond": 1e-3, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(unit, 1.0)
    tot = sum(v[1] for v in agg.values())
    print(f"launches: {len(rows)}; total device time {tot * scale:.1f} us (cold-cache, serialised: compare SHARES)\n")
    print("| kernel | launches | avg us | total us | share | grid | block |")
    print("|---|---:|---:|---:|---:|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {v[0]} | {v[1] * scale / v[0]:.2f} | {v[1] * scale:.1f} | {100 * v[1] / tot:.1f}% | {v[2]} | {v[3]} |")


if __name__ == "__main__":
    main(sys.argv[1])
