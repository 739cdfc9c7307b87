"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python tools/summarize_launches.py launches.csv [first_id last_id | --step K] > profiles/rXX_launch_list_summary.txt"""
import csv
import re
import sys


def step_range(path, step):
    """ID range of train step `step` (0-based): every step starts with the two nchw_to_cl launches (audio, video input)"""
    ids = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum" and "nchw_to_cl" in r["Kernel Name"]:
            ids.append(int(r["ID"]))
    return ids[2 * step], ids[2 * step + 2] - 1


def main():
    path = sys.argv[1]
    if len(sys.argv) > 3 and sys.argv[2] == "--step":
        lo, hi = step_range(path, int(sys.argv[3]))
        sys.argv = sys.argv[:2] + [str(lo), str(hi)]
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 60
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        i = int(r["ID"])
        if not (lo <= i <= hi):
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "second": 1e3}.get(unit, 1e-6)
        name = r["Kernel Name"].replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
        name = re.sub(r"\(.*", "", name)
        name = re.sub(r"<.*", "", name).replace("void ", "")
        rows.append((name, ms))
    tot = sum(ms for _, ms in rows)
    agg = {}
    for n, ms in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += ms
    print(f"total {tot:.1f} ms over {len(rows)} launches (ids {lo}..{hi if hi < 1 << 60 else 'end'})")
    for n, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n:48s} n={c:5d} ms={ms:9.2f} share={ms / tot:.3f}")


if __name__ == "__main__":
    main()
