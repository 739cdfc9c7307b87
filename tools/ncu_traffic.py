"""DRAM traffic of the kernel sets bench.py reports rooflines for, from an `ncu --set full` (or --metrics dram__bytes_*)
capture of ONE train step of bench.py:

    ncu --set full --clock-control none -k regex:'wgrad|split_bf16|sk_kernel' --csv --page raw --log-file raw.csv python bench.py ...
    python tools/ncu_traffic.py raw.csv [first_id last_id] > profiles/ncu_traffic.json

Per bench kernel label: bytes_per_launch = (dram__bytes_read.sum + dram__bytes_write.sum, summed over every launch of the
kernels the label's CUDA-event bracket covers) / (launches of the label's main kernel)."""
import csv
import json
import sys

SETS = {   # bench label -> (main kernel, every kernel inside the bracket)
    "wgrad_bf16_kernel(+reduce)": ("wgrad_bf16_kernel", ("wgrad_bf16_kernel", "wgrad_reduce_kernel", "split_bf16_kernel")),
    "wgrad_bf16_kernel(+split,reduce)": ("wgrad_bf16_kernel", ("wgrad_bf16_kernel", "wgrad_reduce_kernel", "split_bf16_kernel")),
    "sk_kernel": ("sk_kernel", ("sk_kernel",)),
    "conv_halo_kernel": ("conv_halo_kernel", ("conv_halo_kernel",)),
}


def main():
    path = sys.argv[1]
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 60
    with open(path, newline="") as f:
        lines = [l for l in f if l.strip() and not l.startswith("==")]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    ci = {n: i for i, n in enumerate(hdr)}
    per = {}
    if "Metric Name" in ci:      # long format (--metrics ... --csv)
        for r in rows[1:]:
            if len(r) <= ci["Metric Value"] or not r[ci["ID"]].isdigit() or not (lo <= int(r[ci["ID"]]) <= hi):
                continue
            if r[ci["Metric Name"]] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                v = float(r[ci["Metric Value"]].replace(",", ""))
                unit = r[ci["Metric Unit"]].lower()
                v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
                d = per.setdefault(int(r[ci["ID"]]), [r[ci["Kernel Name"]], 0.0])
                d[1] += v
    else:                        # wide format (--page raw --csv): second row holds the units
        units = rows[1]
        for r in rows[2:]:
            if not r or not r[ci["ID"]].isdigit() or not (lo <= int(r[ci["ID"]]) <= hi):
                continue
            tot = 0.0
            for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                v = float(r[ci[m]].replace(",", ""))
                tot += v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(units[ci[m]].lower(), 1)
            per[int(r[ci["ID"]])] = [r[ci["Kernel Name"]], tot]
    out = {}
    for label, (main_k, ks) in SETS.items():
        n_main = sum(1 for k, _ in per.values() if main_k in k)
        tot = sum(b for k, b in per.values() if any(x in k for x in ks))
        if n_main:
            out[label] = {"bytes_per_launch": tot / n_main, "launches": n_main, "total_bytes": tot,
                          "source": f"ncu capture {path.split('/')[-1]}: dram__bytes_read.sum + dram__bytes_write.sum over "
                                    f"{', '.join(ks)} / launches of {main_k}"}
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
