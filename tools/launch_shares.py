"""Share of device time per kernel from an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X.csv`).
    python tools/launch_shares.py gpurun_out/r02_bench_launches.csv [--md]
Times in such a list are cold-cache and serialised: the SHARES are what is compared with the live CUDA-event timings."""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    md = "--md" in sys.argv
    rows = [ln for ln in open(path, errors="replace") if ln.startswith('"')]
    rd = csv.DictReader(rows)
    tot = collections.Counter()
    cnt = collections.Counter()
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)   # -> microseconds
        m = re.search(r"(?:::|\s|^)(k_\w+|[A-Za-z_]\w*)\s*(?:<[^()]*>)?\s*\(", r["Kernel Name"].replace("<unnamed>", "anon"))
        name = m.group(1) if m else r["Kernel Name"][:40]
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    if total == 0:
        print("no gpu__time_duration.sum rows found")
        return
    print(f"{sum(cnt.values())} launches, {total / 1e3:.2f} ms in total" if not md else "| share | kernel | launches | mean us |\n|---|---|---|---|")
    for name, v in tot.most_common():
        if md:
            print(f"| {100 * v / total:.1f} % | `{name}` | {cnt[name]} | {v / cnt[name]:.1f} |")
        else:
            print(f"{100 * v / total:6.1f} %  {name:32s} {cnt[name]:6d} launches  {v / cnt[name]:8.1f} us mean")


if __name__ == "__main__":
    main()
