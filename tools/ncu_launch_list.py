"""Condense an `ncu --metrics gpu__time_duration.sum --csv --log-file X` launch log into the per-launch table kept under
profiles/ (id, ms, share of the captured step, grid, block, kernel).

    GMR_PROFILE_STEP=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches_raw.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline \
        --no-extra-workloads --no-graph
    python tools/ncu_launch_list.py gpurun_out/launches_raw.csv "comment line" > profiles/rNN_ncu_launches_step.csv
"""
import csv
import sys


def main():
    path = sys.argv[1]
    note = sys.argv[2] if len(sys.argv) > 2 else ""
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rows = list(csv.DictReader(lines))
    out = []
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3,
                  "second": 1e3}[unit]
        out.append((ms, r["Grid Size"], r["Block Size"], r["Kernel Name"].replace(",", ";")))
    total = sum(o[0] for o in out)
    if note:
        print("# " + note)
    print("# per-launch times are cold-cache and serialised: compare SHARES.  total %.3f ms over %d launches"
          % (total, len(out)))
    print("id,ms,share,grid,block,kernel")
    for i, (ms, grid, block, name) in enumerate(out):
        print("%d,%.4f,%.4f,%s,%s,%s" % (i, ms, ms / total if total else 0.0, grid, block, name))


if __name__ == "__main__":
    main()
