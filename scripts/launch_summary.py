"""Per-kernel summary of an ncu launch list (scripts/ncu_launches.sh): launches per step, time per step and
share, DRAM bytes and GB/s.  usage: python scripts/launch_summary.py launches.csv [steps_in_capture]"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    rows = list(csv.reader(open(path, errors="replace")))
    hdr = None
    per = collections.defaultdict(lambda: collections.defaultdict(float))
    cnt = collections.Counter()
    for r in rows:
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
                ki, vi, mi, ii = r.index("Kernel Name"), r.index("Metric Value"), r.index("Metric Name"), r.index("ID")
            continue
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("sunet::", "")
        v = float(r[vi].replace(",", ""))
        per[name][r[mi]] += v
        if r[mi].startswith("gpu__time"):
            cnt[name] += 1
    steps = float(sys.argv[2]) if len(sys.argv) > 2 else float(cnt.get("adam_kernel", 1))
    tot = sum(v["gpu__time_duration.sum"] for v in per.values())
    print(f"# {path}: {sum(cnt.values())} launches, {steps:g} steps, {tot / 1e6 / steps:.3f} ms/step (serialised, cold)")
    print(f"{'kernel':40s} {'n/step':>7s} {'ms/step':>8s} {'share':>6s} {'MB/launch':>10s} {'GB/s':>8s}")
    for k, v in sorted(per.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
        t = v["gpu__time_duration.sum"]
        b = v.get("dram__bytes_read.sum", 0.0) + v.get("dram__bytes_write.sum", 0.0)
        print(f"{k[:40]:40s} {cnt[k] / steps:7.1f} {t / 1e6 / steps:8.3f} {100 * t / tot:5.1f}% {b / cnt[k] / 1e6:10.1f} "
              f"{b / t:8.1f}")


def traffic_json(path, out):
    """Average DRAM bytes per launch of the G1 family (conv3_halo2_kernel*, conv_gemm_kernel*) -> bench.py's
    roofline.traffic.  usage: python scripts/launch_summary.py launches.csv --traffic-json out.json"""
    import json
    rows = list(csv.reader(open(path, errors="replace")))
    hdr = None
    tot, n, t = 0.0, 0, 0.0
    for r in rows:
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
                ki, vi, mi = r.index("Kernel Name"), r.index("Metric Value"), r.index("Metric Name")
            continue
        if len(r) <= vi or not ("conv3_halo" in r[ki] or "conv_gemm_kernel" in r[ki]):
            continue
        v = float(r[vi].replace(",", ""))
        if r[mi].startswith("dram__bytes"):
            tot += v
        elif r[mi].startswith("gpu__time"):
            n += 1
            t += v
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import kernel_source_sha          # the stamp bench.py checks before it trusts this file
    json.dump({"kernel_family": "G1 (conv3_halo2_kernel*, conv_gemm_kernel*)", "launches": n,
               "kernel_sha": kernel_source_sha(),
               "dram_bytes_per_launch": tot / n, "avg_launch_ns_under_ncu": t / n,
               "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                         "(scripts/ncu_launches.sh), file " + path.split("/")[-1]}, open(out, "w"), indent=1)


if __name__ == "__main__":
    if len(sys.argv) > 3 and sys.argv[2] == "--traffic-json":
        traffic_json(sys.argv[1], sys.argv[3])
    else:
        main()
