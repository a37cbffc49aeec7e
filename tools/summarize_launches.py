"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: keeps ONE steady-state step (between two
latent_in launches of the reference batch) and prints per-kernel / per-grid shares.
usage: python tools/summarize_launches.py <launches.csv> <out_prefix>   -> <out_prefix>.md + <out_prefix>_step.csv"""
import collections
import csv
import re
import sys


def main():
    src, out = sys.argv[1], sys.argv[2]
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    ii = [i for i, r in enumerate(rows) if "image_in_kernel" in r["Kernel Name"] or "image_patches" in r["Kernel Name"]]
    if ii:      # image pipeline: a step = [image_in (refs), image_in (degraded), ...]
        starts = ii[0::2]
    else:       # latent pipeline: a step = [ref latent_in, main latent_in, ...]
        li = [i for i, r in enumerate(rows) if r["Kernel Name"].startswith(("latent_in", "ir::latent_in"))]
        starts = li[0::2]
    s, e = (starts[-2], starts[-1]) if len(starts) >= 2 else (0, len(rows))
    step = rows[s:e]
    with open(out + "_step.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel", "block", "grid", "duration_ns"])
        for r in step:
            w.writerow([r["ID"], re.sub(r"\(.*", "", r["Kernel Name"]), r["Block Size"], r["Grid Size"], r["Metric Value"]])
    tot = sum(float(r["Metric Value"]) for r in step) / 1e3
    byname, bygrid = collections.Counter(), collections.OrderedDict()
    for r in step:
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("ir::", "").replace("void ", "")
        t = float(r["Metric Value"]) / 1e3
        byname[name] += t
        a = bygrid.setdefault((name, r["Grid Size"]), [0, 0.0])
        a[0] += 1
        a[1] += t
    with open(out + ".md", "w") as f:
        f.write(f"# ncu launch list, one steady-state step ({len(step)} launches, sum of durations {tot / 1e3:.2f} ms)\n\n")
        f.write("Cold-cache, serialised per-launch times (`--clock-control none`): compare SHARES, not absolutes.\n\n")
        f.write("| kernel | us | share |\n|---|---|---|\n")
        for n, t in byname.most_common():
            f.write(f"| {n} | {t:.1f} | {t / tot * 100:.1f}% |\n")
        f.write("\n| kernel | grid | launches | us total | us/launch |\n|---|---|---|---|---|\n")
        for (n, g), (c, t) in sorted(bygrid.items(), key=lambda kv: -kv[1][1])[:60]:
            f.write(f"| {n} | {g} | {c} | {t:.1f} | {t / c:.1f} |\n")
    print(open(out + ".md").read())


if __name__ == "__main__":
    main()
