"""Per-launch summary of an `ncu --set full` capture exported with `ncu -i X.ncu-rep --page raw --csv` (tools/gpu_r2w.sh):
duration, tensor-pipe activity, DRAM bytes, L2 hit rate, registers, grid, top warp-stall reasons.
usage: python tools/ncu_summary.py <raw.csv> [stdout log of tools/ncu_conv.py (layer names in launch order)] [launches per layer = 2] > summary.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
cols = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")
          and "not_issued" not in h]
names = []
if len(sys.argv) > 2:
    for line in open(sys.argv[2]):
        parts = line.split()
        if len(parts) > 3 and parts[2] == "ms":
            names.append(parts[0])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2      # launches per named layer (tools/ncu_conv.py REPS)
labels = [n for n in names for _ in range(reps)]
w = csv.writer(sys.stdout)
w.writerow(["layer (tools/ncu_conv.py)", "kernel"] + ["%s [%s]" % (c, units[ix[c]]) for c in cols if c in ix] + ["top warp stalls (per issue)"])
for li, r in enumerate(data):
    k = r[ix["Kernel Name"]]
    short = k.split("(")[0].replace("void <unnamed>::", "")
    vals = [r[ix[c]] for c in cols if c in ix]
    st = []
    for h in stalls:
        try:
            st.append((float(r[ix[h]].replace(",", "")), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
        except ValueError:
            pass
    st.sort(reverse=True)
    w.writerow([labels[li] if li < len(labels) else "lf_train_step_b32", short] + vals + [" ".join("%s=%.2f" % (n, v) for v, n in st[:4])])
