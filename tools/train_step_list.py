"""Developer script: the kernels of ONE training step, in launch order, from an ncu launch list
(ncu --metrics gpu__time_duration.sum --clock-control none --csv ... python tools/ncu_train_kernels.py):
python tools/train_step_list.py <launches.csv>  -- prints from the last k-NN grid build to the end of that step."""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
body = [(r[ik], float(r[iv].replace(",", ""))) for r in rows if r is not hdr and len(r) > iv and r[iv].replace(",", "").replace(".", "").isdigit()]
unit_ns = "ns" in " ".join(next((r for r in rows if r is not hdr and len(r) > iv), []))
starts = [i for i, (k, _) in enumerate(body) if "knn_grid_build" in k]
step = body[starts[-1]:]
tot = 0.0
for k, v in step:
    us = v / 1e3 if unit_ns else v
    tot += us
    print(f"{us:8.1f} us  {k[:110]}")
print(f"# sum {tot:.1f} us in {len(step)} launches (serialised, cold cache: the side-stream kernel is counted in full)")
