"""DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the individually timed
kernels, from `ncu --set full` reports -> profiles/roofline_traffic.json (read by bench.py).
    python tools/ncu_traffic.py config2=gpurun_out/full_x_c2.ncu-rep config4=gpurun_out/ncuraw_x_c4.csv ..."""
import csv, json, subprocess, sys
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
NAMES = {"k_expand": "k_expand", "Onesweep": "radix_sort", "k_row_sort": "radix_sort", "k_step2_pairs": "k_step2_pairs",
         "k_step3_entries": "step3_numeric", "k_step3_windows": "step3_numeric"}
out = {}
for arg in sys.argv[1:]:
    cfg, rep = arg.split("=")
    # a .ncu-rep, or the `ncu -i ... --page raw --csv` text exported on the GPU box (tools/r02_profiles.sh)
    txt = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    acc, seen = {}, {}
    for r in rows[2:]:
        for pat, name in NAMES.items():
            if pat in r[ix["Kernel Name"]]:
                b = sum(float(r[ix[m]]) * UNIT[units[ix[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
                acc[name] = acc.get(name, 0.0) + b
                seen[name] = seen.get(name, 0) + 1
    # one SpGEMM iteration was captured: the sort's passes add up, the others appear once
    out[cfg] = {k: int(v if k == "radix_sort" else v / seen[k]) for k, v in acc.items()}
json.dump(out, open("profiles/roofline_traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))
