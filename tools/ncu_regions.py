"""Group the SASS of one kernel in an ncu report (--import-source on) into regions of equal execution count:
instructions executed, share of the kernel's instructions and of its stall samples, lanes active.
    python tools/ncu_regions.py gpurun_out/full_x.ncu-rep [--sass FIRST LAST]"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
f = lambda r, k: float(r[ix[k]] or 0)
tot = sum(f(r, 'Instructions Executed') for r in data); tots = sum(f(r, '# Samples') for r in data)
print(f'# {rep}: {tot/1e9:.3f} G warp instructions, {tots:.0f} samples')
if '--sass' in sys.argv:
    a, b = int(sys.argv[sys.argv.index('--sass') + 1]), int(sys.argv[sys.argv.index('--sass') + 2])
    for n in range(a, b + 1):
        r = data[n]
        print(f"{n:4d} {f(r,'Instructions Executed')/1e6:8.1f}M lanes {f(r,'Avg. Threads Executed'):4.1f} samples {r[ix['# Samples']]:>7} | {r[ix['Source']].strip()[:100]}")
    sys.exit(0)
acc = accs = 0; start = 0; prev = None
for n, r in enumerate(data + [None]):
    ie = f(r, 'Instructions Executed') if r else -1
    if prev is not None and (r is None or ie > prev * 1.3 or ie < prev / 1.3):
        lanes = f(data[start], 'Avg. Threads Executed')
        if acc > 0:
            print(f"[{start:3d}-{n-1:3d}] {n-start:3d} instr x {prev/1e6:7.1f}M = {acc/1e6:8.1f}M ({100*acc/tot:4.1f}%) samples {100*accs/tots:4.1f}% lanes {lanes:4.1f} | {data[start][ix['Source']].strip()[:60]}")
        acc = accs = 0; start = n
    if r:
        acc += ie; accs += f(r, '# Samples'); prev = ie
