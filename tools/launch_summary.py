"""Per-kernel times of the LAST SpGEMM iteration in an ncu launch list (tools/launch_list.sh)."""
import csv, sys
for f in sys.argv[1:]:
    rows = [r for r in csv.reader(open(f)) if len(r) > 10 and r[0].isdigit()]
    idx = [i for i, r in enumerate(rows) if 'k_tile_products' in r[4] or 'k_row_window' in r[4]]
    start = idx[-1] - 1 if idx else 0
    print('#', f)
    tot = 0.0
    for r in rows[start:]:
        name = r[4].split('(')[0].replace('void ', '').replace('<unnamed>::', '')[:72]
        t = float(r[-1]) / 1e3
        tot += t
        print(f'   {name:72s} {t:9.1f} us  grid {r[8]:>14s} block {r[7]}')
    print(f'   total {tot / 1e3:.3f} ms')
