"""Print the last N rows of an `ncu --metrics gpu__time_duration.sum --csv` launch list (kernel name, microseconds)."""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
for r in rows[1:][-n:]:
    print(f"{float(r[-1])/1e3:9.1f} us  {r[4][:70]}")
