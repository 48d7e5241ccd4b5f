"""Hot spots of an `ncu --page source --csv` dump (SASS view): top instructions by stall samples, and totals by region."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, isamp, iex, ithr = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed")
data = [(i, r[isrc].strip(), int(r[isamp] or 0), int(r[iex] or 0), r[ithr]) for i, r in enumerate(rows[2:]) if len(r) > iex]
tot_s = sum(d[2] for d in data); tot_e = sum(d[3] for d in data)
print("instructions:", len(data), "total samples", tot_s, "total warp-instr executed", tot_e)
print("--- top by samples")
for d in sorted(data, key=lambda d: -d[2])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{d[0]:5d} {100*d[2]/tot_s:5.1f}% samp {100*d[3]/tot_e:5.2f}% exec thr {d[4]:>5s}  {d[1][:90]}")
# cumulative profile by instruction index buckets
print("--- by index bucket (100 instr)")
for b in range(0, len(data), 100):
    s = sum(d[2] for d in data[b:b+100]); e = sum(d[3] for d in data[b:b+100])
    if s * 50 > tot_s or e * 50 > tot_e:
        print(f"{b:5d}-{b+99:5d}: {100*s/tot_s:5.1f}% samples {100*e/tot_e:5.1f}% exec")
