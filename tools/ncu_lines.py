"""Per-source-line roll-up of an `ncu --page source --csv --print-source cuda,sass` export: samples, warp instructions,
thread instructions and average active threads for the hottest lines (development aid for DESIGN.md's kernel notes)."""
import csv, sys, collections
def main(path, top=40):
    rows = list(csv.reader(open(path)))
    fpath = ""; hdr = None; agg = collections.OrderedDict()
    for r in rows:
        if not r: continue
        if r[0] == "File Path": fpath = r[1].split("/")[-1]; continue
        if r[0] == "Line No": hdr = r; continue
        if r[0] in ("Function Name", "Kernel Name") or hdr is None: continue
        if r[2] != "-": continue          # SASS rows carry an address; source rows have '-'
        try:
            ln = int(r[0])
        except ValueError:
            continue
        g = lambda name: float(r[hdr.index(name)] or 0)
        agg[(fpath, ln)] = (g("# Samples"), g("Instructions Executed"), g("Thread Instructions Executed"), g("stall_long_sb"), r[1][:110])
    tot_s = sum(v[0] for v in agg.values()); tot_i = sum(v[1] for v in agg.values()); tot_t = sum(v[2] for v in agg.values())
    print(f"total samples {tot_s:.0f} warp-inst {tot_i:.3g} thread-inst {tot_t:.3g} avg active {tot_t / max(tot_i, 1):.2f}")
    for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{f}:{ln:4d} samp {100 * v[0] / tot_s:5.1f}% inst {100 * v[1] / tot_i:5.1f}% act {v[2] / max(v[1], 1):5.1f} longsb {100 * v[3] / max(v[0], 1):3.0f}% | {v[4]}")
if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
