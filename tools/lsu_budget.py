"""LSU / shared-memory data-pipe accounting of a kernel from an `ncu --set full --import-source on` report
(the analysis behind profiles/r01_lsu_pipe_analysis.txt, automated):

    python tools/lsu_budget.py gpurun_out/<report>.ncu-rep [kernel-name-substring] [warp-tiles]

Prints the pipe utilisation, the wavefront totals by class (shared loads / stores, hardware "bank conflicts" =
cycles lost to TMA writes + real conflicts, global/local), the per-instruction excess over the ideal wavefront
count, and -- if the number of warp-tiles of the launch is given -- wavefronts per warp-tile."""
import csv
import io
import subprocess
import sys


def page(report, which):
    out = subprocess.run(["ncu", "-i", report, "--page", which, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    report = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    warp_tiles = float(sys.argv[3]) if len(sys.argv) > 3 else None
    raw = page(report, "raw")
    hdr = raw[0]
    ki, di = hdr.index("Kernel Name"), hdr.index("gpu__time_duration.sum")
    rows = [r for r in raw[2:] if want in r[ki]]
    if not rows:
        sys.exit(f"no kernel matching {want!r}")
    r = max(rows, key=lambda x: float(x[di].replace(",", "") or 0))
    get = lambda name: float(r[hdr.index(name)].replace(",", "")) if name in hdr and r[hdr.index(name)] else float("nan")
    print("kernel:", r[ki][:110])
    print(f"duration {get('gpu__time_duration.sum'):.3f} {raw[1][di]}")
    for name in ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
                 "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
                 "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
                 "smsp__issue_active.avg.pct_of_peak_sustained_active",
                 "lts__throughput.avg.pct_of_peak_sustained_elapsed",
                 "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"):
        print(f"  {name:72s} {get(name):8.2f} %")
    sms = get("launch__grid_size") if get("launch__grid_size") <= 160 else 148
    shared = get("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
    total = get("SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg") * 148
    classes = [("shared loads", get("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum")),
               ("shared stores", get("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum")),
               ("shared atomics", get("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum")),
               ("  of which counted as bank conflicts", get("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")),
               ("global / local (lgds)", get("SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg") * 148),
               ("all", total)]
    print("LSU data-pipe wavefronts:")
    for name, v in classes:
        extra = f"   {v / warp_tiles:8.1f} per warp-tile" if warp_tiles else ""
        print(f"  {name:40s} {v / 1e6:10.1f} M{extra}")
    src = page(report, "source")
    for hi, row in enumerate(src):
        if row and row[0] == "Address":
            break
    else:
        return
    h = src[hi]
    ix = {n: i for i, n in enumerate(h)}
    if "L1 Wavefronts Shared" not in ix:
        return
    tot = ideal = 0.0
    excess = []
    for row in src[hi + 1:]:
        if len(row) < len(h):
            continue
        try:
            w, i = float(row[ix["L1 Wavefronts Shared"]] or 0), float(row[ix["L1 Wavefronts Shared Ideal"]] or 0)
        except ValueError:
            continue
        tot += w
        ideal += i
        if w > 1.01 * i and w - i > 1e5:
            excess.append((w - i, w, i, row[ix["Source"]].strip()[:70]))
    print(f"per-instruction counters (last kernel of the report): {tot / 1e6:.1f} M shared wavefronts, {ideal / 1e6:.1f} M ideal;"
          f" hardware total minus this = TMA-write stalls and polling: {(shared - tot) / 1e6:.1f} M")
    for d, w, i, s in sorted(excess, reverse=True)[:12]:
        print(f"  excess {d / 1e6:7.2f} M  ({w / i:4.1f}x)  {s}")


if __name__ == "__main__":
    main()
