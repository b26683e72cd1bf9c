"""Turns an ncu metric log of the bench kernel into profiles/counters.json -- the per-launch counters bench.py's roofline
block divides by the times it measures itself.

    ncu --metrics <METRICS> --clock-control none -k regex:traverse_bvh8_vote --csv --log-file gpurun_out/r02_counters.csv \
        python scripts/one_pass.py passes=2
    python scripts/ncu_counters.py gpurun_out/r02_counters.csv profiles/counters.json

one_pass.py traces the primary set `passes` times, then the random set: the last launch of each group is taken."""
import csv, json, sys

METRICS = ("gpu__time_duration.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,lts__t_sectors.sum,dram__bytes_read.sum,"
           "dram__bytes_write.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,sm__inst_executed.avg.per_cycle_elapsed,"
           "smsp__thread_inst_executed_per_inst_executed.ratio,lts__throughput.avg.pct_of_peak_sustained_elapsed,"
           "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed,sm__warps_active.avg.per_cycle_active")


def main(src, dst, passes=2):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    iid, iname, imet, iunit, ival = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
    launches = {}
    for r in rows[1:]:
        launches.setdefault(int(r[iid]), {"kernel": r[iname]})[r[imet]] = (float(r[ival].replace(",", "")), r[iunit])
    ids = sorted(launches)
    assert len(ids) == 2 * passes, f"expected {2 * passes} launches, found {len(ids)}"
    pick = {"primary": launches[ids[passes - 1]], "random": launches[ids[2 * passes - 1]]}

    def val(l, m, scale=1.0):
        v, unit = l[m]
        mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "usecond": 1e-6, "us": 1e-6, "msecond": 1e-3, "ms": 1e-3, "nsecond": 1e-9, "ns": 1e-9, "second": 1, "s": 1}.get(unit, 1)
        return v * mult * scale

    out = {"source": f"{src}: ncu --metrics ... --clock-control none on scripts/one_pass.py, the last of {passes} launches per ray set (warm caches)",
           "kernel": pick["primary"]["kernel"][:80],
           "ncu_duration_us": {n: round(val(l, "gpu__time_duration.sum") * 1e6, 1) for n, l in pick.items()},
           "thread_inst_per_launch": {n: val(l, "smsp__thread_inst_executed.sum") for n, l in pick.items()},
           "warp_inst_per_launch": {n: val(l, "smsp__inst_executed.sum") for n, l in pick.items()},
           "l2_sectors_per_launch": {n: val(l, "lts__t_sectors.sum") for n, l in pick.items()},
           "dram_bytes_per_launch": {n: val(l, "dram__bytes_read.sum") + val(l, "dram__bytes_write.sum") for n, l in pick.items()},
           "l1_hit_pct": {n: round(val(l, "l1tex__t_sector_hit_rate.pct"), 2) for n, l in pick.items()},
           "l2_hit_pct": {n: round(val(l, "lts__t_sector_hit_rate.pct"), 2) for n, l in pick.items()},
           "l2_throughput_pct_of_peak": {n: round(val(l, "lts__throughput.avg.pct_of_peak_sustained_elapsed"), 2) for n, l in pick.items()},
           "l1_data_pipe_pct_of_peak": {n: round(val(l, "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed"), 2) for n, l in pick.items()},
           "warps_active_per_sm": {n: round(val(l, "sm__warps_active.avg.per_cycle_active"), 2) for n, l in pick.items()}}
    out["dram_bytes_per_launch_avg"] = sum(out["dram_bytes_per_launch"].values()) / 2
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    if len(sys.argv) < 3:
        print(METRICS)
    else:
        main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 2)
