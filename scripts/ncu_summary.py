"""One line per captured kernel from the raw page of an ncu --set full capture
(ncu -i x.ncu-rep --page raw --csv > raw.csv): duration, DRAM traffic, throughput percentages."""
import csv
import sys

F = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
     "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}


def main():
    for path in sys.argv[1:]:
        rows = list(csv.reader(open(path)))
        header, units = rows[0], rows[1]
        col = {n: i for i, n in enumerate(header)}

        def find(*parts):
            return next((n for n in header if all(p in n for p in parts)), None)

        dram_pct = find("dram__throughput.avg.pct_of_peak_sustained_elapsed")
        sm_pct = find("sm__throughput.avg.pct_of_peak_sustained_elapsed")
        sys_pct = find("syslts__t_sector_throughput_aperture_sysmem")
        occ = find("sm__warps_active.avg.pct_of_peak_sustained_active")
        print("# " + path)
        seen = {}
        for r in rows[2:]:
            name = r[col["Kernel Name"]].split("(")[0].replace("void ", "")
            seen[name] = seen.get(name, 0) + 1
            if seen[name] > 2:
                continue

            def val(n):
                return float(r[col[n]].replace(",", "")) * F.get(units[col[n]], 1)

            dur = val("gpu__time_duration.sum")
            rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
            print("%-40s grid %-7s %8.1f us  dram rd %.4f GB wr %.4f GB  (%.0f GB/s)  dram %s%%  sm %s%%  "
                  "sysmem %s%%  regs %s  warps_active %s%%" % (
                      name, r[col["launch__grid_size"]], dur, rd / 1e9, wr / 1e9,
                      (rd + wr) / 1e3 / dur, r[col[dram_pct]] if dram_pct else "?",
                      r[col[sm_pct]] if sm_pct else "?", r[col[sys_pct]] if sys_pct else "?",
                      r[col["launch__registers_per_thread"]], r[col[occ]] if occ else "?"))


if __name__ == "__main__":
    main()
