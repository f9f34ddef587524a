"""DRAM traffic per launch of the sweep kernels from an `ncu --set full` report -> profiles/<name>.json, the file
bench.py reads to fill `roofline.traffic` (dram__bytes_read.sum + dram__bytes_write.sum, per launch).
usage: python tools/ncu_traffic.py report.ncu-rep atoms out.json"""
import csv
import io
import json
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rep, atoms, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i, rd_i, wr_i = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    dur_i = hdr.index("gpu__time_duration.sum")
    kernels = {}
    for r in data:
        key = "density_sweep" if "density_" in r[name_i] else "force_sweep" if "force_" in r[name_i] else None
        if key is None or key in kernels:
            continue
        rd = float(r[rd_i]) * UNIT[units[rd_i]]
        wr = float(r[wr_i]) * UNIT[units[wr_i]]
        kernels[key] = {"kernel": r[name_i], "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes": rd + wr,
                        "duration_under_ncu_us": float(r[dur_i]), "bytes_per_atom": (rd + wr) / atoms}
    json.dump({"atoms": atoms, "report": rep.split("/")[-1], "how": "ncu --set full --clock-control none, one launch each",
               "kernels": kernels}, open(out, "w"), indent=1)
    print(json.dumps(kernels, indent=1))


if __name__ == "__main__":
    main()
