"""DRAM traffic per launch of the kernels bench.py quotes, read from this round's `ncu --set full` captures and written to
profiles/r02_ncu_traffic.json (bench.py's `roofline.traffic` reads that file -- never a literal copied by hand).

    python tools/ncu_traffic.py key=gpurun_out/file.ncu-rep[:kernel-substring] ...

key is the name bench.py looks up (gbm_f32_greeks, gbm_f64_greeks, heston_f32_antithetic, svj_f32_antithetic,
paths_f32, paths_f64_out_f32_state, ...).  The first launch in the report whose kernel name contains the substring
(default: any) is used."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")


def launches(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    for d in data:
        yield {h: (u, v) for h, u, v in zip(hdr, units, d)}


def to_bytes(unit, val):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def main(argv):
    res = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for arg in argv:
        key, rest = arg.split("=", 1)
        path, _, sub = rest.partition(":")
        for col in launches(path):
            name = col["Kernel Name"][1]
            if sub and sub not in name:
                continue
            rd, wr = to_bytes(*col["dram__bytes_read.sum"]), to_bytes(*col["dram__bytes_write.sum"])
            res[key] = {"dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_written": wr, "kernel": name,
                        "grid": col["launch__grid_size"][1], "duration_us_under_ncu": col["gpu__time_duration.sum"][1],
                        "source": f"dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full capture {os.path.basename(path)} "
                                  f"(summary: profiles/r02_ncu_{key}.txt)"}
            break
        else:
            print(f"no launch matching {sub!r} in {path}", file=sys.stderr)
    json.dump(res, open(OUT, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main(sys.argv[1:])
