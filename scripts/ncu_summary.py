"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of counters the roofline uses.
Usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [> profiles/xxx.txt]
       python scripts/ncu_summary.py gpurun_out/prof.ncu-rep --json profiles/<kernel>.json --kernel <name substring>
              --source whisper_ipa_b200/csrc/<file>.cu --shape small/B256/beams1/f16 [--captured "r02, command ..."]
The JSON form is what bench.py's roofline.traffic reads: per-launch DRAM bytes (mean over the captured launches of that
kernel), duration, pipe utilisation, plus the sha256 of the kernel's source file at capture time - bench.py refuses the
numbers once that file has changed."""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print(f"== {r[hdr.index('Kernel Name')][:100]}  grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}")
        for w in WANT:
            for i, h in enumerate(hdr):
                if h == w or h.endswith("." + w):
                    print(f"   {w:75s} {r[i]:>16s} {units[i]}")
                    break
        try:
            rd = float(r[hdr.index("dram__bytes_read.sum")].replace(",", ""))
            wr = float(r[hdr.index("dram__bytes_write.sum")].replace(",", ""))
            ur, uw = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
            print(f"   traffic (read+write)                                                        {rd:.3f} {ur} + {wr:.3f} {uw}")
        except (ValueError, IndexError):
            pass


def _num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}


def write_json(path, out_json, kernel, source, shape, captured):
    import hashlib
    import json
    import os
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    sel = [r for r in rows[2:] if kernel in r[col["Kernel Name"]]]
    if not sel:
        raise SystemExit(f"no launch of a kernel matching {kernel!r} in {path}")

    def mean(metric, scale_units=True):
        i = col.get(metric)
        if i is None:
            return None
        vals = [_num(r[i]) for r in sel]
        vals = [v for v in vals if v is not None]
        if not vals:
            return None
        m = sum(vals) / len(vals)
        return m * UNIT.get(units[i], 1.0) if scale_units else m
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = {
        "kernel": sel[0][col["Kernel Name"]], "launches_captured": len(sel), "grid": sel[0][col["Grid Size"]], "block": sel[0][col["Block Size"]],
        "shape": shape, "captured": captured,
        "source": source, "source_sha16": hashlib.sha256(open(os.path.join(root, source), "rb").read()).hexdigest()[:16],
        "duration_us": mean("gpu__time_duration.sum"),
        "dram_bytes_read": mean("dram__bytes_read.sum"), "dram_bytes_write": mean("dram__bytes_write.sum"),
        "dram_throughput_pct": mean("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", False),
        "tensor_pipe_pct": mean("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", False),
        "sm_throughput_pct": mean("sm__throughput.avg.pct_of_peak_sustained_elapsed", False),
        "warps_active_pct": mean("sm__warps_active.avg.pct_of_peak_sustained_active", False),
        "registers_per_thread": mean("launch__registers_per_thread", False),
        "l2_hit_rate_pct": mean("lts__t_sector_hit_rate.pct", False),
        "note": "under ncu: cold cache, serialised, replayed ~40x; durations are for shares only, never bench values",
    }
    with open(out_json, "w") as f:
        json.dump(d, f, indent=1)
        f.write("\n")
    print(json.dumps(d))


if __name__ == "__main__":
    if "--json" in sys.argv:
        import argparse
        ap = argparse.ArgumentParser()
        ap.add_argument("rep")
        ap.add_argument("--json", required=True)
        ap.add_argument("--kernel", required=True)
        ap.add_argument("--source", required=True)
        ap.add_argument("--shape", required=True)
        ap.add_argument("--captured", default="")
        a = ap.parse_args()
        write_json(a.rep, a.json, a.kernel, a.source, a.shape, a.captured)
    else:
        main(sys.argv[1])
