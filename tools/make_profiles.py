"""Turns the raw artefacts a gpurun call left in gpurun_out/ into the tracked summaries under profiles/:
  python tools/make_profiles.py r1 launches=<csv> rb_fwd=<ncu-rep> rb_bwd=<ncu-rep> wgrad=<ncu-rep> [vq=<ncu-rep>] ...
launches=<csv>   ncu --metrics gpu__time_duration.sum launch list of one eager train_step  -> <round>_launches.csv (copy)
                 + <round>_launch_summary.txt (per-kernel totals and shares)
<name>=<ncu-rep> ncu --set full capture of one kernel -> <round>_ncu_<name>.txt (counters + stall reasons) and an entry in
                 <round>_ncu.json {name: {duration_us, dram_read_bytes, dram_write_bytes, ...}} that bench.py reads for
                 roofline.traffic."""
import csv
import io
import json
import os
import shutil
import subprocess
import sys
from contextlib import redirect_stdout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ncu_summary  # noqa: E402
import summarize_launches  # noqa: E402

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3}


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]
    def val(k):
        try:
            return float(r[hdr.index(k)].replace(",", "")) * UNIT.get(units[hdr.index(k)], 1.0)
        except ValueError:  # "no data" / "n/a": the counter was not collected for this kernel
            return None
    d = {"kernel": r[hdr.index("Kernel Name")], "duration_us": val("gpu__time_duration.sum"),
         "dram_read_bytes": val("dram__bytes_read.sum"), "dram_write_bytes": val("dram__bytes_write.sum"),
         "dram_pct_of_peak": val("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
         "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
         "registers_per_thread": val("launch__registers_per_thread"), "grid": val("launch__grid_size"),
         "block": val("launch__block_size")}
    for k in hdr:
        if k.endswith("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"):
            d["tensor_pipe_active_realtime_pct"] = val(k)
    for k, name in (("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_pipe_active_pct"),
                    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsu_data_pipe_pct"),
                    ("sm__cycles_elapsed.max", "sm_cycles")):
        if k in hdr:
            d[name] = val(k)
    d["traffic_bytes"] = (d["dram_read_bytes"] or 0.0) + (d["dram_write_bytes"] or 0.0)
    return d


def main():
    rnd = sys.argv[1]
    prof = os.path.join(ROOT, "profiles")
    os.makedirs(prof, exist_ok=True)
    jpath = os.path.join(prof, f"{rnd}_ncu.json")
    js = json.load(open(jpath)) if os.path.exists(jpath) else {}
    for arg in sys.argv[2:]:
        name, path = arg.split("=", 1)
        if name.startswith("launches"):
            shutil.copy(path, os.path.join(prof, f"{rnd}_{name}.csv"))
            buf = io.StringIO()
            with redirect_stdout(buf):
                summarize_launches.main(path, 40)
            open(os.path.join(prof, f"{rnd}_{name}_summary.txt"), "w").write(buf.getvalue())
        else:
            buf = io.StringIO()
            with redirect_stdout(buf):
                ncu_summary.main(path)
            open(os.path.join(prof, f"{rnd}_ncu_{name}.txt"), "w").write(buf.getvalue())
            js[name] = raw(path)
    json.dump(js, open(jpath, "w"), indent=1)
    print("wrote", sorted(os.listdir(prof)))


if __name__ == "__main__":
    main()
