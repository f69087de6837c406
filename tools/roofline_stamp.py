"""Write profiles/roofline_traffic.json from an ``ncu --set full`` capture of the search kernel and stamp
it with the SHA-256 of the kernel sources it was captured from (bench.py refuses stale counters).
usage: python tools/roofline_stamp.py report.ncu-rep n_structures n_templates "<how it was captured>" """
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bench import kernel_stamp  # noqa: E402


def main():
    rep, n_structures, n_templates, how = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    d = dict(zip(rows[0], rows[2]))
    units = dict(zip(rows[0], rows[1]))

    def to_bytes(key):
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[key]]
        return float(d[key].replace(",", "")) * scale

    dram = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
    inst = float(d["smsp__inst_executed.sum"].replace(",", ""))
    pairs = n_structures * n_templates
    out = {
        "source": how,
        "kernel": d.get("Kernel Name", "?"),
        "kernel_sha256": kernel_stamp(),
        "search_kernel_dram_bytes_per_launch": int(dram),
        "structures_in_capture": n_structures,
        "search_kernel_dram_bytes_per_structure": int(dram / n_structures),
        "search_kernel_warp_instructions_per_launch": int(inst),
        "pairs_in_capture": pairs,
        "search_kernel_warp_instructions_per_pair": round(inst / pairs, 1),
    }
    (ROOT / "profiles" / "roofline_traffic.json").write_text(json.dumps(out, indent=1) + "\n")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
