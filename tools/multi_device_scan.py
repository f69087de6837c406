"""The one-process whole-box call on real GPUs: ``Matcher.scan_files(paths, devices=[0..N-1])`` and
``scan_to_tsv(..., devices=...)`` over N GPUs of this host vs the same scan on one GPU.
usage: python tools/multi_device_scan.py [n_gpus] [n_files]"""
import io
import os
import shutil
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bench import active_templates, make_workload  # noqa: E402
from enzymm_b200 import jess_run  # noqa: E402


def main():
    n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    n_files = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    workload = make_workload(0, 1024, 400, 1, 8)
    root = tempfile.mkdtemp(prefix="emm_multi_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        texts = [workload.to_pdb(i).encode() for i in range(1024)]
        paths = []
        for i in range(n_files):
            path = os.path.join(root, f"s{i:06d}.pdb")
            with open(path, "wb") as handle:
                handle.write(texts[i % 1024])
            paths.append(path)
        matcher = jess_run.Matcher(templates=active_templates())
        results = {}
        for label, devices in (("1 GPU", None), (f"{n_gpus} GPUs", list(range(n_gpus)))):
            for rep in range(2):
                t0 = time.perf_counter()
                chunks = list(matcher.scan_files(paths, chunk_size=1024, devices=devices))
                dt = time.perf_counter() - t0
            results[label] = chunks
            print(f"scan_files {label}: {n_files / dt:.0f} files/s, {sum(len(r) for _, _, r in chunks)} hits (second call)")
        a, b = results["1 GPU"], results[f"{n_gpus} GPUs"]
        same = len(a) == len(b) and all(x[0] == y[0] and x[2].tobytes() == y[2].tobytes() for x, y in zip(a, b))
        print("chunks in input order with identical hit records:", same)
        tables = {}
        for label, devices in (("1 GPU", None), (f"{n_gpus} GPUs", list(range(n_gpus)))):
            for rep in range(2):
                sink = io.StringIO()
                t0 = time.perf_counter()
                rows = matcher.scan_to_tsv(paths, sink, chunk_size=1024, devices=devices)
                dt = time.perf_counter() - t0
            tables[label] = sink.getvalue()
            print(f"scan_to_tsv {label}: {n_files / dt:.0f} files/s, {rows} rows (second call)")
        print("identical tables:", tables["1 GPU"] == tables[f"{n_gpus} GPUs"])
        matcher.close()
        assert same and tables["1 GPU"] == tables[f"{n_gpus} GPUs"]
    finally:
        shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    main()
