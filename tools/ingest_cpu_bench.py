"""Host-only ingest rates (no GPU): text readers per format, packed corpus write / read.  Numbers from the
BUILD CONTAINER (8 cores) are in profiles/r02_ingest_cpu.txt; the GPU box has 16-32 host threads.
usage: python tools/ingest_cpu_bench.py [n_files=512]"""
import ctypes
import os
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]

from bench import DEFAULT_DIST, active_templates, make_workload  # noqa: E402
from enzymm_b200.library import CompiledLibrary  # noqa: E402
from enzymm_b200.packing import pack_files, read_corpus, write_corpus  # noqa: E402
from enzymm_b200.structures import Molecule, _native_lib  # noqa: E402


def best(fn, reps=5):
    out = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        out = min(out, time.perf_counter() - t0)
    return out


def main():
    from test_cif_ingest import to_cif
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    threads = len(os.sched_getaffinity(0))
    workload = make_workload(0, min(n, 256), 400, 1, threads)
    root = Path(tempfile.mkdtemp(prefix="emm_ingest_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None))
    pdb, cif = [], []
    for i in range(n):
        p = root / f"s{i:05d}.pdb"
        p.write_text(workload.to_pdb(i % workload.n_structures))
        pdb.append(str(p))
    for i in range(min(n, 128)):                    # the Python CIF writer of the tests is slow: fewer files
        c = root / f"s{i:05d}.cif"
        c.write_text(to_cif(Molecule.load(pdb[i]), f"s{i}"))
        cif.append(str(c))
    lib = _native_lib()
    lib.emm_pdb_batch_free.argtypes = [ctypes.c_void_p]
    print(f"{n} PDB files ({os.path.getsize(pdb[0]) / 1e3:.0f} KB each), {len(cif)} mmCIF renderings; {threads} host threads")
    for label, paths in (("PDB", pdb), ("mmCIF", cif)):
        arr = (ctypes.c_char_p * len(paths))(*[p.encode() for p in paths])
        for fn in ("emm_pdb_pack_files", "emm_pdb_load_files"):
            for nt in (1, threads):
                def call():
                    h = ctypes.c_void_p()
                    assert getattr(lib, fn)(arr, ctypes.c_int32(len(paths)), ctypes.c_int32(nt), ctypes.byref(h)) == 0
                    lib.emm_pdb_batch_free(h)
                dt = best(call)
                print(f"  {label:5s} {fn:20s} {nt:2d} thread(s): {len(paths) / dt:8.0f} files/s ({dt / len(paths) * 1e6:6.0f} us/file)")
    templates = active_templates()
    dist = [DEFAULT_DIST[min(t.effective_size, 8)] for t in templates]
    compiled = CompiledLibrary(templates, 2.0, dist, dist)
    corpus = root / "all.emmpack"
    dt = best(lambda: write_corpus(pdb, corpus), 3)
    print(f"  write_corpus ({threads} threads): {n / dt:.0f} files/s, {os.path.getsize(corpus) / n / 1e3:.0f} KB per structure")
    dt = best(lambda: read_corpus(corpus, compiled), 5)
    print(f"  read_corpus  (1 thread; map + classify kinds + expand the class column): {n / dt:.0f} structures/s")
    dt = best(lambda: pack_files(pdb, compiled), 3)
    print(f"  pack_files   ({threads} threads; parse + classify): {n / dt:.0f} files/s")
    for p in pdb + cif + [str(corpus)]:
        os.unlink(p)
    os.rmdir(root)


if __name__ == "__main__":
    main()
