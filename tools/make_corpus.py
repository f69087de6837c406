"""Parse and pack structure files once into packed corpus files for repeated screening.

    python tools/make_corpus.py OUT_PREFIX [--per-file 65536] [--threads N] [--skip-bad] [--use-author] FILE_OR_LIST ...

Arguments ending in ``.txt`` / ``.list`` are read as lists of paths (one per line), everything else is a
structure file (PDB or mmCIF, gzip-compressed or not).  Writes ``OUT_PREFIX.00000.emmpack``, ... with at most
``--per-file`` structures each (a 400-residue structure takes ~120 KB); screen them with
``Matcher.scan_files(sorted(glob("OUT_PREFIX.*.emmpack")))`` or ``scan_to_tsv`` -- no text is parsed again,
and another template library or other thresholds need no new corpus (``enzymm_b200/packing.py``)."""
import argparse
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("out_prefix")
    ap.add_argument("inputs", nargs="+")
    ap.add_argument("--per-file", type=int, default=65536)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--skip-bad", action="store_true", help="warn about unreadable files instead of stopping")
    ap.add_argument("--use-author", action="store_true", help="mmCIF: auth_* identifiers instead of label_*")
    args = ap.parse_args()
    from enzymm_b200.packing import CORPUS_SUFFIX, _query_ids, write_corpus
    paths = []
    for item in args.inputs:
        if item.endswith((".txt", ".list")):
            paths += [line.strip() for line in Path(item).read_text().splitlines() if line.strip()]
        else:
            paths.append(item)
    ids = _query_ids(paths)                      # numbered over the WHOLE input, as load_molecules would
    t0 = time.perf_counter()
    written = 0
    for k, lo in enumerate(range(0, len(paths), args.per_file)):
        out = f"{args.out_prefix}.{k:05d}{CORPUS_SUFFIX}"
        written += write_corpus(paths[lo:lo + args.per_file], out, threads=args.threads, use_author=args.use_author,
                                on_error="skip" if args.skip_bad else "raise", ids=ids[lo:lo + args.per_file])
        print(out, file=sys.stderr)
    dt = time.perf_counter() - t0
    print(f"{written} structures in {dt:.1f} s ({written / max(dt, 1e-9):.0f} files/s)", file=sys.stderr)


if __name__ == "__main__":
    main()
