"""Pack the shipped M-CSA template library (INPUT DATA) into one bundle.

The reference ships 7607 small template files under
``enzymm/jess_templates_20230210/`` (41 MB on disk, mostly block slack).  They are input
data for the matching hot path, not source code, so they are packed verbatim into one
gzip'd text bundle that travels with the repo to the GPU box:

    @@ <relative path>\n<file content>

Paths are sorted so the bundle (and hence template order) is deterministic; the reference's
own order is ``glob.glob`` order (``template.py:1491-1492``), i.e. filesystem dependent.

Run once in the build container:  python tools/pack_templates.py
"""
import gzip
import sys
from pathlib import Path

SRC = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/enzymm/jess_templates_20230210")
DST = Path(__file__).resolve().parent.parent / "enzymm_b200" / "data" / "jess_templates_20230210.bundle.gz"


def main() -> None:
    paths = sorted(SRC.glob("**/*.pdb"))
    with gzip.GzipFile(DST, "wb", compresslevel=9, mtime=0) as out:
        for p in paths:
            out.write(f"@@ {p.relative_to(SRC).as_posix()}\n".encode())
            text = p.read_text()
            if not text.endswith("\n"):
                text += "\n"
            out.write(text.encode())
    print(f"packed {len(paths)} templates -> {DST} ({DST.stat().st_size/1e6:.2f} MB)")


if __name__ == "__main__":
    main()
