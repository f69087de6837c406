"""A/B of the two search-kernel instantiations on the bench batch, with a bit-for-bit comparison.

A launch that does not split pairs (``donate_after < 0``: the default from 8 192 structures up) runs
``emm_search_kernel<..., kDonate=false>``, the instantiation without the split-pair work loop; every
other launch runs the split-capable one the parity tests cover.  This tool searches ONE uploaded
batch with both (``EMM_SPLIT_CAPABLE=1`` forces the capable instantiation on an unsplit launch) and
with splitting on, checks that the hit records are identical byte for byte, and prints the kernel
times.  Then the small-launch cases: the two fixtures staged and a 4-chain assembly searched in
place, ``donate_after=-1`` (plain instantiation) against ``donate_after=1`` (split at every chance).

Meaningful with ``tools/patches/plain_instantiation.patch`` applied (``git apply`` it, ``make -C
enzymm_b200/csrc``): the shipped kernel has one instantiation and ignores ``EMM_SPLIT_CAPABLE``.  The
result of the round-2 run is ``profiles/r02_plain_instantiation.txt``.

usage: python tools/gpu_plain_check.py [n_structures=10000] [out=gpurun_out/r02_plain_check.txt]
"""
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
T0 = time.perf_counter()

from bench import DEFAULT_DIST, active_templates, make_workload  # noqa: E402
from enzymm_b200.engine import Engine  # noqa: E402
from enzymm_b200.library import CompiledLibrary  # noqa: E402
from enzymm_b200.packing import pack_molecules  # noqa: E402
from enzymm_b200.structures import Molecule  # noqa: E402
from enzymm_b200.synth import SynthConfig, generate_chunk  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    out_path = Path(sys.argv[2]) if len(sys.argv) > 2 else ROOT / "gpurun_out" / "r02_plain_check.txt"
    out_path.parent.mkdir(parents=True, exist_ok=True)
    out = open(out_path, "w")

    def say(*a):
        line = " ".join(str(x) for x in a)
        print(line, flush=True)
        out.write(line + "\n")
        out.flush()

    ok = True
    templates = active_templates()
    dists = [DEFAULT_DIST[min(t.effective_size, 8)] for t in templates]
    engine = Engine(CompiledLibrary(templates, 2.0, dists, dists))
    say(f"[{time.perf_counter() - T0:.1f} s] library on device: {len(templates)} templates")

    # small launches first (cheap): fixtures staged, one assembly in place
    golden = ROOT / "tests" / "golden"
    mols = [Molecule.load(golden / "1AMY.pdb"), Molecule.load(golden / "AF-P0DUB6-F1-model_v4.pdb")]
    big = generate_chunk(0, SynthConfig(n_chains=4), templates, 1).to_molecule(0)
    for name, batch, want in (("fixtures (staged)", pack_molecules(mols, engine.compiled), 24),
                              ("4-chain assembly (in place)", pack_molecules([big], engine.compiled), None)):
        plain = engine.query(batch, donate_after=-1)
        split = engine.query(batch, donate_after=1)
        same = plain.tobytes() == split.tobytes() and (want is None or len(plain) == want)
        ok = ok and same
        say(f"{name}: plain instantiation {len(plain)} hits, split at every chance {len(split)} hits, "
            f"identical records: {same}")
    say(f"[{time.perf_counter() - T0:.1f} s] small launches done")

    workload = make_workload(0, n, 400, 1, 8)
    batch = workload.to_packed(engine.compiled)
    sess = engine.session_for(batch.n_atoms, batch.n_structures)
    sess.upload(batch)
    say(f"[{time.perf_counter() - T0:.1f} s] {n} structures uploaded")
    runs = [("plain instantiation, no splitting (new default at this size)", "0", -1),
            ("split-capable instantiation, no splitting (previous default)", "1", -1),
            ("plain instantiation again", "0", -1),
            ("split-capable instantiation again", "1", -1),
            ("split-capable instantiation, splitting on (donate_after=48)", "0", 48)]
    first = None
    for name, capable, donate in runs:
        os.environ["EMM_SPLIT_CAPABLE"] = capable
        sess.clear_timings()
        sess.run(force_prepare=True, donate_after=donate)
        hits = sess.download().copy()
        ms = sess.kernel_ms("search")
        if first is None:
            first = hits
        same = hits.tobytes() == first.tobytes()
        ok = ok and same
        say(f"{name}: search kernel {sum(ms):.2f} ms, {len(hits)} hits, identical to the first run: {same}")
    os.environ["EMM_SPLIT_CAPABLE"] = "0"
    say(f"[{time.perf_counter() - T0:.1f} s] RESULT: {'OK' if ok else 'MISMATCH'}")
    out.close()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
