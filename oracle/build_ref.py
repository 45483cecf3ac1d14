"""TEST INFRASTRUCTURE — recipe that makes the UNMODIFIED reference runnable on the GPU box.

    python -m oracle.build_ref          # authoring container only (needs /root/reference)

The reference is pure Python, so "building" it means: import its sampling path through the stand-in packages of
oracle/ref_shim.py, list every module that import pulled from /root/reference, and copy exactly those files — byte for
byte, tree preserved — into ``oracle/_ref/reference/``.  That directory is git-ignored (no reference source enters the
history) but not gpurun-ignored, so it travels to the GPU box like a built ``.so``; there ``bench.py --impl reference``
times the reference's own ``DFoTVideo`` / ``DFoTVideoPose`` (``cpu_baseline.kind = "reference"``) and falls back to the
oracle port (``kind = "port"``) only when the directory is absent.  `__graft_entry__.build()` runs this recipe whenever
/root/reference exists.
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "oracle", "_ref", "reference")


def build(verbose: bool = True) -> str:
    if not os.path.isdir(os.path.join(SRC, "algorithms", "dfot")):
        raise RuntimeError(f"{SRC} is not present: oracle/_ref can only be populated in the authoring container")
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    os.environ["DFOT_REFERENCE_ROOT"] = SRC
    from oracle import ref_shim
    ref_shim.install()
    from algorithms.dfot.dfot_video import DFoTVideo  # noqa: F401
    from algorithms.dfot.dfot_video_pose import DFoTVideoPose  # noqa: F401
    files = sorted({m.__file__ for m in list(sys.modules.values())
                    if getattr(m, "__file__", None) and m.__file__.startswith(SRC + os.sep)})
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    manifest = {}
    for f in files:
        rel = os.path.relpath(f, SRC)
        out = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(f, out)
        with open(f, "rb") as fh:
            manifest[rel] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(os.path.dirname(DST), "MANIFEST.json"), "w") as fh:
        json.dump(dict(source=SRC, files=manifest), fh, indent=1)
    if verbose:
        print(f"oracle/_ref: {len(files)} reference files copied unmodified to {DST}")
    return DST


if __name__ == "__main__":
    build()
