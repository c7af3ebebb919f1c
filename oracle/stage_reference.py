"""TEST / BENCH INFRASTRUCTURE ONLY -- stages the UNMODIFIED reference modules for the GPU box.

``/root/reference`` exists only in the build container. ``__graft_entry__.build()`` calls :func:`stage` there: the
reference's ``src/`` tree (pure Python, no build step) is packed, byte for byte, into ONE binary archive
``oracle/_ref/gpode_reference_src.zip`` (+ a manifest of SHA-256 sums). ``oracle/_ref/`` is git-ignored -- no reference
source enters the repository history -- but it is not gpurun-ignored, so the archive travels to the GPU box like the
built ``.so``. ``oracle/reference_harness.py`` imports the modules straight from the archive (zipimport; the reference
uses implicit namespace packages), with ``oracle/torchdiffeq_shim`` standing in for the absent torchdiffeq 0.2.0, and
``bench.py --impl reference`` / ``cpu_baseline`` then time the reference's own code (``kind: "reference"``) instead of
the oracle port. The product never reads the archive.
"""
import hashlib
import json
import os
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
ARCHIVE = os.path.join(REF_DIR, "gpode_reference_src.zip")
MANIFEST = os.path.join(REF_DIR, "MANIFEST.json")


def stage(reference_root="/root/reference", force=False):
    """Returns the archive path, or None when the reference tree is not present (GPU box: the prebuilt file is used)."""
    src = os.path.join(reference_root, "src")
    if not os.path.isdir(os.path.join(src, "core")):
        return ARCHIVE if os.path.exists(ARCHIVE) else None
    files = []
    for d, _, names in os.walk(src):
        for n in sorted(names):
            if n.endswith(".py"):
                files.append(os.path.join(d, n))
    files.sort()
    sums = {}
    for f in files:
        with open(f, "rb") as fh:
            sums[os.path.relpath(f, reference_root)] = hashlib.sha256(fh.read()).hexdigest()
    if not force and os.path.exists(ARCHIVE) and os.path.exists(MANIFEST):
        try:
            if json.load(open(MANIFEST)).get("sha256") == sums:
                return ARCHIVE
        except ValueError:
            pass
    os.makedirs(REF_DIR, exist_ok=True)
    with zipfile.ZipFile(ARCHIVE, "w", zipfile.ZIP_DEFLATED) as z:
        dirs = set()
        for f in files:   # explicit directory entries: zipimport needs them to see implicit namespace packages
            rel = os.path.dirname(os.path.relpath(f, reference_root))
            while rel and rel not in dirs:
                dirs.add(rel)
                rel = os.path.dirname(rel)
        for d in sorted(dirs):
            z.writestr(zipfile.ZipInfo(d + "/"), b"")
        for f in files:
            z.write(f, os.path.relpath(f, reference_root))
    with open(MANIFEST, "w") as fh:
        json.dump({"what": "unmodified src/**/*.py of hegdepashupati/gaussian-process-odes, packed by oracle/stage_reference.py",
                   "files": len(files), "sha256": sums}, fh, indent=1)
    return ARCHIVE


if __name__ == "__main__":
    print(stage(force=True))
