"""Stage the UNMODIFIED reference as ONE archive, ``oracle/_ref/reference.zip`` — TEST INFRASTRUCTURE.

The reference is a pure-Python package: there is nothing to compile, so its "build" for the purposes of the
oracle is an importable archive of the sources where they lie under ``/root/reference`` (zipimport reads
packages straight from a zip).  The archive is a build output like ``oracle/_build/libcc_oracle.so``:
git-ignored (no reference source ever enters this repository's history), but it travels to the GPU box with
``gpurun`` snapshots, where

* ``bench.py --impl reference`` then times the reference's own ``CollectiveCrossingEnv.step`` +
  ``GreedyPolicy`` loop (``cpu_baseline.kind = "reference"``; without the archive: the Python port),
* ``tests/test_reference_suite.py`` runs the reference's own test files against the drop-in façade, and
* ``tests/test_oracle_vs_reference.py`` pins the oracle against the reference itself.

Run by ``__graft_entry__.build()`` whenever ``/root/reference`` is present:  ``python -m oracle.stage_ref``.
"""

from __future__ import annotations

import hashlib
import json
import zipfile
from pathlib import Path

REFERENCE = Path("/root/reference")
OUT_DIR = Path(__file__).resolve().parent / "_ref"
ARCHIVE = OUT_DIR / "reference.zip"
MANIFEST = OUT_DIR / "MANIFEST.json"
# what the hot path, its callers and its tests consist of; docs, CI and tooling stay behind
INCLUDE = ("src/collectivecrossing", "src/baseline_policies", "tests", "scripts/run_greedy_policy_demo.py",
           "scripts/run_waiting_policy_demo.py", "examples/training_script.py", "pyproject.toml", "LICENSE", "README.md")
SUFFIXES = {".py", ".json", ".toml", ".md", ""}


def _files():
    for rel in INCLUDE:
        p = REFERENCE / rel
        if p.is_file():
            yield p
        elif p.is_dir():
            for q in sorted(p.rglob("*")):
                if q.is_file() and "__pycache__" not in q.parts and q.suffix in SUFFIXES:
                    yield q


def stage(force: bool = False) -> Path | None:
    """Write the archive (deterministic: sorted entries, fixed timestamps).  Returns its path, or None when
    the reference is not mounted (the GPU box: the archive that travelled with the snapshot is used as is)."""
    if not (REFERENCE / "src" / "collectivecrossing" / "collectivecrossing.py").exists():
        return ARCHIVE if ARCHIVE.exists() else None
    files = list(_files())
    digest = {str(f.relative_to(REFERENCE)): hashlib.sha256(f.read_bytes()).hexdigest() for f in files}
    if not force and ARCHIVE.exists() and MANIFEST.exists():
        try:
            if json.loads(MANIFEST.read_text())["sha256"] == digest:
                return ARCHIVE
        except Exception:  # noqa: BLE001 - stale manifest: rewrite
            pass
    OUT_DIR.mkdir(parents=True, exist_ok=True)
    with zipfile.ZipFile(ARCHIVE, "w", zipfile.ZIP_DEFLATED) as z:
        # explicit directory entries: zipimport finds namespace packages (collectivecrossing/utils has no __init__.py)
        # only through them
        dirs = sorted({str(Path(*f.relative_to(REFERENCE).parts[:k])) + "/" for f in files for k in range(1, len(f.relative_to(REFERENCE).parts))})
        for d in dirs:
            info = zipfile.ZipInfo(d, date_time=(2020, 1, 1, 0, 0, 0))
            info.external_attr = (0o755 << 16) | 0x10
            z.writestr(info, b"")
        for f in files:
            info = zipfile.ZipInfo(str(f.relative_to(REFERENCE)), date_time=(2020, 1, 1, 0, 0, 0))
            info.compress_type = zipfile.ZIP_DEFLATED
            info.external_attr = 0o644 << 16
            z.writestr(info, f.read_bytes())
    MANIFEST.write_text(json.dumps({"source": str(REFERENCE), "files": len(files), "sha256": digest}, indent=1))
    return ARCHIVE


if __name__ == "__main__":
    out = stage(force=True)
    print(f"staged: {out}" if out else "reference not mounted and no archive present")
