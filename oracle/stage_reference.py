"""Stage the UNMODIFIED Python reference for the GPU box.  TEST / BENCH INFRASTRUCTURE ONLY.

`/root/reference` exists only in the dev container.  bench.py's CPU arm times the reference's own Python path
(BASELINE.md section 3), so the package has to travel with the `gpurun` snapshot: this recipe copies
`/root/reference/marlgrid/**/*.py` byte for byte into the git-ignored directory `oracle/_ref/marlgrid/` (listed in
.gitignore, NOT in .gpurunignore -- exactly like the built .so files) and writes a manifest with the sha256 of every file,
which `oracle.reference_bench` re-checks before it times anything.  Nothing is edited, nothing enters git history.

    python -m oracle.stage_reference            # run by __graft_entry__.build() when /root/reference is present
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("MARLGRID_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def stage(verbose=True):
    pkg = os.path.join(SRC, "marlgrid")
    if not os.path.isdir(pkg):
        if verbose:
            print(f"stage_reference: {pkg} not present (GPU box?) -- keeping whatever is staged in {DST}")
        return os.path.isdir(os.path.join(DST, "marlgrid"))
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    manifest = {}
    for root, _dirs, files in os.walk(pkg):
        for fn in sorted(files):
            if not fn.endswith(".py"):
                continue
            src = os.path.join(root, fn)
            rel = os.path.relpath(src, SRC)
            dst = os.path.join(DST, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
            manifest[rel] = sha256(dst)
    head = None
    try:  # the reference commit, for the record
        with open(os.path.join(SRC, ".git", "HEAD")) as f:
            head = f.read().strip()
        if head.startswith("ref:"):
            with open(os.path.join(SRC, ".git", head.split()[1])) as f:
                head = f.read().strip()
    except Exception:  # noqa: BLE001
        pass
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SRC, "commit": head, "files": manifest}, f, indent=1, sort_keys=True)
    if verbose:
        print(f"stage_reference: {len(manifest)} files of the unmodified reference -> {DST} (git-ignored)")
    return True


def verify():
    """True iff oracle/_ref holds exactly the files of its manifest (unmodified since staging)."""
    mf = os.path.join(DST, "MANIFEST.json")
    if not os.path.exists(mf):
        return False
    with open(mf) as f:
        m = json.load(f)
    return all(os.path.exists(os.path.join(DST, rel)) and sha256(os.path.join(DST, rel)) == h for rel, h in m["files"].items())


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
