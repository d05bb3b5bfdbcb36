#!/usr/bin/env python
"""Stage the reference's caller scripts for ONE gpurun call (build container only).

The drop-in contract is that the reference's `test_sr.py` / `test_w.py` run BYTE-UNMODIFIED with
`dropin/models` providing `models`.  The GPU box has no /root/reference, and reference sources are never committed, so
the files are copied into the git-ignored scratch directory `_staged_ref/` (which a gpurun snapshot carries), used by
`tests/test_dropin_scripts.py -m gpu`, and removed again (`--clean`) once the call is over:

    python tools/stage_reference_scripts.py            # copy + verify against tests/golden/reference_scripts_sha256.txt
    gpurun -- 'python -m pytest tests/test_dropin_scripts.py -m gpu -q ...'
    python tools/stage_reference_scripts.py --clean

`--write-hashes` (re)generates the committed hash list from /root/reference, which is how the GPU test proves that what it
ran was the unmodified reference script.
"""
import hashlib
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("MARCONET_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "_staged_ref")
HASHES = os.path.join(ROOT, "tests", "golden", "reference_scripts_sha256.txt")
FILES = ["test_sr.py", "test_w.py", "utils/alphabets.py", "utils/yolo_ocr_xloc.py"]


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def main():
    if "--clean" in sys.argv:
        shutil.rmtree(DST, ignore_errors=True)
        print("removed", DST)
        return
    if not os.path.isfile(os.path.join(REF, "test_sr.py")):
        raise SystemExit(f"reference tree not present at {REF}")
    if "--write-hashes" in sys.argv:
        with open(HASHES, "w") as f:
            for rel in FILES:
                f.write(f"{sha(os.path.join(REF, rel))}  {rel}\n")
        print("wrote", HASHES)
    want = dict(line.split()[::-1] for line in open(HASHES).read().splitlines() if line.strip())
    shutil.rmtree(DST, ignore_errors=True)
    for rel in FILES:
        os.makedirs(os.path.dirname(os.path.join(DST, rel)), exist_ok=True)
        shutil.copyfile(os.path.join(REF, rel), os.path.join(DST, rel))
        assert sha(os.path.join(DST, rel)) == want[rel], f"{rel}: hash differs from {HASHES}"
    print("staged", FILES, "->", DST)


if __name__ == "__main__":
    main()
