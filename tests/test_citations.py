"""Every `file.py:line` citation of the reference in the header, the sources and the docs must name a file that exists
under /root/reference with at least that many lines.  Runs only where the reference is mounted (the build container)."""
import collections
import glob
import os
import re

import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = re.compile(r"([A-Za-z0-9_./…-]*[A-Za-z0-9_]+\.(?:py|java|xml|sd))\s*:\s*(\d+(?:[-–]\d+)?(?:\s*,\s*:?\d+(?:[-–]\d+)?)*)")


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted here")
def test_reference_citations_resolve():
    by_name = collections.defaultdict(list)
    for root, _, files in os.walk(REF):
        if "/.git" in root:
            continue
        for f in files:
            by_name[f].append(os.path.join(root, f))
    lines = {}

    def nlines(p):
        if p not in lines:
            with open(p, errors="ignore") as fh:
                lines[p] = sum(1 for _ in fh)
        return lines[p]

    sources = glob.glob(f"{ROOT}/include/*.h") + glob.glob(f"{ROOT}/marqo_b200/**/*.py", recursive=True) + \
        glob.glob(f"{ROOT}/marqo_b200/csrc/*.cu*") + glob.glob(f"{ROOT}/oracle/*.py") + glob.glob(f"{ROOT}/oracle/*.c") + \
        [f"{ROOT}/DESIGN.md", f"{ROOT}/INTEGRATION.md", f"{ROOT}/README.md"]
    checked, bad = 0, []
    for src in sources:
        with open(src, errors="ignore") as fh:
            text = fh.read()
        for m in PAT.finditer(text):
            path = m.group(1).replace("…/", "").replace("…", "").lstrip("./")
            base = os.path.basename(path)
            cands = by_name.get(base)
            if not cands:
                if glob.glob(f"{ROOT}/**/{base}", recursive=True):
                    continue                                  # a citation of this repository's own file
                bad.append((os.path.relpath(src, ROOT), m.group(0), "no such file in the reference"))
                continue
            narrowed = [c for c in cands if c.endswith(path)] or cands
            top = max(int(x) for x in re.findall(r"\d+", m.group(2)))
            checked += 1
            if not any(top <= nlines(c) for c in narrowed):
                bad.append((os.path.relpath(src, ROOT), m.group(0), f"line {top} past the end"))
    assert checked > 100 and not bad, bad[:20]
