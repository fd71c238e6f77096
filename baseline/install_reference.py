"""Place the reference repository under ``baseline/_ref`` so that it travels to the GPU box.

The upstream tree is a plain directory of Python modules (no setup.py / pyproject), so "installing" it is a copy:
    python baseline/install_reference.py [/root/reference]
``baseline/_ref`` is git-ignored (reference sources never enter this repository's history) but not
gpurun-ignored.  Images, notebooks and datasets are left out; nothing is edited.
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
KEEP = ("configs", "models", "optimizer", "transforms", "util", "tools", "datasets", "main.py", "test.py", "inference.py",
        "requirements.txt", "LICENSE", "README.md")


def install(src: str = "/root/reference", force: bool = False) -> str:
    if os.path.isdir(os.path.join(DEST, "models")) and not force:
        return DEST
    if not os.path.isdir(os.path.join(src, "models")):
        raise RuntimeError(f"reference tree not found at {src}")
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST)
    for name in KEEP:
        s = os.path.join(src, name)
        if os.path.isdir(s):
            shutil.copytree(s, os.path.join(DEST, name), ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.jpg", "*.png"))
        elif os.path.exists(s):
            shutil.copy2(s, os.path.join(DEST, name))
    return DEST


if __name__ == "__main__":
    print(install(sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "/root/reference", force="--force" in sys.argv))
