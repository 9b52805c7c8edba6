"""Puts the UNMODIFIED reference package under baseline/_ref (git-ignored; it travels to the GPU box with gpurun).

    python scripts/install_reference.py

First choice is the contract's `pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps --target
baseline/_ref <copy of /root/reference>`. In this image that fails: the reference's build backend (`hatchling`,
pyproject.toml:1-3) is not installed and there is no index to fetch it from. lotus-data/lotus is a pure-Python package whose
wheel is nothing but its `lotus/` directory, so the fallback places exactly that directory under baseline/_ref — the result a
successful wheel install would have produced. Nothing under baseline/_ref is tracked by git and nothing in the product imports it;
it is used by (a) tests/test_gpu_reference_dropin.py — the reference's own operators over B200VS — and (b) bench.py's
operator-scope reference leg."""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("LOTUS_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def main() -> int:
    if not os.path.isdir(os.path.join(SRC, "lotus")):
        print(f"{SRC} is not here (GPU box?): keeping whatever baseline/_ref already holds")
        return 0
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST, exist_ok=True)
    tmp = tempfile.mkdtemp()
    work = os.path.join(tmp, "ref")
    shutil.copytree(SRC, work, ignore=shutil.ignore_patterns(".git", "docs", "assets", "examples"))
    r = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--find-links", "/opt/wheelhouse",
                        "--no-deps", "--target", DST, work], capture_output=True, text=True)
    how = "pip"
    if r.returncode != 0 or not os.path.isdir(os.path.join(DST, "lotus")):
        how = "package directory (pip failed: " + (r.stderr.strip().splitlines()[-1] if r.stderr.strip() else "?") + ")"
        shutil.copytree(os.path.join(SRC, "lotus"), os.path.join(DST, "lotus"))
    with open(os.path.join(DST, "INSTALLED_FROM.txt"), "w") as f:
        f.write(f"{SRC} via {how}\n")
    shutil.rmtree(tmp, ignore_errors=True)
    print(f"reference installed under baseline/_ref via {how}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
