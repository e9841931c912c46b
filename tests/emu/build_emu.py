"""Build the host-emulation library of the kernel sources (TEST INFRASTRUCTURE ONLY).

g++ -DDONK_EMU compiles donk_b200/csrc/*.cuh as sequential host code (see common.cuh) into
tests/emu/libdonk_emu.so.  Only tests load it.
"""
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
SRC = HERE / "donk_emu.cpp"
OUT = HERE / "libdonk_emu.so"


def build(force: bool = False) -> Path:
    deps = [SRC] + sorted((ROOT / "donk_b200" / "csrc").glob("*.cuh"))
    if not force and OUT.exists() and all(OUT.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return OUT
    cmd = ["g++", "-O2", "-std=c++17", "-DDONK_EMU", "-mfma", "-ffp-contract=off", "-fPIC", "-shared",
           "-x", "c++", str(SRC), "-o", str(OUT)]
    subprocess.run(cmd, check=True, cwd=ROOT)
    return OUT


def backend():
    from donk_b200._lib import Native

    return Native(build(), "donk_emu_", "cpu")


if __name__ == "__main__":
    print(build(force=True))
