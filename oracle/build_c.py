"""Builds the plain-C restatement of the integer / byte / index work (oracle/css_oracle_int.c) with gcc.
TEST INFRASTRUCTURE ONLY: tests/test_oracle_c.py and __graft_entry__.build() call this; nothing in css_b200/ does."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "css_oracle_int.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libcss_oracle_int.so")


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    gcc = shutil.which("gcc") or shutil.which("cc")
    if gcc is None:
        raise RuntimeError("gcc not found")
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = [gcc, "-O2", "-std=c99", "-Wall", "-Wextra", "-ffp-contract=off", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("gcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
