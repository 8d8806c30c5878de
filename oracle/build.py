"""Build recipe for the C restatement (oracle/sgbm_ref.c -> oracle/_build/libsgbm_ref.so).

TEST INFRASTRUCTURE ONLY.  The reference has no compiled sources of its own (pure Python over cv2),
so there is no `oracle/_ref` to compile; the reference's real CPU path is the installed `cv2`.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "sgbm_ref.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libsgbm_ref.so")


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    subprocess.check_call(["gcc", "-O2", "-Wall", "-shared", "-fPIC", "-o", OUT, SRC])
    return OUT


if __name__ == "__main__":
    print(build(force=True))
