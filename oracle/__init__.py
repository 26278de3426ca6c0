"""CPU oracle loader — TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package; rend3_b200/ never does."""
import ctypes
import os
import subprocess

_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_DIR, "libr3_oracle.so")


def build(force: bool = False) -> str:
    srcs = [os.path.join(_DIR, f) for f in ("r3_oracle.c", "r3_oracle_forward.inc", "r3_oracle.h", "Makefile")]
    srcs += [os.path.join(_DIR, "..", "include", f) for f in ("r3_layouts.h", "rend3_b200.h")]
    stale = force or not os.path.exists(LIB_PATH) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs
    )
    if stale and os.path.exists(os.path.join(_DIR, "Makefile")):
        try:
            subprocess.run(["make", "-C", _DIR, "-B", "libr3_oracle.so"], check=True, capture_output=True, text=True)
        except (subprocess.CalledProcessError, FileNotFoundError) as e:  # keep a prebuilt .so usable on boxes without gcc
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(f"oracle build failed: {getattr(e, 'stderr', e)}")
    return LIB_PATH


def load_oracle_backend():
    """A rend3_b200.backend.Backend bound to the oracle's r3o_* entry points."""
    from rend3_b200.backend import Backend

    return Backend(ctypes.CDLL(build()), "r3o_", 0)


def set_threads(n: int) -> int:
    """Pin the oracle's OpenMP thread count (torchrun exports OMP_NUM_THREADS=1); returns the count in effect."""
    lib = ctypes.CDLL(build())
    lib.r3o_set_threads(int(n))
    return int(lib.r3o_get_threads())
