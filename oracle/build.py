"""ORACLE — TEST INFRASTRUCTURE ONLY.  Builds oracle/liboracle.so (the C restatement, oracle/flat_ip_c.c) with
the Makefile next to it.  Called by __graft_entry__.build(), tests/conftest.py and oracle/c_oracle.py; never
by the product package."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def build_oracle(force: bool = False) -> None:
    res = subprocess.run(["make", "-C", HERE] + (["-B"] if force else []), capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + res.stdout + res.stderr)


if __name__ == "__main__":
    build_oracle(force=True)
