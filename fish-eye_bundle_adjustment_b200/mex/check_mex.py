"""Type-check the MEX gateway against include/feba.h with a stub mex.h (no MATLAB in the build
container): ``gcc -fsyntax-only``.  Run by __graft_entry__.build()."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def main() -> int:
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
    if not cc:
        print("check_mex: no C compiler, skipped")
        return 0
    cmd = [cc, "-std=c11", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(HERE, "stub"), os.path.join(HERE, "feba_mex.c")]
    return subprocess.run(cmd).returncode


if __name__ == "__main__":
    sys.exit(main())
