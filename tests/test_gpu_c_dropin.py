"""The drop-in flow from a plain C program (tests/c/dropin_probit.c): compiled against
include/apop_b200.h and the host mirror's header, linked with both shared libraries, run on
the B200.  This is the reference-side binding of INTEGRATION.md exercised without Python."""
import os
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_c_program_registers_fits_and_bootstraps():
    pkg = os.path.join(ROOT, "apophenia_b200")
    exe = os.path.join(tempfile.mkdtemp(), "dropin_probit")
    subprocess.check_call(["gcc", "-O2", "-std=gnu11", "-o", exe, os.path.join(ROOT, "tests", "c", "dropin_probit.c"),
                           f"-L{pkg}", "-lapop_hostmini", "-lapop_b200", "-lm", f"-Wl,-rpath,{pkg}"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(out.stdout, out.stderr)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "OK" in out.stdout and "status 0" in out.stdout
